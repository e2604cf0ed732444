"""Public Tacotron2 API shell (drop-in for models/tacotron2/networks.py).

  needs_postprocessing / truncate_mel / resize_mel     networks.py:39-67
  Tacotron2(checkpoint, n_symbol, decoder_max_step, arabic_in, vowelizer)   :71-253
  Tacotron2Wave(model_sd_path, vocoder_sd, vocoder_config, vowelizer, arabic_in, n_symbol)  :256-426
The acoustic model runs through the C ABI (tacotron2_ms.Tacotron2MS.infer); the alignment-based mel
truncation and the bicubic speed resize of a whole batch are ONE kernel launch and one host read of the
new lengths (ttsb_tacotron2_postprocess) instead of the reference's per-utterance Python loop with a host
sync each; `truncate_mel` / `resize_mel` remain as the module-level functions the
reference exports. The vocoder runs once per batch over the length-masked padded mels.
"""
from typing import List, Optional, Union

import torch
import torch.nn as nn

from ... import _lib, text
from ...text.symbols import EOS_TOKENS, SEPARATOR_TOKEN
from ...utils import get_basic_config
from ...vocoder import load_hifigan
from ...vocoder.hifigan.denoiser import Denoiser
from ..fastpitch.networks import _load_vowelizer, text_collate_fn
from .tacotron2_ms import Tacotron2MS

_NO_POSTPROCESS = ('a', 'i', 'u', 'aa', 'ii', 'uu', 'n', 'm', 'h')


def needs_postprocessing(token: str):
    return token not in _NO_POSTPROCESS


def truncate_mel(mel_spec: torch.Tensor, ps_end):
    """Cut the mel where the attention on the inserted separator first reaches 80 % of its maximum,
    then repeat the last frame three times (networks.py:44-49)."""
    hit = (ps_end >= 0.8 * ps_end.max()).nonzero()
    n_end = int(hit[0]) if hit.numel() else mel_spec.shape[1]
    return torch.nn.functional.pad(mel_spec[:, :n_end], (0, 3), mode='replicate')


def resize_mel(mel: torch.Tensor, rate: Union[int, float] = 1.0, mode: str = 'bicubic'):
    n_f, n_t = mel.shape[-2:]
    n_new = int(1 / rate * n_t)
    if n_new == n_t:
        return mel
    return torch.nn.functional.interpolate(mel[None, None, ...], (n_f, n_new), mode=mode)[0, 0]


def postprocess_batch(mel: torch.Tensor, mel_lens, align: Optional[torch.Tensor], cols: Optional[List[int]],
                      speed: Union[int, float, None]) -> List[torch.Tensor]:
    """mel [B,F,T] (+ lengths, alignments [B,T,L]) -> list of [F,T_b'] mels: utterance b is truncated on alignment
    column cols[b] (< 0 or None: not truncated) and then resized by `speed` (None: not resized), exactly as
    `truncate_mel` then `resize_mel`. On a CUDA tensor: one launch for the batch + one host read of the new lengths."""
    B, F, T = mel.shape
    lens = [int(x) for x in (mel_lens.tolist() if torch.is_tensor(mel_lens) else mel_lens)]
    if speed is None and (cols is None or all(c < 0 for c in cols)):
        return [mel[b, :, :lens[b]] for b in range(B)]
    if not mel.is_cuda:
        # module-level reference functions on host tensors (the wrapper unit tests drive them with a CPU stand-in model)
        out = []
        for b in range(B):
            m = mel[b, :, :lens[b]]
            if cols is not None and cols[b] >= 0:
                m = truncate_mel(m, align[b, :lens[b], cols[b]])
            if speed is not None:
                m = resize_mel(m, rate=speed)
            out.append(m)
        return out
    lib = _lib.load()
    dev = mel.device
    rate = 1.0 if speed is None else float(speed)
    t_cap = int(1.0 / rate * (max(lens) + 3)) + 1 if rate != 1.0 else max(lens) + 3
    mel = mel.to(torch.float32).contiguous()
    lens_d = torch.tensor(lens, dtype=torch.int32, device=dev)
    cols_d = None if cols is None else torch.tensor(cols, dtype=torch.int32, device=dev)
    align_c = None if (align is None or cols is None) else align.to(torch.float32).contiguous()
    L = 0 if align_c is None else align_c.shape[2]
    out = torch.empty(B, F, t_cap, dtype=torch.float32, device=dev)
    out_lens = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ttsb_tacotron2_postprocess(_lib.ptr(mel), _lib.ptr(lens_d), _lib.ptr(align_c), _lib.ptr(cols_d), rate,
                                                  B, F, T, L, t_cap, _lib.ptr(out), _lib.ptr(out_lens),
                                                  _lib.current_stream(dev)))
    new = out_lens.tolist()                 # the one host read of the batch
    return [out[b, :, :new[b]] for b in range(B)]


class Tacotron2(Tacotron2MS):
    def __init__(self, checkpoint: str = None, n_symbol: int = 40, decoder_max_step: int = 3000, arabic_in: bool = True,
                 vowelizer: Optional[str] = None, **kwargs):
        super().__init__(n_symbol=n_symbol, decoder_max_step=decoder_max_step, **kwargs)
        self.n_eos = len(EOS_TOKENS)
        self.arabic_in = arabic_in
        state = None
        if checkpoint is not None:
            state = torch.load(checkpoint, map_location='cpu')
            self.load_state_dict(state['model'])
        self.config = get_basic_config()
        self.vowelizers = {}
        if vowelizer is not None:
            self.vowelizers[vowelizer] = _load_vowelizer(vowelizer, self.config)
        self.default_vowelizer = vowelizer
        self.phon_to_id = None
        if state is not None and 'symbols' in state:
            self.phon_to_id = {phon: i for i, phon in enumerate(state['symbols'])}
        self.eval()

    @property
    def device(self):
        return next(self.parameters()).device

    def _vowelize(self, utterance: str, vowelizer=None):
        vowelizer = self.default_vowelizer if vowelizer is None else vowelizer
        if vowelizer is not None:
            if vowelizer not in self.vowelizers:
                self.vowelizers[vowelizer] = _load_vowelizer(vowelizer, self.config)
            utterance = self.vowelizers[vowelizer].predict(text.buckwalter_to_arabic(utterance))
        return utterance

    def _tokenize(self, utterance: str, vowelizer=None):
        utterance = self._vowelize(utterance, vowelizer)
        if self.arabic_in:
            return text.arabic_to_tokens(utterance)
        return text.buckwalter_to_tokens(utterance)

    def _prepare(self, utterance, vowelizer, postprocess_mel):
        """tokens (+ the extra separator that makes the end of speech visible in the alignment,
        networks.py:133-137) -> (ids tensor, whether to truncate afterwards)."""
        tokens = self._tokenize(utterance, vowelizer)
        process = False
        if postprocess_mel and needs_postprocessing(tokens[-self.n_eos - 1]):
            tokens.insert(-self.n_eos, SEPARATOR_TOKEN)
            process = True
        return torch.LongTensor(text.tokens_to_ids(tokens, self.phon_to_id)), process

    @torch.inference_mode()
    def ttmel_single(self, utterance: str, speaker_id: int = 0, speed: Union[int, float, None] = None, vowelizer=None,
                     postprocess_mel: bool = True):
        ids, process = self._prepare(utterance, vowelizer, postprocess_mel)
        mel, _, align = self.infer(ids[None].to(self.device), torch.LongTensor([speaker_id]).to(self.device))
        col = align.shape[2] - self.n_eos - 1 if process else -1
        return postprocess_batch(mel, [mel.shape[2]], align, [col], speed)[0]   # [F, T]

    @torch.inference_mode()
    def ttmel_batch(self, batch: List[str], speaker_id: int = 0, speed: Union[int, float, None] = None, vowelizer=None,
                    postprocess_mel: bool = True):
        prepared = [self._prepare(line, vowelizer, postprocess_mel) for line in batch]
        padded, lens_sorted, inverse = text_collate_fn([p[0] for p in prepared])
        sids = torch.full((len(batch),), speaker_id, dtype=torch.long)
        mel, mel_lens, align = self.infer(padded.to(self.device), sids.to(self.device), lens_sorted.to(self.device))
        lens_sorted = lens_sorted.tolist()
        order = inverse.tolist()
        cols = [-1] * len(batch)               # per ROW of the sorted batch
        for i, row in enumerate(order):
            if prepared[i][1]:
                cols[row] = lens_sorted[row] - self.n_eos - 1
        done = postprocess_batch(mel, mel_lens, align, cols, speed)
        return [done[row] for row in order]

    def ttmel(self, text_input: Union[str, List[str]], speaker_id: int = 0, speed: Union[int, float, None] = None,
              batch_size: int = 8, vowelizer=None, postprocess_mel: bool = True):
        if isinstance(text_input, str):
            return self.ttmel_single(text_input, speaker_id, speed, vowelizer, postprocess_mel)
        assert isinstance(text_input, list)
        if batch_size == 1:
            return [self.ttmel_single(s, speaker_id, speed, vowelizer, postprocess_mel) for s in text_input]
        mels = []
        for k in range(0, len(text_input), batch_size):
            mels += self.ttmel_batch(text_input[k:k + batch_size], speaker_id, speed, vowelizer, postprocess_mel)
        return mels


class Tacotron2Wave(nn.Module):
    def __init__(self, model_sd_path: str, vocoder_sd: Optional[str] = None, vocoder_config: Optional[str] = None,
                 vowelizer: Optional[str] = None, arabic_in: bool = True, n_symbol: int = 40):
        super().__init__()
        model = Tacotron2(n_symbol=n_symbol, arabic_in=arabic_in, vowelizer=vowelizer)
        model.load_state_dict(torch.load(model_sd_path, map_location='cpu')['model'])
        self.model = model
        if vocoder_sd is None or vocoder_config is None:
            config = get_basic_config()
            vocoder_sd = config.vocoder_state_path
            vocoder_config = config.vocoder_config_path
        self.vocoder = load_hifigan(vocoder_sd, vocoder_config)
        self.denoiser = Denoiser(self.vocoder)
        self.eval()

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, x):
        return x

    @torch.inference_mode()
    def _vocode(self, mel_list: List[torch.Tensor], denoise: float):
        """Mels of different lengths -> one padded, length-masked vocoder call -> CPU waveforms."""
        lens = torch.tensor([m.shape[1] for m in mel_list], dtype=torch.int32)
        t_max = int(lens.max())
        batch = torch.zeros(len(mel_list), mel_list[0].shape[0], t_max, dtype=torch.float32, device=self.device)
        for i, m in enumerate(mel_list):
            batch[i, :, :m.shape[1]] = m
        wav = self.vocoder.run(mel_f32=batch, lens=lens, lens_host=lens.tolist())
        if denoise > 0:
            wav = self.denoiser.denoise_batch(wav, lens.to(wav.device) * self.vocoder.hop, denoise)
        wav = wav.cpu()
        hop = self.vocoder.hop
        return [wav[i, :int(lens[i]) * hop] for i in range(len(mel_list))]

    @torch.inference_mode()
    def tts_single(self, text_input: str, speed: Union[int, float, None] = None, speaker_id: int = 0, denoise: float = 0,
                   vowelizer=None, postprocess_mel: bool = True, return_mel: bool = False):
        mel = self.model.ttmel_single(text_input, speaker_id, speed, vowelizer, postprocess_mel)
        wave = self._vocode([mel], denoise)[0]
        return (wave, mel) if return_mel else wave

    @torch.inference_mode()
    def tts_batch(self, batch: List[str], speed: Union[int, float, None] = None, denoise: float = 0, speaker_id: int = 0,
                  vowelizer=None, postprocess_mel: bool = True, return_mel: bool = False):
        mels = self.model.ttmel_batch(batch, speaker_id, speed, vowelizer, postprocess_mel)
        return self._vocode(mels, denoise)   # return_mel is ignored on the batch path, as in the reference (:343-345)

    def tts(self, text_buckw: Union[str, List[str]], speed: Union[int, float, None] = None, denoise: float = 0.005,
            speaker_id: int = 0, batch_size: int = 8, vowelizer=None, postprocess_mel: bool = True,
            return_mel: bool = False) -> Union[torch.Tensor, List[torch.Tensor]]:
        kw = dict(speed=speed, denoise=denoise, speaker_id=speaker_id, vowelizer=vowelizer,
                  postprocess_mel=postprocess_mel, return_mel=return_mel)
        if isinstance(text_buckw, str):
            return self.tts_single(text_buckw, **kw)
        assert isinstance(text_buckw, list)
        if batch_size == 1:
            return [self.tts_single(s, **kw) for s in text_buckw]
        wavs = []
        for k in range(0, len(text_buckw), batch_size):
            wavs += self.tts_batch(text_buckw[k:k + batch_size], **kw)
        return wavs
