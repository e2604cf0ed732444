"""Public FastPitch API shell (drop-in for models/fastpitch/networks.py).

Same classes, method names, argument names and return types as the reference:
  text_collate_fn                                   networks.py:16-35
  FastPitch(checkpoint, arabic_in, vowelizer)        networks.py:45-253   .ttmel_single/.ttmel_batch/.ttmel
  FastPitch2Wave(model_sd_path, vocoder_sd, ...)     networks.py:256-435  .tts_single/.tts_batch/.tts
Differences, all on the fast side of the boundary: the vocoder runs ONCE per batch over the padded,
length-masked mel batch (instead of once per utterance) and waveforms leave the GPU in one pinned
copy. Results per utterance are those of the reference's per-utterance loop.
"""
from typing import List, Optional, Union

import torch
import torch.nn as nn

from ... import text
from ...utils import get_basic_config
from ...vocoder import load_hifigan
from ...vocoder.hifigan.denoiser import Denoiser
from .fastpitch.model import FastPitch as _FastPitch


def text_collate_fn(batch: List[torch.Tensor]):
    """list of 1-D id tensors -> (zero-padded [B, L_max] sorted by length desc, sorted lengths,
    inverse permutation)."""
    lens = torch.LongTensor([x.numel() for x in batch])
    lens_sorted, order = torch.sort(lens, descending=True)
    padded = torch.zeros(len(batch), int(lens_sorted[0]), dtype=torch.long)
    for row, src in enumerate(order.tolist()):
        padded[row, :batch[src].numel()] = batch[src]
    return padded, lens_sorted, order.argsort()


def pitch_trf(mul: float = 1, add: float = 0):
    def _apply(pitch_pred, enc_mask_sum, mean, std):
        return mul * pitch_pred + add
    return _apply


# The reference's `load_vowelizer` (models/diacritizers/__init__.py:4-12) builds one of two char-level BiLSTM diacritizers
# (shakkala, shakkelha). Those models are outside the hot path this package implements, but the
# text pre-step they plug into is part of the API (networks.py:75-85): any object with `predict(arabic_text) -> str` can be
# registered under the reference's names (e.g. the reference's own model objects) and is then used exactly where the
# reference uses them; predictions are memoised per sentence (the front-end is what a batch waits
# for at 10^4 x real time).
_VOWELIZERS = {}


class _CachedVowelizer:
    def __init__(self, model, max_entries=65536):
        self.model = model
        self.cache = {}
        self.max_entries = max_entries

    def predict(self, utterance: str) -> str:
        hit = self.cache.get(utterance)
        if hit is None:
            hit = self.model.predict(utterance)
            if len(self.cache) >= self.max_entries:
                self.cache.clear()
            self.cache[utterance] = hit
        return hit


def register_vowelizer(name: str, model_or_factory):
    """Make `vowelizer=name` available: `model_or_factory` is an object with predict(str) -> str, or a callable
    (config) -> such an object that is invoked on first use (the reference loads its diacritizers lazily too)."""
    _VOWELIZERS[name] = model_or_factory


def _load_vowelizer(name, config):
    if name not in _VOWELIZERS:
        raise NotImplementedError(
            "vowelizer '%s' is not registered: the diacritizer RNNs are outside the hot path this package implements "
            "(SURVEY.md §2 row 12). Pass already-vocalised text, or plug a model in with "
            "tts_arabic_pytorch_b200.models.fastpitch.networks.register_vowelizer(name, obj_with_predict)" % name)
    entry = _VOWELIZERS[name]
    model = entry if hasattr(entry, 'predict') else entry(config)
    return _CachedVowelizer(model)


class FastPitch(_FastPitch):
    def __init__(self, checkpoint: str, arabic_in: bool = True, vowelizer: Optional[str] = None, **kwargs):
        from . import net_config
        state = torch.load(checkpoint, map_location='cpu')
        cfg = state['config'] if 'config' in state else net_config
        super().__init__(**cfg)
        self.arabic_in = arabic_in
        self.load_state_dict(state['model'])
        self.config = get_basic_config()
        self.vowelizers = {}
        if vowelizer is not None:
            self.vowelizers[vowelizer] = _load_vowelizer(vowelizer, self.config)
        self.default_vowelizer = vowelizer
        self.phon_to_id = None
        if 'symbols' in state:
            self.phon_to_id = {phon: i for i, phon in enumerate(state['symbols'])}
        self.eval()

    @property
    def device(self):
        return next(self.parameters()).device

    # ------------------------------------------------------------------ text -> ids
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
            return text.arabic_to_tokens(utterance, append_space=False)
        return text.buckwalter_to_tokens(utterance, append_space=False)

    def _ids(self, utterance: str, vowelizer=None):
        return torch.LongTensor(text.tokens_to_ids(self._tokenize(utterance, vowelizer), self.phon_to_id))

    @staticmethod
    def _transform(pitch_mul, pitch_add, pitch_transform):
        if (pitch_mul != 1. or pitch_add != 0.) and pitch_transform is None:
            return pitch_trf(pitch_mul, pitch_add)
        return pitch_transform

    # ------------------------------------------------------------------ ids -> mel (batch core)
    @torch.inference_mode()
    def _infer_ids(self, id_list: List[torch.Tensor], speed, speaker_id, pitch_transform, dur_tgt, pitch_tgt,
                   energy_tgt, max_duration, channel_last=False, pad_to: int = 0, frame_len_hook=None):
        """pad_to / frame_len_hook: a shard of a larger batch keeps that batch's padding condition (parallel.py)."""
        padded, _, inverse = text_collate_fn(id_list)
        if pad_to > padded.shape[1]:
            padded = torch.nn.functional.pad(padded, (0, pad_to - padded.shape[1]))
        out = self.infer(padded.to(self.device), pace=speed, speaker=speaker_id, dur_tgt=dur_tgt, pitch_tgt=pitch_tgt,
                         energy_tgt=energy_tgt, pitch_transform=pitch_transform, max_duration=max_duration,
                         return_channel_last=channel_last, frame_len_hook=frame_len_hook)
        return out, inverse

    @torch.inference_mode()
    def ttmel_single(self, utterance: str, speed: float = 1, speaker_id: int = 0, vowelizer=None,
                     pitch_mul: float = 1., pitch_add: float = 0., dur_tgt=None, pitch_tgt=None, energy_tgt=None,
                     pitch_transform=None, max_duration=75):
        ids = self._ids(utterance, vowelizer)
        mel, *_ = self.infer(ids[None].to(self.device), pace=speed, speaker=speaker_id, dur_tgt=dur_tgt,
                             pitch_tgt=pitch_tgt, energy_tgt=energy_tgt,
                             pitch_transform=self._transform(pitch_mul, pitch_add, pitch_transform),
                             max_duration=max_duration)
        return mel[0]   # [F, T]

    @torch.inference_mode()
    def ttmel_batch(self, batch: List[str], speed: float = 1, speaker_id: int = 0, vowelizer=None,
                    pitch_mul: float = 1., pitch_add: float = 0., dur_tgt=None, pitch_tgt=None, energy_tgt=None,
                    pitch_transform=None, max_duration=75):
        id_list = [self._ids(line, vowelizer) for line in batch]
        (mel, dec_lens, *_), inverse = self._infer_ids(id_list, speed, speaker_id,
                                                       self._transform(pitch_mul, pitch_add, pitch_transform),
                                                       dur_tgt, pitch_tgt, energy_tgt, max_duration)
        lens = dec_lens.tolist()            # one host read for the whole batch
        return [mel[row, :, :lens[row]] for row in inverse.tolist()]

    def ttmel(self, text_input: Union[str, List[str]], speed: float = 1, speaker_id: int = 0, batch_size: int = 1,
              vowelizer=None, pitch_mul: float = 1., pitch_add: float = 0.):
        kw = dict(speed=speed, speaker_id=speaker_id, vowelizer=vowelizer, pitch_mul=pitch_mul, pitch_add=pitch_add)
        if isinstance(text_input, str):
            return self.ttmel_single(text_input, **kw)
        assert isinstance(text_input, list)
        if batch_size == 1:
            return [self.ttmel_single(sample, **kw) for sample in text_input]
        mels = []
        for k in range(0, len(text_input), batch_size):
            mels += self.ttmel_batch(text_input[k:k + batch_size], **kw)
        return mels


class FastPitch2Wave(nn.Module):
    def __init__(self, model_sd_path: str, vocoder_sd: Optional[str] = None, vocoder_config: Optional[str] = None,
                 vowelizer: Optional[str] = None, arabic_in: bool = True):
        super().__init__()
        self.model = FastPitch(model_sd_path, arabic_in=arabic_in, vowelizer=vowelizer)
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

    # ------------------------------------------------------------------ ids -> waveforms (batch core)
    @torch.inference_mode()
    def synthesize_ids(self, id_list: List[torch.Tensor], speed=1., speaker_id=0, denoise=0., pitch_transform=None,
                       max_duration=75, to_cpu=True, pad_to: int = 0, frame_len_hook=None, return_padded=False,
                       host_alloc=None):
        """Core used by every tts_* method: padded FastPitch batch -> ONE masked vocoder call ->
        (optional) batched denoiser -> single D2H copy. Returns (list of 1-D waveforms in input order,
        list of [80,T_i] mels)."""
        (mel, dec_lens, _, _, _, mel_cl), inverse = self.model._infer_ids(
            id_list, speed, speaker_id, pitch_transform, None, None, None, max_duration, channel_last=True,
            pad_to=pad_to, frame_len_hook=frame_len_hook)
        host = None
        n_cols = mel_cl.shape[1] * self.vocoder.hop
        if host_alloc is not None:
            # parallel.synthesize (host_shm): this rank's rows of the shared pinned segment, sorted-batch order
            host = host_alloc(mel_cl.shape[0], n_cols)
        elif to_cpu and not return_padded:
            # the waveforms' pinned destination is known up front: finished groups of utterances leave over PCIe while
            # the generator works on the next group (Generator.run, host_out)
            host = torch.empty(mel_cl.shape[0], n_cols, dtype=torch.float32, pin_memory=True)
        wav = self.vocoder.run(mel_cl=mel_cl, lens=dec_lens, host_out=host if denoise <= 0 else None)   # [B, T_max*hop]
        if denoise > 0:
            wav = self.denoiser.denoise_batch(wav, dec_lens * self.vocoder.hop, denoise)
            if host is not None:              # the generator's output was not the final waveform: one copy of the result
                host.copy_(wav, non_blocking=True)
                torch.cuda.current_stream(wav.device).synchronize()
        if return_padded:
            # device-resident form for parallel.synthesize: padded waveforms in SORTED order + what undoes the sort
            return wav, dec_lens * self.vocoder.hop, inverse, mel
        lens = dec_lens.tolist()
        hop = self.vocoder.hop
        order = inverse.tolist()
        if to_cpu:
            # one D2H copy of the padded batch into pinned memory; the per-utterance results are views of it. A fresh
            # pinned tensor per call keeps earlier results valid (the reference returns independent `wav[0].cpu()`
            # tensors) and costs no allocation in steady state: torch's caching host allocator recycles the blocks of
            # results the caller has dropped. (A reused staging buffer + per-utterance clones measured 40 ms of host
            # copies per 256-utterance batch, a quarter of the GPU step.)
            if host is None:
                host = torch.empty(wav.shape, dtype=wav.dtype, pin_memory=True)
                host.copy_(wav, non_blocking=True)
                torch.cuda.current_stream(wav.device).synchronize()
            wavs = [host[row, :lens[row] * hop] for row in order]
        else:
            wavs = [wav[row, :lens[row] * hop] for row in order]
        return (wavs, [mel[row, :, :lens[row]] for row in order])

    @torch.inference_mode()
    def tts_single(self, text_buckw: str, speed: float = 1, speaker_id: int = 0, denoise: float = 0, vowelizer=None,
                   pitch_mul: float = 1., pitch_add: float = 0., return_mel: bool = False):
        ids = self.model._ids(text_buckw, vowelizer)
        wavs, mels = self.synthesize_ids([ids], speed, speaker_id, denoise,
                                         self.model._transform(pitch_mul, pitch_add, None))
        if return_mel:
            return wavs[0], mels[0]
        return wavs[0]

    @torch.inference_mode()
    def tts_batch(self, batch: List[str], speed: float = 1, speaker_id: int = 0, denoise: float = 0, vowelizer=None,
                  pitch_mul: float = 1., pitch_add: float = 0., return_mel: bool = False):
        id_list = [self.model._ids(line, vowelizer) for line in batch]
        wavs, _ = self.synthesize_ids(id_list, speed, speaker_id, denoise,
                                      self.model._transform(pitch_mul, pitch_add, None))
        # the reference evaluates `wav_list, mel_list` without returning it (networks.py:347-348):
        # return_mel has no effect on the batch path, kept for drop-in behaviour
        return wavs

    def tts(self, text_input: Union[str, List[str]], speed: float = 1., denoise: float = 0.005, speaker_id: int = 0,
            batch_size: int = 2, vowelizer=None, pitch_mul: float = 1., pitch_add: float = 0.,
            return_mel: bool = False) -> Union[torch.Tensor, List[torch.Tensor]]:
        """text (str | list[str]) -> waveform(s), 1-D fp32 CPU tensors of n_samples each."""
        kw = dict(speed=speed, speaker_id=speaker_id, denoise=denoise, vowelizer=vowelizer, pitch_mul=pitch_mul,
                  pitch_add=pitch_add, return_mel=return_mel)
        if isinstance(text_input, str):
            return self.tts_single(text_input, **kw)
        assert isinstance(text_input, list)
        if batch_size == 1:
            return [self.tts_single(sample, **kw) for sample in text_input]
        wavs = []
        for k in range(0, len(text_input), batch_size):
            wavs += self.tts_batch(text_input[k:k + batch_size], **kw)
        return wavs
