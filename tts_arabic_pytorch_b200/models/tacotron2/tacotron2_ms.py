"""Multi-speaker Tacotron2 front: parameters under the reference's state_dict names, `infer()` through
the C ABI (ttsb_tacotron2_encode / _decode / _finish).

Drop-in for models.tacotron2.tacotron2_ms.Tacotron2MS.infer (tacotron2_ms.py:278-332). Inference only.
The reference's prenet applies dropout(p=0.5) even in eval mode (torchaudio:283-285), so its output
is stochastic; here the keep-masks are drawn with torch's device RNG per chunk of decoder steps, or
injected through `prenet_masks` (tests, reproducibility).
"""
import ctypes
import warnings
from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn

from ... import _lib
from ..fastpitch.fastpitch.model import _attach


def parameter_spec(n_symbol=40, num_speakers=40, speaker_embedding_dim=128, n_mels=80):
    E, H, P, A, NF, KL = 512, 1024, 256, 128, 32, 31
    M = E + (speaker_embedding_dim if num_speakers > 1 else 0)
    spec = OrderedDict()

    def P_(name, *shape):
        spec[name] = (tuple(shape), False, torch.float32)

    def bn(prefix, n):
        P_(prefix + '.weight', n)
        P_(prefix + '.bias', n)
        spec[prefix + '.running_mean'] = ((n,), True, torch.float32)
        spec[prefix + '.running_var'] = ((n,), True, torch.float32)
        spec[prefix + '.num_batches_tracked'] = ((), True, torch.long)

    P_('embedding.weight', n_symbol, E)
    for i in range(3):
        P_('encoder.convolutions.%d.0.weight' % i, E, E, 5)
        P_('encoder.convolutions.%d.0.bias' % i, E)
        bn('encoder.convolutions.%d.1' % i, E)
    for suf in ('', '_reverse'):
        P_('encoder.lstm.weight_ih_l0' + suf, 4 * E // 2, E)
        P_('encoder.lstm.weight_hh_l0' + suf, 4 * E // 2, E // 2)
        P_('encoder.lstm.bias_ih_l0' + suf, 4 * E // 2)
        P_('encoder.lstm.bias_hh_l0' + suf, 4 * E // 2)
    P_('decoder.prenet.layers.0.weight', P, n_mels)
    P_('decoder.prenet.layers.1.weight', P, P)
    P_('decoder.attention_rnn.weight_ih', 4 * H, P + M)
    P_('decoder.attention_rnn.weight_hh', 4 * H, H)
    P_('decoder.attention_rnn.bias_ih', 4 * H)
    P_('decoder.attention_rnn.bias_hh', 4 * H)
    P_('decoder.attention_layer.query_layer.weight', A, H)
    P_('decoder.attention_layer.memory_layer.weight', A, M)
    P_('decoder.attention_layer.v.weight', 1, A)
    P_('decoder.attention_layer.location_layer.location_conv.weight', NF, 2, KL)
    P_('decoder.attention_layer.location_layer.location_dense.weight', A, NF)
    P_('decoder.decoder_rnn.weight_ih', 4 * H, H + M)
    P_('decoder.decoder_rnn.weight_hh', 4 * H, H)
    P_('decoder.decoder_rnn.bias_ih', 4 * H)
    P_('decoder.decoder_rnn.bias_hh', 4 * H)
    P_('decoder.linear_projection.weight', n_mels, H + M)
    P_('decoder.linear_projection.bias', n_mels)
    P_('decoder.gate_layer.weight', 1, H + M)
    P_('decoder.gate_layer.bias', 1)
    dims = [n_mels, 512, 512, 512, 512, n_mels]
    for i in range(5):
        P_('postnet.convolutions.%d.0.weight' % i, dims[i + 1], dims[i], 5)
        P_('postnet.convolutions.%d.0.bias' % i, dims[i + 1])
        bn('postnet.convolutions.%d.1' % i, dims[i + 1])
    if num_speakers > 1:
        P_('speaker_embedding.weight', num_speakers, speaker_embedding_dim)
    return spec


class Tacotron2MS(nn.Module):
    MAX_GROUP = 64       # utterances per launch train (t2_lstm_cell_body's shared staging)
    STEP_CHUNK = 64      # decoder steps per cooperative launch = between two host reads of the "all finished" flag

    def __init__(self, mask_padding: bool = False, n_mels: int = 80, n_symbol: int = 148, n_frames_per_step: int = 1,
                 num_speakers=40, speaker_embedding_dim=128, decoder_max_step: int = 2000,
                 decoder_early_stopping: bool = True, gate_threshold: float = 0.5, **kwargs):
        super().__init__()
        fixed = dict(symbol_embedding_dim=512, encoder_embedding_dim=512, encoder_n_convolution=3, encoder_kernel_size=5,
                     decoder_rnn_dim=1024, attention_rnn_dim=1024, attention_hidden_dim=128, attention_location_n_filter=32,
                     attention_location_kernel_size=31, prenet_dim=256, postnet_n_convolution=5, postnet_kernel_size=5,
                     postnet_embedding_dim=512)
        for k, v in kwargs.items():
            if k in fixed and v != fixed[k]:
                raise NotImplementedError('%s=%r: only the reference architecture (%r) is implemented' % (k, v, fixed[k]))
        if n_frames_per_step != 1 or n_mels != 80:
            raise NotImplementedError('n_frames_per_step=1 and n_mels=80 only')
        self.n_mels = n_mels
        self.mask_padding = mask_padding
        self.decoder_max_step = decoder_max_step
        self.decoder_early_stopping = decoder_early_stopping
        self.gate_threshold = gate_threshold
        self.num_speakers = num_speakers
        for name, (shape, is_buf, dtype) in parameter_spec(n_symbol, num_speakers, speaker_embedding_dim, n_mels).items():
            init = torch.ones(shape, dtype=dtype) if name.endswith('running_var') else torch.zeros(shape, dtype=dtype)
            _attach(self, name, init, is_buf)
        self._handle = None
        self._handle_key = None
        self._ws = _lib.Workspace()
        self._state = _lib.Workspace()

    def _drop_handle(self):
        if getattr(self, '_handle', None) is not None:
            _lib.load().ttsb_tacotron2_destroy(self._handle)
        self._handle = None
        self._handle_key = None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._drop_handle()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._drop_handle()
        return out

    def _get_handle(self, device):
        if self._handle is not None and self._handle_key == device:
            return self._handle
        self._drop_handle()
        named = {k: v.detach().float().cpu() for k, v in self.state_dict().items() if not k.endswith('num_batches_tracked')}
        table, keep = _lib.tensor_table(named)
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().ttsb_tacotron2_create(table, len(named), device.index or 0, ctypes.byref(handle)))
        del keep
        self._handle, self._handle_key = handle, device
        return handle

    def forward(self, *a, **k):
        raise NotImplementedError('training forward is out of scope; use infer()')

    @torch.no_grad()
    def infer(self, tokens: torch.Tensor, speaker_ids: Optional[torch.Tensor] = None,
              lengths: Optional[torch.Tensor] = None, prenet_masks: Optional[torch.Tensor] = None,
              return_channel_last: bool = False):
        """-> (mel [B,80,T], mel_lengths int32 [B], alignments [B,T,L]) as the reference.
        prenet_masks (optional) [steps, 2, B, 256] bool keep-masks; decoding then runs exactly
        min(steps, decoder_max_step) steps unless every utterance stops earlier."""
        device = self.embedding.weight.device
        if device.type != 'cuda':
            raise RuntimeError('tts_arabic_pytorch_b200 has no CPU path: move the model to a CUDA device')
        lib = _lib.load()
        tokens = tokens.to(device=device, dtype=torch.int64).contiguous()
        B, L = tokens.shape
        if B > self.MAX_GROUP:
            # the LSTM-cell kernels stage gate pre-activations for at most 64 utterances: larger batches run as groups
            # (utterances are independent; frames beyond an utterance's own mel_lengths are not part of the contract)
            parts = []
            for k in range(0, B, self.MAX_GROUP):
                sl = slice(k, k + self.MAX_GROUP)
                parts.append(self.infer(tokens[sl], None if speaker_ids is None else speaker_ids[sl],
                                        None if lengths is None else lengths[sl],
                                        None if prenet_masks is None else prenet_masks[:, :, sl],
                                        return_channel_last))
            t_max = max(p[0].shape[2] for p in parts)

            def cat(i, dim, width):
                return torch.cat([torch.nn.functional.pad(p[i], width(t_max - p[0].shape[2])) for p in parts], dim=0)
            out = (cat(0, 0, lambda d: (0, d)), torch.cat([p[1] for p in parts]), cat(2, 0, lambda d: (0, 0, 0, d)))
            return out + (cat(3, 0, lambda d: (0, 0, 0, d)),) if return_channel_last else out
        if lengths is None:
            lengths = torch.full((B,), L, dtype=torch.int32, device=device)
        lengths = lengths.to(device=device, dtype=torch.int32).contiguous()
        if speaker_ids is None:
            speaker_ids = torch.zeros(B, dtype=torch.int64, device=device)
        speaker_ids = speaker_ids.to(device=device, dtype=torch.int64).contiguous()
        max_steps = self.decoder_max_step if prenet_masks is None else min(self.decoder_max_step, prenet_masks.shape[0])
        with torch.cuda.device(device):
            handle = self._get_handle(device)
            stream = _lib.current_stream(device)
            state = self._state.get(lib.ttsb_tacotron2_state_bytes(handle, B, L, max_steps), device)
            ws = self._ws.get(lib.ttsb_tacotron2_workspace_bytes(handle, B, L, 0), device)
            _lib.check(lib.ttsb_tacotron2_encode(handle, _lib.ptr(tokens), _lib.ptr(lengths), _lib.ptr(speaker_ids), B, L,
                                                 max_steps, _lib.ptr(state), _lib.ptr(ws), ws.numel(), stream))
            done = (ctypes.c_int * 2)(-1, 0)
            step = 0
            while step < max_steps:
                n = min(self.STEP_CHUNK, max_steps - step)
                if prenet_masks is None:
                    masks = (torch.rand(n, 2, B, 256, device=device) >= 0.5).to(torch.uint8)
                else:
                    masks = prenet_masks[step:step + n].to(device=device, dtype=torch.uint8).contiguous()
                _lib.check(lib.ttsb_tacotron2_decode(handle, B, L, max_steps, step, n, _lib.ptr(masks),
                                                     float(self.gate_threshold), int(bool(self.decoder_early_stopping)),
                                                     _lib.ptr(state), done, stream))
                if done[1] & 1:
                    raise IndexError('token id out of range [0, %d)' % self.embedding.weight.shape[0])
                if done[1] & 4:
                    raise IndexError('speaker id out of range [0, %d)' % self.num_speakers)
                step += n
                if self.decoder_early_stopping and done[0] >= 0:
                    break
            if self.decoder_early_stopping and done[0] >= 0:
                T = done[0] + 1
            else:
                T = step
                if T == self.decoder_max_step:
                    warnings.warn('Reached max decoder steps. The generated spectrogram might not cover the whole transcript.')
            ws = self._ws.get(lib.ttsb_tacotron2_workspace_bytes(handle, B, L, T), device)
            mel = torch.empty(B, self.n_mels, T, dtype=torch.float32, device=device)
            mel_lens = torch.empty(B, dtype=torch.int32, device=device)
            align = torch.empty(B, T, L, dtype=torch.float32, device=device)
            mel_cl = torch.empty(B, T, 128, dtype=torch.float16, device=device) if return_channel_last else None
            _lib.check(lib.ttsb_tacotron2_finish(handle, B, L, max_steps, T, _lib.ptr(mel), _lib.ptr(mel_lens), _lib.ptr(align),
                                                 _lib.ptr(mel_cl), _lib.ptr(state), _lib.ptr(ws), ws.numel(), stream))
        # utterances that stopped before T keep counting in the reference only until they finish:
        # mel_lens already holds that count (torchaudio:846-849)
        out = (mel, mel_lens, align)
        return out + (mel_cl,) if return_channel_last else out
