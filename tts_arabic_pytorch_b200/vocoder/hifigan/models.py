"""HiFi-GAN V1 generator front: an nn.Module that owns the parameters under the reference's
state_dict names and forwards through the C ABI (ttsb_hifigan_forward).

Drop-in for vocoder.hifigan.models.Generator (vocoder/hifigan/models.py:86-136):
  Generator(h)                      h = AttrDict of pretrained/hifigan-asc-v1/config.json
  .load_state_dict(ckpt['generator'])   weight-norm parametrised keys, as shipped
  .remove_weight_norm()
  .forward(x)                       [80,T] -> [1,256T]   or   [B,80,T] -> [B,1,256T]
Extra (not in the reference): forward(x, lens=...) masks every layer at each utterance's own
length so a padded batch equals the reference's per-utterance loop.
"""
import ctypes
import os

import torch
import torch.nn as nn
from torch.nn.utils import parametrize
from torch.nn.utils.parametrizations import weight_norm

from ... import _lib
from ...utils.synth import fold_weight_norm


class _ConvParams(nn.Module):
    """Parameter container with the attribute names of a torch conv (`weight`, `bias`); never called."""

    def __init__(self, shape):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(shape))
        self.bias = nn.Parameter(torch.zeros(shape[0] if self.bias_dim == 0 else shape[1]))

    bias_dim = 0


class _ConvTParams(_ConvParams):
    bias_dim = 1


class _ResBlockParams(nn.Module):
    def __init__(self, channels, kernel_size, n_dil):
        super().__init__()
        self.convs1 = nn.ModuleList([weight_norm(_ConvParams((channels, channels, kernel_size))) for _ in range(n_dil)])
        self.convs2 = nn.ModuleList([weight_norm(_ConvParams((channels, channels, kernel_size))) for _ in range(n_dil)])


def _strip(mod):
    if parametrize.is_parametrized(mod, 'weight'):
        parametrize.remove_parametrizations(mod, 'weight')


class Generator(nn.Module):
    def __init__(self, h):
        super().__init__()
        if str(h['resblock']) != '1':
            raise NotImplementedError('only ResBlock1 generators (HiFi-GAN V1) are supported')
        self.h = h
        self.num_kernels = len(h['resblock_kernel_sizes'])
        self.num_upsamples = len(h['upsample_rates'])
        c = h['upsample_initial_channel']
        self.conv_pre = weight_norm(_ConvParams((c, h.get('num_mels', 80), 7)))
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        for u, k in zip(h['upsample_rates'], h['upsample_kernel_sizes']):
            self.ups.append(weight_norm(_ConvTParams((c, c // 2, k))))
            c //= 2
            for rk, dil in zip(h['resblock_kernel_sizes'], h['resblock_dilation_sizes']):
                self.resblocks.append(_ResBlockParams(c, rk, len(dil)))
        self.conv_post = weight_norm(_ConvParams((1, c, 7)))
        self._handle = None
        self._handle_key = None
        self._ws = _lib.Workspace()

    # ------------------------------------------------------------------ reference API
    def remove_weight_norm(self):
        _strip(self.conv_pre)
        _strip(self.conv_post)
        for m in self.ups:
            _strip(m)
        for rb in self.resblocks:
            for m in list(rb.convs1) + list(rb.convs2):
                _strip(m)
        self._drop_handle()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._drop_handle()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._drop_handle()
        return out

    @property
    def hop(self):
        n = 1
        for u in self.h['upsample_rates']:
            n *= u
        return n

    # ------------------------------------------------------------------ C ABI plumbing
    def _drop_handle(self):
        self.__dict__.pop('_graphs', None)     # captured graphs hold the handle's device buffers
        if getattr(self, '_handle', None) is not None:
            _lib.load().ttsb_hifigan_destroy(self._handle)
        self._handle = None
        self._handle_key = None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def _get_handle(self, device):
        if self._handle is not None and self._handle_key == device:
            return self._handle
        self._drop_handle()
        lib = _lib.load()
        h = self.h
        cfg = _lib.HifiganConfig()
        cfg.num_mels = h.get('num_mels', 80)
        cfg.upsample_initial_channel = h['upsample_initial_channel']
        cfg.num_upsamples = self.num_upsamples
        cfg.num_kernels = self.num_kernels
        for i, (u, k) in enumerate(zip(h['upsample_rates'], h['upsample_kernel_sizes'])):
            cfg.upsample_rates[i] = u
            cfg.upsample_kernel_sizes[i] = k
        for j, (rk, dil) in enumerate(zip(h['resblock_kernel_sizes'], h['resblock_dilation_sizes'])):
            if len(dil) != 3:
                raise NotImplementedError('ResBlock1 with 3 dilations expected')
            cfg.resblock_kernel_sizes[j] = rk
            for p, d in enumerate(dil):
                cfg.resblock_dilations[j][p] = d
        folded = fold_weight_norm({k: v.detach().float().cpu() for k, v in self.state_dict().items()})
        table, keep = _lib.tensor_table(folded)
        handle = ctypes.c_void_p()
        _lib.check(lib.ttsb_hifigan_create(ctypes.byref(cfg), table, len(folded), device.index or 0,
                                           ctypes.byref(handle)))
        del keep
        self._handle, self._handle_key = handle, device
        return handle

    def _device(self):
        return self.conv_post.bias.device

    def run(self, mel_f32=None, mel_cl=None, lens=None, lens_host=None, host_out=None):
        """Padded batch -> [B, T*hop] fp32. `mel_f32` [B,80,T] or `mel_cl` [B,T,128] fp16. lens_host: the frame counts as a
        Python list when the caller already has them on the host (FastPitch.infer attaches them to the `dec_lens` it
        returns): the padded batch then runs in chunks at each chunk's own longest utterance (ttsb_hifigan_forward).
        host_out: a pinned [B, T*hop] fp32 host tensor. The batch then runs in groups of utterances (one workspace chunk
        each) and every finished group is copied device -> host on a side stream while the next group computes, so the
        134 MB of a 256-utterance batch leave over PCIe under the generator instead of after it; the call returns once
        the last copy has landed."""
        device = self._device()
        if device.type != 'cuda':
            raise RuntimeError('tts_arabic_pytorch_b200 has no CPU path: move the vocoder to a CUDA device '
                               '(Generator.cuda()) before calling it')
        lib = _lib.load()
        src = mel_f32 if mel_f32 is not None else mel_cl
        if mel_f32 is not None:
            mel_f32 = mel_f32.to(device=device, dtype=torch.float32).contiguous()
            B, _, T = mel_f32.shape
        else:
            assert mel_cl.dtype == torch.float16 and mel_cl.is_contiguous() and mel_cl.device == device
            B, T, _ = mel_cl.shape
        if lens is not None:
            if lens_host is None:
                lens_host = getattr(lens, 'host_list', None)
            lens = lens.to(device=device, dtype=torch.int32).contiguous()
        h_lens = None
        if lens is not None and lens_host is not None and len(lens_host) == B:
            h_lens = (ctypes.c_int32 * B)(*[int(x) for x in lens_host])
        with torch.cuda.device(device):
            handle = self._get_handle(device)
            wav = torch.empty(B, T * self.hop, dtype=torch.float32, device=device)
            nbytes = lib.ttsb_hifigan_workspace_bytes(handle, B, T)
            ws = self._ws.get(nbytes, device)
            group = B
            if host_out is not None:
                assert host_out.shape == wav.shape and host_out.dtype == torch.float32 and host_out.device.type == 'cpu'
                group = max(1, min(B, int(os.environ.get('TTSB_D2H_GROUP_FRAMES', '32768')) // max(T, 1)))
            if group >= B and host_out is None:
                _lib.check(lib.ttsb_hifigan_forward(handle, _lib.ptr(mel_f32), _lib.ptr(mel_cl), _lib.ptr(lens), h_lens, B, T,
                                                    _lib.ptr(wav), _lib.ptr(ws), nbytes, _lib.current_stream(device)))
            else:
                copy_stream = self.__dict__.get('_copy_stream')
                if copy_stream is None or copy_stream.device != device:
                    copy_stream = self.__dict__['_copy_stream'] = torch.cuda.Stream(device=device)
                main = torch.cuda.current_stream(device)
                for b0 in range(0, B, group):
                    b1 = min(B, b0 + group)
                    sub_lens = None if h_lens is None else (ctypes.c_int32 * (b1 - b0))(*h_lens[b0:b1])
                    _lib.check(lib.ttsb_hifigan_forward(
                        handle, _lib.ptr(mel_f32[b0:b1]) if mel_f32 is not None else None,
                        _lib.ptr(mel_cl[b0:b1]) if mel_cl is not None else None,
                        _lib.ptr(lens[b0:b1]) if lens is not None else None, sub_lens, b1 - b0, T, _lib.ptr(wav[b0:b1]),
                        _lib.ptr(ws), nbytes, _lib.current_stream(device)))
                    done = torch.cuda.Event()
                    done.record(main)
                    copy_stream.wait_event(done)
                    with torch.cuda.stream(copy_stream):
                        host_out[b0:b1].copy_(wav[b0:b1], non_blocking=True)
                wav.record_stream(copy_stream)
                copy_stream.synchronize()
        del src
        return wav

    # ------------------------------------------------------------------ CUDA graphs (small batches)
    def capture_graph(self, mel, lens=None):
        """Captures the generator's launch train for this input SHAPE into a CUDA graph and returns `replay(mel=None,
        lens=None) -> wav [B, T*hop]`: one graph launch instead of ~62 kernel launches, which is what a single utterance is
        bound by (BASELINE config 2: the launches of a 512-frame utterance take longer to issue than to run).
        `replay` copies a new mel (same shape) into the captured input buffer first; the returned tensor is the graph's
        static output and is overwritten by the next replay. One graph per (B, T, with-lens) is kept."""
        device = self._device()
        if device.type != 'cuda':
            raise RuntimeError('tts_arabic_pytorch_b200 has no CPU path: move the vocoder to a CUDA device')
        if mel.dim() == 2:
            mel = mel[None]
        key = (tuple(mel.shape), lens is not None)
        cache = self.__dict__.setdefault('_graphs', {})
        if key not in cache:
            static_mel = mel.to(device=device, dtype=torch.float32).contiguous().clone()
            static_lens = None if lens is None else lens.to(device=device, dtype=torch.int32).contiguous().clone()
            self.run(mel_f32=static_mel, lens=static_lens)      # warm-up outside capture: handle, workspace, attributes
            torch.cuda.synchronize(device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_wav = self.run(mel_f32=static_mel, lens=static_lens)
            cache[key] = (graph, static_mel, static_lens, static_wav)
        graph, static_mel, static_lens, static_wav = cache[key]

        def replay(new_mel=None, new_lens=None):
            if new_mel is not None:
                static_mel.copy_(new_mel if new_mel.dim() == 3 else new_mel[None], non_blocking=True)
            if new_lens is not None and static_lens is not None:
                static_lens.copy_(new_lens, non_blocking=True)
            graph.replay()
            return static_wav
        return replay

    def _drop_graphs(self):
        self.__dict__.pop('_graphs', None)

    def forward(self, x, lens=None):
        if x.dim() == 2:                       # [80,T] -> [1, 256T]   (hifigan/models.py:111-127)
            return self.run(mel_f32=x[None], lens=lens)
        return self.run(mel_f32=x, lens=lens)[:, None, :]
