"""FastPitch acoustic model front: parameters under the reference's state_dict names, `infer()`
through the C ABI (ttsb_fastpitch_encode / _condition / _decode).

Drop-in for models.fastpitch.fastpitch.model.FastPitch.infer (model.py:351-409). Only inference is
provided; `forward` (training) raises. The training-only aligner tensors (`attention.*`) are kept
as inert parameters so that strict checkpoint loading works.
"""
import ctypes
from collections import OrderedDict

import torch
import torch.nn as nn

from .... import _lib


def parameter_spec(cfg):
    """name -> (shape, is_buffer) for every entry of the reference module's state_dict."""
    D = cfg['symbols_embedding_dim']
    nm = cfg['n_mel_channels']
    spec = OrderedDict()

    def P(name, *shape):
        spec[name] = (tuple(shape), False)

    def fft(prefix, n_layers, d_head, d_inner, k, n_embed):
        if n_embed:
            P(prefix + '.word_emb.weight', n_embed, D)
        spec[prefix + '.pos_emb.inv_freq'] = ((D // 2,), True)
        for i in range(n_layers):
            p = '%s.layers.%d.' % (prefix, i)
            P(p + 'dec_attn.qkv_net.weight', 3 * d_head, D)
            P(p + 'dec_attn.qkv_net.bias', 3 * d_head)
            P(p + 'dec_attn.o_net.weight', D, d_head)
            P(p + 'dec_attn.layer_norm.weight', D)
            P(p + 'dec_attn.layer_norm.bias', D)
            P(p + 'pos_ff.CoreNet.0.weight', d_inner, D, k)
            P(p + 'pos_ff.CoreNet.0.bias', d_inner)
            P(p + 'pos_ff.CoreNet.2.weight', D, d_inner, k)
            P(p + 'pos_ff.CoreNet.2.bias', D)
            P(p + 'pos_ff.layer_norm.weight', D)
            P(p + 'pos_ff.layer_norm.bias', D)

    def predictor(prefix, filt, k, n_layers):
        for i in range(n_layers):
            p = '%s.layers.%d.' % (prefix, i)
            P(p + 'conv.weight', filt, D if i == 0 else filt, k)
            P(p + 'conv.bias', filt)
            P(p + 'norm.weight', filt)
            P(p + 'norm.bias', filt)
        P(prefix + '.fc.weight', 1, filt)
        P(prefix + '.fc.bias', 1)

    if cfg['in_fft_n_heads'] != 1 or cfg['out_fft_n_heads'] != 1:
        raise NotImplementedError('single-head FFT blocks only (reference net_config)')
    fft('encoder', cfg['in_fft_n_layers'], cfg['in_fft_d_head'], cfg['in_fft_conv1d_filter_size'],
        cfg['in_fft_conv1d_kernel_size'], cfg['n_symbols'])
    if cfg['n_speakers'] > 1:
        P('speaker_emb.weight', cfg['n_speakers'], D)
    predictor('duration_predictor', cfg['dur_predictor_filter_size'], cfg['dur_predictor_kernel_size'],
              cfg['dur_predictor_n_layers'])
    fft('decoder', cfg['out_fft_n_layers'], cfg['out_fft_d_head'], cfg['out_fft_conv1d_filter_size'],
        cfg['out_fft_conv1d_kernel_size'], 0)
    predictor('pitch_predictor', cfg['pitch_predictor_filter_size'], cfg['pitch_predictor_kernel_size'],
              cfg['pitch_predictor_n_layers'])
    P('pitch_emb.weight', D, cfg.get('pitch_conditioning_formants', 1), cfg['pitch_embedding_kernel_size'])
    P('pitch_emb.bias', D)
    spec['pitch_mean'] = ((1,), True)
    spec['pitch_std'] = ((1,), True)
    if cfg['energy_conditioning']:
        predictor('energy_predictor', cfg['energy_predictor_filter_size'], cfg['energy_predictor_kernel_size'],
                  cfg['energy_predictor_n_layers'])
        P('energy_emb.weight', D, 1, cfg['energy_embedding_kernel_size'])
        P('energy_emb.bias', D)
    P('proj.weight', nm, cfg['out_fft_output_size'])
    P('proj.bias', nm)
    # training-only aligner, inert here
    for name, shp in [('attention.query_proj.0.conv', (2 * nm, nm, 3)), ('attention.query_proj.2.conv', (nm, 2 * nm, 1)),
                      ('attention.query_proj.4.conv', (nm, nm, 1)), ('attention.attn_proj', (1, 80, 1, 1)),
                      ('attention.key_proj.0.conv', (2 * D, D, 3)), ('attention.key_proj.2.conv', (80, 2 * D, 1))]:
        P(name + '.weight', *shp)
        P(name + '.bias', shp[0])
    return spec


class _Node(nn.Module):
    """Plain container; children / parameters are attached by dotted path."""


def _attach(root, dotted, tensor, is_buffer):
    parts = dotted.split('.')
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class FastPitch(nn.Module):
    def __init__(self, **cfg):
        super().__init__()
        self.cfg = dict(cfg)
        self.cfg.setdefault('pitch_conditioning_formants', 1)
        if self.cfg['pitch_conditioning_formants'] != 1:
            raise NotImplementedError('one pitch formant (reference net_config)')
        for name, (shape, is_buf) in parameter_spec(self.cfg).items():
            _attach(self, name, torch.zeros(shape), is_buf)
        D = self.cfg['symbols_embedding_dim']
        inv = 1 / (10000 ** (torch.arange(0.0, D, 2.0) / D))      # transformer.py:37-39
        self.encoder.pos_emb.inv_freq.copy_(inv)
        self.decoder.pos_emb.inv_freq.copy_(inv)
        self.energy_conditioning = bool(self.cfg['energy_conditioning'])
        self._handle = None
        self._handle_key = None
        self._ws = _lib.Workspace()
        self._state = _lib.Workspace()

    # ------------------------------------------------------------------ plumbing
    def _drop_handle(self):
        if getattr(self, '_handle', None) is not None:
            _lib.load().ttsb_fastpitch_destroy(self._handle)
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
        c = self.cfg
        if (c['in_fft_d_head'] != c['out_fft_d_head'] or c['in_fft_conv1d_filter_size'] != c['out_fft_conv1d_filter_size']
                or c['in_fft_conv1d_kernel_size'] != c['out_fft_conv1d_kernel_size']):
            raise NotImplementedError('encoder/decoder FFT blocks must share shapes')
        filt = {c['dur_predictor_filter_size'], c['pitch_predictor_filter_size']}
        kern = {c['dur_predictor_kernel_size'], c['pitch_predictor_kernel_size']}
        nlay = {c['dur_predictor_n_layers'], c['pitch_predictor_n_layers']}
        if c['energy_conditioning']:
            filt.add(c['energy_predictor_filter_size']); kern.add(c['energy_predictor_kernel_size'])
            nlay.add(c['energy_predictor_n_layers'])
        if len(filt) != 1 or len(kern) != 1 or nlay != {2}:
            raise NotImplementedError('predictors must share filter/kernel size and have 2 layers')
        cfg = _lib.FastpitchConfig()
        cfg.n_mel_channels = c['n_mel_channels']
        cfg.n_symbols = c['n_symbols']
        cfg.d_model = c['symbols_embedding_dim']
        cfg.n_layers_enc = c['in_fft_n_layers']
        cfg.n_layers_dec = c['out_fft_n_layers']
        cfg.d_head = c['in_fft_d_head']
        cfg.d_inner = c['in_fft_conv1d_filter_size']
        cfg.conv_kernel = c['in_fft_conv1d_kernel_size']
        cfg.pred_filter = filt.pop()
        cfg.pred_kernel = kern.pop()
        cfg.energy_conditioning = int(bool(c['energy_conditioning']))
        cfg.n_speakers = c['n_speakers']
        cfg.speaker_emb_weight = float(c['speaker_emb_weight'])
        named = {k: v.detach().float().cpu() for k, v in self.state_dict().items() if not k.startswith('attention.')}
        table, keep = _lib.tensor_table(named)
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().ttsb_fastpitch_create(ctypes.byref(cfg), table, len(named), device.index or 0,
                                                     ctypes.byref(handle)))
        del keep
        self._handle, self._handle_key = handle, device
        return handle

    def forward(self, *a, **k):
        raise NotImplementedError('training forward is out of scope; use infer()')

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def infer(self, inputs, pace=1.0, dur_tgt=None, pitch_tgt=None, energy_tgt=None, pitch_transform=None,
              max_duration=75, speaker=0, return_channel_last=False, taps=None, frame_len_hook=None):
        """Same contract as the reference (model.py:351-409): returns
        (mel_out [B,80,T], dec_lens [B] int64, dur_pred [B,L], pitch_pred [B,1,L], energy_pred [B,L] | None).
        With return_channel_last=True a 6th item is appended: the fp16 [B,T,128] mel for the vocoder.
        frame_len_hook (sharded batches, parallel.synthesize): maps this batch's max(dec_lens) to the frame count to
        decode with (>= it), so that a shard can keep the padded-frame condition of the global batch."""
        device = self.proj.weight.device
        if device.type != 'cuda':
            raise RuntimeError('tts_arabic_pytorch_b200 has no CPU path: move the model to a CUDA device')
        lib = _lib.load()
        ids = inputs.to(device=device, dtype=torch.int64).contiguous()
        B, L = ids.shape
        if L == 0:
            raise ValueError('empty token tensor')
        pad = ids == self.cfg['padding_idx']
        f32 = dict(dtype=torch.float32, device=device)
        # speaker (model.py:358-362): an int for the whole batch or a tensor of B ids; single-speaker checkpoints
        # ignore it like the reference (speaker_emb is None)
        spk, spk_ids = -1, None
        if self.cfg['n_speakers'] > 1:
            if torch.is_tensor(speaker) and speaker.numel() > 1:
                spk_ids = speaker.to(device=device, dtype=torch.int64).reshape(-1).contiguous()
                if spk_ids.numel() != B:
                    raise ValueError('speaker tensor must hold one id per utterance')
            else:
                spk = int(speaker)
                if not 0 <= spk < self.cfg['n_speakers']:
                    raise IndexError('speaker id %d out of range [0, %d)' % (spk, self.cfg['n_speakers']))
        with torch.cuda.device(device):
            handle = self._get_handle(device)
            stream = _lib.current_stream(device)
            state = self._state.get(lib.ttsb_fastpitch_state_bytes(handle, B, L), device)
            nb = lib.ttsb_fastpitch_workspace_bytes(handle, B, L, 0)
            ws = self._ws.get(nb, device)
            log_dur = torch.empty(B, L, **f32)
            pitch = torch.empty(B, L, **f32)
            _lib.check(lib.ttsb_fastpitch_encode(handle, _lib.ptr(ids), B, L, spk, _lib.ptr(spk_ids), _lib.ptr(log_dur),
                                                 _lib.ptr(pitch), _lib.ptr(state), _lib.ptr(ws), ws.numel(), stream))
            if taps is not None:            # parity tests: the encoder output before conditioning (model.py:364)
                enc = torch.empty(B, L, self.cfg['symbols_embedding_dim'], dtype=torch.float16, device=device)
                _lib.check(lib.ttsb_fastpitch_read_enc_out(handle, B, L, _lib.ptr(state), _lib.ptr(enc), stream))
                taps['enc_out'] = enc.float()
            pitch_pred = pitch[:, None, :]
            if pitch_transform is not None:                         # model.py:373-380
                if float(self.pitch_std[0]) == 0.0:
                    mean, std = 218.14, 67.24
                else:
                    mean, std = self.pitch_mean[0], self.pitch_std[0]
                pitch_pred = pitch_transform(pitch_pred, (~pad).sum(dim=1), mean, std)
            pitch_in = (pitch_pred if pitch_tgt is None else pitch_tgt).to(**f32).reshape(B, L).contiguous()
            e_tgt = None if energy_tgt is None else energy_tgt.to(**f32).reshape(B, L).contiguous()
            d_tgt = None if dur_tgt is None else dur_tgt.to(**f32).reshape(B, L).contiguous()
            dur_pred = torch.empty(B, L, **f32)
            energy_pred = torch.empty(B, L, **f32) if (self.energy_conditioning and e_tgt is None) else None
            dec_lens = torch.empty(B, dtype=torch.int64, device=device)
            summary = torch.empty(2 + B, dtype=torch.int32, device=device)
            _lib.check(lib.ttsb_fastpitch_condition(handle, B, L, _lib.ptr(log_dur), _lib.ptr(pitch_in), _lib.ptr(e_tgt),
                                                    _lib.ptr(d_tgt), float(pace), float(max_duration),
                                                    _lib.ptr(dur_pred), _lib.ptr(energy_pred), _lib.ptr(dec_lens),
                                                    _lib.ptr(summary), _lib.ptr(state), _lib.ptr(ws), ws.numel(), stream))
            # THE host sync of the call — the reference's own (model.py:76: max(dec_lens) sizes the decoder batch). The
            # input checks ride on it: ids were validated on the device by encode (no extra reductions or syncs)
            summ = summary.tolist()
            T, status = summ[0], summ[1]
            # every utterance's frame count came with the same read: the vocoder uses it to run each chunk of the padded
            # batch at the chunk's own longest utterance (Generator.run)
            dec_lens.host_list = summ[2:]
            if status & 1:
                raise IndexError('token id out of range [0, %d)' % self.cfg['n_symbols'])
            if status & 4:
                raise IndexError('speaker id out of range [0, %d)' % self.cfg['n_speakers'])
            if status & 2:
                raise ValueError('padding ids must be trailing and every utterance non-empty')
            if T <= 0:
                raise RuntimeError('all predicted durations are zero')
            if frame_len_hook is not None:
                T = max(T, int(frame_len_hook(T)))
            nb = lib.ttsb_fastpitch_workspace_bytes(handle, B, L, T)
            ws = self._ws.get(nb, device)
            mel = torch.empty(B, self.cfg['n_mel_channels'], T, **f32)
            mel_cl = torch.empty(B, T, 128, dtype=torch.float16, device=device) if return_channel_last else None
            _lib.check(lib.ttsb_fastpitch_decode(handle, B, L, T, _lib.ptr(mel), _lib.ptr(mel_cl), _lib.ptr(state),
                                                 _lib.ptr(ws), ws.numel(), stream))
        out = (mel, dec_lens, dur_pred, pitch_pred, energy_pred)
        return out + (mel_cl,) if return_channel_last else out
