"""Deterministic synthetic checkpoints in the reference's own formats.

There are no pretrained weights in the build container, so parity tests, the
golden fixtures and bench.py all run on seeded random weights. The generator below depends on
nothing but torch's CPU RNG, so the SAME tensors are produced here, on the GPU box, and inside
`oracle/make_golden.py` (where they are loaded into the real reference modules).

Formats written (SURVEY.md §5 "Checkpoint / resume"):
  FastPitch : {'model': state_dict, 'config': net_config}       models/fastpitch/networks.py:52-60
  HiFi-GAN  : {'generator': state_dict with weight-norm params}  vocoder/__init__.py:15-16
"""
import math
from collections import OrderedDict

import torch

FASTPITCH_CONFIG = {
    'n_mel_channels': 80, 'n_symbols': 148, 'padding_idx': 0, 'symbols_embedding_dim': 384,
    'in_fft_n_layers': 6, 'in_fft_n_heads': 1, 'in_fft_d_head': 64,
    'in_fft_conv1d_kernel_size': 3, 'in_fft_conv1d_filter_size': 1536, 'in_fft_output_size': 384,
    'p_in_fft_dropout': 0.1, 'p_in_fft_dropatt': 0.1, 'p_in_fft_dropemb': 0.0,
    'out_fft_n_layers': 6, 'out_fft_n_heads': 1, 'out_fft_d_head': 64,
    'out_fft_conv1d_kernel_size': 3, 'out_fft_conv1d_filter_size': 1536, 'out_fft_output_size': 384,
    'p_out_fft_dropout': 0.1, 'p_out_fft_dropatt': 0.1, 'p_out_fft_dropemb': 0.0,
    'dur_predictor_kernel_size': 3, 'dur_predictor_filter_size': 256,
    'p_dur_predictor_dropout': 0.1, 'dur_predictor_n_layers': 2,
    'pitch_predictor_kernel_size': 3, 'pitch_predictor_filter_size': 256,
    'p_pitch_predictor_dropout': 0.1, 'pitch_predictor_n_layers': 2,
    'pitch_embedding_kernel_size': 3, 'n_speakers': 1, 'speaker_emb_weight': 1.0,
    'energy_predictor_kernel_size': 3, 'energy_predictor_filter_size': 256,
    'p_energy_predictor_dropout': 0.1, 'energy_predictor_n_layers': 2,
    'energy_conditioning': True, 'energy_embedding_kernel_size': 3,
}

HIFIGAN_CONFIG = {
    'resblock': '1', 'upsample_rates': [8, 8, 2, 2], 'upsample_kernel_sizes': [16, 16, 4, 4],
    'upsample_initial_channel': 512, 'resblock_kernel_sizes': [3, 7, 11],
    'resblock_dilation_sizes': [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    'num_mels': 80, 'sampling_rate': 22050, 'hop_size': 256, 'n_fft': 1024, 'win_size': 1024,
}


class _Rng:
    def __init__(self, seed):
        self.g = torch.Generator(device='cpu')
        self.g.manual_seed(seed)

    def normal(self, shape, std):
        return torch.randn(shape, generator=self.g, dtype=torch.float32) * std


def hifigan_state_dict(seed=1234, cfg=None, weight_norm=True, gain=1.0):
    """HiFi-GAN V1 generator weights. `weight_norm=True` emits the parametrised keys
    (`...parametrizations.weight.original0/1`) that the reference loader expects; False emits the
    folded `.weight` keys (what remove_weight_norm() leaves)."""
    cfg = cfg or HIFIGAN_CONFIG
    r = _Rng(seed)
    sd = OrderedDict()

    def put(name, w, b):
        if weight_norm:
            # w = g * v / ||v|| with the norm over every dim but 0 (torch weight_norm default dim=0)
            v = w
            g = v.flatten(1).norm(dim=1).reshape([-1] + [1] * (v.dim() - 1))
            sd[name + '.bias'] = b
            sd[name + '.parametrizations.weight.original0'] = g
            sd[name + '.parametrizations.weight.original1'] = v
        else:
            sd[name + '.weight'] = w
            sd[name + '.bias'] = b

    c0 = cfg['upsample_initial_channel']
    nm = cfg['num_mels']
    # mel inputs live around -5 +- 2 (log-mel); keep conv_pre output O(1)
    put('conv_pre', r.normal((c0, nm, 7), 0.35 / math.sqrt(nm * 7)), r.normal((c0,), 0.1))
    cin = c0
    for i, (u, k) in enumerate(zip(cfg['upsample_rates'], cfg['upsample_kernel_sizes'])):
        cout = cin // 2
        # each output sample sees k/u taps of cin channels
        put('ups.%d' % i, r.normal((cin, cout, k), 1.4 / math.sqrt(cin * k / u)), r.normal((cout,), 0.05))
        for j, (rk, dils) in enumerate(zip(cfg['resblock_kernel_sizes'], cfg['resblock_dilation_sizes'])):
            idx = i * len(cfg['resblock_kernel_sizes']) + j
            for p in range(len(dils)):
                std = gain / math.sqrt(cout * rk)
                put('resblocks.%d.convs1.%d' % (idx, p), r.normal((cout, cout, rk), std), r.normal((cout,), 0.05))
                put('resblocks.%d.convs2.%d' % (idx, p), r.normal((cout, cout, rk), std), r.normal((cout,), 0.05))
        cin = cout
    put('conv_post', r.normal((1, cin, 7), 0.2 / math.sqrt(cin * 7)), r.normal((1,), 0.01))
    return sd


def fold_weight_norm(sd):
    """state_dict with weight-norm parametrisation (new `parametrizations.weight.original0/1` or old
    `weight_g/weight_v` keys) -> plain `.weight` keys. Pure tensor math, no nn.Module."""
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith('.parametrizations.weight.original1') or k.endswith('.weight_v'):
            base = k[:-len('.parametrizations.weight.original1')] if 'parametrizations' in k else k[:-len('.weight_v')]
            gk = base + ('.parametrizations.weight.original0' if 'parametrizations' in k else '.weight_g')
            g = sd[gk].float()
            vf = v.float()
            norm = vf.flatten(1).norm(dim=1).reshape([-1] + [1] * (vf.dim() - 1))
            out[base + '.weight'] = vf * (g / norm)
        elif k.endswith('.parametrizations.weight.original0') or k.endswith('.weight_g'):
            continue
        else:
            out[k] = v
    return out


def fastpitch_state_dict(seed=1234, cfg=None, dur_mode='const4'):
    """FastPitch weights with every key of the reference module's state_dict (including the
    training-only `attention.*` aligner tensors, which strict loading requires: SURVEY.md §2 row 4).

    dur_mode 'const4': duration head bias = ln 5, weight = 0 -> every valid token gets exactly 4
    frames (SURVEY.md §8c calibration); 'random': small random head -> varied durations."""
    cfg = cfg or FASTPITCH_CONFIG
    r = _Rng(seed)
    sd = OrderedDict()
    D = cfg['symbols_embedding_dim']

    def lin(name, n_out, n_in, bias=True, gain=1.0):
        sd[name + '.weight'] = r.normal((n_out, n_in), gain / math.sqrt(n_in))
        if bias:
            sd[name + '.bias'] = r.normal((n_out,), 0.05)

    def conv(name, n_out, n_in, k, gain=1.0):
        sd[name + '.weight'] = r.normal((n_out, n_in, k), gain / math.sqrt(n_in * k))
        sd[name + '.bias'] = r.normal((n_out,), 0.05)

    def ln(name, n):
        sd[name + '.weight'] = 1.0 + r.normal((n,), 0.1)
        sd[name + '.bias'] = r.normal((n,), 0.1)

    def fft(prefix, n_layers, d_head, d_inner, k, embed):
        if embed:
            sd[prefix + '.word_emb.weight'] = r.normal((cfg['n_symbols'], D), 0.7)
        sd[prefix + '.pos_emb.inv_freq'] = 1 / (10000 ** (torch.arange(0.0, D, 2.0) / D))
        for i in range(n_layers):
            p = '%s.layers.%d' % (prefix, i)
            lin(p + '.dec_attn.qkv_net', 3 * d_head, D, gain=1.5)
            lin(p + '.dec_attn.o_net', D, d_head, bias=False)
            ln(p + '.dec_attn.layer_norm', D)
            conv(p + '.pos_ff.CoreNet.0', d_inner, D, k, gain=1.2)
            conv(p + '.pos_ff.CoreNet.2', D, d_inner, k, gain=1.2)
            ln(p + '.pos_ff.layer_norm', D)

    def predictor(prefix, filt, k, n_layers):
        for i in range(n_layers):
            conv('%s.layers.%d.conv' % (prefix, i), filt, D if i == 0 else filt, k, gain=1.3)
            ln('%s.layers.%d.norm' % (prefix, i), filt)
        lin(prefix + '.fc', 1, filt)

    fft('encoder', cfg['in_fft_n_layers'], cfg['in_fft_d_head'], cfg['in_fft_conv1d_filter_size'],
        cfg['in_fft_conv1d_kernel_size'], True)
    if cfg['n_speakers'] > 1:
        sd['speaker_emb.weight'] = r.normal((cfg['n_speakers'], D), 0.3)
    predictor('duration_predictor', cfg['dur_predictor_filter_size'], cfg['dur_predictor_kernel_size'],
              cfg['dur_predictor_n_layers'])
    fft('decoder', cfg['out_fft_n_layers'], cfg['out_fft_d_head'], cfg['out_fft_conv1d_filter_size'],
        cfg['out_fft_conv1d_kernel_size'], False)
    predictor('pitch_predictor', cfg['pitch_predictor_filter_size'], cfg['pitch_predictor_kernel_size'],
              cfg['pitch_predictor_n_layers'])
    conv('pitch_emb', D, 1, cfg['pitch_embedding_kernel_size'], gain=0.5)
    sd['pitch_mean'] = torch.zeros(1)
    sd['pitch_std'] = torch.zeros(1)
    if cfg['energy_conditioning']:
        predictor('energy_predictor', cfg['energy_predictor_filter_size'],
                  cfg['energy_predictor_kernel_size'], cfg['energy_predictor_n_layers'])
        conv('energy_emb', D, 1, cfg['energy_embedding_kernel_size'], gain=0.5)
    lin('proj', cfg['n_mel_channels'], D)
    sd['proj.bias'] = sd['proj.bias'] - 5.0   # log-mel range

    if dur_mode == 'const4':
        sd['duration_predictor.fc.weight'] = torch.zeros_like(sd['duration_predictor.fc.weight'])
        sd['duration_predictor.fc.bias'] = torch.full((1,), math.log(5.0))
    elif dur_mode == 'random':
        sd['duration_predictor.fc.weight'] = sd['duration_predictor.fc.weight'] * 0.35
        sd['duration_predictor.fc.bias'] = torch.full((1,), math.log(4.3))
    else:
        raise ValueError(dur_mode)

    # training-only aligner (models/fastpitch/fastpitch/attention.py:85-133), shapes as constructed
    # by FastPitch.__init__ (model.py:234-236): ConvAttention(80, 0, 384, use_query_proj=True,
    # align_query_enc_type='3xconv')
    nm = cfg['n_mel_channels']
    att = [('attention.key_proj.0.conv', (2 * D, D, 3)), ('attention.key_proj.2.conv', (80, 2 * D, 1)),
           ('attention.query_proj.0.conv', (2 * nm, nm, 3)), ('attention.query_proj.2.conv', (nm, 2 * nm, 1)),
           ('attention.query_proj.4.conv', (nm, nm, 1)), ('attention.attn_proj', (1, 80, 1, 1))]
    for name, shp in att:
        sd[name + '.weight'] = r.normal(shp, 0.02)
        sd[name + '.bias'] = torch.zeros(shp[0])
    return sd


def tacotron2_state_dict(seed=1236, n_symbol=40, num_speakers=40, gate_bias=-3.0):
    """Tacotron2MS weights with every key of the reference module (tacotron2_ms.py:119-207 +
    torchaudio _Encoder/_Decoder/_Postnet). `gate_bias` keeps the stop gate shut for synthetic tests
    (decoding runs a fixed number of steps)."""
    r = _Rng(seed)
    sd = OrderedDict()
    E, H, S, P, Amem, Ahid, NF, KL, M = 512, 1024, 128, 256, 640, 128, 32, 31, 80

    def bn(p, n):
        sd[p + '.weight'] = 1.0 + r.normal((n,), 0.1)
        sd[p + '.bias'] = r.normal((n,), 0.1)
        sd[p + '.running_mean'] = r.normal((n,), 0.1)
        sd[p + '.running_var'] = 1.0 + r.normal((n,), 0.1).abs()
        sd[p + '.num_batches_tracked'] = torch.tensor(100)

    sd['embedding.weight'] = r.normal((n_symbol, E), 0.5)
    for i in range(3):
        sd['encoder.convolutions.%d.0.weight' % i] = r.normal((E, E, 5), 1.2 / math.sqrt(E * 5))
        sd['encoder.convolutions.%d.0.bias' % i] = r.normal((E,), 0.05)
        bn('encoder.convolutions.%d.1' % i, E)
    for suf in ('', '_reverse'):
        sd['encoder.lstm.weight_ih_l0' + suf] = r.normal((4 * E // 2, E), 1.0 / math.sqrt(E))
        sd['encoder.lstm.weight_hh_l0' + suf] = r.normal((4 * E // 2, E // 2), 1.0 / math.sqrt(E // 2))
        sd['encoder.lstm.bias_ih_l0' + suf] = r.normal((4 * E // 2,), 0.05)
        sd['encoder.lstm.bias_hh_l0' + suf] = r.normal((4 * E // 2,), 0.05)
    sd['decoder.prenet.layers.0.weight'] = r.normal((P, M), 1.0 / math.sqrt(M))
    sd['decoder.prenet.layers.1.weight'] = r.normal((P, P), 1.0 / math.sqrt(P))
    for name, n_in in (('decoder.attention_rnn', P + Amem), ('decoder.decoder_rnn', H + Amem)):
        sd[name + '.weight_ih'] = r.normal((4 * H, n_in), 1.0 / math.sqrt(n_in))
        sd[name + '.weight_hh'] = r.normal((4 * H, H), 1.0 / math.sqrt(H))
        sd[name + '.bias_ih'] = r.normal((4 * H,), 0.05)
        sd[name + '.bias_hh'] = r.normal((4 * H,), 0.05)
        if name == 'decoder.attention_rnn':
            A = 'decoder.attention_layer.'
            sd[A + 'query_layer.weight'] = r.normal((Ahid, H), 1.0 / math.sqrt(H))
            sd[A + 'memory_layer.weight'] = r.normal((Ahid, Amem), 1.0 / math.sqrt(Amem))
            sd[A + 'v.weight'] = r.normal((1, Ahid), 3.0 / math.sqrt(Ahid))
            sd[A + 'location_layer.location_conv.weight'] = r.normal((NF, 2, KL), 1.0 / math.sqrt(2 * KL))
            sd[A + 'location_layer.location_dense.weight'] = r.normal((Ahid, NF), 1.0 / math.sqrt(NF))
    sd['decoder.linear_projection.weight'] = r.normal((M, H + Amem), 1.0 / math.sqrt(H + Amem))
    sd['decoder.linear_projection.bias'] = r.normal((M,), 0.1)
    sd['decoder.gate_layer.weight'] = r.normal((1, H + Amem), 0.2 / math.sqrt(H + Amem))
    sd['decoder.gate_layer.bias'] = torch.full((1,), float(gate_bias))
    dims = [M, 512, 512, 512, 512, M]
    for i in range(5):
        sd['postnet.convolutions.%d.0.weight' % i] = r.normal((dims[i + 1], dims[i], 5), 1.0 / math.sqrt(dims[i] * 5))
        sd['postnet.convolutions.%d.0.bias' % i] = r.normal((dims[i + 1],), 0.05)
        bn('postnet.convolutions.%d.1' % i, dims[i + 1])
    if num_speakers > 1:
        sd['speaker_embedding.weight'] = r.normal((num_speakers, S), 0.3)
    return sd


def write_checkpoints(directory, seed=1234, dur_mode='const4'):
    """Writes fastpitch.pth + hifigan.pth (+ config.json) in the reference's formats; returns paths."""
    import json
    import os
    os.makedirs(directory, exist_ok=True)
    fp = os.path.join(directory, 'fastpitch.pth')
    hg = os.path.join(directory, 'hifigan.pth')
    cj = os.path.join(directory, 'config.json')
    torch.save({'model': fastpitch_state_dict(seed, dur_mode=dur_mode), 'config': dict(FASTPITCH_CONFIG)}, fp)
    torch.save({'generator': hifigan_state_dict(seed + 1)}, hg)
    with open(cj, 'w') as f:
        json.dump(HIFIGAN_CONFIG, f)
    # Tacotron2 checkpoint format: {'model': state_dict} (models/tacotron2/networks.py:85-98)
    torch.save({'model': tacotron2_state_dict(seed + 2)}, os.path.join(directory, 'tacotron2.pth'))
    return fp, hg, cj
