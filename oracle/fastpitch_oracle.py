"""FastPitch.infer, functional CPU restatement (test oracle).

Follows models/fastpitch/fastpitch/model.py (infer :351-409, regulate_len :68-90,
TemporalPredictor :129-133, ConvReLUNorm :54-57) and transformer.py (FFTransformer.forward
:207-225, TransformerLayer :172-177, MultiHeadAttn._forward :113-160, PositionwiseConvFF._forward
:72-90, PositionalEmbedding :41-48). Dropout is inactive in eval mode and omitted.
Weights are a flat dict with the reference state_dict keys.
"""
import torch
import torch.nn.functional as F


def _pos_emb(inv_freq, n, dtype):
    # transformer.py:41-48 : cat[sin(pos*inv_freq), cos(pos*inv_freq)]
    pos = torch.arange(n, dtype=dtype)
    s = pos[:, None] * inv_freq[None, :].to(dtype)
    return torch.cat([s.sin(), s.cos()], dim=1)[None]


def _attention(w, p, x, mask, d_head):
    # transformer.py:113-160, n_head == 1, post-LN
    qkv = F.linear(x, w[p + '.qkv_net.weight'], w[p + '.qkv_net.bias'])
    q, k, v = torch.chunk(qkv, 3, dim=2)
    score = torch.bmm(q, k.transpose(1, 2)) * (1.0 / d_head ** 0.5)
    key_pad = ~mask.squeeze(2)                                   # [B,S] True at padding
    score = score.masked_fill(key_pad[:, None, :], float('-inf'))
    prob = F.softmax(score, dim=2)
    vec = torch.bmm(prob, v)
    out = F.linear(vec, w[p + '.o_net.weight'])
    D = x.shape[-1]
    return F.layer_norm(x + out, (D,), w[p + '.layer_norm.weight'], w[p + '.layer_norm.bias'])


def _conv_ff(w, p, x, k):
    # transformer.py:83-88 : conv -> relu -> conv (no mask in between), post-LN
    t = x.transpose(1, 2)
    t = F.conv1d(t, w[p + '.CoreNet.0.weight'], w[p + '.CoreNet.0.bias'], padding=k // 2)
    t = F.relu(t)
    t = F.conv1d(t, w[p + '.CoreNet.2.weight'], w[p + '.CoreNet.2.bias'], padding=k // 2)
    t = t.transpose(1, 2)
    D = x.shape[-1]
    return F.layer_norm(x + t, (D,), w[p + '.layer_norm.weight'], w[p + '.layer_norm.bias'])


def fft_stack(w, prefix, x, mask, n_layers, d_head, k, conditioning=0, taps=None):
    # transformer.py:216-225
    D = x.shape[-1]
    pe = _pos_emb(w[prefix + '.pos_emb.inv_freq'], x.shape[1], x.dtype) * mask
    out = x + pe + conditioning
    if taps is not None:
        taps[prefix + '.in'] = out
    for i in range(n_layers):
        p = '%s.layers.%d' % (prefix, i)
        out = _attention(w, p + '.dec_attn', out, mask, d_head) * mask   # :173-174
        out = _conv_ff(w, p + '.pos_ff', out, k) * mask                  # :175-176
        if taps is not None:
            taps['%s.layer%d' % (prefix, i)] = out
    return out


def temporal_predictor(w, prefix, x, mask, n_layers, k):
    # model.py:129-133 with ConvReLUNorm :54-57
    out = (x * mask).transpose(1, 2)
    for i in range(n_layers):
        p = '%s.layers.%d' % (prefix, i)
        out = F.relu(F.conv1d(out, w[p + '.conv.weight'], w[p + '.conv.bias'], padding=k // 2))
        c = out.shape[1]
        out = F.layer_norm(out.transpose(1, 2), (c,), w[p + '.norm.weight'], w[p + '.norm.bias']).transpose(1, 2)
    out = out.transpose(1, 2)
    return F.linear(out, w[prefix + '.fc.weight'], w[prefix + '.fc.bias']) * mask


def regulate_len(durations, enc_out, pace=1.0):
    # model.py:68-90 (mel_max_len=None); written as a gather instead of the one-hot matmul
    reps = (durations.float() / pace + 0.5).long()
    dec_lens = reps.sum(dim=1)
    T = int(dec_lens.max())
    B, L, D = enc_out.shape
    out = enc_out.new_zeros(B, T, D)
    for b in range(B):
        idx = torch.repeat_interleave(torch.arange(L), reps[b])
        out[b, :idx.numel()] = enc_out[b, idx]
    return out, dec_lens


def fastpitch_infer(w, cfg, ids, pace=1.0, dur_tgt=None, pitch_tgt=None, energy_tgt=None,
                    pitch_transform=None, max_duration=75, speaker=0, dtype=torch.float32, taps=None):
    """Returns (mel [B,80,T], dec_lens [B], dur_pred [B,L], pitch_pred [B,1,L], energy_pred [B,L]);
    model.py:351-409."""
    w = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in w.items()}
    if cfg['n_speakers'] > 1:
        spk = w['speaker_emb.weight'][torch.full((ids.shape[0],), speaker, dtype=torch.long)][:, None, :]
        spk = spk * cfg['speaker_emb_weight']                      # :355-361
    else:
        spk = 0
    mask = (ids != cfg['padding_idx'])[:, :, None]                 # transformer.py:214
    x = w['encoder.word_emb.weight'][ids]
    enc = fft_stack(w, 'encoder', x, mask, cfg['in_fft_n_layers'], cfg['in_fft_d_head'],
                    cfg['in_fft_conv1d_kernel_size'], spk, taps)
    if taps is not None:
        taps['enc_out'] = enc
    log_dur = temporal_predictor(w, 'duration_predictor', enc, mask, cfg['dur_predictor_n_layers'],
                                 cfg['dur_predictor_kernel_size']).squeeze(-1)
    dur_pred = torch.clamp(torch.exp(log_dur) - 1, 0, max_duration)   # :368
    pitch_pred = temporal_predictor(w, 'pitch_predictor', enc, mask, cfg['pitch_predictor_n_layers'],
                                    cfg['pitch_predictor_kernel_size']).permute(0, 2, 1)
    if pitch_transform is not None:                                 # :373-380
        if w['pitch_std'][0] == 0.0:
            mean, std = 218.14, 67.24
        else:
            mean, std = w['pitch_mean'][0], w['pitch_std'][0]
        pitch_pred = pitch_transform(pitch_pred, mask.sum(dim=(1, 2)), mean, std)
    pk = cfg['pitch_embedding_kernel_size']
    pin = pitch_pred if pitch_tgt is None else pitch_tgt.to(dtype)
    enc = enc + F.conv1d(pin, w['pitch_emb.weight'], w['pitch_emb.bias'], padding=(pk - 1) // 2).transpose(1, 2)
    if cfg['energy_conditioning']:                                  # :389-397
        ek = cfg['energy_embedding_kernel_size']
        if energy_tgt is None:
            energy_pred = temporal_predictor(w, 'energy_predictor', enc, mask, cfg['energy_predictor_n_layers'],
                                             cfg['energy_predictor_kernel_size']).squeeze(-1)
            ein = energy_pred.unsqueeze(1)
        else:
            energy_pred = None
            ein = energy_tgt.to(dtype)
        enc = enc + F.conv1d(ein, w['energy_emb.weight'], w['energy_emb.bias'], padding=(ek - 1) // 2).transpose(1, 2)
    else:
        energy_pred = None
    if taps is not None:
        taps['enc_cond'] = enc
    reg, dec_lens = regulate_len(dur_pred if dur_tgt is None else dur_tgt, enc, pace)   # :401-403
    T = reg.shape[1]
    dmask = (torch.arange(T)[None, :] < dec_lens[:, None])[:, :, None]   # transformer.py:26-31,210
    dec = fft_stack(w, 'decoder', reg, dmask, cfg['out_fft_n_layers'], cfg['out_fft_d_head'],
                    cfg['out_fft_conv1d_kernel_size'], 0, taps)
    mel = F.linear(dec, w['proj.weight'], w['proj.bias']).permute(0, 2, 1)   # :406-408
    return mel, dec_lens, dur_pred, pitch_pred, energy_pred
