"""Tacotron2MS.infer, functional CPU restatement (test oracle).

Follows models/tacotron2/tacotron2_ms.py:278-332 and the torchaudio 2.11.0 classes it instantiates
(torchaudio/models/tacotron2.py — a third-party dependency that is NOT vendored in the reference:
_Encoder.forward :396-418, _Prenet.forward :273-285, _Attention :203-255, _LocationLayer :150-168,
_Decoder.decode :611-684, _Decoder.infer :779-866, _Postnet.forward :330-346).

The reference's prenet applies dropout(p=0.5, training=True) on EVERY call, so two reference runs
differ; parity is only definable with injected masks (SURVEY.md §7 hard part 5): `prenet_masks`
[steps, 2, B, prenet_dim] holds the keep-masks (already scaled: 0.0 or 2.0). `make_golden_tacotron2.py`
injects the same masks into the real reference by patching F.dropout.
"""
import torch
import torch.nn.functional as F


def _bn_eval(x, w, p, eps=1e-5):
    return F.batch_norm(x, w[p + '.running_mean'], w[p + '.running_var'], w[p + '.weight'], w[p + '.bias'],
                        training=False, eps=eps)


def _lstm_dir(x, lens, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM over packed sequences: x [B,L,I], per-utterance lengths; outputs beyond
    an utterance's length are zero (pad_packed_sequence)."""
    B, L, _ = x.shape
    H = w_hh.shape[1]
    out = x.new_zeros(B, L, H)
    for b in range(B):
        n = int(lens[b])
        h = x.new_zeros(H)
        c = x.new_zeros(H)
        steps = range(n - 1, -1, -1) if reverse else range(n)
        for t in steps:
            g = w_ih @ x[b, t] + b_ih + w_hh @ h + b_hh
            i, f, gg, o = g.chunk(4)           # torch gate order: input, forget, cell, output
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            out[b, t] = h
    return out


def encoder(w, x, lens):
    # torchaudio:396-418 (eval: dropout inactive)
    for i in range(3):
        p = 'encoder.convolutions.%d' % i
        k = w[p + '.0.weight'].shape[2]
        x = F.relu(_bn_eval(F.conv1d(x, w[p + '.0.weight'], w[p + '.0.bias'], padding=(k - 1) // 2), w, p + '.1'))
    x = x.transpose(1, 2)
    fwd = _lstm_dir(x, lens, w['encoder.lstm.weight_ih_l0'], w['encoder.lstm.weight_hh_l0'],
                    w['encoder.lstm.bias_ih_l0'], w['encoder.lstm.bias_hh_l0'], False)
    bwd = _lstm_dir(x, lens, w['encoder.lstm.weight_ih_l0_reverse'], w['encoder.lstm.weight_hh_l0_reverse'],
                    w['encoder.lstm.bias_ih_l0_reverse'], w['encoder.lstm.bias_hh_l0_reverse'], True)
    return torch.cat([fwd, bwd], dim=2)


def _lstm_cell(x, h, c, w, p):
    g = F.linear(x, w[p + '.weight_ih'], w[p + '.bias_ih']) + F.linear(h, w[p + '.weight_hh'], w[p + '.bias_hh'])
    i, f, gg, o = g.chunk(4, dim=1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c), c


def tacotron2_infer(w, tokens, speaker_ids=None, lengths=None, prenet_masks=None, max_steps=3000,
                    gate_threshold=0.5, early_stopping=True, dtype=torch.float32):
    """-> (mel_postnet [B,80,T], mel_lengths int32 [B], alignments [B,T,L]); tacotron2_ms.py:278-332."""
    w = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in w.items()}
    B, L = tokens.shape
    if lengths is None:
        lengths = torch.full((B,), L, dtype=torch.long)
    if speaker_ids is None:
        speaker_ids = torch.zeros(B, dtype=torch.long)
    emb = w['embedding.weight'][tokens].transpose(1, 2)
    memory = encoder(w, emb, lengths)                                        # [B,L,512]
    if 'speaker_embedding.weight' in w:                                      # :315-320
        spk = w['speaker_embedding.weight'][speaker_ids][:, None, :].expand(-1, L, -1)
        memory = torch.cat([memory, spk], dim=2)
    pad = torch.arange(L)[None, :] >= lengths[:, None]                       # True at padding
    A = 'decoder.attention_layer.'
    processed_memory = F.linear(memory, w[A + 'memory_layer.weight'])        # torchaudio:600-605
    n_mels = w['decoder.linear_projection.weight'].shape[0]
    H = w['decoder.attention_rnn.weight_hh'].shape[1]
    ah, ac = memory.new_zeros(B, H), memory.new_zeros(B, H)
    dh, dc = memory.new_zeros(B, H), memory.new_zeros(B, H)
    aw, awc = memory.new_zeros(B, L), memory.new_zeros(B, L)
    ctx = memory.new_zeros(B, memory.shape[2])
    frame = memory.new_zeros(B, n_mels)
    mel_lens = torch.zeros(B, dtype=torch.int32)
    finished = torch.zeros(B, dtype=torch.bool)
    mels, aligns = [], []
    kloc = w[A + 'location_layer.location_conv.weight'].shape[2]
    for step in range(max_steps):
        x = frame
        for li in range(2):                                                  # prenet, torchaudio:283-285
            x = F.relu(F.linear(x, w['decoder.prenet.layers.%d.weight' % li]))
            if prenet_masks is not None:
                x = x * prenet_masks[step, li].to(dtype)
            else:
                x = F.dropout(x, 0.5, training=True)
        ah, ac = _lstm_cell(torch.cat([x, ctx], dim=1), ah, ac, w, 'decoder.attention_rnn')   # :654-656
        loc = F.conv1d(torch.stack([aw, awc], dim=1), w[A + 'location_layer.location_conv.weight'],
                       padding=(kloc - 1) // 2).transpose(1, 2)
        loc = F.linear(loc, w[A + 'location_layer.location_dense.weight'])
        q = F.linear(ah, w[A + 'query_layer.weight'])[:, None, :]
        energies = F.linear(torch.tanh(q + loc + processed_memory), w[A + 'v.weight']).squeeze(2)
        energies = energies.masked_fill(pad, float('-inf'))
        aw = F.softmax(energies, dim=1)
        ctx = torch.bmm(aw[:, None, :], memory).squeeze(1)
        awc = awc + aw
        dh, dc = _lstm_cell(torch.cat([ah, ctx], dim=1), dh, dc, w, 'decoder.decoder_rnn')    # :666-669
        hc = torch.cat([dh, ctx], dim=1)
        frame = F.linear(hc, w['decoder.linear_projection.weight'], w['decoder.linear_projection.bias'])
        gate = F.linear(hc, w['decoder.gate_layer.weight'], w['decoder.gate_layer.bias']).squeeze(1)
        mels.append(frame)
        aligns.append(aw)
        mel_lens[~finished] += 1                                             # torchaudio:846-849
        finished |= torch.sigmoid(gate) > gate_threshold
        if early_stopping and bool(finished.all()):
            break
    mel = torch.stack(mels, dim=2)                                           # [B,80,T]
    x = mel
    for i in range(5):                                                       # postnet, torchaudio:330-346
        p = 'postnet.convolutions.%d' % i
        k = w[p + '.0.weight'].shape[2]
        x = _bn_eval(F.conv1d(x, w[p + '.0.weight'], w[p + '.0.bias'], padding=(k - 1) // 2), w, p + '.1')
        if i < 4:
            x = torch.tanh(x)
    return mel + x, mel_lens, torch.stack(aligns, dim=1)
