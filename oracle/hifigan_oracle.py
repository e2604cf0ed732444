"""HiFi-GAN V1 generator, functional CPU restatement (test oracle).

Follows vocoder/hifigan/models.py: Generator.forward :111-127, ResBlock1.forward :46-53,
get_padding :18-19, LRELU_SLOPE :11; weight-norm folding as done by remove_weight_norm()
(vocoder/__init__.py:19). Weights are a flat dict with the folded `.weight/.bias` keys.
"""
import torch
import torch.nn.functional as F

STAGE_SLOPE = 0.1     # LRELU_SLOPE, models.py:11
FINAL_SLOPE = 0.01    # F.leaky_relu default, models.py:123


def _same_pad(k, d):
    return (k * d - d) // 2   # models.py:18-19


def resblock1(w, prefix, x, k, dilations):
    """models.py:46-53: for each dilation: x = x + conv_d1(lrelu(conv_d(lrelu(x))))"""
    for p, d in enumerate(dilations):
        t = F.leaky_relu(x, STAGE_SLOPE)
        t = F.conv1d(t, w['%s.convs1.%d.weight' % (prefix, p)], w['%s.convs1.%d.bias' % (prefix, p)],
                     padding=_same_pad(k, d), dilation=d)
        t = F.leaky_relu(t, STAGE_SLOPE)
        t = F.conv1d(t, w['%s.convs2.%d.weight' % (prefix, p)], w['%s.convs2.%d.bias' % (prefix, p)],
                     padding=_same_pad(k, 1))
        x = x + t
    return x


def generator_forward(w, cfg, mel, dtype=torch.float32, taps=None):
    """mel [B,80,T] (or [80,T]) -> wav [B,1,256T] (or [1,256T]); models.py:111-127.
    `taps` (optional dict) receives intermediate tensors for layer-level tests."""
    w = {k: v.to(dtype) for k, v in w.items()}
    x = mel.to(dtype)
    squeeze = x.dim() == 2
    if squeeze:
        x = x[None]
    x = F.conv1d(x, w['conv_pre.weight'], w['conv_pre.bias'], padding=3)
    if taps is not None:
        taps['conv_pre'] = x
    nk = len(cfg['resblock_kernel_sizes'])
    for i, (u, k) in enumerate(zip(cfg['upsample_rates'], cfg['upsample_kernel_sizes'])):
        x = F.leaky_relu(x, STAGE_SLOPE)
        x = F.conv_transpose1d(x, w['ups.%d.weight' % i], w['ups.%d.bias' % i], stride=u, padding=(k - u) // 2)
        if taps is not None:
            taps['ups.%d' % i] = x
        acc = None
        for j, (rk, dil) in enumerate(zip(cfg['resblock_kernel_sizes'], cfg['resblock_dilation_sizes'])):
            r = resblock1(w, 'resblocks.%d' % (i * nk + j), x, rk, dil)
            acc = r if acc is None else acc + r
        x = acc / nk
        if taps is not None:
            taps['stage.%d' % i] = x
    x = F.leaky_relu(x, FINAL_SLOPE)
    x = F.conv1d(x, w['conv_post.weight'], w['conv_post.bias'], padding=3)
    x = torch.tanh(x)
    return x[0] if squeeze else x


def vocode_batch(w, cfg, mel, lens, dtype=torch.float32):
    """What FastPitch2Wave.tts_batch does with a padded mel batch (models/fastpitch/networks.py:
    340-345): slice each utterance to its own length and run the generator UNBATCHED.
    Returns a list of 1-D waveforms."""
    out = []
    for b in range(mel.shape[0]):
        out.append(generator_forward(w, cfg, mel[b, :, :int(lens[b])], dtype)[0])
    return out
