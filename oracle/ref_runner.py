"""Runs the UNMODIFIED reference modules vendored under oracle/_ref/ (oracle/make_ref.py) the way the reference's own
FastPitch2Wave.tts_batch does (models/fastpitch/networks.py:322-350): FastPitch.infer on the padded batch, then
Generator.forward once per utterance on mel[b, :, :dec_len]. TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.exists(os.path.join(_HERE, '_ref', 'fastpitch', 'model.py')) and \
        os.path.exists(os.path.join(_HERE, '_ref', 'hifigan', 'models.py'))


def build_models(fastpitch_sd, fastpitch_config, hifigan_sd, hifigan_config, device='cpu', dtype=torch.float32):
    """(FastPitch, Generator) of the reference, eval mode, weight norm removed as load_hifigan does (vocoder/__init__.py:3-20).
    `hifigan_sd` is the checkpoint's 'generator' dict WITH weight-norm parametrisations (the reference's format)."""
    from oracle._ref.fastpitch.model import FastPitch
    from oracle._ref.hifigan.env import AttrDict
    from oracle._ref.hifigan.models import Generator
    fp = FastPitch(**fastpitch_config)
    fp.load_state_dict(fastpitch_sd)
    voc = Generator(AttrDict(hifigan_config))
    voc.load_state_dict(hifigan_sd)
    voc.remove_weight_norm()
    return fp.eval().to(device=device, dtype=dtype), voc.eval().to(device=device, dtype=dtype)


@torch.no_grad()
def step(fp, voc, ids, batched_vocoder=False):
    """One reference step: returns (audio samples produced, list of waveforms)."""
    mel, dec_lens, *_ = fp.infer(ids)
    hop = 256
    if batched_vocoder:
        wav = voc(mel)
        return int(dec_lens.sum()) * hop, [wav]
    wavs = [voc(mel[b, :, :int(n)]) for b, n in enumerate(dec_lens.tolist())]
    return sum(int(w.numel()) for w in wavs), wavs
