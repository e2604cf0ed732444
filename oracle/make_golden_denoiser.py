"""Mints tests/golden/denoiser_small.npz by running the REAL reference Denoiser (/root/reference/vocoder/hifigan/
denoiser.py, imported in-process together with the reference Generator; build container only) on the seeded
synthetic HiFi-GAN checkpoint: the bias spectrum it derives from the zero-mel response, and its output for two
seeded waveforms at two strengths. The fixture pins oracle/denoiser_oracle.py (tests/test_oracle.py) and, through
the GPU tests, csrc/denoiser.cu.   Usage: python oracle/make_golden_denoiser.py
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
os.chdir(REF)

from tts_arabic_pytorch_b200.utils import synth  # noqa: E402


def main():
    from vocoder.hifigan.denoiser import Denoiser
    from vocoder.hifigan.env import AttrDict
    from vocoder.hifigan.models import Generator
    g = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    g.load_state_dict(synth.hifigan_state_dict(1235))
    g.eval()
    g.remove_weight_norm()
    d = Denoiser(g)
    gen = torch.Generator().manual_seed(7)
    a0 = torch.tanh(torch.randn(1, 20 * 256, generator=gen) * 0.3)
    a1 = torch.tanh(torch.randn(1, 7 * 256, generator=gen) * 0.1)
    out = {'bias_spec': d.bias_spec.numpy(), 'audio0': a0.numpy(), 'audio1': a1.numpy()}
    with torch.no_grad():
        for name, a in (('0', a0), ('1', a1)):
            for s in (0.005, 0.1):
                out['out%s_s%g' % (name, s)] = d(a, s).numpy()
    path = os.path.join(REPO, 'tests', 'golden', 'denoiser_small.npz')
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
