"""Mints tests/golden/tacotron2_small.npz from the REAL reference (Tacotron2MS + torchaudio) with the
synthetic weights of utils/synth.py and injected prenet dropout masks (the reference's prenet dropout
is always on, so masks are injected by patching torchaudio's F.dropout for the [B,256] prenet calls).
Build-container only."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
os.chdir(REF)
from tts_arabic_pytorch_b200.utils import synth  # noqa: E402


def main():
    import torchaudio.models.tacotron2 as ta
    from models.tacotron2.tacotron2_ms import Tacotron2MS
    warnings.simplefilter('ignore')
    steps = 24
    sd = synth.tacotron2_state_dict(1236)
    m = Tacotron2MS(n_symbol=40, decoder_max_step=steps)
    m.load_state_dict(sd)
    m.eval()
    g = torch.Generator().manual_seed(0)
    tokens = torch.randint(1, 40, (3, 14), generator=g)
    lens = torch.tensor([14, 10, 5])
    tokens[1, 10:] = 0
    tokens[2, 5:] = 0
    spk = torch.tensor([0, 3, 5])
    masks = (torch.rand(steps, 2, 3, 256, generator=g) > 0.5).float() * 2.0
    cnt = {'i': 0}
    orig = F.dropout

    def fake(x, p=0.5, training=True, inplace=False):
        if training and p == 0.5 and tuple(x.shape) == (3, 256):
            i = cnt['i']
            cnt['i'] += 1
            return x * masks[i // 2, i % 2]
        return orig(x, p, training, inplace)

    ta.F.dropout = fake
    with torch.no_grad():
        mel, mel_lens, align = m.infer(tokens, spk, lens)
    ta.F.dropout = orig
    out = os.path.join(REPO, 'tests', 'golden', 'tacotron2_small.npz')
    np.savez_compressed(out, tokens=tokens.numpy(), lengths=lens.numpy(), speaker_ids=spk.numpy(),
                        masks=(masks > 0).numpy(), mel=mel.numpy(), mel_lengths=mel_lens.numpy(), alignments=align.numpy())
    print(out, os.path.getsize(out), mel.shape, mel_lens.tolist())


if __name__ == '__main__':
    main()
