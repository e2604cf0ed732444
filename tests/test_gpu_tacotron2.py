"""Tacotron2MS.infer on the GPU (BASELINE config 4 path) against the golden fixture minted from the
reference with injected prenet dropout masks, and against the CPU oracle at config-4 shape."""
import os

import numpy as np
import pytest
import torch

from tests import tolerances as tol
from tts_arabic_pytorch_b200.utils import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def taco():
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device')
    from tts_arabic_pytorch_b200.models.tacotron2.tacotron2_ms import Tacotron2MS
    m = Tacotron2MS(n_symbol=40, decoder_max_step=64)
    m.load_state_dict(synth.tacotron2_state_dict(1236))
    return m.eval().cuda()


def test_tacotron2_golden_injected_masks(taco, golden_dir):
    g = np.load(os.path.join(golden_dir, 'tacotron2_small.npz'))
    masks = torch.from_numpy(g['masks'])
    mel, mel_lens, align = taco.infer(torch.from_numpy(g['tokens']), torch.from_numpy(g['speaker_ids']),
                                      torch.from_numpy(g['lengths']), prenet_masks=masks)
    assert mel.shape == g['mel'].shape == (3, 80, 24)
    assert mel_lens.dtype == torch.int32 and mel_lens.tolist() == g['mel_lengths'].tolist()
    assert align.shape == g['alignments'].shape
    a = align.cpu().numpy()
    assert float(np.abs(a[1, :, 10:]).max()) == 0.0 and float(np.abs(a[2, :, 5:]).max()) == 0.0
    assert np.abs(a - g['alignments']).max() < tol.T2_ALIGN_ABS
    assert np.abs(mel.cpu().numpy() - g['mel']).max() < tol.T2_MEL_LINF


def test_tacotron2_vs_oracle_config4_shape(taco):
    """config 4: B=8, L=64, fixed decoder steps (gate held shut by the synthetic bias); 48 steps here so the
    CPU oracle stays in seconds."""
    from oracle import tacotron2_oracle as t2o
    gen = torch.Generator().manual_seed(4)
    tokens = torch.randint(1, 40, (8, 64), generator=gen)
    lens = torch.full((8,), 64, dtype=torch.long)
    masks = torch.rand(48, 2, 8, 256, generator=gen) > 0.5
    sd = synth.tacotron2_state_dict(1236)
    ref_mel, ref_lens, ref_al = t2o.tacotron2_infer(sd, tokens, torch.zeros(8, dtype=torch.long), lens,
                                                    prenet_masks=masks.float() * 2.0, max_steps=48)
    mel, mel_lens, align = taco.infer(tokens, torch.zeros(8, dtype=torch.long), lens, prenet_masks=masks)
    assert mel_lens.tolist() == ref_lens.tolist() == [48] * 8
    assert np.abs(align.cpu().numpy() - ref_al.numpy()).max() < tol.T2_ALIGN_ABS
    assert np.abs(mel.cpu().numpy() - ref_mel.numpy()).max() < tol.T2_MEL_LINF


def test_tacotron2_early_stop_and_random_masks(taco):
    """With the gate forced open after a few steps every utterance stops; output lengths follow the
    reference's bookkeeping (torchaudio:846-852); default masks come from the device RNG."""
    sd = synth.tacotron2_state_dict(1236, gate_bias=6.0)      # sigmoid(6) > 0.5 at step 0
    from tts_arabic_pytorch_b200.models.tacotron2.tacotron2_ms import Tacotron2MS
    m = Tacotron2MS(n_symbol=40, decoder_max_step=64)
    m.load_state_dict(sd)
    m = m.eval().cuda()
    tokens = torch.randint(1, 40, (2, 9))
    mel, mel_lens, align = m.infer(tokens)
    assert mel.shape == (2, 80, 1) and mel_lens.tolist() == [1, 1] and align.shape == (2, 1, 9)
    assert bool(torch.isfinite(mel).all())
