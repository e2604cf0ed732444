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


def test_persistent_decoder_equals_per_step_launches(taco):
    """The cooperative persistent kernel runs the same phase bodies as the six per-step launches: identical results."""
    import subprocess
    import sys
    import os
    gen = torch.Generator().manual_seed(11)
    tokens = torch.randint(1, 40, (5, 17), generator=gen)
    lens = torch.tensor([17, 15, 11, 6, 2])
    for b, n in enumerate(lens.tolist()):
        tokens[b, n:] = 0
    masks = torch.rand(40, 2, 5, 256, generator=gen) > 0.5
    spk = torch.tensor([0, 1, 2, 3, 4])
    mel, mel_lens, align = taco.infer(tokens, spk, lens, prenet_masks=masks)
    # the legacy path lives behind an environment switch read once per process
    code = '''
import sys, torch, numpy as np
sys.path.insert(0, %r)
from tts_arabic_pytorch_b200.models.tacotron2.tacotron2_ms import Tacotron2MS
from tts_arabic_pytorch_b200.utils import synth
m = Tacotron2MS(n_symbol=40, decoder_max_step=64); m.load_state_dict(synth.tacotron2_state_dict(1236)); m = m.eval().cuda()
d = torch.load(%r)
mel, ml, al = m.infer(d["tokens"], d["spk"], d["lens"], prenet_masks=d["masks"])
torch.save({"mel": mel.cpu(), "ml": ml.cpu(), "al": al.cpu()}, %r)
'''
    import tempfile
    tmp = tempfile.mkdtemp()
    torch.save({'tokens': tokens, 'spk': spk, 'lens': lens, 'masks': masks}, os.path.join(tmp, 'in.pt'))
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TTSB_T2_PERSISTENT='0')
    r = subprocess.run([sys.executable, '-c', code % (repo, os.path.join(tmp, 'in.pt'), os.path.join(tmp, 'out.pt'))],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = torch.load(os.path.join(tmp, 'out.pt'))
    assert mel_lens.cpu().tolist() == ref['ml'].tolist()
    assert float((mel.cpu() - ref['mel']).abs().max()) < 1e-5
    assert float((align.cpu() - ref['al']).abs().max()) < 1e-6


def test_batches_above_64_run_as_groups_and_bad_ids_raise(taco):
    gen = torch.Generator().manual_seed(12)
    tokens = torch.randint(1, 40, (70, 6), generator=gen)
    masks = torch.rand(4, 2, 70, 256, generator=gen) > 0.5
    mel, mel_lens, align = taco.infer(tokens, prenet_masks=masks)
    assert mel.shape == (70, 80, 4) and align.shape == (70, 4, 6) and mel_lens.tolist() == [4] * 70
    m2, _, _ = taco.infer(tokens[64:], prenet_masks=masks[:, :, 64:])
    assert float((mel[64:] - m2).abs().max()) < 1e-5
    bad = tokens[:2].clone()
    bad[1, 2] = 40
    with pytest.raises(IndexError):
        taco.infer(bad, prenet_masks=masks[:, :, :2])
    with pytest.raises(IndexError):
        taco.infer(tokens[:2], torch.tensor([0, 40]), prenet_masks=masks[:, :, :2])
    mel3, *_ = taco.infer(tokens[:2], prenet_masks=masks[:, :, :2])       # the device stays usable
    assert bool(torch.isfinite(mel3).all())
