"""Tacotron2.ttmel_batch / ttmel_single wrapper logic (models/tacotron2/networks.py:123-208 of the reference: separator
insertion, text_collate_fn sorting, alignment-based truncate_mel + replicate pad, bicubic resize_mel) against
tests/golden/tacotron2_wrapper.npz, which oracle/make_golden_r2.py minted by running the REAL reference wrapper over a
deterministic stand-in for Tacotron2MS.infer (oracle/t2_wrapper_stub.py). The same stand-in drives this package's
wrapper here, so every difference is the wrapper's. Runs on the CPU: the wrapper's own arithmetic is host-side (and, on a
GPU, `tests/test_gpu_t2_post.py` holds the device kernel to the same fixture)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.t2_wrapper_stub import stub_infer_outputs


@pytest.fixture(scope='module')
def fixture(golden_dir):
    return (np.load(os.path.join(golden_dir, 'tacotron2_wrapper.npz')),
            json.load(open(os.path.join(golden_dir, 'tacotron2_wrapper.json'), encoding='utf-8')))


@pytest.fixture(scope='module')
def wrapper():
    from tts_arabic_pytorch_b200.models.tacotron2.networks import Tacotron2
    m = Tacotron2(checkpoint=None, n_symbol=40, arabic_in=False)
    calls = []

    def fake_infer(tokens, speaker_ids=None, lengths=None, **kw):
        calls.append((tokens.clone().cpu(), None if lengths is None else lengths.clone().cpu()))
        return stub_infer_outputs(tokens.cpu(), None if lengths is None else lengths.cpu(), 100 + len(calls))

    m.infer = fake_infer
    m._calls = calls
    return m


def test_batch_wrapper_matches_reference(fixture, wrapper):
    g, meta = fixture
    lines = meta['lines']
    for case in meta['cases']:
        name = case['name']
        if not name.startswith('batch'):
            continue
        wrapper._calls.clear()
        mels = wrapper.ttmel_batch(list(lines), **case['kw'])
        # the acoustic model saw exactly the reference's padded, length-sorted token batch (separator inserted)
        assert wrapper._calls[0][0].tolist() == g[name + '_tokens'].tolist()
        assert wrapper._calls[0][1].tolist() == g[name + '_lengths'].tolist()
        assert len(mels) == case['n']
        for i, m in enumerate(mels):
            ref = g['%s_mel%d' % (name, i)]
            assert tuple(m.shape) == ref.shape, (name, i)
            assert np.abs(m.cpu().numpy() - ref).max() < 1e-5, (name, i)


def test_single_wrapper_matches_reference(fixture, wrapper):
    g, meta = fixture
    line = meta['lines'][1]
    for case in meta['cases']:
        name = case['name']
        if not name.startswith('single'):
            continue
        wrapper._calls.clear()
        m = wrapper.ttmel_single(line, **case['kw'])
        assert wrapper._calls[0][0].tolist() == g[name + '_tokens'].tolist()
        ref = g[name + '_mel']
        assert tuple(m.shape) == ref.shape
        assert np.abs(m.cpu().numpy() - ref).max() < 1e-5


def test_truncate_and_resize_primitives():
    """Edge behaviour of the two helpers (networks.py:44-67): first index reaching 80 % of the maximum, three
    replicated frames, identity resize when the rounded length does not change."""
    from tts_arabic_pytorch_b200.models.tacotron2.networks import resize_mel, truncate_mel
    mel = torch.arange(40, dtype=torch.float32).reshape(4, 10)
    ps = torch.tensor([0, 0, .1, .2, .5, .79, .81, 1.0, .9, .3])
    out = truncate_mel(mel, ps)
    assert out.shape == (4, 6 + 3) and torch.equal(out[:, :6], mel[:, :6]) and torch.equal(out[:, 6:], mel[:, 5:6].expand(4, 3))
    assert resize_mel(mel, rate=1.0) is mel
    assert resize_mel(mel, rate=0.99).shape == (4, 10)      # int(10 / 0.99) == 10: returned unchanged
    assert resize_mel(mel, rate=0.5).shape == (4, 20)
