"""The batched Tacotron2 post-processing kernel (ttsb_tacotron2_postprocess: truncate_mel + replicate pad + bicubic
resize_mel for a whole batch in one launch) against the reference wrapper's outputs in tests/golden/tacotron2_wrapper.npz
(minted by oracle/make_golden_r2.py from models/tacotron2/networks.py:44-67,123-208 of the reference)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.t2_wrapper_stub import stub_infer_outputs

pytestmark = pytest.mark.gpu


def test_device_postprocessing_matches_reference_wrapper(golden_dir):
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device')
    from tts_arabic_pytorch_b200 import _lib
    from tts_arabic_pytorch_b200.models.tacotron2.networks import Tacotron2
    g = np.load(os.path.join(golden_dir, 'tacotron2_wrapper.npz'))
    meta = json.load(open(os.path.join(golden_dir, 'tacotron2_wrapper.json'), encoding='utf-8'))
    m = Tacotron2(checkpoint=None, n_symbol=40, arabic_in=False).cuda()
    calls = []

    def fake_infer(tokens, speaker_ids=None, lengths=None, **kw):
        calls.append(1)
        mel, lens, align = stub_infer_outputs(tokens.cpu(), None if lengths is None else lengths.cpu(), 100 + len(calls))
        return mel.cuda(), lens.cuda(), align.cuda()

    m.infer = fake_infer
    lib = _lib.load()
    for case in meta['cases']:
        name = case['name']
        calls.clear()
        n0 = lib.ttsb_launch_count()
        if name.startswith('batch'):
            mels = m.ttmel_batch(list(meta['lines']), **case['kw'])
            refs = [g['%s_mel%d' % (name, i)] for i in range(case['n'])]
        else:
            mels = [m.ttmel_single(meta['lines'][1], **case['kw'])]
            refs = [g[name + '_mel']]
        launches = lib.ttsb_launch_count() - n0
        assert launches == (0 if name == 'batch_raw' else 1), (name, launches)     # one launch per batch, none if nothing to do
        for i, (a, r) in enumerate(zip(mels, refs)):
            assert a.is_cuda and tuple(a.shape) == r.shape, (name, i, tuple(a.shape), r.shape)
            assert np.abs(a.cpu().numpy() - r).max() < 2e-5, (name, i)
