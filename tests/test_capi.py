"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads here (no GPU) and
exports every symbol declared in include/ttsb200.h; the product fails loudly without CUDA."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(REPO, 'include', 'ttsb200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ttsb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from tts_arabic_pytorch_b200 import _lib
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), 'symbol %s declared in ttsb200.h but not exported' % name
    assert sorted(_lib.EXPORTS) == declared, 'ctypes signature table out of sync with the header'
    assert lib.ttsb_version() == 1


def test_runtime_switch_validation():
    from tts_arabic_pytorch_b200 import _lib
    lib = _lib.load()
    assert lib.ttsb_set_desc_mode(7) != 0
    assert b'desc mode' in lib.ttsb_last_error()
    mode = lib.ttsb_get_desc_mode()
    assert lib.ttsb_set_desc_mode(mode) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback():
    from tts_arabic_pytorch_b200.models.fastpitch.fastpitch.model import FastPitch
    from tts_arabic_pytorch_b200.utils import synth
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
    g = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    with pytest.raises(RuntimeError, match='no CPU path'):
        g(torch.zeros(80, 8))
    m = FastPitch(**synth.FASTPITCH_CONFIG)
    with pytest.raises(RuntimeError, match='no CPU path'):
        m.infer(torch.ones(1, 4, dtype=torch.long))


def test_checkpoint_formats_round_trip(tmp_path):
    """The loaders ingest exactly the reference's checkpoint formats."""
    from tts_arabic_pytorch_b200.utils import synth
    from tts_arabic_pytorch_b200.vocoder import load_hifigan
    fp, hg, cj = synth.write_checkpoints(str(tmp_path))
    g = load_hifigan(hg, cj)
    keys = set(g.state_dict().keys())
    assert 'conv_pre.weight' in keys and not any('parametrizations' in k for k in keys)
    folded = synth.fold_weight_norm(synth.hifigan_state_dict(1235))
    assert torch.allclose(g.state_dict()['ups.1.weight'], folded['ups.1.weight'], atol=1e-6)
    ck = torch.load(fp, map_location='cpu')
    assert set(ck.keys()) == {'model', 'config'}
