import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def hifigan_weights():
    """Folded (weight-norm removed) synthetic HiFi-GAN weights, same seed as the golden fixtures."""
    from tts_arabic_pytorch_b200.utils import synth
    return synth.fold_weight_norm(synth.hifigan_state_dict(1235))


@pytest.fixture(scope='session')
def fastpitch_weights_const4():
    from tts_arabic_pytorch_b200.utils import synth
    return synth.fastpitch_state_dict(1234, dur_mode='const4')


@pytest.fixture(scope='session')
def fastpitch_weights_random():
    from tts_arabic_pytorch_b200.utils import synth
    return synth.fastpitch_state_dict(1234, dur_mode='random')
