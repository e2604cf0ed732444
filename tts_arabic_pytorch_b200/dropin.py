"""Registers this package's modules under the reference's absolute import names so that code written
against nipponjo/tts-arabic-pytorch (`from models.fastpitch import FastPitch2Wave`,
`from vocoder import load_hifigan`, `import text`, `from utils import get_basic_config`) runs unmodified.

    import tts_arabic_pytorch_b200.dropin as dropin; dropin.install()
"""
import importlib
import sys

_ALIASES = {
    'models': 'tts_arabic_pytorch_b200.models',
    'models.fastpitch': 'tts_arabic_pytorch_b200.models.fastpitch',
    'models.fastpitch.networks': 'tts_arabic_pytorch_b200.models.fastpitch.networks',
    'models.fastpitch.fastpitch': 'tts_arabic_pytorch_b200.models.fastpitch.fastpitch',
    'models.fastpitch.fastpitch.model': 'tts_arabic_pytorch_b200.models.fastpitch.fastpitch.model',
    'models.tacotron2': 'tts_arabic_pytorch_b200.models.tacotron2',
    'models.tacotron2.networks': 'tts_arabic_pytorch_b200.models.tacotron2.networks',
    'models.tacotron2.tacotron2_ms': 'tts_arabic_pytorch_b200.models.tacotron2.tacotron2_ms',
    'vocoder': 'tts_arabic_pytorch_b200.vocoder',
    'vocoder.hifigan': 'tts_arabic_pytorch_b200.vocoder.hifigan',
    'vocoder.hifigan.models': 'tts_arabic_pytorch_b200.vocoder.hifigan.models',
    'vocoder.hifigan.env': 'tts_arabic_pytorch_b200.vocoder.hifigan.env',
    'vocoder.hifigan.denoiser': 'tts_arabic_pytorch_b200.vocoder.hifigan.denoiser',
    'text': 'tts_arabic_pytorch_b200.text',
    'text.symbols': 'tts_arabic_pytorch_b200.text.symbols',
    'utils': 'tts_arabic_pytorch_b200.utils',
}


def install(force=False):
    for alias, target in _ALIASES.items():
        if alias in sys.modules and not force:
            if sys.modules[alias].__name__ != target:
                raise RuntimeError("module '%s' is already imported from elsewhere; call install() first" % alias)
            continue
        sys.modules[alias] = importlib.import_module(target)
