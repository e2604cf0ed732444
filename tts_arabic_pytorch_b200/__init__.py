"""B200-native batched text -> mel -> waveform inference (FastPitch + HiFi-GAN), drop-in for the
hot path of nipponjo/tts-arabic-pytorch.

Layout mirrors the reference import paths so that `tts_arabic_pytorch_b200.dropin.install()` can
alias them (`models.fastpitch`, `vocoder`, ...):

    tts_arabic_pytorch_b200.models.fastpitch   FastPitch, FastPitch2Wave, net_config
    tts_arabic_pytorch_b200.vocoder            load_hifigan
    tts_arabic_pytorch_b200.vocoder.hifigan    models.Generator, denoiser.Denoiser, env.AttrDict
    tts_arabic_pytorch_b200.csrc               CUDA kernels + C ABI (include/ttsb200.h)
"""
__version__ = '0.1.0'
