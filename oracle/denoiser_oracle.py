"""TEST INFRASTRUCTURE — CPU restatement of the reference HiFi-GAN bias denoiser
(vocoder/hifigan/denoiser.py:29-72), pinned to the real reference through tests/golden/denoiser_small.npz
(oracle/make_golden_denoiser.py). Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

The reference builds torchaudio.transforms.Spectrogram(1024, hop_length=256, win_length=1024, power=None) and
InverseSpectrogram with the same geometry (denoiser.py:43-48); their defaults are a periodic hann window,
center=True, pad_mode='reflect', onesided, normalized=False — restated here on torch.stft / torch.istft.
"""
import torch

N_FFT, HOP = 1024, 256


def bias_spectrum(zero_mel_audio):
    """denoiser.py:51-64: magnitude spectrum of the vocoder's zero-mel response, first frame only -> [1,513,1]."""
    spec = _stft(zero_mel_audio.reshape(1, -1).float()).abs()
    return spec[:, :, 0][:, :, None]


def _stft(audio):
    return torch.stft(audio, N_FFT, hop_length=HOP, win_length=N_FFT, window=torch.hann_window(N_FFT),
                      center=True, pad_mode='reflect', normalized=False, onesided=True, return_complex=True)


def denoise(audio, bias_spec, strength=0.1):
    """denoiser.py:66-72. audio [1,N] fp32 -> [1, 256 * (N // 256)]."""
    spec = _stft(audio.float())
    mag = torch.clamp(spec.abs() - bias_spec * strength, min=0.0)
    den = mag * torch.exp(1j * spec.angle())
    return torch.istft(den, N_FFT, hop_length=HOP, win_length=N_FFT, window=torch.hann_window(N_FFT), center=True,
                       normalized=False, onesided=True)
