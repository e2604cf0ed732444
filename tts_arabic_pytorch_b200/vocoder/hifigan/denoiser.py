"""HiFi-GAN bias denoiser (drop-in for vocoder/hifigan/denoiser.py:29-72).

Spectral subtraction: STFT(1024, hop 256, hann, centre/reflect) of the waveform, subtract
`strength x` the magnitude spectrum the vocoder emits for an all-zero mel (first frame only),
clamp at 0, inverse STFT with the original phase. The transforms run through torch.stft/istft
(cuFFT) — this component is adjacent to the hot path, not yet a custom kernel.
"""
import torch
import torch.nn as nn


class Denoiser(nn.Module):
    def __init__(self, hifigan, filter_length=1024, n_overlap=4, win_length=1024, mode='zeros', **infer_kw):
        super().__init__()
        self.n_fft = filter_length
        self.hop = filter_length // n_overlap
        self.win_length = win_length
        self.mode = mode
        self._vocoder = [hifigan]          # not registered: the vocoder is owned by the caller
        self.register_buffer('window', torch.hann_window(win_length), persistent=False)
        # the reference computes the bias spectrum eagerly on the vocoder's device; on a CPU-resident
        # module that is impossible here (no CPU path), so it is materialised on first use
        self.register_buffer('bias_spec', torch.zeros(1, filter_length // 2 + 1, 1))
        self._bias_ready = False

    def _stft(self, audio):
        return torch.stft(audio.float(), self.n_fft, hop_length=self.hop, win_length=self.win_length,
                          window=self.window.to(audio.device), center=True, pad_mode='reflect', normalized=False,
                          onesided=True, return_complex=True)

    def _istft(self, spec):
        return torch.istft(spec, self.n_fft, hop_length=self.hop, win_length=self.win_length,
                           window=self.window.to(spec.device), center=True, normalized=False, onesided=True)

    @torch.no_grad()
    def _ensure_bias(self, device):
        if self._bias_ready and self.bias_spec.device == device:
            return
        voc = self._vocoder[0]
        init = {'zeros': torch.zeros, 'normal': torch.randn}[self.mode]
        mel = init((1, 80, 88), dtype=torch.float32, device=device)   # denoiser.py:51
        bias_audio = voc(mel).float().reshape(1, -1)
        spec = self._stft(bias_audio).abs()
        self.bias_spec = spec[:, :, 0][:, :, None].to(device)
        self._bias_ready = True

    @torch.no_grad()
    def forward(self, audio, strength=0.1):
        """audio [1,N] (or [B,N], all rows full length) -> denoised, same shape."""
        self._ensure_bias(audio.device)
        spec = self._stft(audio)
        mag = torch.clamp(spec.abs() - self.bias_spec * strength, min=0.0)
        return self._istft(torch.polar(mag, spec.angle()))

    @torch.no_grad()
    def denoise_batch(self, wav, n_samples, strength):
        """Padded batch [B,N_max] with per-utterance sample counts: each utterance is processed at its
        own length (reflect padding and overlap-add envelope depend on it), result re-padded."""
        out = torch.zeros_like(wav)
        for b, n in enumerate(n_samples.tolist()):
            if n > 0:
                out[b, :n] = self.forward(wav[b:b + 1, :n], strength)[0, :n]
        return out
