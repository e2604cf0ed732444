"""HiFi-GAN bias denoiser (drop-in for vocoder/hifigan/denoiser.py:29-72).

Spectral subtraction: STFT(1024, hop 256, hann, centre/reflect) of the waveform, subtract
`strength x` the magnitude spectrum the vocoder emits for an all-zero mel (first frame only),
clamp at 0, inverse STFT with the original phase.

On a CUDA tensor the whole padded batch goes through the C ABI in one call (`ttsb_denoiser_forward`,
csrc/denoiser.cu: one CTA per frame keeps the frame in shared memory from the reflect-padded window load
through both 1024-point transforms, a second kernel does the overlap-add), each utterance at its own
length like the reference's per-utterance calls. `_forward_torch` (torch.stft / torch.istft) is the
load-time path for the bias spectrum and the cross-check of the GPU tests.
"""
import torch
import torch.nn as nn

from ... import _lib


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
        self._ws = _lib.Workspace()

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
    def _forward_torch(self, audio, strength=0.1):
        """The reference formulation (denoiser.py:66-72) on torch.stft / torch.istft; audio [B,N], rows full length."""
        self._ensure_bias(audio.device)
        spec = self._stft(audio)
        mag = torch.clamp(spec.abs() - self.bias_spec * strength, min=0.0)
        return self._istft(torch.polar(mag, spec.angle()))

    @torch.no_grad()
    def forward(self, audio, strength=0.1):
        """audio [1,N] (or [B,N], all rows full length) -> denoised, same shape."""
        if not audio.is_cuda:
            raise RuntimeError('Denoiser.forward: no CPU path in this build; move the module and its input to CUDA')
        squeeze = audio.dim() == 1
        a = audio[None] if squeeze else audio
        n = torch.full((a.shape[0],), a.shape[1], dtype=torch.int32, device=a.device)
        out = self.denoise_batch(a, n, strength)
        return out[0] if squeeze else out

    @torch.no_grad()
    def denoise_batch(self, wav, n_samples, strength):
        """Padded batch [B,N_max] with per-utterance sample counts: each utterance is processed at its
        own length (reflect padding and overlap-add envelope depend on it), samples beyond it are zero.
        One C-ABI call (two kernel launches) for the whole batch."""
        if not wav.is_cuda:
            raise RuntimeError('Denoiser.denoise_batch: no CPU path in this build')
        self._ensure_bias(wav.device)
        lib = _lib.load()
        wav = wav.float().contiguous()
        B, n_max = wav.shape
        n32 = torch.as_tensor(n_samples).to(device=wav.device, dtype=torch.int32).contiguous()
        bias = self.bias_spec.reshape(-1).float().contiguous()
        out = torch.empty_like(wav)
        nbytes = lib.ttsb_denoiser_workspace_bytes(B, n_max)
        ws = self._ws.get(nbytes, wav.device)
        _lib.check(lib.ttsb_denoiser_forward(_lib.ptr(wav), _lib.ptr(n32), B, n_max, _lib.ptr(bias), float(strength),
                                             _lib.ptr(out), _lib.ptr(ws), nbytes, _lib.current_stream(wav.device)))
        return out
