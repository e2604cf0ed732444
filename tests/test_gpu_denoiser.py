"""Parity of the CUDA denoiser (csrc/denoiser.cu through ttsb_denoiser_forward) against the reference formulation
(vocoder/hifigan/denoiser.py:66-72: torchaudio Spectrogram -> magnitude subtraction -> InverseSpectrogram, i.e.
torch.stft / torch.istft with n_fft = win = 1024, hop 256, periodic hann, center/reflect) evaluated once per utterance
like models/fastpitch/networks.py:343-344 does."""
import numpy as np
import pytest
import torch

from tts_arabic_pytorch_b200.utils import synth

pytestmark = pytest.mark.gpu

# fp32 transforms on both sides; the radix-2 shared-memory FFT and cuFFT differ by a few ulp per butterfly stage
DENOISE_ABS_TOL = 2e-5


@pytest.fixture(scope='module')
def denoiser():
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device (they never fall back to the CPU)')
    from tts_arabic_pytorch_b200.vocoder.hifigan.denoiser import Denoiser
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
    g = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    g.load_state_dict(synth.hifigan_state_dict(1235))
    g.eval()
    g.remove_weight_norm()
    g = g.cuda()
    return Denoiser(g).cuda()


def test_bias_spectrum_is_the_first_stft_frame_of_the_zero_mel_response(denoiser):
    dev = torch.device('cuda:0')
    denoiser._ensure_bias(dev)
    assert denoiser.bias_spec.shape == (1, 513, 1)
    assert bool(torch.isfinite(denoiser.bias_spec).all()) and float(denoiser.bias_spec.max()) > 0


@pytest.mark.parametrize('strength', [0.005, 0.1, 5.0])
def test_batched_cuda_denoiser_matches_per_utterance_torch(denoiser, strength):
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(3)
    lens = [40 * 256, 17 * 256, 3 * 256, 0, 40 * 256, 5 * 256]
    n_max = max(lens)
    wav = torch.zeros(len(lens), n_max)
    for b, n in enumerate(lens):
        wav[b, :n] = torch.tanh(torch.randn(n, generator=g) * 0.3)
    wav = wav.to(dev)
    out = denoiser.denoise_batch(wav, torch.tensor(lens), strength)
    torch.cuda.synchronize()
    assert out.shape == wav.shape and bool(torch.isfinite(out).all())
    for b, n in enumerate(lens):
        assert float(out[b, n:].abs().max()) == 0.0 if n < n_max else True
        if n == 0:
            continue
        ref = denoiser._forward_torch(wav[b:b + 1, :n], strength)[0]
        assert ref.numel() == n
        assert float((out[b, :n] - ref).abs().max()) < DENOISE_ABS_TOL * max(1.0, float(ref.abs().max()))


def test_forward_keeps_the_reference_call_shape(denoiser):
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(4)
    audio = torch.tanh(torch.randn(1, 24 * 256, generator=g) * 0.2).to(dev)     # vocoder output [1, N]
    y = denoiser(audio, 0.005)
    assert y.shape == audio.shape
    ref = denoiser._forward_torch(audio, 0.005)
    assert float((y - ref).abs().max()) < DENOISE_ABS_TOL
    # strength 0 must be the identity up to transform round-off (STFT/ISTFT with a hann window at 75 % overlap)
    y0 = denoiser(audio, 0.0)
    assert float((y0 - audio).abs().max()) < 1e-5
    with pytest.raises(RuntimeError, match='no CPU path'):
        denoiser(audio.cpu(), 0.005)


def test_cuda_denoiser_matches_reference_golden(denoiser, golden_dir):
    """The fixture holds outputs of the REAL reference Denoiser (oracle/make_golden_denoiser.py): same bias spectrum,
    same waveforms in, the CUDA kernels must reproduce its waveforms; the two utterances go through as one padded
    batch with their own lengths."""
    import os
    g = np.load(os.path.join(golden_dir, 'denoiser_small.npz'))
    dev = torch.device('cuda:0')
    denoiser._ensure_bias(dev)
    own_bias = denoiser.bias_spec.detach().cpu().numpy()
    # the bias spectrum computed from OUR generator's zero-mel response against the reference's (fp16 vocoder path)
    assert np.abs(own_bias - g['bias_spec']).max() < 2e-2 * max(1.0, float(np.abs(g['bias_spec']).max()))
    saved = denoiser.bias_spec
    try:
        denoiser.bias_spec = torch.from_numpy(g['bias_spec']).to(dev)        # isolate the transform kernels
        n0, n1 = g['audio0'].shape[1], g['audio1'].shape[1]
        wav = torch.zeros(2, n0)
        wav[0] = torch.from_numpy(g['audio0'][0])
        wav[1, :n1] = torch.from_numpy(g['audio1'][0])
        for s in (0.005, 0.1):
            out = denoiser.denoise_batch(wav.to(dev), torch.tensor([n0, n1]), s).cpu().numpy()
            assert np.abs(out[0] - g['out0_s%g' % s][0]).max() < DENOISE_ABS_TOL
            assert np.abs(out[1, :n1] - g['out1_s%g' % s][0]).max() < DENOISE_ABS_TOL
            assert np.abs(out[1, n1:]).max() == 0.0
    finally:
        denoiser.bias_spec = saved
