"""The reference-facing Python API on the GPU: FastPitch2Wave.tts / FastPitch.ttmel with text in,
CPU waveforms out (models/fastpitch/networks.py:197-253, 352-435), against the CPU oracle."""
import numpy as np
import pytest
import torch

from tests import tolerances as tol
from tts_arabic_pytorch_b200.utils import synth

pytestmark = pytest.mark.gpu

LINES = [">als~alAmu Ealaykum yA Sadiyqiy",
         "marHabAF bikum",
         ">aHrazat muntaxabAtu lbarAziyli fawzan"]


@pytest.fixture(scope='module')
def tts_model(tmp_path_factory):
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device')
    from tts_arabic_pytorch_b200.models.fastpitch import FastPitch2Wave
    d = tmp_path_factory.mktemp('ckpt')
    fp, hg, cj = synth.write_checkpoints(str(d), seed=1234)
    return FastPitch2Wave(fp, vocoder_sd=hg, vocoder_config=cj, arabic_in=False).cuda()


def _oracle(lines):
    from oracle import fastpitch_oracle as fpo
    from oracle import hifigan_oracle as hgo
    from tts_arabic_pytorch_b200 import text
    from tts_arabic_pytorch_b200.models.fastpitch.networks import text_collate_fn
    ids = [torch.LongTensor(text.tokens_to_ids(text.buckwalter_to_tokens(l, append_space=False))) for l in lines]
    padded, _, inverse = text_collate_fn(ids)
    fsd = synth.fastpitch_state_dict(1234)
    gsd = synth.fold_weight_norm(synth.hifigan_state_dict(1235))
    mel, dec_lens, *_ = fpo.fastpitch_infer(fsd, synth.FASTPITCH_CONFIG, padded)
    wavs = hgo.vocode_batch(gsd, synth.HIFIGAN_CONFIG, mel, dec_lens)
    return [wavs[r] for r in inverse.tolist()], [mel[r, :, :int(dec_lens[r])] for r in inverse.tolist()]


def _rel_rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / (np.sqrt(np.mean(b ** 2)) + 1e-12))


def test_tts_batch_matches_reference_semantics(tts_model):
    ref_wavs, ref_mels = _oracle(LINES)
    wavs = tts_model.tts(LINES, batch_size=3, denoise=0)
    assert isinstance(wavs, list) and len(wavs) == 3
    for w, r in zip(wavs, ref_wavs):
        assert w.device.type == 'cpu' and w.dtype == torch.float32 and w.dim() == 1
        assert w.shape == r.shape
        assert _rel_rms(w.numpy(), r.numpy()) < tol.E2E_WAV_REL_RMS
    mels = tts_model.model.ttmel(LINES, batch_size=3)
    for m, r in zip(mels, ref_mels):
        assert m.shape == r.shape and m.device.type == 'cuda'
        assert float((m.cpu() - r).abs().max()) < tol.MEL_LINF


def test_tts_single_string_and_small_batches(tts_model):
    ref_wavs, _ = _oracle(LINES[:1])      # batch of one: no padding anywhere
    w = tts_model.tts(LINES[0], denoise=0)
    assert isinstance(w, torch.Tensor) and w.dim() == 1 and w.device.type == 'cpu'
    assert _rel_rms(w.numpy(), ref_wavs[0].numpy()) < tol.E2E_WAV_REL_RMS
    w2, mel = tts_model.tts(LINES[0], denoise=0, return_mel=True)
    assert mel.shape[0] == 80 and w2.numel() == mel.shape[1] * 256
    # batch_size=2 over 3 lines: two padded batches, results in input order
    ws = tts_model.tts(LINES, batch_size=2, denoise=0)
    assert [x.numel() for x in ws] == [x.numel() for x in tts_model.tts(LINES, batch_size=1, denoise=0)]


def test_denoiser_default_strength_runs_and_stays_close(tts_model):
    w0 = tts_model.tts(LINES[1], denoise=0)
    w1 = tts_model.tts(LINES[1])          # denoise=0.005 default (networks.py:355)
    assert w1.shape == w0.shape and bool(torch.isfinite(w1).all())
    assert _rel_rms(w1.numpy(), w0.numpy()) < 0.2


def test_speed_changes_length(tts_model):
    n1 = tts_model.tts(LINES[0], denoise=0, speed=1.0).numel()
    n2 = tts_model.tts(LINES[0], denoise=0, speed=2.0).numel()
    assert n2 == n1 // 2      # const-4 durations: round(4/2) = 2 frames per token


def test_tacotron2wave_api(tmp_path_factory):
    """Tacotron2Wave.tts (models/tacotron2/networks.py:347-426): text in, 1-D fp32 CPU waveforms out, mel
    post-processing (separator insertion + alignment-based truncation) and speed resize on the way."""
    import os
    from tts_arabic_pytorch_b200.models.tacotron2 import Tacotron2Wave
    d = tmp_path_factory.mktemp('ckpt_t2')
    fp, hg, cj = synth.write_checkpoints(str(d), seed=1234)
    m = Tacotron2Wave(os.path.join(str(d), 'tacotron2.pth'), vocoder_sd=hg, vocoder_config=cj, arabic_in=False).cuda()
    m.model.decoder_max_step = 40          # synthetic weights never raise the stop gate
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        w = m.tts(LINES[0], denoise=0)
        ws = m.tts(LINES[:2], batch_size=2, denoise=0, speed=1.25)
        mel = m.model.ttmel(LINES[1], postprocess_mel=False)
    assert isinstance(w, torch.Tensor) and w.dim() == 1 and w.device.type == 'cpu' and w.numel() % 256 == 0
    assert bool(torch.isfinite(w).all()) and float(w.abs().max()) <= 1.0
    assert len(ws) == 2 and all(x.dim() == 1 and x.numel() % 256 == 0 for x in ws)
    assert mel.shape == (80, 40) and mel.device.type == 'cuda'
