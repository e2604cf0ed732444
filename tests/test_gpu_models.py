"""Parity of the CUDA path (through the Python fronts -> C ABI -> sm_100a kernels) against the
golden fixtures minted from the reference and against the CPU oracle on seeded inputs."""
import os

import numpy as np
import pytest
import torch

from tests import tolerances as tol
from tts_arabic_pytorch_b200.utils import synth

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device (they never fall back to the CPU)')
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def vocoder():
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
    g = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    g.load_state_dict(synth.hifigan_state_dict(1235))
    g.eval()
    g.remove_weight_norm()
    return g.to(_dev())


def _fastpitch(dur_mode):
    from tts_arabic_pytorch_b200.models.fastpitch.fastpitch.model import FastPitch
    m = FastPitch(**synth.FASTPITCH_CONFIG)
    m.load_state_dict(synth.fastpitch_state_dict(1234, dur_mode=dur_mode))
    return m.eval().to(_dev())


@pytest.fixture(scope='module')
def fp_const4():
    return _fastpitch('const4')


@pytest.fixture(scope='module')
def fp_random():
    return _fastpitch('random')


def _rel_rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / (np.sqrt(np.mean(b ** 2)) + 1e-12))


def test_hifigan_golden_batched(vocoder, golden_dir):
    g = np.load(os.path.join(golden_dir, 'hifigan_small.npz'))
    mel = torch.from_numpy(g['mel']).to(_dev())
    wav = vocoder(mel)
    assert wav.shape == (2, 1, 24 * 256) and wav.dtype == torch.float32
    w = wav.cpu().numpy()
    assert np.isfinite(w).all()
    assert _rel_rms(w, g['wav_batched']) < tol.WAV_REL_RMS
    assert np.abs(w - g['wav_batched']).max() < tol.WAV_LINF


def test_hifigan_unbatched_input_and_per_utterance_masking(vocoder, golden_dir):
    g = np.load(os.path.join(golden_dir, 'hifigan_small.npz'))
    mel = torch.from_numpy(g['mel']).to(_dev())
    # reference call shape: [80,T] -> [1, 256 T]   (hifigan/models.py:111-127)
    w1 = vocoder(mel[1, :, :17])
    assert w1.shape == (1, 17 * 256)
    assert _rel_rms(w1.cpu().numpy(), g['wav_unbatched_len17']) < tol.WAV_REL_RMS
    # the same utterance inside a padded batch with lens must equal its own unbatched result
    wb = vocoder(mel, lens=torch.tensor([24, 17]))
    assert _rel_rms(wb[1, 0, :17 * 256].cpu().numpy(), g['wav_unbatched_len17']) < tol.WAV_REL_RMS
    assert float(wb[1, 0, 17 * 256:].abs().max()) == 0.0
    assert _rel_rms(wb[0, 0].cpu().numpy(), g['wav_batched'][0, 0]) < tol.WAV_REL_RMS


def test_hifigan_vs_oracle_config2_shape(vocoder, hifigan_weights):
    """BASELINE config 2 shape class (batch 1, synthetic log-mel), shortened to 96 frames so the
    CPU oracle stays in seconds; crosses several 128-row tiles at every stage."""
    from oracle import hifigan_oracle as hgo
    gen = torch.Generator().manual_seed(0)
    mel = torch.clamp(torch.randn(1, 80, 96, generator=gen) * 2 - 5, -11.5129, 2.0)
    ref = hgo.generator_forward(hifigan_weights, synth.HIFIGAN_CONFIG, mel).numpy()
    out = vocoder(mel.to(_dev())).cpu().numpy()
    assert out.shape == ref.shape == (1, 1, 96 * 256)
    assert _rel_rms(out, ref) < tol.WAV_REL_RMS
    assert np.abs(out - ref).max() < tol.WAV_LINF


def test_fastpitch_golden_const4(fp_const4, golden_dir):
    g = np.load(os.path.join(golden_dir, 'fastpitch_const4.npz'))
    ids = torch.from_numpy(g['ids'])
    mel, dec_lens, dur, pitch, energy = fp_const4.infer(ids)
    assert dec_lens.dtype == torch.int64 and dec_lens.tolist() == g['dec_lens'].tolist()
    assert mel.shape == g['mel'].shape and pitch.shape == g['pitch_pred'].shape
    assert np.abs(dur.cpu().numpy() - g['dur_pred']).max() < tol.SCALAR_ABS
    assert np.abs(pitch.cpu().numpy() - g['pitch_pred']).max() < tol.SCALAR_ABS
    assert np.abs(energy.cpu().numpy() - g['energy_pred']).max() < tol.SCALAR_ABS
    m = mel.cpu().numpy()
    # padded frames hold proj.bias in the reference too, so the whole tensor is comparable
    assert np.abs(m - g['mel']).max() < tol.MEL_LINF


def test_fastpitch_golden_teacher_forced_and_transform(fp_random, golden_dir):
    g = np.load(os.path.join(golden_dir, 'fastpitch_random.npz'))
    ids = torch.from_numpy(g['ids'])
    dur_tgt = torch.from_numpy(g['dur_tgt'])
    mel, dec_lens, *_ = fp_random.infer(ids, dur_tgt=dur_tgt)
    assert dec_lens.tolist() == g['dec_lens_tf'].tolist()          # includes zero-length tokens
    assert np.abs(mel.cpu().numpy() - g['mel_tf']).max() < tol.MEL_LINF
    trf = lambda p, n, mean, std: 1.1 * p + 0.2                    # noqa: E731
    mel2, dec_lens2, dur2, pitch2, _ = fp_random.infer(ids, pace=0.9, pitch_transform=trf)
    assert np.abs(dur2.cpu().numpy() - g['dur_pred']).max() < tol.SCALAR_ABS
    assert np.abs(pitch2.cpu().numpy() - g['pitch_pred']).max() < tol.SCALAR_ABS
    # free-running durations are a step function of a float: frame counts may differ by a rounding
    # flip on individual tokens (SURVEY.md §7 hard part 4), totals stay within a few frames
    assert np.abs(dec_lens2.cpu().numpy() - g['dec_lens']).max() <= 3


def test_fastpitch_vs_oracle_config3_shape(fp_const4, fastpitch_weights_const4):
    """BASELINE config 3 shape class (no pads, ids < 40), B=4 x L=128 -> T=512."""
    from oracle import fastpitch_oracle as fpo
    gen = torch.Generator().manual_seed(0)
    ids = torch.randint(1, 40, (4, 128), generator=gen)
    ref_mel, ref_lens, ref_dur, ref_pitch, ref_energy = fpo.fastpitch_infer(fastpitch_weights_const4,
                                                                          synth.FASTPITCH_CONFIG, ids)
    mel, dec_lens, dur, pitch, energy = fp_const4.infer(ids)
    assert dec_lens.tolist() == ref_lens.tolist() == [512] * 4
    assert np.abs(mel.cpu().numpy() - ref_mel.numpy()).max() < tol.MEL_LINF
    assert np.abs(pitch.cpu().numpy() - ref_pitch.numpy()).max() < tol.SCALAR_ABS
    assert np.abs(energy.cpu().numpy() - ref_energy.numpy()).max() < tol.SCALAR_ABS


def test_end_to_end_golden(fp_const4, vocoder, golden_dir):
    g = np.load(os.path.join(golden_dir, 'e2e_small.npz'))
    ids = torch.from_numpy(g['ids'])
    mel, dec_lens, _, _, _, mel_cl = fp_const4.infer(ids, return_channel_last=True)
    assert dec_lens.tolist() == g['dec_lens'].tolist()
    wav = vocoder.run(mel_cl=mel_cl, lens=dec_lens)
    for b, key in enumerate(['wav0', 'wav1']):
        n = int(dec_lens[b]) * 256
        w = wav[b, :n].cpu().numpy()
        assert _rel_rms(w, g[key]) < tol.E2E_WAV_REL_RMS
        assert np.abs(w - g[key]).max() < tol.E2E_WAV_LINF
        assert float(wav[b, n:].abs().max() if n < wav.shape[1] else 0.0) == 0.0


def test_tcgen05_and_simt_paths_agree(vocoder):
    """The SIMT check kernels share packing and epilogue with the tcgen05 kernel; both must give the
    same waveform up to accumulation order."""
    from tts_arabic_pytorch_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(3)
    mel = torch.clamp(torch.randn(2, 80, 40, generator=gen) * 2 - 5, -11.5129, 2.0).to(_dev())
    lens = torch.tensor([40, 29])
    impl0 = lib.ttsb_get_conv_impl()
    try:
        _lib.check(lib.ttsb_set_conv_impl(0))
        a = vocoder(mel, lens=lens).cpu().numpy()
        _lib.check(lib.ttsb_set_conv_impl(1))
        b = vocoder(mel, lens=lens).cpu().numpy()
    finally:
        lib.ttsb_set_conv_impl(impl0)
    assert _rel_rms(a, b) < tol.WAV_REL_RMS


def test_full_size_round_trip_properties(fp_const4, vocoder):
    """BASELINE-size sanity without an oracle run: B=32 x L=128 (config 3). Properties that hold at
    any size: const-4 durations give exactly 4L frames; waveforms are finite, in [-1,1], zero
    beyond each utterance; the batch is permutation-equivariant (utterances are independent)."""
    gen = torch.Generator().manual_seed(5)
    ids = torch.randint(1, 40, (32, 128), generator=gen)
    mel, dec_lens, _, _, _, mel_cl = fp_const4.infer(ids, return_channel_last=True)
    assert dec_lens.tolist() == [512] * 32
    wav = vocoder.run(mel_cl=mel_cl, lens=dec_lens)
    assert wav.shape == (32, 512 * 256)
    assert bool(torch.isfinite(wav).all()) and float(wav.abs().max()) <= 1.0
    perm = torch.randperm(32, generator=gen)
    mel_p, _, _, _, _, mel_cl_p = fp_const4.infer(ids[perm], return_channel_last=True)
    wav_p = vocoder.run(mel_cl=mel_cl_p, lens=dec_lens)
    assert float((wav_p - wav[perm.to(wav.device)]).abs().max()) < 1e-3


def test_mixed_length_batch_config5_shape(fp_const4, vocoder, fastpitch_weights_const4, hifigan_weights):
    """BASELINE config 5 shape class on one GPU: 48 utterances of 64..256 phonemes, sorted by length and
    zero-padded like text_collate_fn (models/fastpitch/networks.py:16-35). Size-independent properties plus an
    oracle spot check of the shortest utterance:
      * const-4 durations: dec_len = 4 x (number of non-pad ids), exactly
      * every waveform is exactly zero beyond its own length (per-utterance zero padding in every layer)
      * the vocoder treats an utterance inside the padded batch like the reference's per-utterance call:
        its samples equal those of the same mel run alone
      * the whole batch equals the reference arithmetic for one utterance within the end-to-end tolerance"""
    from oracle import fastpitch_oracle as fpo
    from oracle import hifigan_oracle as hgo
    gen = torch.Generator().manual_seed(0)
    lens = sorted(torch.randint(64, 257, (48,), generator=gen).tolist(), reverse=True)
    ids = torch.zeros(48, lens[0], dtype=torch.long)
    for b, n in enumerate(lens):
        ids[b, :n] = torch.randint(1, 40, (n,), generator=gen)
    mel, dec_lens, _, _, _, mel_cl = fp_const4.infer(ids, return_channel_last=True)
    assert dec_lens.tolist() == [4 * n for n in lens]
    wav = vocoder.run(mel_cl=mel_cl, lens=dec_lens)
    assert wav.shape == (48, 4 * lens[0] * 256) and bool(torch.isfinite(wav).all())
    for b in (0, 17, 47):
        n = 4 * lens[b] * 256
        if n < wav.shape[1]:
            assert float(wav[b, n:].abs().max()) == 0.0
        alone = vocoder(mel[b, :, :4 * lens[b]])           # reference call shape: [80,T] -> [1, 256 T]
        assert float((alone[0] - wav[b, :n]).abs().max()) < 1e-3
    # oracle spot check: the shortest utterance, evaluated by the reference arithmetic inside the same padded batch
    # (FastPitch is not batch-invariant, SURVEY.md §7 hard part 3) is too slow on CPU at this size for all 48, so
    # the padded batch is reduced to the two shortest rows, which keeps the "has right padding" condition of row 47
    j = max(i for i in range(47) if lens[i] > lens[47])     # a strictly longer partner: row 47 keeps padded frames
    sub = ids[[j, 47]][:, :min(lens[j] + 1, ids.shape[1])]
    ref_mel, ref_lens, *_ = fpo.fastpitch_infer(fastpitch_weights_const4, synth.FASTPITCH_CONFIG, sub)
    t47 = int(ref_lens[1])
    assert t47 == 4 * lens[47]
    assert float((mel[47, :, :t47].cpu() - ref_mel[1, :, :t47]).abs().max()) < tol.MEL_LINF
    ref_wav = hgo.vocode_batch(hifigan_weights, synth.HIFIGAN_CONFIG, ref_mel[1:2], ref_lens[1:2])[0].numpy()
    w = wav[47, :t47 * 256].cpu().numpy()
    assert _rel_rms(w, ref_wav) < tol.E2E_WAV_REL_RMS
