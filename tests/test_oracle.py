"""Pins the CPU oracle (oracle/*.py) against fixtures minted from the real reference code
(oracle/make_golden.py). fp32 on both sides, so the tolerance only has to absorb summation-order
differences (gather vs one-hot matmul, bmm vs flat ops): 2e-4 absolute on O(1)..O(5) values."""
import os

import numpy as np
import torch

from oracle import fastpitch_oracle as fpo
from oracle import hifigan_oracle as hgo
from tts_arabic_pytorch_b200.utils import synth

TOL = 2e-4


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_hifigan_oracle_matches_reference(golden_dir, hifigan_weights):
    g = _load(golden_dir, 'hifigan_small.npz')
    mel = torch.from_numpy(g['mel'])
    wav = hgo.generator_forward(hifigan_weights, synth.HIFIGAN_CONFIG, mel)
    assert wav.shape == (2, 1, 24 * 256)
    assert np.abs(wav.numpy() - g['wav_batched']).max() < TOL
    # unbatched [80,T] input returns [1, 256T] (hifigan/models.py:111-127 with a 2-D tensor)
    wav_u = hgo.generator_forward(hifigan_weights, synth.HIFIGAN_CONFIG, mel[1, :, :17])
    assert wav_u.shape == (1, 17 * 256)
    assert np.abs(wav_u.numpy() - g['wav_unbatched_len17']).max() < TOL


def test_weight_norm_fold_is_identity_on_direction_times_gain():
    sd = synth.hifigan_state_dict(7)
    folded = synth.fold_weight_norm(sd)
    v = sd['ups.0.parametrizations.weight.original1']
    # the synthetic g equals ||v|| so folding must return v itself
    assert torch.allclose(folded['ups.0.weight'], v, atol=1e-6)
    # old-style keys fold the same way
    old = {'c.weight_g': sd['ups.0.parametrizations.weight.original0'] * 2, 'c.weight_v': v, 'c.bias': sd['ups.0.bias']}
    assert torch.allclose(synth.fold_weight_norm(old)['c.weight'], 2 * v, atol=1e-5)


def test_fastpitch_oracle_const4(golden_dir, fastpitch_weights_const4):
    g = _load(golden_dir, 'fastpitch_const4.npz')
    ids = torch.from_numpy(g['ids'])
    taps = {}
    mel, dec_lens, dur, pitch, energy = fpo.fastpitch_infer(fastpitch_weights_const4, synth.FASTPITCH_CONFIG, ids, taps=taps)
    assert dec_lens.tolist() == g['dec_lens'].tolist() == [80, 52, 28]
    assert np.abs(taps['enc_out'].numpy() - g['enc_out']).max() < TOL
    assert np.abs(dur.numpy() - g['dur_pred']).max() < TOL
    assert np.abs(pitch.numpy() - g['pitch_pred']).max() < TOL
    assert np.abs(energy.numpy() - g['energy_pred']).max() < TOL
    assert mel.shape == g['mel'].shape
    assert np.abs(mel.numpy() - g['mel']).max() < 5 * TOL


def test_fastpitch_oracle_random_durations_pace_and_transform(golden_dir, fastpitch_weights_random):
    g = _load(golden_dir, 'fastpitch_random.npz')
    ids = torch.from_numpy(g['ids'])
    trf = lambda p, n, mean, std: 1.1 * p + 0.2   # noqa: E731
    mel, dec_lens, dur, pitch, energy = fpo.fastpitch_infer(fastpitch_weights_random, synth.FASTPITCH_CONFIG, ids,
                                                            pace=0.9, pitch_transform=trf)
    assert dec_lens.tolist() == g['dec_lens'].tolist()
    assert np.abs(dur.numpy() - g['dur_pred']).max() < TOL
    assert np.abs(pitch.numpy() - g['pitch_pred']).max() < TOL
    assert np.abs(mel.numpy() - g['mel']).max() < 5 * TOL
    # teacher-forced durations (model.py:401-403), including zero-length tokens
    mel_t, dec_lens_t, *_ = fpo.fastpitch_infer(fastpitch_weights_random, synth.FASTPITCH_CONFIG, ids,
                                                dur_tgt=torch.from_numpy(g['dur_tgt']))
    assert dec_lens_t.tolist() == g['dec_lens_tf'].tolist()
    assert np.abs(mel_t.numpy() - g['mel_tf']).max() < 5 * TOL


def test_config1_line0_plumbing(golden_dir, fastpitch_weights_const4):
    """BASELINE config 1: one utterance of data/infer_text.txt, FastPitch only, CPU."""
    g = _load(golden_dir, 'config1_line0.npz')
    ids = torch.from_numpy(g['ids'])[None]
    mel, dec_lens, *_ = fpo.fastpitch_infer(fastpitch_weights_const4, synth.FASTPITCH_CONFIG, ids)
    assert dec_lens.tolist() == g['dec_lens'].tolist() == [4 * ids.shape[1]]
    assert np.abs(mel.numpy() - g['mel']).max() < 5 * TOL


def test_end_to_end_oracle(golden_dir, fastpitch_weights_const4, hifigan_weights):
    g = _load(golden_dir, 'e2e_small.npz')
    ids = torch.from_numpy(g['ids'])
    mel, dec_lens, *_ = fpo.fastpitch_infer(fastpitch_weights_const4, synth.FASTPITCH_CONFIG, ids)
    wavs = hgo.vocode_batch(hifigan_weights, synth.HIFIGAN_CONFIG, mel, dec_lens)
    for b, key in enumerate(['wav0', 'wav1']):
        assert wavs[b].shape == g[key].shape == (int(dec_lens[b]) * 256,)
        assert np.abs(wavs[b].numpy() - g[key]).max() < 5 * TOL


def test_fastpitch_is_not_batch_invariant(fastpitch_weights_const4):
    """SURVEY.md §7 hard part 3: an utterance inside a padded batch differs from the same utterance
    alone (conv-FF leaks the first padded position), and the difference depends only on whether
    there is at least one padded position. The CUDA path must reproduce the padded-batch numbers."""
    torch.manual_seed(1)
    ids = torch.randint(1, 40, (2, 16))
    ids[1, 9:] = 0
    cfg = synth.FASTPITCH_CONFIG
    t_b, t_s, t_w = {}, {}, {}
    fpo.fastpitch_infer(fastpitch_weights_const4, cfg, ids, taps=t_b)
    fpo.fastpitch_infer(fastpitch_weights_const4, cfg, ids[1:, :9], taps=t_s)
    wide = torch.zeros(2, 40, dtype=torch.long)
    wide[:, :16] = ids
    fpo.fastpitch_infer(fastpitch_weights_const4, cfg, wide, taps=t_w)
    a, s, ww = t_b['enc_out'][1, :9], t_s['enc_out'][0], t_w['enc_out'][1, :9]
    assert (a - s).abs().max() > 1e-2          # batch-variant
    assert (a - ww).abs().max() < 1e-4         # but independent of HOW MUCH padding


def test_tacotron2_oracle_matches_reference_with_injected_masks(golden_dir):
    """BASELINE config 4 path (Tacotron2MS.infer + torchaudio decoder), prenet dropout masks injected on
    both sides, gate held shut so all utterances run the full 24 steps."""
    from oracle import tacotron2_oracle as t2o
    g = _load(golden_dir, 'tacotron2_small.npz')
    sd = synth.tacotron2_state_dict(1236)
    masks = torch.from_numpy(g['masks']).float() * 2.0
    mel, mel_lens, align = t2o.tacotron2_infer(sd, torch.from_numpy(g['tokens']), torch.from_numpy(g['speaker_ids']),
                                               torch.from_numpy(g['lengths']), prenet_masks=masks, max_steps=24)
    assert mel_lens.tolist() == g['mel_lengths'].tolist() == [24, 24, 24]
    assert mel.shape == g['mel'].shape == (3, 80, 24)
    assert np.abs(mel.numpy() - g['mel']).max() < TOL
    assert np.abs(align.numpy() - g['alignments']).max() < TOL
    # attention never looks at padded tokens
    assert float(align[1, :, 10:].abs().max()) == 0.0 and float(align[2, :, 5:].abs().max()) == 0.0


def test_denoiser_oracle_matches_reference_fixture(golden_dir, hifigan_weights):
    """oracle/denoiser_oracle.py against outputs of the real reference Denoiser (torchaudio Spectrogram /
    InverseSpectrogram, vocoder/hifigan/denoiser.py:29-72) minted by oracle/make_golden_denoiser.py."""
    import numpy as np
    import torch
    from oracle import denoiser_oracle as dno
    from oracle import hifigan_oracle as hgo
    from tts_arabic_pytorch_b200.utils import synth
    g = np.load(os.path.join(golden_dir, 'denoiser_small.npz'))
    # bias spectrum: zero-mel response of the (oracle) generator, first STFT frame (denoiser.py:51-64)
    zero_audio = hgo.generator_forward(hifigan_weights, synth.HIFIGAN_CONFIG, torch.zeros(1, 80, 88))
    bias = dno.bias_spectrum(zero_audio)
    assert bias.shape == (1, 513, 1)
    assert np.abs(bias.numpy() - g['bias_spec']).max() < 2e-4 * max(1.0, float(np.abs(g['bias_spec']).max()))
    ref_bias = torch.from_numpy(g['bias_spec'])
    for name in ('0', '1'):
        audio = torch.from_numpy(g['audio' + name])
        for s in (0.005, 0.1):
            out = dno.denoise(audio, ref_bias, s).numpy()
            ref = g['out%s_s%g' % (name, s)]
            assert out.shape == ref.shape
            assert np.abs(out - ref).max() < 1e-5


def test_vendored_reference_modules_agree_with_the_oracle_port():
    """oracle/_ref (oracle/make_ref.py: byte-for-byte copies of the reference's FastPitch / HiFi-GAN modules, git-ignored)
    is what bench.py times as the reference arm; where it is present it must be the unmodified files and agree with the
    restatement the parity tests use."""
    import hashlib
    import json
    import os

    import pytest
    import torch
    from oracle import fastpitch_oracle as fpo
    from oracle import hifigan_oracle as hgo
    from oracle import ref_runner
    from tts_arabic_pytorch_b200.utils import synth

    if not ref_runner.available():
        pytest.skip('oracle/_ref not built (python oracle/make_ref.py needs /root/reference)')
    root = os.path.join(os.path.dirname(os.path.abspath(ref_runner.__file__)), '_ref')
    manifest = json.load(open(os.path.join(root, 'MANIFEST.json')))
    for rel, meta in manifest.items():
        with open(os.path.join(root, rel), 'rb') as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == meta['sha256'], rel
        src = os.path.join('/root/reference', meta['source'])
        if os.path.exists(src):
            with open(src, 'rb') as fh:
                assert hashlib.sha256(fh.read()).hexdigest() == meta['sha256'], 'differs from the reference: ' + rel
    fsd, gsd = synth.fastpitch_state_dict(1234), synth.hifigan_state_dict(1235)
    fp, voc = ref_runner.build_models(fsd, synth.FASTPITCH_CONFIG, gsd, synth.HIFIGAN_CONFIG)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(1, 40, (2, 14), generator=g)
    ids[1, 9:] = 0
    n, wavs = ref_runner.step(fp, voc, ids)
    mel, lens, *_ = fpo.fastpitch_infer(fsd, synth.FASTPITCH_CONFIG, ids)
    port = hgo.vocode_batch(synth.fold_weight_norm(gsd), synth.HIFIGAN_CONFIG, mel, lens)
    assert n == sum(int(w.numel()) for w in port)
    for a, b in zip(wavs, port):
        assert float((a.flatten() - b.flatten()).abs().max()) < 2e-4
