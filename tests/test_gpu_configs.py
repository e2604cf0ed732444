"""Parity at the BASELINE.json config sizes — round-1 VERDICT "next round" item 1:

  C1  line 0 of data/infer_text.txt: the reference's golden mel through the CUDA FastPitch.infer
  C2  Generator, mel [1,80,512] -> [1,1,131072] against the oracle at FULL size
  C3  FastPitch2Wave shape: B=32 x 128 phonemes, mel of all 32 rows and waveforms of 2 utterances against the oracle
  C4  Tacotron2MS.infer: B=8, L=64, 256 decoder steps with injected prenet masks against the oracle
  (C5 lives in tests/test_gpu_parallel.py: sharded batch == one padded batch)

plus the branches of FastPitch.infer that had no pinned fixture in round 1: encoder output tap, multi-speaker
conditioning, pitch_tgt / energy_tgt (tests/golden/fastpitch_multispk.npz, minted from the reference by
oracle/make_golden_r2.py), per-utterance speaker tensors and the device-side input validation."""
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


def _rel_rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / (np.sqrt(np.mean(b ** 2)) + 1e-12))


def _fastpitch(cfg, sd):
    from tts_arabic_pytorch_b200.models.fastpitch.fastpitch.model import FastPitch
    m = FastPitch(**cfg)
    m.load_state_dict(sd)
    return m.eval().to(_dev())


@pytest.fixture(scope='module')
def fp_const4():
    return _fastpitch(synth.FASTPITCH_CONFIG, synth.fastpitch_state_dict(1234, dur_mode='const4'))


@pytest.fixture(scope='module')
def vocoder():
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
    g = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    g.load_state_dict(synth.hifigan_state_dict(1235))
    g.eval()
    g.remove_weight_norm()
    return g.to(_dev())


def test_config1_reference_golden_through_cuda_infer(fp_const4, golden_dir):
    """The fixture holds the REAL reference's FastPitch.infer output for line 0 of data/infer_text.txt."""
    g = np.load(os.path.join(golden_dir, 'config1_line0.npz'))
    ids = torch.from_numpy(g['ids']).long()[None]
    mel, dec_lens, *_ = fp_const4.infer(ids)
    assert dec_lens.tolist() == g['dec_lens'].tolist() == [4 * ids.shape[1]]
    assert mel.shape == g['mel'].shape
    assert np.abs(mel.cpu().numpy() - g['mel']).max() < tol.MEL_LINF


def test_config2_generator_full_size_vs_oracle(vocoder, hifigan_weights):
    """BASELINE config 2 at its real size: mel [1,80,512] -> 131072 samples."""
    from oracle import hifigan_oracle as hgo
    gen = torch.Generator().manual_seed(0)
    mel = torch.clamp(torch.randn(1, 80, 512, generator=gen) * 2 - 5, -11.5129, 2.0)
    ref = hgo.generator_forward(hifigan_weights, synth.HIFIGAN_CONFIG, mel).numpy()
    out = vocoder(mel.to(_dev())).cpu().numpy()
    assert out.shape == ref.shape == (1, 1, 131072)
    assert _rel_rms(out, ref) < tol.WAV_REL_RMS
    assert np.abs(out - ref).max() < tol.WAV_LINF


def test_config3_b32_mel_and_waveforms_vs_oracle(fp_const4, vocoder, fastpitch_weights_const4, hifigan_weights):
    """BASELINE config 3 at its real size: 32 x 128 phonemes -> 512 frames each. Mel of every row and the waveforms
    of two utterances (the vocoder oracle costs ~1 s of CPU per utterance) against the reference arithmetic."""
    from oracle import fastpitch_oracle as fpo
    from oracle import hifigan_oracle as hgo
    gen = torch.Generator().manual_seed(0)
    ids = torch.randint(1, 40, (32, 128), generator=gen)
    ref_mel, ref_lens, _, ref_pitch, ref_energy = fpo.fastpitch_infer(fastpitch_weights_const4, synth.FASTPITCH_CONFIG, ids)
    mel, dec_lens, _, pitch, energy, mel_cl = fp_const4.infer(ids, return_channel_last=True)
    assert dec_lens.tolist() == ref_lens.tolist() == [512] * 32
    assert np.abs(mel.cpu().numpy() - ref_mel.numpy()).max() < tol.MEL_LINF
    assert np.abs(pitch.cpu().numpy() - ref_pitch.numpy()).max() < tol.SCALAR_ABS
    assert np.abs(energy.cpu().numpy() - ref_energy.numpy()).max() < tol.SCALAR_ABS
    wav = vocoder.run(mel_cl=mel_cl, lens=dec_lens)
    assert wav.shape == (32, 131072)
    for b in (3, 29):
        ref_wav = hgo.vocode_batch(hifigan_weights, synth.HIFIGAN_CONFIG, ref_mel[b:b + 1], ref_lens[b:b + 1])[0].numpy()
        w = wav[b].cpu().numpy()
        assert _rel_rms(w, ref_wav) < tol.E2E_WAV_REL_RMS
        assert np.abs(w - ref_wav).max() < tol.E2E_WAV_LINF


def test_config4_tacotron2_256_steps_vs_oracle():
    """BASELINE config 4 at its real size: B=8, L=64, 256 decoder steps (the synthetic gate bias keeps the stop gate
    shut), injected prenet keep-masks on both sides (the reference's prenet dropout is always on)."""
    from oracle import tacotron2_oracle as t2o
    from tts_arabic_pytorch_b200.models.tacotron2.tacotron2_ms import Tacotron2MS
    _dev()
    steps = 256
    sd = synth.tacotron2_state_dict(1236)
    m = Tacotron2MS(n_symbol=40, decoder_max_step=steps)
    m.load_state_dict(sd)
    m = m.eval().cuda()
    gen = torch.Generator().manual_seed(4)
    tokens = torch.randint(1, 40, (8, 64), generator=gen)
    lens = torch.full((8,), 64, dtype=torch.long)
    masks = torch.rand(steps, 2, 8, 256, generator=gen) > 0.5
    spk = torch.zeros(8, dtype=torch.long)
    ref_mel, ref_lens, ref_al = t2o.tacotron2_infer(sd, tokens, spk, lens, prenet_masks=masks.float() * 2.0, max_steps=steps)
    mel, mel_lens, align = m.infer(tokens, spk, lens, prenet_masks=masks)
    assert mel.shape == (8, 80, steps) and mel_lens.tolist() == ref_lens.tolist() == [steps] * 8
    assert np.abs(align.cpu().numpy() - ref_al.numpy()).max() < tol.T2_ALIGN_ABS
    assert np.abs(mel.cpu().numpy() - ref_mel.numpy()).max() < tol.T2_MEL_LINF


def test_encoder_output_tap_matches_reference_golden(fp_const4, golden_dir):
    """a4 (FFTransformer) in isolation: `enc_out` of the reference's `self.encoder(inputs)` on a padded batch."""
    g = np.load(os.path.join(golden_dir, 'fastpitch_const4.npz'))
    taps = {}
    fp_const4.infer(torch.from_numpy(g['ids']), taps=taps)
    enc = taps['enc_out'].cpu().numpy()
    assert enc.shape == g['enc_out'].shape
    # post-LayerNorm activations, O(1); padded positions are exactly zero on both sides
    assert np.abs(enc - g['enc_out']).max() < tol.MEL_LINF
    pad = g['ids'] == 0
    assert float(np.abs(enc[pad]).max()) == 0.0


@pytest.fixture(scope='module')
def fp_multispk():
    cfg = dict(synth.FASTPITCH_CONFIG, n_speakers=3, speaker_emb_weight=0.7)
    return _fastpitch(cfg, synth.fastpitch_state_dict(4321, cfg=cfg, dur_mode='const4'))


def test_multispeaker_conditioning_golden(fp_multispk, golden_dir):
    g = np.load(os.path.join(golden_dir, 'fastpitch_multispk.npz'))
    ids = torch.from_numpy(g['ids'])
    mel, dec_lens, dur, pitch, energy = fp_multispk.infer(ids, speaker=2)
    assert dec_lens.tolist() == g['dec_lens_spk2'].tolist()
    assert np.abs(mel.cpu().numpy() - g['mel_spk2']).max() < tol.MEL_LINF
    assert np.abs(dur.cpu().numpy() - g['dur_spk2']).max() < tol.SCALAR_ABS
    assert np.abs(pitch.cpu().numpy() - g['pitch_spk2']).max() < tol.SCALAR_ABS
    assert np.abs(energy.cpu().numpy() - g['energy_spk2']).max() < tol.SCALAR_ABS
    # a different speaker changes the result (the embedding is really applied)
    mel0, *_ = fp_multispk.infer(ids, speaker=0)
    assert float((mel0 - mel).abs().max()) > 10 * tol.MEL_LINF


def test_pitch_and_energy_targets_golden(fp_multispk, golden_dir):
    g = np.load(os.path.join(golden_dir, 'fastpitch_multispk.npz'))
    ids = torch.from_numpy(g['ids'])
    pitch_tgt, energy_tgt = torch.from_numpy(g['pitch_tgt']), torch.from_numpy(g['energy_tgt'])
    mel, dec_lens, _, _, energy = fp_multispk.infer(ids, speaker=1, pitch_tgt=pitch_tgt)
    assert dec_lens.tolist() == g['dec_lens_spk1_ptgt'].tolist()
    assert np.abs(mel.cpu().numpy() - g['mel_spk1_ptgt']).max() < tol.MEL_LINF
    assert np.abs(energy.cpu().numpy() - g['energy_spk1_ptgt']).max() < tol.SCALAR_ABS
    # energy_tgt: the reference's own infer() raises UnboundLocalError on this branch; the fixture composes the
    # reference modules as model.py:355-408 does (oracle/make_golden_r2.py). energy_pred is None here.
    mel2, dec_lens2, _, _, energy2 = fp_multispk.infer(ids, speaker=1, pitch_tgt=pitch_tgt, energy_tgt=energy_tgt)
    assert energy2 is None and dec_lens2.tolist() == g['dec_lens_spk1_petgt'].tolist()
    assert np.abs(mel2.cpu().numpy() - g['mel_spk1_petgt']).max() < tol.MEL_LINF


def test_per_utterance_speaker_tensor(fp_multispk, golden_dir):
    """The reference broadcasts `speaker` over the batch (model.py:358-359), so a tensor of B ids conditions each
    utterance on its own speaker; row b must equal the scalar-speaker run for that id."""
    g = np.load(os.path.join(golden_dir, 'fastpitch_multispk.npz'))
    ids = torch.from_numpy(g['ids'])
    spk = torch.tensor([2, 0, 1])
    mel, dec_lens, *_ = fp_multispk.infer(ids, speaker=spk)
    for b in range(3):
        mel_b, lens_b, *_ = fp_multispk.infer(ids, speaker=int(spk[b]))
        n = int(lens_b[b])
        assert int(dec_lens[b]) == n
        assert float((mel[b, :, :n] - mel_b[b, :, :n]).abs().max()) < 1e-3


def test_input_validation_raises_like_the_reference(fp_multispk, fp_const4):
    ids = torch.randint(1, 40, (2, 9))
    with pytest.raises(IndexError):
        bad = ids.clone()
        bad[1, 3] = synth.FASTPITCH_CONFIG['n_symbols']          # nn.Embedding raises IndexError in the reference
        fp_const4.infer(bad)
    with pytest.raises(IndexError):
        bad = ids.clone()
        bad[0, 0] = -1
        fp_const4.infer(bad)
    with pytest.raises(IndexError):
        fp_multispk.infer(ids, speaker=3)
    with pytest.raises(IndexError):
        fp_multispk.infer(ids, speaker=torch.tensor([0, 7]))
    with pytest.raises(ValueError):                               # documented contract: padding is trailing
        bad = ids.clone()
        bad[0, 4] = 0
        fp_const4.infer(bad)
    # the device stays usable after a rejected batch
    mel, dec_lens, *_ = fp_const4.infer(ids)
    assert dec_lens.tolist() == [36, 36] and bool(torch.isfinite(mel).all())


def test_dropin_install_runs_reference_style_client_code(tmp_path):
    """`inference.py`-style usage against the reference's absolute import names (inference.py:6-10, 27-39, 55-58)."""
    import subprocess
    import sys
    fp, hg, cj = synth.write_checkpoints(str(tmp_path), seed=1234)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = '''
import sys
sys.path.insert(0, %r)
import tts_arabic_pytorch_b200.dropin as dropin
dropin.install()
import text
from utils import get_basic_config
from vocoder import load_hifigan
from models.fastpitch import FastPitch2Wave
from models.fastpitch.networks import text_collate_fn
model = FastPitch2Wave(%r, vocoder_sd=%r, vocoder_config=%r, arabic_in=False).cuda()
wavs = model.tts([">als~alAmu Ealaykum", "marHabAF"], speed=1.0, denoise=0.005, batch_size=2)
assert len(wavs) == 2 and all(w.dim() == 1 and w.device.type == "cpu" for w in wavs)
print("ok", [int(w.numel()) for w in wavs])
''' % (repo, fp, hg, cj)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().startswith('ok')


def test_generator_cuda_graph_replay_matches_stream_launches(vocoder):
    """Small batches are launch-bound; a captured graph of the generator's launch train must reproduce the stream
    launches bit for bit, also on new inputs of the same shape."""
    gen = torch.Generator().manual_seed(9)
    mel = torch.clamp(torch.randn(1, 80, 64, generator=gen) * 2 - 5, -11.5129, 2.0).to(_dev())
    ref = vocoder(mel).clone()
    replay = vocoder.capture_graph(mel)
    out = replay()
    torch.cuda.synchronize()
    assert torch.equal(out.view(-1), ref.view(-1))
    mel2 = torch.clamp(torch.randn(1, 80, 64, generator=gen) * 2 - 5, -11.5129, 2.0).to(_dev())
    out2 = replay(mel2).clone()
    assert torch.equal(out2.view(-1), vocoder(mel2).view(-1))
    assert vocoder.capture_graph(mel2) is not None and len(vocoder._graphs) == 1     # same shape: same graph


def test_generator_runs_chunks_at_their_own_length_bit_exactly(fp_const4):
    """Mixed-length batches (BASELINE config 5): with the frame counts on the host the generator processes the padded
    batch in chunks that each run at their own longest utterance — same samples, no padded frames computed."""
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
    os.environ['TTSB_HIFIGAN_CHUNK_FRAMES'] = '600'          # read at handle creation: several chunks in this test
    try:
        g = Generator(AttrDict(synth.HIFIGAN_CONFIG))
        g.load_state_dict(synth.hifigan_state_dict(1235))
        g.eval()
        g.remove_weight_norm()
        g = g.to(_dev())
        gen = torch.Generator().manual_seed(3)
        lens = [200, 150, 120, 64, 10, 1]
        mel = torch.clamp(torch.randn(6, 80, 200, generator=gen) * 2 - 5, -11.5129, 2.0).to(_dev())
        lt = torch.tensor(lens)
        full = g.run(mel_f32=mel, lens=lt)
        ragged = g.run(mel_f32=mel, lens=lt, lens_host=lens)
        assert torch.equal(full, ragged)
        # channel-last input straight from FastPitch (strided chunks of the caller's [B, T, 128] tensor)
        ids = torch.zeros(5, 40, dtype=torch.long)
        for b, n in enumerate([40, 33, 20, 7, 2]):
            ids[b, :n] = torch.randint(1, 40, (n,), generator=gen)
        _, dec_lens, _, _, _, mel_cl = fp_const4.infer(ids, return_channel_last=True)
        assert dec_lens.host_list == [160, 132, 80, 28, 8]
        a = g.run(mel_cl=mel_cl, lens=dec_lens)                       # uses dec_lens.host_list
        b = g.run(mel_cl=mel_cl, lens=dec_lens.clone())               # no host copy: whole batch at T
        assert torch.equal(a, b)
    finally:
        os.environ.pop('TTSB_HIFIGAN_CHUNK_FRAMES', None)
