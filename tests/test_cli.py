"""inference.py-compatible command line (reference inference.py:21-107): wav writer round trip and argument handling on
the CPU; the end-to-end run (list file -> wavs + index.html) on the GPU."""
import os

import pytest
import torch

from tts_arabic_pytorch_b200 import inference
from tts_arabic_pytorch_b200.utils import synth


def test_wav_writer_round_trip_and_scipy_reads_it(tmp_path):
    w = torch.randn(3001).clamp(-1, 1)
    p = str(tmp_path / 'a.wav')
    inference.write_wav_f32(p, w)
    r, rate = inference.read_wav_f32(p)
    assert rate == 22050 and torch.equal(r, w)
    import scipy.io.wavfile as wf
    sr, d = wf.read(p)
    assert sr == 22050 and d.dtype.name == 'float32' and float(abs(d - w.numpy()).max()) == 0.0


def test_cpu_flag_is_rejected_loudly(tmp_path):
    with pytest.raises(RuntimeError, match='CUDA device only'):
        inference.main(['--cpu', '--list', str(tmp_path / 'none.txt')])


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device')
    fp, hg, cj = synth.write_checkpoints(str(tmp_path), seed=1234)
    lst = tmp_path / 'list.txt'
    lines = [">als~alAmu Ealaykum yA Sadiyqiy", "marHabAF bikum", ">aHrazat muntaxabAtu lbarAziyli fawzan"]
    lst.write_text('\n'.join(lines) + '\n', encoding='utf-8')
    out = tmp_path / 'res'
    n = inference.main(['--list', str(lst), '--checkpoint', fp, '--vocoder_sd', hg, '--vocoder_config', cj,
                        '--out_dir', str(out), '--batch_size', '2', '--denoise', '0.005', '--speed', '1.0'])
    assert n == 3
    for i in range(3):
        w, rate = inference.read_wav_f32(str(out / 'wavs' / ('static%d.wav' % i)))
        assert rate == 22050 and w.numel() > 0 and w.numel() % 256 == 0 and bool(torch.isfinite(w).all())
    html = (out / 'index.html').read_text(encoding='utf-8')
    assert html.count('<audio') == 3 and 'wavs/static2.wav' in html
