"""Runs the HiFi-GAN generator alone on a synthetic padded mel batch (for ncu launch lists / captures and quick
CUDA-event timing).   python tools/run_vocoder.py [--batch 64] [--frames 512] [--reps 3]
One pass = 1 pack_mel launch + the generator's launches (62 with conv_pair on the C <= 64 stages), so
`ncu -s $((2*63)) -c 63` captures exactly the third pass."""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--frames', type=int, default=512)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--profile-last', action='store_true', help='cudaProfilerStart/Stop around the last pass (ncu --profile-from-start off)')
    a = ap.parse_args()
    import torch
    from tts_arabic_pytorch_b200 import _lib
    from tts_arabic_pytorch_b200.utils import synth
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
    dev = torch.device('cuda:0')
    voc = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    voc.load_state_dict(synth.hifigan_state_dict(1235))
    voc.remove_weight_norm()
    voc = voc.eval().to(dev)
    g = torch.Generator().manual_seed(0)
    mel = torch.clamp(torch.randn(a.batch, 80, a.frames, generator=g) * 2 - 5, -11.5129, 2.0).to(dev)
    lib = _lib.load()
    times = []
    for i in range(a.reps):
        n0 = lib.ttsb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if a.profile_last and i == a.reps - 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        e0.record()
        wav = voc(mel)
        e1.record()
        torch.cuda.synchronize()
        if a.profile_last and i == a.reps - 1:
            torch.cuda.profiler.stop()
        times.append(e0.elapsed_time(e1))
        launches = lib.ttsb_launch_count() - n0
    fl = 614.1e6 * a.batch * a.frames
    best = min(times)
    print('vocoder B=%d T=%d: %d launches per pass, best %.2f ms (%.1f TFLOP/s algorithmic), all: %s' % (
        a.batch, a.frames, launches, best, fl / best / 1e9, ' '.join('%.2f' % t for t in times)))
    assert bool(torch.isfinite(wav).all())


if __name__ == '__main__':
    main()
