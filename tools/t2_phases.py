"""Per-phase cycle counts of the persistent Tacotron2 decoder (CTA 0; see t2_decoder_persistent_kernel):
python tools/t2_phases.py [batch] [steps]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    from tts_arabic_pytorch_b200 import _lib
    from tts_arabic_pytorch_b200.models.tacotron2.tacotron2_ms import Tacotron2MS
    from tts_arabic_pytorch_b200.utils import synth
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    lib = _lib.load()
    m = Tacotron2MS(n_symbol=40, decoder_max_step=steps)
    m.load_state_dict(synth.tacotron2_state_dict(1236))
    m = m.eval().cuda()
    tok = torch.randint(1, 40, (B, 64))
    import warnings
    warnings.simplefilter('ignore')
    m.infer(tok)
    tl = torch.zeros(257 * 128, dtype=torch.int64, device='cuda')
    lib.ttsb_debug_set_timeline(_lib.ptr(tl))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.infer(tok)
    e1.record()
    torch.cuda.synchronize()
    lib.ttsb_debug_set_timeline(None)
    v = tl[256 * 128:256 * 128 + 8].tolist()
    names = ['A attention LSTM', 'barrier', 'B attention', 'barrier', 'C decoder LSTM', 'barrier', 'D proj+prenet', 'barrier']
    tot = sum(v)
    print('persistent decoder, B=%d, %d steps: %.1f us per step (whole infer: %.2f ms)' % (B, steps, e0.elapsed_time(e1) * 1e3 / steps, e0.elapsed_time(e1)))
    for n, c in zip(names, v):
        print('  %-18s %8.0f cycles per step  (%4.1f %%)' % (n, c / steps, 100.0 * c / max(tot, 1)))


if __name__ == '__main__':
    main()
