"""Prints the in-kernel clock64() timeline of conv_tc_kernel CTAs for one microbench layer
(debug aid; see tl_mark in csrc/conv_tc.cu).  python tools/timeline.py s1_128_k11_d5 [batch]"""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tools'))
from bench_conv import LAYERS  # noqa: E402


def main():
    import torch
    from tts_arabic_pytorch_b200 import _lib
    lib = _lib.load()
    want = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    for name, kind, cin, cout, k, dil, stride, rpf, use_res in LAYERS:
        if name != want:
            continue
        T = 512 * rpf
        wshape = (cout, cin, k) if kind == 0 else (cin, cout, k)
        w = torch.randn(wshape, generator=g) * 0.05
        bias = torch.randn(cout, generator=g) * 0.1
        h = ctypes.c_void_p()
        _lib.check(lib.ttsb_conv1d_create(kind, cin, cout, k, dil, stride, ctypes.c_void_p(w.data_ptr()),
                                          ctypes.c_void_p(bias.data_ptr()), 0, ctypes.byref(h)))
        cpad = lib.ttsb_conv1d_cin_pad(h)
        n_out = cout * (stride if kind == 1 else 1)
        x = (torch.randn(B, T, cpad, generator=g) * 0.5).half().to(dev)
        res = (torch.randn(B, T, n_out, generator=g) * 0.5).half().to(dev) if use_res else None
        out = torch.empty(B, T, n_out, dtype=torch.float16, device=dev)
        for _ in range(3):
            _lib.check(lib.ttsb_conv1d_forward(h, _lib.ptr(x), B, T, _lib.ptr(res), 0.1, None, _lib.ptr(out), None))
        torch.cuda.synchronize()
        tl = torch.zeros(256 * 128, dtype=torch.int64, device=dev)
        lib.ttsb_debug_set_timeline(_lib.ptr(tl))
        _lib.check(lib.ttsb_conv1d_forward(h, _lib.ptr(x), B, T, _lib.ptr(res), 0.1, None, _lib.ptr(out), None))
        torch.cuda.synchronize()
        lib.ttsb_debug_set_timeline(None)
        t = tl.cpu().view(256, 128)
        t0 = int(t[:, 0][t[:, 0] > 0].min())
        print('layer %s (conv_tc2): cycles relative to each CTA start; per tile: mma[gotTMEM gotA issued] '
              'epi[wait seen drained done] panel_issued' % name)
        for cta in [0, 1, 147, 148, 200, 255]:
            row = t[cta]
            if int(row[0]) == 0:
                continue
            base = int(row[0])
            print('cta %3d start@%d setup=%d' % (cta, base - t0, int(row[1]) - base))
            for i in range(7):
                v = [int(row[8 + i * 8 + k]) - base if int(row[8 + i * 8 + k]) > 0 else -1 for k in range(8)]
                print('    tile %d  mma %6d %6d %6d | epi %6d %6d %6d %6d | panel %6d' % (i, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]))
                d = [int(row[64 + i * 8 + k]) - base if int(row[64 + i * 8 + k]) > 0 else -1 for k in range(7)]
                w = int(row[64 + i * 8 + 7])
                print('            epi detail (lean: seen | chunk 1: start, rows + acc loaded, math, tile released, stored | end): %s | issuer waited %d cycles in %d of the weight stages'
                      % (' '.join('%d' % a for a in d), w // 1000, w % 1000))


if __name__ == '__main__':
    main()
