"""Per-layer micro-benchmark of the conv primitive at vocoder/FastPitch shapes (B utterances of
T=512 frames): CUDA-event time per launch, algorithmic TFLOP/s and activation GB/s.
  python tools/bench_conv.py [--batch 4] [--iters 20] [--only substr] [--json out.json]
"""
import argparse
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

# name, kind, cin, cout, k, dil, stride, rows_per_frame, residual
LAYERS = [
    ('pre_80x512_k7', 0, 80, 512, 7, 1, 1, 1, False),
    ('ups0_512x256_s8', 1, 512, 256, 16, 1, 8, 1, False),
    ('s0_256_k3_d1', 0, 256, 256, 3, 1, 1, 8, True),
    ('s0_256_k7_d3', 0, 256, 256, 7, 3, 1, 8, True),
    ('s0_256_k11_d5', 0, 256, 256, 11, 5, 1, 8, True),
    ('ups1_256x128_s8', 1, 256, 128, 16, 1, 8, 8, False),
    ('s1_128_k3_d1', 0, 128, 128, 3, 1, 1, 64, True),
    ('s1_128_k3_nores', 0, 128, 128, 3, 1, 1, 64, False),
    ('s1_128_k7_nores', 0, 128, 128, 7, 3, 1, 64, False),
    ('s1_128_k7_d3', 0, 128, 128, 7, 3, 1, 64, True),
    ('s1_128_k11_d5', 0, 128, 128, 11, 5, 1, 64, True),
    ('ups2_128x64_s2', 1, 128, 64, 4, 1, 2, 64, False),
    ('s2_64_k3_d1', 0, 64, 64, 3, 1, 1, 128, True),
    ('s2_64_k11_d5', 0, 64, 64, 11, 5, 1, 128, True),
    ('ups3_64x32_s2', 1, 64, 32, 4, 1, 2, 128, False),
    ('s3_32_k3_d1', 0, 32, 32, 3, 1, 1, 256, True),
    ('s3_32_k11_d5', 0, 32, 32, 11, 5, 1, 256, True),
    ('ff1_384x1536_k3', 0, 384, 1536, 3, 1, 1, 1, False),
    ('ff2_1536x384_k3', 0, 1536, 384, 3, 1, 1, 1, True),
    ('qkv_384x192', 0, 384, 192, 1, 1, 1, 1, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--frames', type=int, default=512)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--only', default='')
    ap.add_argument('--json', default='')
    a = ap.parse_args()
    import torch
    from tts_arabic_pytorch_b200 import _lib
    lib = _lib.load()
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    out_rows = []
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for name, kind, cin, cout, k, dil, stride, rpf, use_res in LAYERS:
        if a.only and a.only not in name:
            continue
        B, T = a.batch, a.frames * rpf
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
        times = []
        for _ in range(a.iters):
            flush.zero_()                       # evict activations from L2 between launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.ttsb_conv1d_forward(h, _lib.ptr(x), B, T, _lib.ptr(res), 0.1, None, _lib.ptr(out), None))
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
        times.sort()
        us = times[len(times) // 2]
        taps = k if kind == 0 else 2
        flops = 2.0 * B * T * n_out * taps * cin
        act_bytes = B * T * (cpad + n_out * (2 if use_res else 1)) * 2
        row = {'layer': name, 'rows': B * T, 'us': us, 'tflops': flops / us / 1e6, 'act_GBs': act_bytes / us / 1e3}
        out_rows.append(row)
        print('%-20s rows=%8d  %9.1f us  %7.1f TFLOP/s  %7.1f GB/s(act)' % (name, B * T, us, row['tflops'], row['act_GBs']))
        lib.ttsb_conv1d_destroy(h)
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(out_rows, f, indent=1)


if __name__ == '__main__':
    main()
