"""Prints the in-kernel clock64() timeline of conv_pair_kernel CTAs (debug aid; see tlp_mark in
csrc/conv_pair.cu).   python tools/timeline_pair.py C k dil [batch]"""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

NAMES = ['x_issue', 'x_land', 'xform', 'c1_rdy', 'c1_iss', 'c2_rdy', 'c2_iss', 'acc1', 'tt_free', 'mid_done',
         'fin_wait', 'acc2', 'fin_done', 'mid_ld0', 'mid_sts', 'mid_fnc']


def main():
    import torch
    from tts_arabic_pytorch_b200 import _lib
    lib = _lib.load()
    FWD = lib.ttsb_convpair_forward_act if os.environ.get('TTSB_PROBE_ACT', '1') == '1' else lib.ttsb_convpair_forward   # timing only
    C, k, dil = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    rpf = {64: 128, 32: 256}.get(C, 64)
    T = 512 * rpf
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    w1 = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).half().float()
    w2 = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).half().float()
    b1 = torch.randn(C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    h = ctypes.c_void_p()
    _lib.check(lib.ttsb_convpair_create(C, k, dil, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), 0, ctypes.byref(h)))
    plan = (ctypes.c_int * 8)()
    _lib.check(lib.ttsb_convpair_plan(h, plan))
    x = (torch.randn(B, T, C, generator=g) * 0.5).half().to(dev)
    out = torch.empty_like(x)
    for _ in range(2):
        _lib.check(FWD(h, _lib.ptr(x), B, T, None, 0.1, _lib.ptr(out), None))
    torch.cuda.synchronize()
    tl = torch.zeros(256 * 128, dtype=torch.int64, device=dev)
    lib.ttsb_debug_set_timeline(_lib.ptr(tl))
    _lib.check(FWD(h, _lib.ptr(x), B, T, None, 0.1, _lib.ptr(out), None))
    torch.cuda.synchronize()
    lib.ttsb_debug_set_timeline(None)
    t = tl.cpu().view(256, 128)
    print('pair C=%d k=%d dil=%d B=%d plan=%s: cycles relative to CTA start' % (C, k, dil, B, list(plan)))
    print('          ' + ' '.join('%8s' % n for n in NAMES))
    for cta in [0, 1, 77, 147]:
        row = t[cta]
        if int(row[0]) == 0:
            continue
        base = int(row[0])
        print('cta %3d' % cta)
        for i in range(7):
            v = [int(row[8 + i * 16 + j]) - base if int(row[8 + i * 16 + j]) > 0 else -1 for j in range(16)]
            print('  item %d  ' % i + ' '.join('%8d' % a for a in v))
        d = [int(row[120 + j]) - base if int(row[120 + j]) > 0 else -1 for j in range(7)]
        print('  final epilogue of item 4: acc seen %d | steady chunk: start %d, rows + acc loaded %d, math %d, tile released %d, stored %d | end %d' % tuple(d))


if __name__ == '__main__':
    main()
