"""GPU probe: every conv-site shape of both models through ttsb_conv1d_*, for each kernel variant
(SIMT check kernel, tcgen05 kernel with A-descriptor modes 0..3), against torch fp32 on the same
fp16-rounded operands. Each variant runs in its own subprocess with a timeout so a trap or a hang
in one variant cannot take the others (or the box) down.

  python tools/probe_conv.py            -> gpurun_out/probe_conv.json + summary on stdout
  python tools/probe_conv.py --child simt|tc0|tc1|tc2|tc3
"""
import ctypes
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

# (name, kind, cin, cout, k, dil, stride)
CASES = [
    ('lin384x192', 0, 384, 192, 1, 1, 1),
    ('lin64x384', 0, 64, 384, 1, 1, 1),
    ('ff1_384x1536_k3', 0, 384, 1536, 3, 1, 1),
    ('ff2_1536x384_k3', 0, 1536, 384, 3, 1, 1),
    ('pred_384x256_k3', 0, 384, 256, 3, 1, 1),
    ('convpre_80x512_k7', 0, 80, 512, 7, 1, 1),
    ('rb256_k3_d1', 0, 256, 256, 3, 1, 1),
    ('rb256_k11_d5', 0, 256, 256, 11, 5, 1),
    ('rb128_k7_d3', 0, 128, 128, 7, 3, 1),
    ('rb64_k11_d5', 0, 64, 64, 11, 5, 1),
    ('rb32_k3_d1', 0, 32, 32, 3, 1, 1),
    ('rb32_k11_d5', 0, 32, 32, 11, 5, 1),
    ('ups512x256_k16_s8', 1, 512, 256, 16, 1, 8),
    ('ups256x128_k16_s8', 1, 256, 128, 16, 1, 8),
    ('ups128x64_k4_s2', 1, 128, 64, 4, 1, 2),
    ('ups64x32_k4_s2', 1, 64, 32, 4, 1, 2),
]


def child(variant):
    import torch
    import torch.nn.functional as F
    from tts_arabic_pytorch_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = _lib.load()
    if variant == 'simt':
        _lib.check(lib.ttsb_set_conv_impl(1))
    elif variant in ('v1', 'v2'):
        _lib.check(lib.ttsb_set_conv_impl(0))
        _lib.check(lib.ttsb_set_desc_mode(0))
        _lib.check(lib.ttsb_set_tc_version(int(variant[1])))
    else:
        _lib.check(lib.ttsb_set_conv_impl(0))
        _lib.check(lib.ttsb_set_desc_mode(int(variant[2:])))
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    results = {}
    B, T = 3, 700      # several row tiles per utterance and a ragged tail; persistent CTAs wrap around
    for name, kind, cin, cout, k, dil, stride in CASES:
        wshape = (cout, cin, k) if kind == 0 else (cin, cout, k)
        w = (torch.randn(wshape, generator=g) / (cin * (k if kind == 0 else k / stride)) ** 0.5).half().float()
        bias = torch.randn(cout, generator=g) * 0.1
        h = ctypes.c_void_p()
        _lib.check(lib.ttsb_conv1d_create(kind, cin, cout, k, dil, stride, ctypes.c_void_p(w.data_ptr()),
                                          ctypes.c_void_p(bias.data_ptr()), 0, ctypes.byref(h)))
        cpad = lib.ttsb_conv1d_cin_pad(h)
        x = torch.zeros(B, T, cpad, dtype=torch.float16)
        x[:, :, :cin] = torch.randn(B, T, cin, generator=g).half()
        lens = torch.tensor([T, T - 37, 129], dtype=torch.int32)
        x[1, T - 37:] = 0
        x[2, 129:] = 0
        n_out = cout * (stride if kind == 1 else 1)
        res = torch.randn(B, T, n_out, generator=g).half()
        xd, rd, ld = x.to(dev), res.to(dev), lens.to(dev)
        out = torch.full((B, T, n_out), float('nan'), dtype=torch.float16, device=dev)
        # reference on the same fp16-rounded operands, fp32 math
        xr = xd[:, :, :cin].float().transpose(1, 2)
        wd, bd = w.to(dev), bias.to(dev)
        if kind == 0:
            ref = F.conv1d(xr, wd, bd, padding=(k - 1) // 2 * dil, dilation=dil).transpose(1, 2)
        else:
            ref = F.conv_transpose1d(xr, wd, bd, stride=stride, padding=stride // 2).transpose(1, 2)
            ref = ref.reshape(B, T, n_out)      # [B, T*s, cout] == [B, T, s*cout]
        ref = ref + rd.float()
        ref = torch.where(ref > 0, ref, ref * 0.1)
        mask = (torch.arange(T, device=dev)[None, :] < ld[:, None])[:, :, None]
        ref = ref * mask
        t0 = time.time()
        st = lib.ttsb_conv1d_forward(h, _lib.ptr(xd), B, T, _lib.ptr(rd), 0.1, _lib.ptr(ld), _lib.ptr(out), None)
        err = None
        if st != 0:
            err = lib.ttsb_last_error().decode()
        try:
            torch.cuda.synchronize()
        except Exception as e:   # trap / illegal instruction: context is gone, stop this child
            results[name] = {'ok': False, 'error': 'sync failed: %s' % e}
            break
        if err:
            results[name] = {'ok': False, 'error': err}
            continue
        o = out.float()
        diff = (o - ref).abs()
        nan = int(torch.isnan(o).sum())
        scale = float(ref.abs().max())
        maxerr = float(torch.nan_to_num(diff, nan=1e9).max())
        results[name] = {'ok': bool(nan == 0 and maxerr <= 4e-3 * max(scale, 1.0)), 'max_err': maxerr, 'ref_max': scale,
                         'nan': nan, 'ms': (time.time() - t0) * 1e3}
        lib.ttsb_conv1d_destroy(h)
    flag = ctypes.c_int(0)
    try:
        lib.ttsb_device_error_flag(ctypes.byref(flag))
    except Exception:
        pass
    print('RESULT ' + json.dumps({'variant': variant, 'device_flag': flag.value, 'cases': results}))


def main():
    out_dir = os.path.join(REPO, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    summary = {}
    variants = sys.argv[1:] if len(sys.argv) > 1 else ['simt', 'v2', 'v1']
    for variant in variants:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', variant], capture_output=True,
                               text=True, timeout=240)
            line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
            if line:
                summary[variant] = json.loads(line[-1][7:])
            else:
                summary[variant] = {'crashed': True, 'rc': r.returncode, 'stderr': r.stderr[-1500:], 'stdout': r.stdout[-500:]}
        except subprocess.TimeoutExpired:
            summary[variant] = {'timeout': True}
    with open(os.path.join(out_dir, 'probe_conv.json'), 'w') as f:
        json.dump(summary, f, indent=1)
    for v, s in summary.items():
        if 'cases' not in s:
            print(v, 'FAILED TO RUN', json.dumps(s)[:600])
            continue
        ok = [n for n, c in s['cases'].items() if c.get('ok')]
        bad = {n: (c.get('max_err'), c.get('error')) for n, c in s['cases'].items() if not c.get('ok')}
        print('%s: %d/%d ok, device_flag=%s, bad=%s' % (v, len(ok), len(CASES), s.get('device_flag'), bad))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--child':
        child(sys.argv[2])
    else:
        main()
