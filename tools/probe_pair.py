"""GPU probe + micro-benchmark of the fused ResBlock1 step (csrc/conv_pair.cu) through ttsb_convpair_*:
every (C, k, dilation) of the generator's C <= 64 stages against torch fp32 on the same fp16-rounded
operands (intermediate rounded to fp16 like the kernel does), then CUDA-event timing at stage size.
Runs in a child process with a timeout so a trap or hang cannot take the box down.

  python tools/probe_pair.py [--bench] [--batch 16]     -> gpurun_out/probe_pair.json + summary
"""
import ctypes
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

# (C, rows per mel frame)
STAGES = [(64, 128), (32, 256)]
KS = [3, 7, 11]
DILS = [1, 3, 5]


def reference(x, lens, w1, b1, w2, b2, k, dil, slope):
    """x: [B,T,C] fp16 on device (rows >= lens are zero). fp32 math, fp16 rounding where the kernel rounds."""
    import torch
    import torch.nn.functional as F
    B, T, C = x.shape
    xr = x.float().transpose(1, 2)
    mask = (torch.arange(T, device=x.device)[None, :] < lens[:, None])[:, None, :]
    lx = torch.where(xr > 0, xr, xr * slope).half().float()
    t = F.conv1d(lx, w1, b1, padding=(k - 1) // 2 * dil, dilation=dil)
    t = torch.where(t > 0, t, t * slope) * mask
    t = t.half().float()
    y = F.conv1d(t, w2, b2, padding=(k - 1) // 2) + xr
    y = y * mask
    return y.transpose(1, 2)


def child(bench, batch):
    import torch
    from tts_arabic_pytorch_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = _lib.load()
    FWD = lib.ttsb_convpair_forward_act if os.environ.get('TTSB_PROBE_ACT', '1') == '1' else lib.ttsb_convpair_forward   # timing only
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    results = {}
    for C, rpf in STAGES:
        for k in KS:
            for dil in DILS:
                name = 'c%d_k%d_d%d' % (C, k, dil)
                w1 = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).half().float()
                w2 = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).half().float()
                b1 = torch.randn(C, generator=g) * 0.1
                b2 = torch.randn(C, generator=g) * 0.1
                h = ctypes.c_void_p()
                _lib.check(lib.ttsb_convpair_create(C, k, dil, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), 0,
                                                    ctypes.byref(h)))
                plan = (ctypes.c_int * 8)()
                _lib.check(lib.ttsb_convpair_plan(h, plan))
                plan = list(plan)
                if not plan[0]:
                    results[name] = {'ok': True, 'fused': False, 'plan': plan}
                    lib.ttsb_convpair_destroy(h)
                    continue
                # ragged batch: several tiles per utterance, a tail tile, a short utterance, persistent CTAs wrap
                B, T = 3, 20011
                lens = torch.tensor([T, T - 371, 129], dtype=torch.int32)
                x = torch.randn(B, T, C, generator=g).half()
                x[1, T - 371:] = 0
                x[2, 129:] = 0
                xd, ld = x.to(dev), lens.to(dev)
                out = torch.full((B, T, C), float('nan'), dtype=torch.float16, device=dev)
                st = lib.ttsb_convpair_forward(h, _lib.ptr(xd), B, T, _lib.ptr(ld), 0.1, _lib.ptr(out), None)
                if st != 0:
                    results[name] = {'ok': False, 'error': lib.ttsb_last_error().decode(), 'plan': plan}
                    lib.ttsb_convpair_destroy(h)
                    continue
                try:
                    torch.cuda.synchronize()
                except Exception as e:
                    results[name] = {'ok': False, 'error': 'sync failed: %s' % e, 'plan': plan}
                    break
                ref = reference(xd, ld, w1.to(dev), b1.to(dev), w2.to(dev), b2.to(dev), k, dil, 0.1)
                o = out.float()
                nan = int(torch.isnan(o).sum())
                diff = torch.nan_to_num((o - ref).abs(), nan=1e9)
                maxerr = float(diff.max())
                scale = float(ref.abs().max())
                r = {'ok': bool(nan == 0 and maxerr <= 4e-3 * max(scale, 1.0)), 'fused': True, 'max_err': maxerr,
                     'ref_max': scale, 'nan': nan, 'plan': plan}
                if not r['ok']:
                    bad = (diff > 4e-3 * max(scale, 1.0)).nonzero()
                    r['n_bad'] = int(bad.shape[0])
                    r['first_bad'] = bad[:6].tolist()
                    r['bad_rows_mod'] = sorted(set(int(v) % plan[1] for v in bad[:2000, 1].tolist()))[:40]
                if bench and r['ok']:
                    Bb, Tb = batch, 512 * rpf
                    xb = (torch.randn(Bb, Tb, C, generator=g) * 0.5).half().to(dev)
                    ob = torch.empty_like(xb)
                    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
                    for _ in range(2):
                        _lib.check(FWD(h, _lib.ptr(xb), Bb, Tb, None, 0.1, _lib.ptr(ob), None))
                    torch.cuda.synchronize()
                    times = []
                    for _ in range(7):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        _lib.check(FWD(h, _lib.ptr(xb), Bb, Tb, None, 0.1, _lib.ptr(ob), None))
                        e1.record()
                        torch.cuda.synchronize()
                        times.append(e0.elapsed_time(e1) * 1e3)
                    times.sort()
                    us = times[len(times) // 2]
                    r['us'] = us
                    r['rows'] = Bb * Tb
                    r['tflops'] = 2 * 2.0 * Bb * Tb * C * C * k / us / 1e6
                    r['hbm_GBs'] = 2.0 * Bb * Tb * C * 2 / us / 1e3     # x in + x' out
                results[name] = r
                lib.ttsb_convpair_destroy(h)
    flag = ctypes.c_int(0)
    try:
        lib.ttsb_device_error_flag(ctypes.byref(flag))
    except Exception:
        pass
    print('RESULT ' + json.dumps({'device_flag': flag.value, 'cases': results}))


def main():
    if '--child' in sys.argv:
        batch = int(sys.argv[sys.argv.index('--batch') + 1]) if '--batch' in sys.argv else 16
        child('--bench' in sys.argv, batch)
        return
    out_dir = os.path.join(REPO, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--child'] + sys.argv[1:], capture_output=True,
                           text=True, timeout=420)
        line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
        s = json.loads(line[-1][7:]) if line else {'crashed': True, 'rc': r.returncode, 'stderr': r.stderr[-2000:],
                                                   'stdout': r.stdout[-500:]}
    except subprocess.TimeoutExpired:
        s = {'timeout': True}
    with open(os.path.join(out_dir, 'probe_pair.json'), 'w') as f:
        json.dump(s, f, indent=1)
    if 'cases' not in s:
        print('FAILED TO RUN', json.dumps(s)[:2500])
        return
    print('device_flag=%d' % s['device_flag'])
    for n, c in s['cases'].items():
        if not c.get('fused', True):
            print('%-12s not fused' % n)
            continue
        line = '%-12s %s max_err=%.2e plan=%s' % (n, 'ok ' if c.get('ok') else 'BAD', c.get('max_err', -1), c.get('plan'))
        if 'us' in c:
            line += '  %8.1f us %7.1f TFLOP/s %7.1f GB/s(x in + x out)' % (c['us'], c['tflops'], c['hbm_GBs'])
        if not c.get('ok'):
            line += ' ' + json.dumps({k: v for k, v in c.items() if k in ('error', 'n_bad', 'first_bad', 'bad_rows_mod', 'nan')})
        print(line)


if __name__ == '__main__':
    main()
