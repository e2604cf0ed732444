"""Per-launch floors of the HiFi-GAN generator against an ncu launch list (no GPU needed).

    python tools/layer_floors.py profiles/r01_s43_launches_b64.csv 32768 > profiles/r01_layer_floors_b64.txt

For each of the 60 generator launches of one chunk (argument 2 = mel frames in the chunk) it prints the measured
duration next to three lower bounds computed from the layer shape (vocoder/hifigan/models.py:111-127, 46-53):

  mma   tensor-pipe issue time: (#tcgen05.mma per launch) x (cycles per M=128, K=16 instruction, measured in
        isolation on B200: profiles/r01_s21_mma_rate.txt) / (148 SMs x clock)
  hbm   compulsory DRAM traffic (activations in + out, residual, MRF accumulator) / measured copy bandwidth

and, for information, `w TB/s`: the L2 -> shared-memory rate the launch sustains just for its weight tiles when they
are streamed once per 128-row tile (layers whose packed weights do not fit next to the panel ring; CTA pairs share
each tile through TMA multicast, which halves it). ncu reports lts__throughput of 20-25 % for the k = 11 launches
(profiles/r01_s52_ncu_full_b16.csv), so this stream is not what bounds them.

ncu durations are serialised and cold-cache; they are used for shares and ratios, not as bench numbers.
"""
import csv
import re
import sys

SMS = 148
CLK = 1.75e9                 # SM clock under ncu for these launches (1.64-1.80 GHz in profiles/r01_s52_ncu_full_b16.csv)
HBM = 6458.1e9               # MEASURED_PEAKS.json hbm_gbs
MMA_CYC_SW128 = {32: 42, 64: 48, 128: 64, 256: 128}   # cycles per tcgen05.mma M=128 K=16 by N, 128-byte operand rows
MMA_CYC_SW64 = {32: 69, 64: 77, 128: 93, 256: 150}    # 64-byte operand rows (C = 32 layers)


def mma_cycles(n, c_in):
    table = MMA_CYC_SW64 if c_in < 64 else MMA_CYC_SW128
    per = 0
    while n > 0:
        step = min(n, 256)
        per += table[max(32, step)]
        n -= step
    return per


def layer(name, rows, c_in, c_out, taps, extra_rw=0, resident=False, fused=1):
    """One launch: `fused` convs of `taps` taps each over `rows` output rows."""
    tiles = rows / 128.0
    n_mma = tiles * fused * taps * (max(c_in, 16) / 16.0)
    t_mma = n_mma * mma_cycles(c_out, c_in) / (SMS * CLK)
    w_bytes = fused * taps * c_in * c_out * 2
    w_stream = 0.0 if resident else tiles * w_bytes / 2      # bytes; CTA pairs multicast each tile
    act = rows * (c_in + c_out) * 2 + extra_rw
    t_hbm = act / HBM
    flops = 2.0 * rows * c_in * c_out * taps * fused
    return dict(name=name, flops=flops, mma=t_mma, wbytes=w_stream, hbm=t_hbm)


def generator_layers(frames):
    out = [layer('conv_pre 80->512 k7', frames, 80, 512, 7)]
    ch = [512, 256, 128, 64, 32]
    rows_per_frame = [8, 64, 128, 256]
    up_taps = [2, 2, 2, 2]      # polyphase: k16 s8 and k4 s2 both touch two input rows per output row
    for s in range(4):
        c_in, c = ch[s], ch[s + 1]
        rows = frames * rows_per_frame[s]
        # the up-sampler reads rows/stride input rows of c_in channels
        stride = 8 if s < 2 else 2
        up = layer('ups%d %d->%d' % (s, c_in, c), rows, c_in, c, up_taps[s])
        up['hbm'] = (rows / stride * c_in + rows * c * (1 if ACT_CHAIN else (2 if c > 64 else 1))) * 2 / HBM
        out.append(up)
        for k in (3, 7, 11):
            for j, d in enumerate((1, 3, 5)):
                last = j == 2
                mrf = rows * c * 2 * (1 if k == 3 else 2) if last else 0      # MRF accumulator: write (k=3) / read + write
                if c > 64:
                    out.append(layer('s%d C%d k%d d%d conv1' % (s, c, k, d), rows, c, c, k))
                    # conv2 reads the residual and writes the sum: once, activated, in the round-2 data flow
                    # (csrc/hifigan.cu); round 1 wrote the raw sum AND its leaky-relu
                    out.append(layer('s%d C%d k%d d1 conv2' % (s, c, k), rows, c, c, k,
                                     extra_rw=rows * c * 2 * (1 if ACT_CHAIN else 2) + mrf))
                else:
                    pair = layer('s%d C%d k%d d%d pair' % (s, c, k, d), rows, c, c, k, extra_rw=mrf, resident=True, fused=2)
                    if c == 32 and k >= 5 and TT_PAIR:
                        # conv2 contracts two taps per K = 64 group on 128-byte rows (conv_pair.cu, tt_pair): the same
                        # number of instructions at 42 instead of 69 cycles
                        n_each = rows / 128.0 * k * 2
                        pair['mma'] = n_each * (MMA_CYC_SW64[32] + MMA_CYC_SW128[32]) / (SMS * CLK)
                    out.append(pair)
    out.append(dict(name='conv_post 32->1 k7 + tanh', flops=2.0 * frames * 256 * 32 * 7, mma=0.0, wbytes=0.0,
                    hbm=frames * 256 * (32 * 2 + 4) / HBM))
    return out


ACT_CHAIN = True     # round-2 data flow; pass a third argument `r1` for round-1 launch lists
TT_PAIR = True       # round-2 C = 32 conv2 (see generator_layers); off with `r1`


def main():
    global ACT_CHAIN, TT_PAIR
    path, frames = sys.argv[1], int(sys.argv[2])
    if len(sys.argv) > 3 and sys.argv[3] == 'r1':
        ACT_CHAIN = False
        TT_PAIR = False
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))][1:]
    durs = [(re.sub(r'\(.*', '', r[4]).replace('void ttsb::', '').replace('ttsb::', ''),
             float(r[-1].replace(',', '')) * 1e-9) for r in rows]
    # align on the generator's own launch train: [pack_mel,] conv_pre ... conv_post_tanh (tools/run_vocoder.py
    # --profile-last captures exactly one pass)
    post = [i for i, (n, _) in enumerate(durs) if n.startswith('conv_post_tanh')]
    layers = generator_layers(frames)
    if post:
        first = post[-1] - len(layers) + 1
        if first < 0:           # the capture began after the pass did: keep the launches it has, right-aligned
            layers = layers[-first:]
            first = 0
        durs = durs[first:post[-1] + 1]
    durs = durs[:len(layers)]
    print('generator launches of one %d-frame chunk: measured (ncu, serialised) vs floors at %.2f GHz' % (frames, CLK / 1e9))
    print('%-28s %-24s %8s %8s %8s %7s %8s %7s  %s' % ('layer', 'kernel', 'meas us', 'mma us', 'hbm us', 'x floor',
                                                      'TFLOP/s', 'w TB/s', 'binding floor'))
    tot = dict(meas=0.0, mma=0.0, hbm=0.0, floor=0.0, flops=0.0)
    stage = {}
    for L, (kname, t) in zip(layers, durs):
        fl = {k: L[k] for k in ('mma', 'hbm')}
        bind = max(fl, key=fl.get)
        floor = fl[bind]
        print('%-28s %-24s %8.1f %8.1f %8.1f %7.2f %8.0f %7.1f  %s' % (L['name'], kname[:24], t * 1e6, L['mma'] * 1e6,
                                                                      L['hbm'] * 1e6, t / floor, L['flops'] / t / 1e12,
                                                                      L['wbytes'] / t / 1e12, bind))
        tot['meas'] += t
        tot['floor'] += floor
        tot['flops'] += L['flops']
        for k in fl:
            tot[k] += fl[k]
        key = L['name'].split(' ')[0]
        if key.startswith('ups'):
            key = 's' + key[3]
        a = stage.setdefault(key, [0.0, 0.0, 0.0])
        a[0] += t
        a[1] += floor
        a[2] += L['flops']
    print()
    print('%-12s %9s %7s %9s %8s %9s' % ('stage', 'meas us', 'share', 'floor us', 'x floor', 'TFLOP/s'))
    for key, (t, f, fl) in stage.items():
        print('%-12s %9.0f %6.1f%% %9.0f %8.2f %9.0f' % (key, t * 1e6, 100 * t / tot['meas'], f * 1e6, t / f, fl / t / 1e12))
    print('%-12s %9.0f %6s  %9.0f %8.2f %9.0f' % ('total', tot['meas'] * 1e6, '', tot['floor'] * 1e6,
                                                  tot['meas'] / tot['floor'], tot['flops'] / tot['meas'] / 1e12))
    print('sum of per-launch floors: mma %.0f us, hbm %.0f us' % (tot['mma'] * 1e6, tot['hbm'] * 1e6))


if __name__ == '__main__':
    main()
