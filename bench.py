#!/usr/bin/env python
"""Benchmarks of the batched text -> mel -> waveform path on synthetic inputs (BASELINE.json), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config target|c2|c3|c4|c5]
  python bench.py --impl reference ...      # the reference's CPU PyTorch arithmetic (oracle port) on the host cores

  target  FastPitch2Wave end to end, 256 x 128-phoneme utterances per GPU (the metric's configuration; weak scaling)
  c2      HiFi-GAN Generator alone, mel [1,80,512] (latency-bound: reported with and without CUDA-graph replay)
  c3      FastPitch2Wave end to end, 32 x 128 phonemes per GPU
  c4      Tacotron2Wave: autoregressive decode (256 steps) + HiFi-GAN, batch 8 (replicas only: one batch per GPU)
  c5      FastPitch2Wave, 512 utterances of 64..256 phonemes sharded over the N GPUs (strong scaling, parallel.synthesize)

Prints ONE JSON line on rank 0 (contract in the task statement): `value` = device-resident throughput (inputs in HBM ->
waveforms in HBM), `e2e` = the same work through the public API (`parallel.synthesize` -> `FastPitch2Wave.synthesize_ids`)
with host token ids in and host waveforms out inside the timed region, `roofline` for the dominant kernel class (the
tcgen05 conv launches of the HiFi-GAN generator) plus `roofline.per_kernel` (CUDA-event time per stage from the library's
stage profiler against algorithmic FLOPs and bytes), `cpu_baseline` (oracle on this box's host cores, bounded sample) and
`gpu_eager_baseline` (the same reference arithmetic as PyTorch eager on this GPU: cuDNN / cuBLAS library kernels).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

HOP = 256
SR = 22050.0
FRAMES_PER_PHONEME = 4                      # const-4 duration head (SURVEY.md §8c calibration)
VOCODER_FLOP_PER_FRAME = 614.1e6
# dram__bytes_read.sum + dram__bytes_write.sum over the generator launches of one 32768-frame pass, `ncu --set full`
# (the capture named below); bench.py cannot read DRAM counters itself, so `roofline.traffic` = this per-frame figure x
# the frames of a step. Re-captured whenever the generator's data flow changes.
VOCODER_DRAM_BYTES_PER_FRAME = 54.45e9 / 32768
VOCODER_DRAM_SOURCE = 'profiles/r02_final_ncu_full_b64.csv (ncu --set full on the final round-2 kernels, 33.12 GB read + 21.33 GB written per 32768-frame pass, per frame x frames)'

# csrc/common.cuh ProfTag
PROF_TAGS = ['untagged', 'voc.conv_pre', 'voc.ups0', 'voc.s0', 'voc.ups1', 'voc.s1', 'voc.ups2', 'voc.s2', 'voc.ups3',
             'voc.s3', 'voc.conv_post', 'voc.pack_mel', 'fp.embed', 'fp.qkv_o', 'fp.attention', 'fp.ffn', 'fp.predictors',
             'fp.pitch_energy_durations', 'fp.regulate', 'fp.proj']


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=5)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--config', default='target', choices=['target', 'c2', 'c3', 'c4', 'c5'])
    p.add_argument('--batch', type=int, default=0, help='utterances per GPU per step (0 = the config default)')
    p.add_argument('--phonemes', type=int, default=128)
    p.add_argument('--cpu-sample', type=int, default=0, help='utterances per CPU-baseline step (0 = auto)')
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--no-eager-baseline', action='store_true')
    p.add_argument('--no-per-kernel', action='store_true')
    p.add_argument('--deliver', default='auto', choices=['auto', 'nccl_host', 'host_shm'],
                   help='how e2e results reach rank 0 host memory at N > 1')
    return p.parse_args()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def usable_cores():
    """Host threads this process may actually use: affinity mask, capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        q, p = open('/sys/fs/cgroup/cpu.max').read().split()
        if q != 'max':
            n = max(1, min(n, int(float(q) / float(p) + 0.5)))
    except Exception:
        pass
    return n


def load_peaks():
    try:
        return json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------------------------
# algorithmic work per stage (the denominators of roofline.per_kernel)
# ------------------------------------------------------------------------------------------------------------------
def vocoder_stage_model(frames):
    """tag -> (FLOPs, compulsory HBM bytes) for `frames` mel frames through the HiFi-GAN V1 generator
    (vocoder/hifigan/models.py:111-127, 46-53; FLOPs = 2 MAC as SURVEY.md §8d; bytes = fp16 activations of the data flow
    in csrc/hifigan.cu: one stored tensor per ResBlock input, conv pairs fused for C <= 64, MRF accumulator in fp16)."""
    ch = [512, 256, 128, 64, 32]
    rpf = [8, 64, 128, 256]
    strides = [8, 8, 2, 2]
    out = {'voc.conv_pre': (2.0 * frames * 80 * 512 * 7, frames * (128 + 512) * 2.0),
           'voc.conv_post': (2.0 * frames * 256 * 32 * 7, frames * 256 * (32 * 2 + 4.0))}
    for s in range(4):
        c_in, c = ch[s], ch[s + 1]
        rows = frames * rpf[s]
        out['voc.ups%d' % s] = (2.0 * rows * c_in * c * 2, (rows / strides[s] * c_in + rows * c) * 2.0)
        flops = sum(2.0 * rows * c * c * k * 6 for k in (3, 7, 11))
        passes_per_pair = 2 if c <= 64 else 5            # fused: x in, x' out; two launches: x, t, t, x(res), x'
        byts = rows * c * 2.0 * (9 * passes_per_pair + 2)  # + the MRF accumulator's two extra read passes
        out['voc.s%d' % s] = (flops, byts)
    return out


def fastpitch_stage_model(b, l, t, cfg):
    """tag -> (FLOPs, bytes) of FastPitch.infer for b utterances, l tokens, t frames each (SURVEY.md §8d per-token figures)."""
    d, dh, di, k = cfg['symbols_embedding_dim'], cfg['in_fft_d_head'], cfg['in_fft_conv1d_filter_size'], 3
    layers_rows = cfg['in_fft_n_layers'] * b * l + cfg['out_fft_n_layers'] * b * t
    att = 0.0
    for n_layers, s in ((cfg['in_fft_n_layers'], l), (cfg['out_fft_n_layers'], t)):
        att += n_layers * b * s * (4.0 * s * dh)          # QK^T + PV
    pf = cfg['dur_predictor_filter_size']
    n_pred = 3 if cfg['energy_conditioning'] else 2
    return {
        'fp.qkv_o': (layers_rows * 2.0 * (d * 3 * dh + dh * d), layers_rows * (d + 3 * dh + dh + 2 * d) * 2.0),
        'fp.attention': (att, layers_rows * (3 * dh + dh) * 2.0),
        'fp.ffn': (layers_rows * 2.0 * 2 * d * di * k, layers_rows * (d + di + di + 2 * d) * 2.0),
        'fp.predictors': (n_pred * b * l * 2.0 * (d * pf * 3 + pf * pf * 3 + pf), n_pred * b * l * (d + 2 * pf) * 2.0),
        'fp.regulate': (0.0, b * t * d * 2 * 2.0),
        'fp.proj': (b * t * 2.0 * d * 80, b * t * (d * 2 + 80 * 4 + 128 * 2.0)),
        'fp.embed': (0.0, b * l * d * 2.0),
        'fp.pitch_energy_durations': (0.0, 2 * b * l * d * 2 * 2.0),
    }


def per_kernel_table(ms_by_tag, model, peaks, steps):
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_hbm = peaks.get('hbm_gbs', 6650.0)
    rows = {}
    total = sum(ms_by_tag.values()) or 1.0
    for tag, ms in ms_by_tag.items():
        if ms <= 0 or tag == 'untagged':
            continue
        fl, by = model.get(tag, (0.0, 0.0))
        sec = ms * 1e-3 / steps
        tf = fl / sec / 1e12
        gb = by / sec / 1e9
        rows[tag] = {'ms_per_step': round(ms / steps, 4), 'share': round(ms / total, 4),
                     'tflops': round(tf, 1), 'tensor_frac': round(tf / peak_tf, 4),
                     'hbm_gbs': round(gb, 1), 'hbm_frac': round(gb / peak_hbm, 4),
                     'bound': 'tensor' if tf / peak_tf >= gb / peak_hbm else 'hbm'}
    return rows


# ------------------------------------------------------------------------------------------------------------------
# reference arithmetic (oracle port): CPU arm and the eager-GPU baseline
# ------------------------------------------------------------------------------------------------------------------
_REF_MODELS = {}


def reference_kind():
    """'reference' when the unmodified reference modules are vendored under oracle/_ref (oracle/make_ref.py; they travel to
    the GPU box with the snapshot), else 'port' (the oracle restatement)."""
    from oracle import ref_runner
    return 'reference' if ref_runner.available() and os.environ.get('TTSB_BENCH_REF', '1') != '0' else 'port'


def reference_step(fsd, gsd_folded, ids, dtype=None, batched_vocoder=False):
    """The reference's path: FastPitch.infer on the padded batch, then the generator once per utterance
    (models/fastpitch/networks.py:322-350) — or once over the padded batch (`batched_vocoder`, what a user would write to
    help cuDNN; padded frames then differ from the per-utterance result, so it is a throughput-only variant).
    Runs the UNMODIFIED reference modules (oracle/_ref, same seeded synthetic checkpoints) when they are present, else the
    oracle restatement over the state dicts passed in."""
    import torch
    if reference_kind() == 'reference':
        from oracle import ref_runner
        from tts_arabic_pytorch_b200.utils import synth
        key = (str(ids.device), dtype or torch.float32)
        if key not in _REF_MODELS:
            sds = _REF_MODELS.get('state_dicts')
            if sds is None:
                sds = (synth.fastpitch_state_dict(1234), synth.hifigan_state_dict(1235))
            _REF_MODELS.clear()                      # one resident copy: the eager-GPU leg walks dtypes one after another
            _REF_MODELS['state_dicts'] = sds
            _REF_MODELS[key] = ref_runner.build_models(sds[0], synth.FASTPITCH_CONFIG, sds[1], synth.HIFIGAN_CONFIG,
                                                       device=ids.device, dtype=key[1])
        fp, voc = _REF_MODELS[key]
        n, wavs = ref_runner.step(fp, voc, ids, batched_vocoder)
        return n, (wavs[0] if batched_vocoder else wavs)
    from oracle import fastpitch_oracle as fpo
    from oracle import hifigan_oracle as hgo
    from tts_arabic_pytorch_b200.utils import synth
    dtype = dtype or torch.float32
    with torch.no_grad():
        mel, dec_lens, *_ = _fp_infer_on(fsd, synth.FASTPITCH_CONFIG, ids, dtype)
        if batched_vocoder:
            wav = hgo.generator_forward(gsd_folded, synth.HIFIGAN_CONFIG, mel, dtype)
            return int(dec_lens.sum()) * HOP, wav
        wavs = hgo.vocode_batch(gsd_folded, synth.HIFIGAN_CONFIG, mel, dec_lens, dtype)
    return sum(int(w.numel()) for w in wavs), wavs


def _fp_infer_on(fsd, cfg, ids, dtype):
    """oracle fastpitch_infer with its host-created index tensors placed on the ids' device (it is written for the CPU)."""
    import torch
    from oracle import fastpitch_oracle as fpo
    if ids.device.type == 'cpu':
        return fpo.fastpitch_infer(fsd, cfg, ids, dtype=dtype)
    prev = torch.get_default_device() if hasattr(torch, 'get_default_device') else None
    torch.set_default_device(ids.device)
    try:
        return fpo.fastpitch_infer(fsd, cfg, ids, dtype=dtype)
    finally:
        torch.set_default_device(prev if prev is not None else 'cpu')


def eager_gpu_baseline(dev, fsd, gsd_folded, ids_dev, budget_s=60.0):
    """The reference arithmetic as PyTorch eager on this GPU (cuDNN convs, cuBLAS GEMMs): fp32 (TF32 convs, torch's
    default) and fp16, vocoder per utterance as the reference does and batched. Bounded sample, CUDA events."""
    import torch
    out = {'kind': reference_kind(), 'library': 'PyTorch eager %s (cuDNN/cuBLAS)' % torch.__version__,
           'sample': '%d x %d phonemes' % tuple(ids_dev.shape), 'unit': 'samples/s'}
    t_start = time.perf_counter()
    for name, dtype in (('fp16', torch.float16), ('fp32', torch.float32)):
        w_f = {k: v.to(dev) for k, v in fsd.items()}
        w_g = {k: v.to(dev) for k, v in gsd_folded.items()}
        for variant, batched in (('batched_vocoder', True), ('per_utterance_vocoder', False)):
            if time.perf_counter() - t_start > budget_s:
                out['%s_%s' % (name, variant)] = None
                continue
            try:
                n, _ = reference_step(w_f, w_g, ids_dev, dtype, batched)          # warm-up (cuDNN autotune, allocator)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                reps = 2
                for _ in range(reps):
                    n, _ = reference_step(w_f, w_g, ids_dev, dtype, batched)
                e1.record()
                torch.cuda.synchronize()
                out['%s_%s' % (name, variant)] = n * reps / (e0.elapsed_time(e1) * 1e-3)
            except Exception as exc:                                           # e.g. out of memory at a large sample
                out['%s_%s' % (name, variant)] = 'failed: %s' % type(exc).__name__
                torch.cuda.empty_cache()
        del w_f, w_g
        torch.cuda.empty_cache()
    vals = [v for k, v in out.items() if isinstance(v, float)]
    out['value'] = max(vals) if vals else None        # the strongest library configuration is the one to beat
    return out


def cpu_baseline(fsd, gsd_folded, ids_host, cores, cpu_sample, label):
    import torch
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    reference_step(fsd, gsd_folded, ids_host[:1].clone())
    per_utt = time.perf_counter() - t0
    bs = cpu_sample or max(1, min(ids_host.shape[0], 32, int(6.0 / max(per_utt, 1e-3))))
    ids = ids_host[:bs].clone()
    t0 = time.perf_counter()
    n, reps = 0, 0
    while time.perf_counter() - t0 < 15.0 and reps < 16:
        n += reference_step(fsd, gsd_folded, ids)[0]
        reps += 1
    dt = time.perf_counter() - t0
    return {'value': n / dt, 'unit': 'samples/s', 'cores': cores, 'kind': reference_kind(),
            'sample': '%d x %d utterances x %d phonemes, fp32 torch CPU (%s)' % (reps, bs, ids.shape[1], label)}


def run_reference(args):
    import torch
    from tts_arabic_pytorch_b200.utils import synth
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = min(usable_cores(), 32)
    torch.set_num_threads(cores)
    gsd = synth.fold_weight_norm(synth.hifigan_state_dict(1235))
    gen = torch.Generator().manual_seed(0)
    cfgname = args.config
    kind = 'port'      # c4 (Tacotron2 lives in torchaudio, not vendored): the oracle restatement
    if cfgname == 'c2':
        from oracle import hifigan_oracle as hgo
        mel = torch.clamp(torch.randn(1, 80, 512, generator=gen) * 2 - 5, -11.5129, 2.0)

        if reference_kind() == 'reference':
            from oracle import ref_runner
            _, ref_voc = ref_runner.build_models(synth.fastpitch_state_dict(1234), synth.FASTPITCH_CONFIG,
                                                 synth.hifigan_state_dict(1235), synth.HIFIGAN_CONFIG)

            def step():
                with torch.no_grad():
                    return int(ref_voc(mel).numel())
            kind = 'reference'
        else:
            def step():
                with torch.no_grad():
                    return int(hgo.generator_forward(gsd, synth.HIFIGAN_CONFIG, mel).numel())
        workload = 'HiFi-GAN Generator, mel [1,80,512] -> 131072 samples, host CPU'
        sample = 'the whole configuration'
    elif cfgname == 'c4':
        from oracle import hifigan_oracle as hgo
        from oracle import tacotron2_oracle as t2o
        sd = synth.tacotron2_state_dict(1236)
        bs, steps_dec = (args.cpu_sample or 2), 64
        tokens = torch.randint(1, 40, (bs, 64), generator=gen)
        masks = (torch.rand(steps_dec, 2, bs, 256, generator=gen) > 0.5).float() * 2.0

        def step():
            with torch.no_grad():
                mel, lens, _ = t2o.tacotron2_infer(sd, tokens, torch.zeros(bs, dtype=torch.long),
                                                   torch.full((bs,), 64, dtype=torch.long), prenet_masks=masks,
                                                   max_steps=steps_dec)
                return sum(int(w.numel()) for w in hgo.vocode_batch(gsd, synth.HIFIGAN_CONFIG, mel, lens))
        workload = 'Tacotron2Wave, 64 tokens, bounded sample of %d utterances x %d decoder steps on the host CPU' % (bs, steps_dec)
        sample = '%d utterances x %d decoder steps per step' % (bs, steps_dec)
    else:
        fsd = synth.fastpitch_state_dict(1234)
        bs = args.cpu_sample or 4
        if cfgname == 'c5':
            lens = sorted(torch.randint(64, 257, (bs,), generator=gen).tolist(), reverse=True)
            ids = torch.zeros(bs, lens[0], dtype=torch.long)
            for b, n in enumerate(lens):
                ids[b, :n] = torch.randint(1, 40, (n,), generator=gen)
            desc = 'mixed 64..256-phoneme'
        else:
            ids = torch.randint(1, 40, (bs, args.phonemes), generator=gen)
            desc = '%d-phoneme' % args.phonemes

        kind = reference_kind()

        def step():
            return reference_step(fsd, gsd, ids)[0]
        workload = ('FastPitch2Wave end-to-end, %s synthetic utterances (4 frames/phoneme), bounded sample of %d utterances '
                    'per step on the host CPU' % (desc, bs))
        sample = '%d utterances per step' % bs
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        n += step()
    dt = time.perf_counter() - t0
    v = n / dt
    line = {
        'impl': 'reference', 'metric': 'audio_samples_per_sec', 'value': v, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': workload, 'name': cfgname, 'phonemes': args.phonemes, 'sample_batch': args.cpu_sample or None},
        'rtf': (dt / (n / SR)),
        'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': kind,
                         'sample': '%d steps x %s' % (args.steps, sample)},
        'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def setup(args):
    import torch
    import torch.distributed as dist
    from tts_arabic_pytorch_b200 import _lib
    c = Ctx()
    c.world = int(os.environ.get('WORLD_SIZE', '1'))
    c.rank = int(os.environ.get('RANK', '0'))
    c.local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (impl=ours) needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(c.local)
    c.dev = torch.device('cuda', c.local)
    if c.world > 1:
        dist.init_process_group('nccl', device_id=c.dev)
    c.lib = _lib.load()
    c.peaks = load_peaks()
    return c


def barrier(c):
    import torch
    import torch.distributed as dist
    if c.world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(c, fn, steps):
    """EXACTLY `steps` calls between two barriers + synchronize, CUDA events on the current stream, max over ranks."""
    import torch
    import torch.distributed as dist
    barrier(c)
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    n0 = c.lib.ttsb_launch_count()
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    barrier(c)
    ms = s.elapsed_time(e)
    launches = c.lib.ttsb_launch_count() - n0
    if c.world > 1:
        t = torch.tensor([ms], device=c.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms, launches


def profile_stages(c, fn, steps):
    n = c.lib.ttsb_prof_n_tags()
    assert n == len(PROF_TAGS), 'bench.py PROF_TAGS is out of date with csrc/common.cuh ProfTag'
    import torch
    torch.cuda.synchronize()
    c.lib.ttsb_prof_enable(1)
    for _ in range(steps):
        fn()
    buf = (ctypes.c_double * n)()
    from tts_arabic_pytorch_b200 import _lib
    _lib.check(c.lib.ttsb_prof_collect(buf, n))
    c.lib.ttsb_prof_enable(0)
    return {PROF_TAGS[i]: float(buf[i]) for i in range(n)}


def check_device_flag(c):
    from tts_arabic_pytorch_b200 import _lib
    flag = ctypes.c_int(0)
    _lib.check(c.lib.ttsb_device_error_flag(ctypes.byref(flag)))
    if flag.value != 0:
        raise RuntimeError('device error flag %d' % flag.value)


def make_tts(c, tmpdir):
    """FastPitch2Wave built from the synthetic checkpoints through the public constructor (checkpoint files and all)."""
    from tts_arabic_pytorch_b200.models.fastpitch import FastPitch2Wave
    from tts_arabic_pytorch_b200.utils import synth
    fp, hg, cj = synth.write_checkpoints(tmpdir, seed=1234)
    return FastPitch2Wave(fp, vocoder_sd=hg, vocoder_config=cj, arabic_in=False).to(c.dev)


def run_fastpitch2wave(args, c, cfgname):
    """target / c3 / c5."""
    import tempfile
    import torch
    import torch.distributed as dist
    from tts_arabic_pytorch_b200 import parallel
    from tts_arabic_pytorch_b200.utils import synth

    tmp = tempfile.mkdtemp(prefix='ttsb_bench_%d_' % c.rank)
    tts = make_tts(c, tmp)
    fp, voc = tts.model, tts.vocoder
    gen = torch.Generator().manual_seed(1000)
    strong = cfgname == 'c5'
    if strong:
        total = args.batch or 512
        lens = torch.randint(64, 257, (total,), generator=torch.Generator().manual_seed(0)).tolist()
        id_list = [torch.randint(1, 40, (n,), generator=gen) for n in lens]
        workload = ('FastPitch2Wave end-to-end, %d utterances of 64..256 phonemes (4 frames/phoneme), sharded over %d GPU(s) '
                    'by parallel.synthesize (snake deal + global padding condition)' % (total, c.world))
    else:
        per_gpu = args.batch or (256 if cfgname == 'target' else 32)
        total = per_gpu * c.world
        id_list = [torch.randint(1, 40, (args.phonemes,), generator=gen) for _ in range(total)]
        workload = ('FastPitch2Wave end-to-end, batch %d x %d phonemes per GPU (4 frames/phoneme -> %d frames -> %d samples '
                    'per utterance), HiFi-GAN V1' % (per_gpu, args.phonemes, args.phonemes * 4, args.phonemes * 4 * HOP))
    lengths = [int(x.numel()) for x in id_list]
    shards, pad_to = parallel.plan_shards(lengths, c.world)
    mine = shards[c.rank]
    total_samples = sum(lengths) * FRAMES_PER_PHONEME * HOP
    my_frames = sum(lengths[i] for i in mine) * FRAMES_PER_PHONEME
    l_max = max(lengths)
    t_global = l_max * FRAMES_PER_PHONEME

    # device-resident arm: this rank's padded id batch already in HBM
    from tts_arabic_pytorch_b200.models.fastpitch.networks import text_collate_fn
    padded, _, _ = text_collate_fn([id_list[i] for i in mine])
    if pad_to[c.rank] > padded.shape[1]:
        padded = torch.nn.functional.pad(padded, (0, pad_to[c.rank] - padded.shape[1]))
    ids_dev = padded.to(c.dev)
    hook = (lambda t: t + 1 if t < t_global else t) if c.world > 1 else None
    voc_ev = []

    def step_device():
        mel, dec_lens, _, _, _, mel_cl = fp.infer(ids_dev, return_channel_last=True, frame_len_hook=hook)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wav = voc.run(mel_cl=mel_cl, lens=dec_lens)
        e1.record()
        voc_ev.append((e0, e1))
        if c.world > 1:
            parallel.gather_waveforms(wav, dec_lens * HOP, dst=0)
        return wav

    deliver = args.deliver
    if deliver == 'auto':
        deliver = 'host_shm' if c.world > 1 else 'nccl_host'

    def step_e2e():
        return parallel.synthesize(tts, id_list, deliver=deliver)

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    voc_ev.clear()
    sampler = ClockSampler(c.local)
    if c.rank == 0:
        sampler.start()
    ms_dev, launches = timed(c, step_device, args.steps)
    voc_ms = sum(a.elapsed_time(b) for a, b in voc_ev)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(c, step_e2e, args.steps)
    clocks = sampler.stop() if c.rank == 0 else None
    stage_ms = None
    if not args.no_per_kernel:
        stage_ms = profile_stages(c, _profiled_step(fp, voc, ids_dev, hook), min(args.steps, 3))
    check_device_flag(c)
    imbalance = None
    if c.world > 1:
        fr = torch.tensor([my_frames], dtype=torch.int64, device=c.dev)
        frs = [torch.empty_like(fr) for _ in range(c.world)]
        dist.all_gather(frs, fr)
        imbalance = [int(f[0]) for f in frs]
    if c.rank != 0:
        if c.world > 1:
            dist.destroy_process_group()
        return

    value = total_samples * args.steps / (ms_dev * 1e-3)
    e2e_value = total_samples * args.steps / (ms_e2e * 1e-3)
    peaks = c.peaks
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_hbm = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback (B200_PROFILING.md sustained)'
    voc_flops = VOCODER_FLOP_PER_FRAME * my_frames * args.steps
    achieved_tf = voc_flops / (voc_ms * 1e-3) / 1e12 if voc_ms > 0 else 0.0
    voc_traffic = VOCODER_DRAM_BYTES_PER_FRAME * my_frames
    hbm_gbs = voc_traffic * args.steps / (voc_ms * 1e-3) / 1e9 if voc_ms > 0 else 0.0
    roof = {'kernel': 'conv_tc2_kernel + conv_pair_kernel (tcgen05 row-GEMM-with-taps; all HiFi-GAN generator launches of a '
                      'step, rank 0)', 'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
            'frac': achieved_tf / peak_tf, 'peak_source': peak_src, 'traffic': voc_traffic,
            'traffic_source': VOCODER_DRAM_SOURCE, 'flops_per_step': voc_flops / args.steps,
            'ms_per_step': voc_ms / args.steps, 'share_of_step': voc_ms / ms_dev if ms_dev > 0 else None,
            'hbm': {'achieved': hbm_gbs, 'peak': peak_hbm, 'unit': 'GB/s', 'frac': hbm_gbs / peak_hbm}}
    if stage_ms is not None:
        model = vocoder_stage_model(my_frames)
        b_loc = len(mine)
        model.update(fastpitch_stage_model(b_loc, int(ids_dev.shape[1]), t_global, synth.FASTPITCH_CONFIG))
        roof['per_kernel'] = per_kernel_table(stage_ms, model, peaks, min(args.steps, 3))
        roof['per_kernel_note'] = ('CUDA events recorded by the library before each stage on the launching stream in a '
                                   'separate pass (ttsb_prof_*); FLOPs/bytes = algorithmic figures of SURVEY.md 8d for this '
                                   "rank's batch; fractions of the measured sustained bf16 peak and measured copy bandwidth")
    h2d = sum(lengths[i] for i in mine) * 8
    d2h = total_samples * 4 if c.world == 1 or deliver != 'host_shm' else my_frames * HOP * 4
    line = {
        'metric': 'audio_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': c.world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
        'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'fp16', 'data': 'synthetic',
        'config': {'workload': workload, 'name': cfgname, 'global_batch': total, 'phonemes': args.phonemes if not strong else '64..256',
                   'parallelism': 'dp%d' % c.world,
                   'l2': 'per-step activations (GBs) exceed the 126 MB L2; only the 28 MB of packed weights stay resident',
                   'conv_impl': 'tcgen05' if c.lib.ttsb_get_conv_impl() == 0 else 'simt'},
        'rtf': (ms_dev * 1e-3 / args.steps) / (total_samples / SR),
        'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps, 'api': 'parallel.synthesize -> FastPitch2Wave.synthesize_ids',
                'deliver': deliver,
                'note': 'host id tensors in, host waveforms of EVERY utterance on rank 0 out; %s' % (
                    'each rank copies its rows device -> shared pinned host memory over its own PCIe link (d2h bytes are per rank)'
                    if deliver == 'host_shm' else 'rank 0 copies every row device -> host')},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roof,
    }
    if imbalance is not None:
        line['config']['frames_per_rank'] = imbalance
    if c.world == 1 and cfgname in ('target', 'c3'):
        fsd = synth.fastpitch_state_dict(1234)
        gsd_f = synth.fold_weight_norm(synth.hifigan_state_dict(1235))
        ids_host = torch.stack([id_list[i] for i in mine])
        if not args.no_eager_baseline:
            sample = ids_host[:min(64, ids_host.shape[0])].to(c.dev)
            line['gpu_eager_baseline'] = eager_gpu_baseline(c.dev, fsd, gsd_f, sample)
        if not args.no_cpu_baseline:
            cores = min(usable_cores(), 32)     # torch CPU convs at batch 1 stop scaling well before this
            line['cpu_baseline'] = cpu_baseline(fsd, gsd_f, ids_host, cores, args.cpu_sample, 'bounded sample')
    if c.world > 1:
        dist.destroy_process_group()
    print(json.dumps(line))


def _profiled_step(fp, voc, ids_dev, hook):
    def f():
        mel, dec_lens, _, _, _, mel_cl = fp.infer(ids_dev, return_channel_last=True, frame_len_hook=hook)
        return voc.run(mel_cl=mel_cl, lens=dec_lens)
    return f



def run_c2(args, c):
    """HiFi-GAN Generator alone at batch 1 x 512 frames: launch-latency territory — with and without graph replay."""
    import torch
    from tts_arabic_pytorch_b200.utils import synth
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
    voc = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    voc.load_state_dict(synth.hifigan_state_dict(1235))
    voc.remove_weight_norm()
    voc = voc.eval().to(c.dev)
    gen = torch.Generator().manual_seed(c.rank)
    b = args.batch or 1
    mel_host = torch.clamp(torch.randn(b, 80, 512, generator=gen) * 2 - 5, -11.5129, 2.0).pin_memory()
    mel_dev = mel_host.to(c.dev)
    wav_host = torch.empty(b, 512 * HOP, dtype=torch.float32).pin_memory()
    samples = b * 512 * HOP * c.world

    def step_device():
        return voc(mel_dev)

    graph = getattr(voc, 'capture_graph', None)
    replay = graph(mel_dev) if graph is not None else None

    def step_e2e():
        m = mel_host.to(c.dev, non_blocking=True)
        w = voc(m)
        wav_host.copy_(w.view(b, -1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(c.local)
    if c.rank == 0:
        sampler.start()
    ms_dev, launches = timed(c, step_device, args.steps)
    ms_graph = None
    if replay is not None:
        for _ in range(3):
            replay()
        ms_graph, _ = timed(c, replay, args.steps)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(c, step_e2e, args.steps)
    clocks = sampler.stop() if c.rank == 0 else None
    stage_ms = None if args.no_per_kernel else profile_stages(c, step_device, min(args.steps, 3))
    check_device_flag(c)
    if c.rank != 0:
        return
    best_ms = min(ms_dev, ms_graph) if ms_graph is not None else ms_dev
    peaks = c.peaks
    peak_tf = peaks.get('bf16_tflops', 1590.0)
    fl = VOCODER_FLOP_PER_FRAME * b * 512
    tf = fl * args.steps / (best_ms * 1e-3) / 1e12
    roof = {'kernel': 'HiFi-GAN generator launches (tcgen05 conv kernels), one utterance: 4 .. 2048 row tiles per launch over '
                      '148 SMs — occupancy- and launch-latency-bound, not a roofline regime', 'bound': 'tensor',
            'achieved': tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': tf / peak_tf,
            'peak_source': 'measured (MEASURED_PEAKS.json bf16_tflops, burst)' if peaks else 'fallback (B200_PROFILING.md)',
            'traffic': None}
    if stage_ms is not None:
        roof['per_kernel'] = per_kernel_table(stage_ms, vocoder_stage_model(b * 512), peaks, min(args.steps, 3))
    line = {
        'metric': 'audio_samples_per_sec', 'value': samples * args.steps / (best_ms * 1e-3), 'unit': 'samples/s',
        'n_gpus': c.world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': best_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp16', 'data': 'synthetic',
        'config': {'workload': 'HiFi-GAN Generator alone, mel [%d,80,512] -> %d samples per GPU (BASELINE config 2)' % (b, b * 512 * HOP),
                   'name': 'c2', 'global_batch': b * c.world, 'frames': 512, 'parallelism': 'replicas x%d' % c.world,
                   'l2': 'one utterance: every tensor fits the 126 MB L2 (weights + activations < 40 MB)'},
        'rtf': (best_ms * 1e-3 / args.steps) / (samples / SR),
        'latency_ms': {'stream_launches': ms_dev / args.steps, 'cuda_graph_replay': None if ms_graph is None else ms_graph / args.steps},
        'e2e': {'value': samples * args.steps / (ms_e2e * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': b * 80 * 512 * 4,
                'd2h_bytes_per_step': b * 512 * HOP * 4, 'ms_per_step': ms_e2e / args.steps, 'api': 'Generator.forward'},
        'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof,
    }
    print(json.dumps(line))


def run_c4(args, c):
    """Tacotron2Wave: autoregressive decode (fixed 256 steps: the synthetic gate stays shut) + HiFi-GAN, batch 8."""
    import torch
    from tts_arabic_pytorch_b200.models.tacotron2.tacotron2_ms import Tacotron2MS
    from tts_arabic_pytorch_b200.utils import synth
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
    steps_dec = 256
    b = args.batch or 8
    m = Tacotron2MS(n_symbol=40, decoder_max_step=steps_dec)
    m.load_state_dict(synth.tacotron2_state_dict(1236))
    m = m.eval().to(c.dev)
    voc = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    voc.load_state_dict(synth.hifigan_state_dict(1235))
    voc.remove_weight_norm()
    voc = voc.eval().to(c.dev)
    gen = torch.Generator().manual_seed(c.rank)
    tok_host = torch.randint(1, 40, (b, 64), generator=gen).pin_memory()
    tok_dev = tok_host.to(c.dev)
    lens = torch.full((b,), 64, dtype=torch.int32, device=c.dev)
    spk = torch.zeros(b, dtype=torch.int64, device=c.dev)
    wav_host = torch.empty(b, steps_dec * HOP, dtype=torch.float32).pin_memory()
    samples = b * steps_dec * HOP * c.world
    dec_ev = []
    import warnings
    warnings.simplefilter('ignore')

    def step_device():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mel, mel_lens, _, mel_cl = m.infer(tok_dev, spk, lens, return_channel_last=True)
        e1.record()
        dec_ev.append((e0, e1))
        return voc.run(mel_cl=mel_cl, lens=mel_lens)

    def step_e2e():
        t = tok_host.to(c.dev, non_blocking=True)
        mel, mel_lens, _, mel_cl = m.infer(t, spk, lens, return_channel_last=True)
        w = voc.run(mel_cl=mel_cl, lens=mel_lens)
        wav_host.copy_(w, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    dec_ev.clear()
    sampler = ClockSampler(c.local)
    if c.rank == 0:
        sampler.start()
    ms_dev, launches = timed(c, step_device, args.steps)
    t2_ms = sum(a.elapsed_time(b_) for a, b_ in dec_ev)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(c, step_e2e, args.steps)
    clocks = sampler.stop() if c.rank == 0 else None
    check_device_flag(c)
    if c.rank != 0:
        return
    peaks = c.peaks
    peak_hbm = peaks.get('hbm_gbs', 6650.0)
    us_per_step = t2_ms * 1e3 / (args.steps * steps_dec)
    w_bytes = 37.7e6                          # fp16 decoder weights touched per step
    gbs = w_bytes / (us_per_step * 1e-6) / 1e9
    line = {
        'metric': 'audio_samples_per_sec', 'value': samples * args.steps / (ms_dev * 1e-3), 'unit': 'samples/s',
        'n_gpus': c.world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp16 weights / fp32 state', 'data': 'synthetic',
        'config': {'workload': 'Tacotron2Wave: encoder + %d autoregressive decoder steps + postnet + HiFi-GAN, batch %d x 64 tokens '
                               'per GPU (BASELINE config 4)' % (steps_dec, b), 'name': 'c4', 'global_batch': b * c.world,
                   'parallelism': 'replicas x%d' % c.world, 'decoder_steps': steps_dec,
                   'l2': 'the 37.7 MB of decoder weights stay in the 126 MB L2 between steps'},
        'rtf': (ms_dev * 1e-3 / args.steps) / (samples / SR),
        'decoder_us_per_step': us_per_step,
        'e2e': {'value': samples * args.steps / (ms_e2e * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': b * 64 * 8,
                'd2h_bytes_per_step': b * steps_dec * HOP * 4, 'ms_per_step': ms_e2e / args.steps, 'api': 'Tacotron2MS.infer + Generator.run'},
        'gpu_launches': int(launches), 'clocks': clocks,
        'roofline': {'kernel': 'Tacotron2 decoder step (LSTM cells + attention + projection; Tacotron2MS.infer encode..finish)',
                     'bound': 'hbm', 'achieved': gbs, 'peak': peak_hbm, 'unit': 'GB/s', 'frac': gbs / peak_hbm,
                     'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if peaks else 'fallback',
                     'traffic': None, 'bytes_per_decoder_step': w_bytes,
                     'note': 'weight-bandwidth floor 37.7 MB per step (SURVEY.md 8d); the weights are L2-resident, so this is a '
                             'latency figure expressed against the HBM roof'},
    }
    print(json.dumps(line))


def run_ours(args):
    c = setup(args)
    if args.config == 'c2':
        run_c2(args, c)
    elif args.config == 'c4':
        run_c4(args, c)
    else:
        run_fastpitch2wave(args, c, args.config)
    if c.world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
