#!/usr/bin/env python
"""Headline benchmark: FastPitch + HiFi-GAN end-to-end audio samples/s on synthetic fixed-length
phoneme batches (BASELINE.json: batch 256 x 128 phonemes per GPU -> 512 frames -> 131072 samples
per utterance), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 256] [--phonemes 128]
  python bench.py --impl reference ...      # the reference's CPU PyTorch arithmetic (oracle port)

Prints ONE JSON line on rank 0 (contract in the task statement): `value` = device-resident
throughput (ids on the GPU -> waveforms on the GPU), `e2e` = the same work through the public API
with pinned-host ids in and pinned-host waveforms out inside the timed region, `roofline` for the
dominant kernel (conv_tc_kernel: all HiFi-GAN launches of a step, tensor-bound), `cpu_baseline`
(oracle on this box's host cores, bounded sample).
"""
import argparse
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
FRAMES_PER_PHONEME = 4                      # const-4 duration head (SURVEY.md §8c calibration)
VOCODER_FLOP_PER_FRAME = 614.1e6
# dram__bytes_read.sum + dram__bytes_write.sum over the generator launches of one 32768-frame chunk, from the
# `ncu --set full` capture profiles/r01_s24_ncu_full_b64.csv (70.64 GB / 32768 frames); bench.py cannot read DRAM
# counters itself, so `roofline.traffic` = this per-frame figure x the frames of a step
VOCODER_DRAM_BYTES_PER_FRAME = 70.64e9 / 32768
FASTPITCH_FLOP_PER_UTT_128 = 28.8e9         # SURVEY.md §8d (L=128 -> T=512)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=5)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--batch', type=int, default=256, help='utterances per GPU per step')
    p.add_argument('--phonemes', type=int, default=128)
    p.add_argument('--cpu-sample', type=int, default=0, help='utterances per CPU-baseline step (0 = auto)')
    p.add_argument('--no-cpu-baseline', action='store_true')
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


def cpu_reference_step(fsd, gsd_folded, ids):
    """The reference's CPU path, restated: FastPitch.infer on the padded batch, then the generator once
    per utterance (models/fastpitch/networks.py:322-350). Returns number of audio samples."""
    import torch
    from oracle import fastpitch_oracle as fpo
    from oracle import hifigan_oracle as hgo
    from tts_arabic_pytorch_b200.utils import synth
    with torch.no_grad():
        mel, dec_lens, *_ = fpo.fastpitch_infer(fsd, synth.FASTPITCH_CONFIG, ids)
        wavs = hgo.vocode_batch(gsd_folded, synth.HIFIGAN_CONFIG, mel, dec_lens)
    return sum(int(w.numel()) for w in wavs)


def run_reference(args):
    import torch
    from tts_arabic_pytorch_b200.utils import synth
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = min(usable_cores(), 32)
    torch.set_num_threads(cores)
    fsd = synth.fastpitch_state_dict(1234)
    gsd = synth.fold_weight_norm(synth.hifigan_state_dict(1235))
    gen = torch.Generator().manual_seed(0)
    bs = args.cpu_sample or 4
    ids = torch.randint(1, 40, (bs, args.phonemes), generator=gen)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_step(fsd, gsd, ids[:1])
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        n += cpu_reference_step(fsd, gsd, ids)
    dt = time.perf_counter() - t0
    v = n / dt
    line = {
        'impl': 'reference', 'metric': 'audio_samples_per_sec', 'value': v, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': 'FastPitch2Wave end-to-end, %d-phoneme synthetic utterances (4 frames/phoneme), '
                               'bounded sample of %d utterances per step on the host CPU' % (args.phonemes, bs),
                   'phonemes': args.phonemes, 'sample_batch': bs},
        'rtf': (dt / (n / 22050.0)),
        'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d steps x %d utterances x %d phonemes' % (args.steps, bs, args.phonemes)},
        'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tts_arabic_pytorch_b200 import _lib
    from tts_arabic_pytorch_b200.models.fastpitch.fastpitch.model import FastPitch
    from tts_arabic_pytorch_b200.utils import synth
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (impl=ours) needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    fsd = synth.fastpitch_state_dict(1234)
    gsd = synth.hifigan_state_dict(1235)
    fp = FastPitch(**synth.FASTPITCH_CONFIG)
    fp.load_state_dict(fsd)
    fp = fp.eval().to(dev)
    voc = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    voc.load_state_dict(gsd)
    voc.remove_weight_norm()
    voc = voc.eval().to(dev)

    B, L = args.batch, args.phonemes
    T = L * FRAMES_PER_PHONEME
    gen = torch.Generator().manual_seed(1000 + rank)
    ids_host = torch.randint(1, 40, (B, L), generator=gen).pin_memory()
    ids_dev = ids_host.to(dev)
    wav_host = torch.empty(B, T * HOP, dtype=torch.float32).pin_memory()
    samples_per_step = B * T * HOP

    voc_ms = []
    # N > 1: the only collective of the path — deliver every rank's waveforms to rank 0 over NCCL
    gather_bufs = None
    if world > 1 and rank == 0:
        gather_bufs = [torch.empty(B, T * HOP, dtype=torch.float32, device=dev) for _ in range(world)]

    def deliver(wav):
        if world > 1:
            dist.gather(wav, gather_bufs, dst=0)
        return wav

    def step_device():
        mel, dec_lens, _, _, _, mel_cl = fp.infer(ids_dev, return_channel_last=True)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        wav = voc.run(mel_cl=mel_cl, lens=dec_lens)
        e1.record()
        voc_ms.append((e0, e1))
        return deliver(wav)

    def step_e2e():
        ids = ids_host.to(dev, non_blocking=True)
        mel, dec_lens, _, _, _, mel_cl = fp.infer(ids, return_channel_last=True)
        wav = deliver(voc.run(mel_cl=mel_cl, lens=dec_lens))
        wav_host.copy_(wav, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return wav

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        n0 = lib.ttsb_launch_count()
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        launches = lib.ttsb_launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, launches

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    voc_ms.clear()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(step_device, args.steps)
    voc_total_ms = sum(a.elapsed_time(b) for a, b in voc_ms)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    flag = _lib.ctypes.c_int(0)
    _lib.check(lib.ttsb_device_error_flag(_lib.ctypes.byref(flag)))
    if flag.value != 0:
        raise RuntimeError('device error flag %d' % flag.value)
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    total_samples = samples_per_step * world
    value = total_samples * args.steps / (ms_dev * 1e-3)
    e2e_value = total_samples * args.steps / (ms_e2e * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback (B200_PROFILING.md sustained)'
    voc_flops = VOCODER_FLOP_PER_FRAME * B * T * args.steps
    achieved_tf = voc_flops / (voc_total_ms * 1e-3) / 1e12 if voc_total_ms > 0 else 0.0
    voc_traffic = VOCODER_DRAM_BYTES_PER_FRAME * B * T            # bytes per step (rank 0)
    peak_hbm = peaks.get('hbm_gbs', 6650.0)
    hbm_gbs = voc_traffic * args.steps / (voc_total_ms * 1e-3) / 1e9 if voc_total_ms > 0 else 0.0
    line = {
        'metric': 'audio_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'fp16', 'data': 'synthetic',
        'config': {'workload': 'FastPitch2Wave end-to-end, batch %d x %d phonemes per GPU (4 frames/phoneme -> %d '
                               'frames -> %d samples per utterance), HiFi-GAN V1' % (B, L, T, T * HOP),
                   'global_batch': B * world, 'phonemes': L, 'frames': T, 'parallelism': 'dp%d' % world,
                   'l2': 'per-step activations (GBs) exceed the 126 MB L2; only the 28 MB of packed weights stay resident',
                   'conv_impl': 'tcgen05' if lib.ttsb_get_conv_impl() == 0 else 'simt',
                   'desc_mode': lib.ttsb_get_desc_mode()},
        'rtf': (ms_dev * 1e-3 / args.steps) / (total_samples / 22050.0),
        'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': B * L * 8,
                'd2h_bytes_per_step': B * T * HOP * 4, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'kernel': 'conv_tc2_kernel + conv_pair_kernel (tcgen05 row-GEMM-with-taps; all HiFi-GAN generator '
                               'launches of a step, rank 0)', 'bound': 'tensor',
                     'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf,
                     'peak_source': peak_src, 'traffic': voc_traffic,
                     'traffic_source': 'profiles/r01_s24_ncu_full_b64.csv (ncu --set full, dram read+write, per frame x frames)',
                     'flops_per_step': voc_flops / args.steps, 'ms_per_step': voc_total_ms / args.steps,
                     'share_of_step': voc_total_ms / ms_dev if ms_dev > 0 else None,
                     'hbm': {'achieved': hbm_gbs, 'peak': peak_hbm, 'unit': 'GB/s', 'frac': hbm_gbs / peak_hbm}},
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = min(usable_cores(), 32)     # torch CPU convs at batch 1 stop scaling well before this
        torch.set_num_threads(cores)
        gsd_f = synth.fold_weight_norm(gsd)
        t0 = time.perf_counter()
        cpu_reference_step(fsd, gsd_f, ids_host[:1].clone())
        per_utt = time.perf_counter() - t0
        bs = args.cpu_sample or max(1, min(4, int(6.0 / max(per_utt, 1e-3))))
        ids = ids_host[:bs].clone()
        t0 = time.perf_counter()
        n, reps = 0, 0
        while time.perf_counter() - t0 < 12.0 and reps < 16:
            n += cpu_reference_step(fsd, gsd_f, ids)
            reps += 1
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': n / dt, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                                'sample': '%d x %d utterances x %d phonemes, fp32 torch CPU' % (reps, bs, L)}
    print(json.dumps(line))


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
