"""The bench contract that can be checked without a GPU: the reference arm (`bench.py --impl reference`, the
reference's CPU arithmetic through the oracle) prints ONE JSON line with the keys the driver reads, under torchrun
only rank 0 prints, and the default arm fails loudly when there is no CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(REPO, 'bench.py')


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=env,
                          cwd=REPO)


def test_reference_arm_prints_one_contract_line():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-sample', '1', '--phonemes', '16'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'audio_samples_per_sec' and d['unit'] == 'samples/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['data'] == 'synthetic'
    assert d['value'] > 0 and d['steps'] == 1 and d['n_gpus'] == 1
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1   # 'reference' where oracle/_ref is built
    assert d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']
    # 16 phonemes x 4 frames x 256 samples per utterance, one utterance, one step
    assert abs(d['value'] * d['ms_per_step'] * 1e-3 - 16 * 4 * 256) < 1.0


def test_reference_arm_other_ranks_stay_silent():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-sample', '1', '--phonemes', '16', '--gpus', '2'],
             env_extra={'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_default_arm_has_no_cpu_fallback():
    r = _run(['--steps', '1', '--warmup', '1', '--batch', '1'])
    assert r.returncode != 0
    assert 'no CPU fallback' in r.stderr
