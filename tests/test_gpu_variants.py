"""Agreement between the kernel variants behind the runtime switches (each read once per process, so every variant
runs in its own subprocess): the fused ResBlock step against the two-launch path, TMA-store epilogues against the
LDS + STG path, the mma.sync attention against the fp32-FMA one, CTA-pair weight multicast on / off."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests import tolerances as tol

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HELPER = os.path.join(REPO, 'tests', 'helpers', 'run_variant.py')


def _run(tmp_path, name, env_extra):
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device')
    out = os.path.join(str(tmp_path), name + '.npz')
    env = dict(os.environ)
    for k in ('TTSB_PAIR', 'TTSB_TMA_OUT', 'TTSB_ATTENTION', 'TTSB_CLUSTER', 'TTSB_PAIR_SMEM_RES', 'TTSB_ACT_CHAIN', 'TTSB_EPI_ACT',
              'TTSB_EPI_TMA_IN', 'TTSB_EPI_RING', 'TTSB_PAIR_TT2', 'TTSB_PAIR_W2X2', 'TTSB_PAIR_ACT_ONLY', 'TTSB_OCC2'):
        env.pop(k, None)
    env.update(env_extra)
    r = subprocess.run([sys.executable, HELPER, out], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


def _rel_rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / (np.sqrt(np.mean(b ** 2)) + 1e-12))


@pytest.fixture(scope='module')
def default_run(tmp_path_factory):
    return _run(tmp_path_factory.mktemp('variants'), 'default', {})


def test_fused_resblock_step_agrees_with_two_launch_path(default_run, tmp_path):
    other = _run(tmp_path, 'nopair', {'TTSB_PAIR': '0'})
    assert other['dec_lens'].tolist() == default_run['dec_lens'].tolist()
    # same fp16 rounding points except lrelu(fp16(x)) vs fp16(lrelu(x)) on the pair inputs
    assert _rel_rms(default_run['wav'], other['wav']) < tol.WAV_REL_RMS
    assert np.abs(default_run['wav'] - other['wav']).max() < tol.WAV_LINF


def test_tma_store_epilogue_is_bit_exact(default_run, tmp_path):
    other = _run(tmp_path, 'notma', {'TTSB_TMA_OUT': '0'})
    assert np.array_equal(default_run['mel'], other['mel'])
    assert np.array_equal(default_run['wav'], other['wav'])


def test_attention_kernels_agree(default_run, tmp_path):
    """default = tcgen05 (attention_tc.cu: both contractions on the 5th-generation tensor cores) against the two older
    generations kept as cross-checks: mma.sync m16n8k16 and plain fp32 FMA."""
    for name in ('mma', 'simt'):
        other = _run(tmp_path, name, {'TTSB_ATTENTION': name})
        assert other['dec_lens'].tolist() == default_run['dec_lens'].tolist()
        assert np.abs(default_run['mel'] - other['mel']).max() < tol.MEL_LINF


def test_weight_multicast_and_transform_placement_do_not_change_results(default_run, tmp_path):
    a = _run(tmp_path, 'nocluster', {'TTSB_CLUSTER': '1'})
    assert np.array_equal(default_run['wav'], a['wav'])
    b = _run(tmp_path, 'globalres', {'TTSB_PAIR_SMEM_RES': '0'})      # residual from global memory instead of the x panel
    assert np.array_equal(default_run['wav'], b['wav'])
    c = _run(tmp_path, 'rawchain', {'TTSB_ACT_CHAIN': '0'})           # round-1 data flow: raw + activated copies
    assert _rel_rms(default_run['wav'], c['wav']) < tol.WAV_REL_RMS


def test_round2_epilogue_and_pair_forms_are_bit_exact(default_run, tmp_path):
    """The round-2 kernel forms change schedules and data movement, not arithmetic: the act-only and TMA-in epilogues of
    conv_tc2 against the round-1 lean epilogue, a single look-ahead slot against two, conv_pair's two-taps-per-row conv2
    (same K = 16 steps in the same order), two items per W2 ring pass and the act-only final epilogue — every one must
    reproduce the default run's mel and waveform exactly."""
    for name, env in (('lean_epilogues', {'TTSB_EPI_ACT': '0', 'TTSB_EPI_TMA_IN': '0'}),
                      ('ring1', {'TTSB_EPI_RING': '1'}),
                      ('tt64', {'TTSB_PAIR_TT2': '0'}),
                      ('tt128_all_k', {'TTSB_PAIR_TT2': '2'}),
                      ('w2_one_item', {'TTSB_PAIR_W2X2': '0'}),
                      ('pair_lean_final', {'TTSB_PAIR_ACT_ONLY': '0'})):
        other = _run(tmp_path, name, env)
        assert np.array_equal(default_run['mel'], other['mel']), name
        assert np.array_equal(default_run['wav'], other['wav']), name
