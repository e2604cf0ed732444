"""Parity of the fused ResBlock1 step (csrc/conv_pair.cu, one launch for
x + conv2(lrelu(conv1(lrelu(x))))  — vocoder/hifigan/models.py:46-53) through the C ABI (ttsb_convpair_*) against
plain torch fp32 on the same fp16-rounded operands, for every (channels, kernel, dilation) the generator's
C <= 64 stages use, on ragged batches with tile-boundary and empty-utterance edge cases."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# fp16 storage of x, of the intermediate and of the output: a few ulp(fp16) of O(1..4) values
PAIR_ABS_TOL = 4e-3


def _dev():
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device (they never fall back to the CPU)')
    return torch.device('cuda:0')


def _reference(x, lens, w1, b1, w2, b2, k, dil, slope):
    B, T, C = x.shape
    xr = x.float().transpose(1, 2)
    mask = (torch.arange(T, device=x.device)[None, :] < lens[:, None])[:, None, :]
    lx = torch.where(xr > 0, xr, xr * slope).half().float()
    t = F.conv1d(lx, w1, b1, padding=(k - 1) // 2 * dil, dilation=dil)
    t = (torch.where(t > 0, t, t * slope) * mask).half().float()
    y = (F.conv1d(t, w2, b2, padding=(k - 1) // 2) + xr) * mask
    return y.transpose(1, 2)


def _make(C, k, dil, seed):
    from tts_arabic_pytorch_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    w1 = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).half().float()
    w2 = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).half().float()
    b1 = torch.randn(C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    h = ctypes.c_void_p()
    _lib.check(lib.ttsb_convpair_create(C, k, dil, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), 0,
                                        ctypes.byref(h)))
    plan = (ctypes.c_int * 8)()
    _lib.check(lib.ttsb_convpair_plan(h, plan))
    return lib, h, list(plan), (w1, b1, w2, b2), g


def _run_case(C, k, dil, T, lens, seed=0, act=False):
    """act=True: the generator's production form — tensors stored ACTIVATED (lrelu(x) in, lrelu(out) out), the
    residual add inverts the activation (ttsb_convpair_forward_act)."""
    from tts_arabic_pytorch_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    dev = _dev()
    lib, h, plan, (w1, b1, w2, b2), g = _make(C, k, dil, seed)
    try:
        assert plan[0] == 1, 'no fused plan for C=%d k=%d d=%d' % (C, k, dil)
        B = len(lens)
        x = torch.randn(B, T, C, generator=g).half()
        for b, n in enumerate(lens):
            x[b, n:] = 0
        xd = x.to(dev)
        ld = torch.tensor(lens, dtype=torch.int32, device=dev)
        out = torch.full((B, T, C), float('nan'), dtype=torch.float16, device=dev)
        if act:
            # operand of the kernel: fp16(lrelu(x)); its exact inverse is the x the reference sees
            lxd = torch.where(xd > 0, xd.float(), xd.float() * 0.1).half()
            xd = torch.where(lxd > 0, lxd.float(), lxd.float() * 10.0)        # fp32, exactly what the epilogue recovers
            _lib.check(lib.ttsb_convpair_forward_act(h, _lib.ptr(lxd), B, T, _lib.ptr(ld), 0.1, _lib.ptr(out), None))
        else:
            _lib.check(lib.ttsb_convpair_forward(h, _lib.ptr(xd), B, T, _lib.ptr(ld), 0.1, _lib.ptr(out), None))
        torch.cuda.synchronize()
        ref = _reference(xd, ld, w1.to(dev), b1.to(dev), w2.to(dev), b2.to(dev), k, dil, 0.1)
        if act:
            ref = torch.where(ref > 0, ref, ref * 0.1)
        o = out.float()
        assert not torch.isnan(o).any()
        scale = max(1.0, float(ref.abs().max()))
        assert float((o - ref).abs().max()) <= PAIR_ABS_TOL * scale
        for b, n in enumerate(lens):      # rows beyond an utterance stay exactly zero (per-utterance padding)
            assert float(o[b, n:].abs().max()) == 0.0 if n < T else True
        flag = ctypes.c_int(0)
        _lib.check(lib.ttsb_device_error_flag(ctypes.byref(flag)))
        assert flag.value == 0
    finally:
        lib.ttsb_convpair_destroy(h)


@pytest.mark.parametrize('C', [32, 64])
@pytest.mark.parametrize('k', [3, 7, 11])
@pytest.mark.parametrize('dil', [1, 3, 5])
def test_pair_matches_torch_on_ragged_batch(C, k, dil):
    # several tiles per utterance, a ragged tail tile, one utterance shorter than a tile; 3*ceil(4000/118) work
    # items < 148 CTAs would never wrap a persistent CTA, so the second utterance count makes them wrap
    _run_case(C, k, dil, 4000, [4000, 3629, 129, 4000, 2500, 3999])


@pytest.mark.parametrize('C', [32, 64])
@pytest.mark.parametrize('k', [3, 7, 11])
@pytest.mark.parametrize('dil', [1, 3, 5])
def test_pair_on_activated_tensors_matches_torch(C, k, dil):
    _run_case(C, k, dil, 4000, [4000, 3629, 129, 4000, 2500, 3999], seed=1, act=True)


@pytest.mark.parametrize('C,k,dil', [(32, 11, 5), (64, 3, 1), (64, 11, 5)])
def test_pair_edge_shapes(C, k, dil):
    m_out = 128 - (k - 1)
    _run_case(C, k, dil, 5, [5, 0, 3])                         # shorter than every halo; an empty utterance
    _run_case(C, k, dil, m_out, [m_out, m_out - 1])            # exactly one tile
    _run_case(C, k, dil, m_out + 1, [m_out + 1, 1])            # one row spills into a second tile
    _run_case(C, k, dil, 40 * m_out, [40 * m_out] * 8 + [17])  # persistent CTAs wrap (321 items over 148 CTAs)
    _run_case(C, k, dil, m_out + 1, [m_out + 1, 1], act=True)
    _run_case(C, k, dil, 40 * m_out, [40 * m_out] * 8 + [17], act=True)


def test_pair_plan_reports_fallback_for_wide_layers():
    """C = 128 has no fused plan (4*C TMEM columns fit, the shared-memory budget does not): the generator must
    then take the two-launch path instead of failing."""
    lib, h, plan, _, _ = _make(128, 3, 1, 0)
    try:
        assert plan[0] == 0
    finally:
        lib.ttsb_convpair_destroy(h)
