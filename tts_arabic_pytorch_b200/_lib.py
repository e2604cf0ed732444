"""ctypes binding of the C-ABI library (include/ttsb200.h).

This is the whole "PyTorch extension": torch only provides device memory (`tensor.data_ptr()`)
and the current stream; every kernel lives in libttsb200.so. There is no CPU fallback — if the
library cannot be loaded, importing callers get a RuntimeError.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# TTSB_LIB: an alternative build of the same C ABI (tools/build_variant.py builds one from any git ref) for same-box A/B
# measurements — boxes differ by a few per cent, so two kernels are only comparable inside one GPU session
LIB_PATH = os.environ.get('TTSB_LIB') or os.path.join(_HERE, 'libttsb200.so')
_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_float, c_size_t, c_int64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_int64


class TensorDesc(ctypes.Structure):
    _fields_ = [('name', ctypes.c_char_p), ('h_data', ctypes.POINTER(ctypes.c_float)), ('ndim', c_int),
                ('shape', c_int64 * 4)]


class HifiganConfig(ctypes.Structure):
    _fields_ = [('num_mels', c_int), ('upsample_initial_channel', c_int), ('num_upsamples', c_int),
                ('upsample_rates', c_int * 8), ('upsample_kernel_sizes', c_int * 8), ('num_kernels', c_int),
                ('resblock_kernel_sizes', c_int * 8), ('resblock_dilations', (c_int * 3) * 8)]


class FastpitchConfig(ctypes.Structure):
    _fields_ = [('n_mel_channels', c_int), ('n_symbols', c_int), ('d_model', c_int), ('n_layers_enc', c_int),
                ('n_layers_dec', c_int), ('d_head', c_int), ('d_inner', c_int), ('conv_kernel', c_int),
                ('pred_filter', c_int), ('pred_kernel', c_int), ('energy_conditioning', c_int),
                ('n_speakers', c_int), ('speaker_emb_weight', c_float)]


_SIGNATURES = {
    'ttsb_last_error': (ctypes.c_char_p, []),
    'ttsb_version': (c_int, []),
    'ttsb_set_conv_impl': (c_int, [c_int]),
    'ttsb_set_desc_mode': (c_int, [c_int]),
    'ttsb_set_tc_version': (c_int, [c_int]),
    'ttsb_get_conv_impl': (c_int, []),
    'ttsb_get_desc_mode': (c_int, []),
    'ttsb_launch_count': (c_int64, []),
    'ttsb_device_error_flag': (c_int, [ctypes.POINTER(c_int)]),
    'ttsb_debug_set_timeline': (c_int, [c_void_p]),
    'ttsb_prof_enable': (c_int, [c_int]),
    'ttsb_prof_n_tags': (c_int, []),
    'ttsb_prof_collect': (c_int, [ctypes.POINTER(ctypes.c_double), c_int]),
    'ttsb_hifigan_create': (c_int, [ctypes.POINTER(HifiganConfig), ctypes.POINTER(TensorDesc), c_int, c_int,
                                    ctypes.POINTER(c_void_p)]),
    'ttsb_hifigan_destroy': (None, [c_void_p]),
    'ttsb_hifigan_hop': (c_int, [c_void_p]),
    'ttsb_hifigan_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int]),
    'ttsb_hifigan_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    'ttsb_fastpitch_create': (c_int, [ctypes.POINTER(FastpitchConfig), ctypes.POINTER(TensorDesc), c_int, c_int,
                                      ctypes.POINTER(c_void_p)]),
    'ttsb_fastpitch_destroy': (None, [c_void_p]),
    'ttsb_fastpitch_state_bytes': (c_size_t, [c_void_p, c_int, c_int]),
    'ttsb_fastpitch_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int]),
    'ttsb_fastpitch_encode': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    'ttsb_fastpitch_read_enc_out': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'ttsb_fastpitch_condition': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                         c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                         c_void_p]),
    'ttsb_fastpitch_decode': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    'ttsb_tacotron2_create': (c_int, [ctypes.POINTER(TensorDesc), c_int, c_int, ctypes.POINTER(c_void_p)]),
    'ttsb_tacotron2_destroy': (None, [c_void_p]),
    'ttsb_tacotron2_state_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int]),
    'ttsb_tacotron2_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int]),
    'ttsb_tacotron2_encode': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    'ttsb_tacotron2_decode': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_float, c_int, c_void_p,
                                      ctypes.POINTER(c_int), c_void_p]),
    'ttsb_tacotron2_finish': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    'ttsb_tacotron2_postprocess': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_double, c_int, c_int, c_int, c_int, c_int,
                                           c_void_p, c_void_p, c_void_p]),
    'ttsb_conv1d_create': (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                   ctypes.POINTER(c_void_p)]),
    'ttsb_conv1d_destroy': (None, [c_void_p]),
    'ttsb_conv1d_cin_pad': (c_int, [c_void_p]),
    'ttsb_conv1d_forward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p,
                                    c_void_p]),
    'ttsb_denoiser_workspace_bytes': (c_size_t, [c_int, c_int]),
    'ttsb_denoiser_forward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    'ttsb_convpair_create': (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     ctypes.POINTER(c_void_p)]),
    'ttsb_convpair_destroy': (None, [c_void_p]),
    'ttsb_convpair_plan': (c_int, [c_void_p, ctypes.POINTER(c_int)]),
    'ttsb_convpair_forward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p]),
    'ttsb_convpair_forward_act': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p]),
}
EXPORTS = tuple(sorted(_SIGNATURES))


def load(build_if_missing=True):
    """Returns the loaded ctypes library; builds it with nvcc if it is absent and nvcc exists."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError('libttsb200.so is missing (%s); run `python -m tts_arabic_pytorch_b200.build`' % LIB_PATH)
            from . import build as _build
            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError here == header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(status):
    if status != 0:
        msg = load().ttsb_last_error()
        raise RuntimeError('ttsb200 error %d: %s' % (status, msg.decode('utf-8', 'replace') if msg else '?'))


def tensor_table(named):
    """dict name -> CPU fp32 contiguous torch tensor  =>  (ctypes array, keepalive list)."""
    import torch
    keep = []
    arr = (TensorDesc * len(named))()
    for i, (name, t) in enumerate(named.items()):
        t = t.detach().to(device='cpu', dtype=torch.float32).contiguous()
        keep.append(t)
        nb = name.encode()
        keep.append(nb)
        arr[i].name = nb
        arr[i].h_data = ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))
        arr[i].ndim = t.dim()
        for d in range(t.dim()):
            arr[i].shape[d] = t.shape[d]
    return arr, keep


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Workspace:
    """Grow-only device byte buffer (the C ABI never allocates on the hot path)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        import torch
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf
