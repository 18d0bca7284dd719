"""ctypes binding of libseldfeat.so -- the only route from Python to the CUDA kernels.

The signatures mirror include/seldfeat.h one to one.  There is deliberately no fallback: if the
library is missing or a call fails, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SELD_LIB', os.path.join(_HERE, 'libseldfeat.so'))   # SELD_LIB: developer A/B builds

SELD_OK = 0
SELD_EINVAL, SELD_EUNSUPPORTED, SELD_ESHORT, SELD_ECUDA, SELD_ENOMEM = -1, -2, -3, -4, -5

_c_float_p = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.c_int64

# name -> (restype, argtypes); must list every symbol include/seldfeat.h declares
SIGNATURES = {
    'seld_plan_create': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _c_float_p, _c_float_p,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]),
    'seld_plan_destroy': (None, [ctypes.c_void_p]),
    'seld_num_frames': (_i64, [ctypes.c_void_p, _i64]),
    'seld_logmel_iv_f32': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i64, ctypes.c_int, _i64, _i64, _i64,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    'seld_logmel_iv_i16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i64, ctypes.c_int, _i64, _i64, _i64,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    'seld_logmel_iv_f32_host': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i64, ctypes.c_int, _i64,
                                               ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    'seld_logmel_iv_i16_host': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i64, ctypes.c_int, _i64,
                                               ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    'seld_logmel_f32': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i64, ctypes.c_int, _i64, _i64, _i64,
                                       ctypes.c_void_p, ctypes.c_void_p]),
    'seld_num_frames_mic': (_i64, [ctypes.c_void_p, _i64]),
    'seld_workspace_bytes': (ctypes.c_size_t, [ctypes.c_void_p, _i64, ctypes.c_int]),
    'seld_logmel_gcc_f32': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i64, ctypes.c_int, _i64, _i64, _i64,
                                           ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.c_void_p]),
    'seld_mic_spectrogram_f32': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i64, ctypes.c_int, _i64, _i64, _i64,
                                                ctypes.c_void_p, ctypes.c_void_p]),
    'seld_logmel_gcc_from_spectra_f32': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i64, ctypes.c_int, _i64,
                                                        ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                        ctypes.c_void_p]),
    'seld_scalar_f32': (ctypes.c_int, [ctypes.c_void_p, _i64, ctypes.c_int, _i64, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                                       ctypes.c_void_p]),
    'seld_scalar_wav2img_f32': (ctypes.c_int, [ctypes.c_void_p, _i64, ctypes.c_int, _i64, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    'seld_foa_rotate_f32': (ctypes.c_int, [ctypes.c_void_p, _i64, ctypes.c_int, _i64, _i64, _i64, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    'seld_wavmix_order': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _i64,
                                         ctypes.c_void_p]),
    'seld_wavmix_f32': (ctypes.c_int, [ctypes.c_void_p, _i64, ctypes.c_int, _i64, _i64, _i64, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_void_p]),
    'seld_launch_count': (ctypes.c_uint64, []),
    'seld_last_cuda_error': (ctypes.c_int, []),
    'seld_strerror': (ctypes.c_char_p, [ctypes.c_int]),
    'seld_version': (ctypes.c_char_p, []),
}

_lib = None


class SeldError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().seld_strerror(code).decode()
        if code == SELD_ECUDA:
            msg += ' (cudaError %d)' % lib().seld_last_cuda_error()
        super().__init__('%s: %s [%d]' % (where, msg, code))


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libseldfeat.so is missing (%s). Build it with `python -m pseldnets_b200.build`; '
                'there is no CPU or PyTorch fallback for this path.' % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code, where):
    if code != SELD_OK:
        raise SeldError(code, where)


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """Context that makes `device` current for the launch -- free when it already is (the one-process-per-GPU case)."""
    import torch
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)
