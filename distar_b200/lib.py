"""ctypes binding of libdistar_b200.so — the only way the Python host code reaches the GPU kernels.

There is deliberately NO fallback: if the shared library is missing or a call fails this module raises.
(The reference has no FFI of its own; this file is the stub INTEGRATION.md shows a maintainer.)
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdistar_b200.so')

_c = ctypes
_vp, _i, _i64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float

_SIGNATURES = {
    'dsb_last_error': (_c.c_char_p, []),
    'dsb_version': (_i, []),
    'dsb_launch_count': (_i64, []),
    'dsb_scatter_connection_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'dsb_scatter_connection_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'dsb_return_scan': (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'dsb_categorical_stats_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    'dsb_categorical_stats_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    'dsb_sample_categorical': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    'dsb_split_bf16': (_i, [_vp, _vp, _vp, _i64, _vp]),
    'dsb_upsample_bilinear2x_fwd': (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    'dsb_upsample_bilinear2x_bwd': (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    'dsb_gemm_bf16_split': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    'dsb_sumsq_partials': (_i, []),
    'dsb_grad_norm': (_i, [_vp, _i64, _vp, _vp, _vp]),
    'dsb_adam_step': (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _f, _f, _i, _vp, _vp, _vp]),
}

EXPORTS = sorted(_SIGNATURES)
_lib = None


class DsbError(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed to load it) and declare every prototype."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DsbError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(there is no CPU fallback)' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def launch_count() -> int:
    return int(load().dsb_launch_count())


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'distar_b200 kernels need contiguous CUDA tensors'
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    lib = load()
    conv = [(_ptr(a) if isinstance(a, torch.Tensor) else a) for a in args]
    rc = getattr(lib, name)(*conv, _stream())
    if rc != 0:
        raise DsbError('%s failed (%d): %s' % (name, rc, lib.dsb_last_error().decode()))
