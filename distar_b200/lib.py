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
    'dsb_entity_features': (_i, [_vp] * 5 + [_i, _vp, _vp, _i, _i64, _vp, _vp]),
    'dsb_spatial_stem_fwd': (_i, [_vp] * 12 + [_i] * 5 + [_vp]),
    'dsb_spatial_stem_bwd': (_i, [_vp] * 10 + [_i] + [_vp] * 3 + [_i] * 4 + [_vp]),
    'dsb_return_scan': (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'dsb_categorical_stats_fwd': (_i, [_vp] * 10 + [_i64, _i, _vp]),
    'dsb_categorical_stats_bwd': (_i, [_vp] * 11 + [_i64, _i, _vp]),
    'dsb_su_sample_step': (_i, [_vp] * 17 + [_i, _i, _i, _f, _vp]),
    'dsb_sample_categorical': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    'dsb_split_bf16': (_i, [_vp, _vp, _vp, _i64, _vp]),
    'dsb_upshift9_fwd': (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp]),
    'dsb_upshift9_bwd': (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    'dsb_upsample_bilinear2x_fwd': (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    'dsb_upsample_bilinear2x_bwd': (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    'dsb_upsample_bilinear2x_nhwc_fwd': (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    'dsb_upsample_bilinear2x_nhwc_bwd': (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    'dsb_gemm_bf16_split': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    'dsb_attn_softmax_fwd': (_i, [_vp, _vp, _i, _vp, _vp, _i64, _i, _vp]),
    'dsb_attn_softmax_bwd': (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i64, _i, _vp]),
    'dsb_layernorm_supported': (_i, [_i]),
    'dsb_layernorm_fwd': (_i, [_vp] * 9 + [_i64, _i, _f, _vp]),
    'dsb_layernorm_bwd_blocks': (_i, [_i64]),
    'dsb_layernorm_bwd': (_i, [_vp] * 7 + [_i, _i64, _i, _vp]),
    'dsb_lstm_cell_fwd': (_i, [_vp] * 13 + [_i, _i, _f, _vp]),
    'dsb_lstm_cell_bwd': (_i, [_vp] * 18 + [_i, _i, _vp]),
    'dsb_upconv_fwd': (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    'dsb_upconv_bwd': (_i, [_vp, _i, _vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    'dsb_proj9_fwd': (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    'dsb_proj9_bwd': (_i, [_vp] * 5 + [_i64, _i, _vp]),
    'dsb_maxpool2_nhwc_fwd': (_i, [_vp] * 5 + [_i64, _i, _i, _i, _vp]),
    'dsb_maxpool2_nhwc_bwd': (_i, [_vp] * 3 + [_i64, _i, _i, _i, _vp]),
    'dsb_gate_update_fwd': (_i, [_vp] * 8 + [_i64, _vp]),
    'dsb_gate_update_bwd': (_i, [_vp] * 9 + [_i64, _vp]),
    'dsb_relu_bwd_split_blocks': (_i, [_i64, _i]),
    'dsb_relu_bwd_split': (_i, [_vp, _vp, _i] + [_vp] * 4 + [_i, _i64, _i, _vp]),
    'dsb_su_prefix_mean_fwd': (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp]),
    'dsb_su_prefix_mean_bwd': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _vp]),
    'dsb_su_lstm_fwd': (_i, [_vp] * 12 + [_i64, _i, _vp]),
    'dsb_su_lstm_bwd': (_i, [_vp] * 16 + [_i64, _i, _vp]),
    'dsb_su_logits_fwd': (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i64, _i, _i, _vp]),
    'dsb_su_logits_bwd': (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp]),
    'dsb_expand_ragged': (_i, [_vp] * 5 + [_i64, _i, _i, _i, _c.c_double, _i, _vp]),
    'dsb_sequence_mask': (_i, [_vp, _i, _vp, _i64, _i, _vp]),
    'dsb_unpack_planes': (_i, [_vp] * 7 + [_i64, _vp]),
    'dsb_lstm_seq_fwd': (_i, [_vp] * 15 + [_i, _i, _i, _f, _vp]),
    'dsb_lstm_seq_bwd': (_i, [_vp] * 21 + [_i, _i, _i, _vp]),
    'dsb_bo_tokens': (_i, [_vp, _vp, _i, _vp, _i64, _i, _i, _i, _vp]),
    'dsb_ln_small_supported': (_i, [_i]),
    'dsb_ln_small_fwd': (_i, [_vp] * 5 + [_i64, _i, _f, _vp]),
    'dsb_ln_small_bwd': (_i, [_vp] * 7 + [_i64, _i, _vp]),
    'dsb_attn_small_fwd': (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    'dsb_attn_small_bwd': (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp]),
    'dsb_pack_pair': (_i, [_vp, _i, _i64, _i, _i64, _vp, _vp, _i, _vp]),
    'dsb_glu_gate_fwd': (_i, [_vp, _vp, _vp, _i64, _vp]),
    'dsb_glu_gate_bwd': (_i, [_vp] * 5 + [_i64, _vp]),
    'dsb_onehot_linear_fwd': (_i, [_vp] * 4 + [_i64, _i, _i, _i64, _i64, _i, _i, _vp, _vp]),
    'dsb_onehot_linear_bwd': (_i, [_vp] * 5 + [_i64, _i, _i, _i64, _i64, _i, _vp]),
    'dsb_target_unit_fwd': (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i, _f, _vp]),
    'dsb_target_unit_bwd': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i64, _i, _f, _vp]),
    'dsb_colsum_pair': (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    'dsb_sumsq_partials': (_i, []),
    'dsb_grad_norm': (_i, [_vp, _i64, _vp, _vp, _f, _vp]),
    'dsb_adam_step': (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _f, _f, _f, _i, _vp, _vp, _vp, _vp]),
}



class GemmArgs(ctypes.Structure):
    """mirror of dsb_gemm_args (include/distar_b200.h)"""
    _fields_ = [('a_hi', _vp), ('a_lo', _vp), ('b_hi', _vp), ('b_lo', _vp),
                ('a_rows', _i64), ('a_cols', _i64), ('b_rows', _i64), ('b_cols', _i64),
                ('a_mn', _c.c_int32), ('b_mn', _c.c_int32),
                ('a_col_base', _c.c_int32), ('a_col_inner', _c.c_int32), ('a_row_outer', _c.c_int32),
                ('a_row_inner', _c.c_int32),
                ('b_col_base', _c.c_int32), ('b_col_inner', _c.c_int32), ('b_row_outer', _c.c_int32),
                ('b_row_inner', _c.c_int32),
                ('bias', _vp), ('alpha', _f), ('relu', _c.c_int32), ('terms', _c.c_int32),
                ('c', _vp), ('c_rows', _i64), ('c_cols', _i64), ('c_hi', _vp), ('c_lo', _vp),
                ('m', _i64), ('n', _c.c_int32), ('k', _c.c_int32),
                ('batch', _c.c_int32), ('inner', _c.c_int32), ('splits', _c.c_int32),
                ('c_row_outer', _c.c_int32), ('c_row_inner', _c.c_int32), ('c_row_split', _c.c_int32),
                ('c_col_base', _c.c_int32), ('c_col_inner', _c.c_int32),
                ('residual', _vp), ('bn', _c.c_int32),
                ('a_conv', _c.c_int32), ('b_conv', _c.c_int32), ('conv_h', _c.c_int32), ('conv_w', _c.c_int32),
                ('conv_c', _c.c_int32), ('conv_taps', _c.c_int32), ('conv_imgs', _i64), ('c_accumulate', _c.c_int32), ('mc', _c.c_int32),
                ('a_exact', _c.c_int32), ('b_exact', _c.c_int32), ('relu_mask', _vp), ('colsum', _vp)]


_SIGNATURES['dsb_gemm_ex'] = (_i, [ctypes.POINTER(GemmArgs), _vp])
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
    # raw handle of torch's current stream; torch.cuda.current_stream() builds a Stream object (~15 us per call, which
    # adds up over the ~4000 launches of a learner step)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


# Measurement hook (bench.py): when GEMM_TRACE is a list, every tcgen05 GEMM launch is bracketed by CUDA events on the launching
# stream and (event0, event1, algorithmic flop, MMA products per element) is appended to it.
GEMM_TRACE = None


def _trace_begin():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _trace_end(e0, flops, products):
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    GEMM_TRACE.append((e0, e1, flops, products))


def gemm_ex(**kw):
    """dsb_gemm_ex with tensors given by keyword (a_hi, a_lo, b_hi, b_lo, bias, c, c_hi, c_lo) plus the integer fields."""
    lib = _lib or load()
    e0 = _trace_begin() if GEMM_TRACE is not None else None
    fields = {}
    for k, v in kw.items():
        if v is None:
            continue
        fields[k] = _ptr(v) if isinstance(v, torch.Tensor) else v
    a, b = kw['a_hi'], kw['b_hi']
    if a.dim() == 2:
        fields['a_rows'], fields['a_cols'] = a.shape
    if b.dim() == 2:
        fields['b_rows'], fields['b_cols'] = b.shape
    c = kw.get('c')
    fields['c_rows'], fields['c_cols'] = (c if c is not None else kw['c_hi']).shape
    g = GemmArgs(**fields)
    rc = lib.dsb_gemm_ex(ctypes.byref(g), _stream())
    if rc != 0:
        raise DsbError('dsb_gemm_ex failed (%d): %s' % (rc, lib.dsb_last_error().decode()))
    if e0 is not None:
        terms = kw.get('terms', 3)
        products = 1 if terms == 1 else 3 - (1 if kw.get('a_exact') else 0) - (1 if kw.get('b_exact') else 0)
        _trace_end(e0, 2.0 * kw['m'] * kw['n'] * kw['k'] * max(1, kw.get('batch', 1)), products)


def int_array(values):
    return (_c.c_int * len(values))(*values)


def ptr_array(tensors):
    """host array of device pointers (kept alive by the caller holding the returned object and the tensors)"""
    arr = (_vp * len(tensors))(*[_ptr(t) for t in tensors])
    return arr


def call(name: str, *args):
    lib = _lib or load()
    conv = [(_ptr(a) if isinstance(a, torch.Tensor) else a) for a in args]
    if GEMM_TRACE is not None and name == 'dsb_gemm_bf16_split':
        e0 = _trace_begin()
        rc = getattr(lib, name)(*conv, _stream())
        M, N, K, terms = args[8], args[9], args[10], args[11]
        _trace_end(e0, 2.0 * M * N * K, 3 if terms == 3 else 1)
        if rc != 0:
            raise DsbError('%s failed (%d): %s' % (name, rc, lib.dsb_last_error().decode()))
        return
    rc = getattr(lib, name)(*conv, _stream())
    if rc != 0:
        raise DsbError('%s failed (%d): %s' % (name, rc, lib.dsb_last_error().decode()))
