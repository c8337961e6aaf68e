"""ctypes binding of include/nmfb200.h -- one Python function per C entry point, nothing else.
The library is required: there is no CPU or PyTorch fallback behind this module."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnmfb200.so")

OK, EINVAL, EDIM, ECUDA, ENCCL, ENOMEM, ESTATE, ENOTSUP, ENUMERIC = range(9)
UNIQUE_ID_BYTES = 128


class NmfResult(ctypes.Structure):
    _fields_ = [
        ("niters", ctypes.c_int64),
        ("converged", ctypes.c_int32),
        ("engine", ctypes.c_int32),
        ("objvalue", ctypes.c_double),
        ("last_dev", ctypes.c_double),
        ("solve_ms", ctypes.c_double),
        ("upload_ms", ctypes.c_double),
        ("coordinate_updates", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("hot_kernel_ms", ctypes.c_double),
        ("hot_kernel_launches", ctypes.c_int64),
        ("sub_iterations", ctypes.c_int64),
        ("tolg_final", ctypes.c_double),
    ]


TRACE_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                            ctypes.c_double, ctypes.c_double)

# name -> (restype, argtypes); kept in one table so tests can check it against the header
_vp, _i, _i64, _f, _d, _cp = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_char_p
_u64 = ctypes.c_uint64


def _solve_sig(ct):
    return (_i, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, ct, ct, ct, _i, _i, _i, ctypes.POINTER(NmfResult)])


def _solve_cd_sig(ct):
    return (_i, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, ct, ct, ct, _i, _i, _u64, _i, _i, _i, ctypes.POINTER(NmfResult)])


def _solve_alspgrad_sig(ct):
    return (_i, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, ct, ct, _i, _i, _i, ctypes.POINTER(NmfResult)])


SIGNATURES = {
    "nmfb200_version": (_i, []),
    "nmfb200_status_string": (_cp, [_i]),
    "nmfb200_create": (_i, [ctypes.POINTER(_vp), _i, _i]),
    "nmfb200_destroy": (_i, [_vp]),
    "nmfb200_last_error": (_cp, [_vp]),
    "nmfb200_set_stream": (_i, [_vp, _vp]),
    "nmfb200_set_option": (_i, [_vp, _cp, _cp]),
    "nmfb200_set_trace": (_i, [_vp, TRACE_FN, _vp]),
    "nmfb200_set_X_f32": (_i, [_vp, _vp, _i64, _i64, _i64, _i]),
    "nmfb200_set_X_f64": (_i, [_vp, _vp, _i64, _i64, _i64, _i]),
    "nmfb200_set_X_dev_f32": (_i, [_vp, _vp, _i64, _i64, _i64, _i]),
    "nmfb200_set_X_dev_f64": (_i, [_vp, _vp, _i64, _i64, _i64, _i]),
    "nmfb200_set_X_csc_f32": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _i, _i]),
    "nmfb200_set_X_csc_f64": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _i, _i]),
    "nmfb200_solve_multmse_f32": _solve_sig(_f),
    "nmfb200_solve_multmse_f64": _solve_sig(_d),
    "nmfb200_solve_multmse_batched_f32": (_i, [_vp, _vp, _i64, _vp, _i64, _i64, ctypes.c_int32, _i64, _f, _f, _f, _i, _i,
                                               ctypes.POINTER(NmfResult)]),
    "nmfb200_solve_multdiv_f32": _solve_sig(_f),
    "nmfb200_solve_multdiv_f64": _solve_sig(_d),
    "nmfb200_solve_greedycd_f32": _solve_sig(_f),
    "nmfb200_solve_greedycd_f64": _solve_sig(_d),
    "nmfb200_solve_projals_f32": _solve_sig(_f),
    "nmfb200_solve_projals_f64": _solve_sig(_d),
    "nmfb200_solve_cd_f32": _solve_cd_sig(_f),
    "nmfb200_solve_cd_f64": _solve_cd_sig(_d),
    "nmfb200_solve_alspgrad_f32": _solve_alspgrad_sig(_f),
    "nmfb200_solve_alspgrad_f64": _solve_alspgrad_sig(_d),
    "nmfb200_mul_X_f32": (_i, [_vp, _i, _vp, _i64, _i64, _vp, _i64]),
    "nmfb200_mul_X_f64": (_i, [_vp, _i, _vp, _i64, _i64, _vp, _i64]),
    "nmfb200_randinit_f32": (_i, [_vp, _vp, _i64, _vp, _i64, _i64, _u64, _i64, _i64, _i, _i, _i]),
    "nmfb200_randinit_f64": (_i, [_vp, _vp, _i64, _vp, _i64, _i64, _u64, _i64, _i64, _i, _i, _i]),
    "nmfb200_rsvd_f32": (_i, [_vp, _i64, _u64, _vp, _i64, _vp, _vp, _i64]),
    "nmfb200_rsvd_f64": (_i, [_vp, _i64, _u64, _vp, _i64, _vp, _vp, _i64]),
    "nmfb200_nndsvd_f32": (_i, [_vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _u64, _i]),
    "nmfb200_nndsvd_f64": (_i, [_vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _u64, _i]),
    "nmfb200_comm_unique_id": (_i, [_vp]),
    "nmfb200_shard_geometry": (_i, [_i64, _i, _i, ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "nmfb200_comm_init": (_i, [_vp, _i, _i, _vp]),
    "nmfb200_comm_destroy": (_i, [_vp]),
}

_lib = None


class LibraryMissing(ImportError):
    pass


def load() -> ctypes.CDLL:
    """dlopen libnmfb200.so (built by nmf.jl_b200/build.py).  Raises loudly if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} is missing: the CUDA library is the product and has no fallback. "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'`."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
