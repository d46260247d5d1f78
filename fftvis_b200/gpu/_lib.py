"""ctypes binding of the C ABI declared in include/fftvis_b200.h.

There is no CPU fallback: if the CUDA library has not been built, or no GPU is present, every
entry point of the GPU backend raises.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_void_p
from functools import lru_cache
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent.parent / "libfftvis_b200.so"


class FVError(RuntimeError):
    """A non-zero status from libfftvis_b200 (message from ``fv_last_error_string``)."""


class fv_beam(ctypes.Structure):
    _fields_ = [
        ("kind", c_int32), ("is_power", c_int32), ("diameter", c_double), ("table", c_void_p),
        ("nza", c_int32), ("naz", c_int32), ("az_wrap_period", c_int32), ("az_pad", c_int32),
        ("az0", c_double), ("daz", c_double), ("za0", c_double), ("dza", c_double),
        ("order", c_int32), ("freq_offset", c_int32), ("spline_pad", c_int32), ("reserved", c_int32),
    ]


class fv_epilogue(ctypes.Structure):
    _fields_ = [
        ("out", c_void_p), ("out_stride_b", c_int64), ("out_stride_p", c_int64),
        ("pmap", c_int32 * 4), ("kmap", c_void_p), ("conj_flag", c_void_p),
        ("accumulate", c_int32),
    ]


_SIGNATURES = {
    "fv_last_error_string": (c_char_p, []),
    "fv_version": (c_int, []),
    "fv_device_count": (c_int, [POINTER(c_int)]),
    "fv_launch_count": (c_int64, []),
    "fv_memcpy2d_async": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "fv_kernel_params": (c_int, [c_double, c_double, c_int, POINTER(c_int), POINTER(c_double)]),
    "fv_next235even": (c_int64, [c_int64]),
    "fv_rotate_cut_scratch_bytes": (c_int64, [c_int64]),
    "fv_rotate_cut": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int64, POINTER(c_double),
                              POINTER(c_double), POINTER(c_double), c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                              c_void_p, c_void_p, c_void_p]),
    "fv_inplace_rot": (c_int, [c_int, POINTER(c_double), c_void_p, c_int64, c_void_p]),
    "fv_weights": (c_int, [c_int, c_int, POINTER(fv_beam), POINTER(fv_beam), c_void_p, c_void_p,
                           c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64, c_void_p, c_int64,
                           c_void_p, c_void_p, c_void_p]),
    "fv_weights_basis": (c_int, [c_int, c_int, POINTER(fv_beam), c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_int, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "fv_coherency": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "fv_tiles_create": (c_int, [POINTER(c_void_p), c_void_p]),
    "fv_tiles_destroy": (c_int, [c_void_p]),
    "fv_tiles_supported": (c_int, [POINTER(fv_beam), c_int]),
    "fv_tiles_sort": (c_int, [c_void_p, c_int, POINTER(fv_beam), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int64]),
    "fv_weights_tiled": (c_int, [c_void_p, c_int, c_int, POINTER(fv_beam), c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_int, c_int64, c_void_p, c_int64, c_void_p]),
    "fv_plan_create": (c_int, [POINTER(c_void_p), c_void_p]),
    "fv_plan_destroy": (c_int, [c_void_p]),
    "fv_plan_set_timing": (c_int, [c_void_p, c_int]),
    "fv_plan_reset_timing": (c_int, [c_void_p]),
    "fv_plan_stage_ms": (c_int, [c_void_p, c_int, POINTER(c_double), POINTER(c_int64)]),
    "fv_plan_bytes": (c_int64, [c_void_p]),
    "fv_nufft2d1": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                            POINTER(c_double), c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                            c_int64, c_double, c_double, POINTER(fv_epilogue)]),
    "fv_modeset_create": (c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_int64, c_int]),
    "fv_modeset_destroy": (c_int, [c_void_p]),
    "fv_nufft2d1_fused": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                                  POINTER(c_double), c_int, c_int, c_void_p, c_void_p, c_double, c_double,
                                  POINTER(fv_epilogue)]),
    "fv_plan_set_option": (c_int, [c_void_p, c_char_p, c_int64]),
    "fv_plan_last_geometry": (c_int, [c_void_p, POINTER(c_int64)]),
    "fv_nufft3": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                          POINTER(c_double), c_void_p, c_void_p, c_void_p, c_int64,
                          POINTER(c_double), POINTER(c_double), c_int, c_int, c_void_p, c_double,
                          c_double, POINTER(fv_epilogue)]),
    "fv_minmax": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                          POINTER(c_double)]),
    "fv_direct_sum": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                              c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double), c_int, c_int,
                              c_void_p, POINTER(fv_epilogue), c_void_p]),
    "fv_basis_contract": (c_int, [c_int, c_void_p, c_int, c_int64, c_void_p, c_int64, c_int, c_int64,
                                  c_int64, c_int, c_int, c_void_p, c_void_p, POINTER(fv_epilogue),
                                  c_void_p]),
    "fv_basis_contract_all": (c_int, [c_int, c_void_p, c_int, c_int64, c_void_p, c_int64, c_int, c_int64,
                                      c_int64, c_void_p, c_void_p, POINTER(fv_epilogue), c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)
STAGES = ("zero", "spread", "fft", "gather", "deconv", "interp")


@lru_cache(maxsize=1)
def lib() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        raise FVError(
            f"{LIB_PATH} is missing: build it with `python -m fftvis_b200.csrc.build` "
            "(fftvis_b200 has no CPU fallback)"
        )
    handle = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    return handle


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().fv_last_error_string().decode(errors="replace")
        raise FVError(f"{what or 'libfftvis_b200'} failed (status {status}): {msg}")


def require_gpu() -> int:
    """Number of CUDA devices; raises FVError (never falls back) when there is none."""
    n = c_int(0)
    check(lib().fv_device_count(ctypes.byref(n)), "fv_device_count")
    return n.value


def doubles(values):
    arr = (c_double * len(values))(*[float(v) for v in values])
    return arr


def make_epilogue(out_ptr: int, stride_b: int, stride_p: int, pmap=(0, 1, 2, 3), kmap_ptr: int = 0,
                  conj_ptr: int = 0, accumulate: bool = False) -> fv_epilogue:
    e = fv_epilogue()
    e.out = out_ptr
    e.out_stride_b = stride_b
    e.out_stride_p = stride_p
    for i in range(4):
        e.pmap[i] = pmap[i]
    e.kmap = kmap_ptr or None
    e.conj_flag = conj_ptr or None
    e.accumulate = 1 if accumulate else 0
    return e
