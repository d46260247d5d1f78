"""GPU non-uniform FFTs: the function-level API of the reference's GPU slot.

Fills ``gpu_nufft2d`` / ``gpu_nufft3d`` (stubs at /root/reference/src/fftvis/gpu/nufft.py:11-98,
same positional arguments) and adds ``gpu_nufft2d_type1`` (the stub set lacks it; CPU analogue
/root/reference/src/fftvis/cpu/nufft.py:120-175).  Host numpy arrays in, host numpy arrays out;
the arithmetic is libfftvis_b200's CUDA kernels + cuFFT.  Sign convention exp(+i ...).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib

_RDT = {1: torch.float32, 2: torch.float64}
_CDT = {1: torch.complex64, 2: torch.complex128}


class ModeSet:
    """Integer modes (m1, m2) of one beam pair's baselines, bucketed by m1 on the host once
    (``fv_modeset``): the target set of the fused type-1 transform."""

    def __init__(self, m1, m2, n_modes: int):
        m1 = np.ascontiguousarray(m1, dtype=np.int32)
        m2 = np.ascontiguousarray(m2, dtype=np.int32)
        if m1.shape != m2.shape or m1.ndim != 1:
            raise ValueError("m1 and m2 must be 1-D arrays of equal length")
        self.nk = int(m1.size)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().fv_modeset_create(ctypes.byref(h), m1.ctypes.data, m2.ctypes.data, self.nk,
                                                int(n_modes)), "fv_modeset_create")
        self._h = h

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().fv_modeset_destroy(self._h)
                self._h = None
        except Exception:
            pass


class NufftPlan:
    """Owner of an ``fv_plan`` (cuFFT plan cache + work grids) bound to one device and stream."""

    def __init__(self, device=None, stream: torch.cuda.Stream | None = None):
        _lib.require_gpu()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.device):
            self.stream = stream or torch.cuda.current_stream()
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().fv_plan_create(ctypes.byref(h), self.stream.cuda_stream), "fv_plan_create")
        self._h = h
        for env, opt in (("FV_T1_ROWS", "t1_rows"), ("FV_T1_COLS", "t1_cols"), ("FV_T3_VX", "t3_vx"),
                         ("FV_T3_VY", "t3_vy"), ("FV_T3_VZ", "t3_vz"), ("FV_T3_THRX", "t3_thrx"),
                         ("FV_T3_THRY", "t3_thry"), ("FV_T3_THRZ", "t3_thrz"), ("FV_T3_HALF", "t3_half"),
                         ("FV_T3_MINBY", "t3_minby"), ("FV_T3_MINBZ", "t3_minbz")):   # tuning experiments
            if os.environ.get(env):
                self.set_option(opt, int(os.environ[env]))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().fv_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_timing(self, enable: bool):
        _lib.check(_lib.lib().fv_plan_set_timing(self._h, int(enable)))

    def reset_timing(self):
        _lib.check(_lib.lib().fv_plan_reset_timing(self._h))

    def stage_times(self) -> dict:
        """{stage: (milliseconds, launches)} since the last reset (synchronises)."""
        out = {}
        for i, name in enumerate(_lib.STAGES):
            ms, cnt = ctypes.c_double(0), ctypes.c_int64(0)
            _lib.check(_lib.lib().fv_plan_stage_ms(self._h, i, ctypes.byref(ms), ctypes.byref(cnt)))
            out[name] = (ms.value, cnt.value)
        return out

    def last_type3_geometry(self) -> dict | None:
        """Grid geometry of the last type-3 transform of this plan (None before the first)."""
        g = (ctypes.c_int64 * 12)()
        _lib.check(_lib.lib().fv_plan_last_geometry(self._h, g))
        if g[0] == 0:
            return None
        dim = int(g[0])
        nf, ng = [int(g[2 + d]) for d in range(dim)], [int(g[5 + d]) for d in range(dim)]
        return dict(dim=dim, w=int(g[1]), nf=nf, ng=ng, G1=int(np.prod(nf)), G2=int(np.prod(ng)),
                    tiles=bool(g[8]), own_fft=bool(g[9]), sub_batch=int(g[10]), ntr=int(g[11]))

    def bytes(self) -> int:
        return int(_lib.lib().fv_plan_bytes(self._h))

    # ---- device-level calls (torch tensors on self.device) --------------------------------
    def type1(self, prec, bx, by, n_dev, scale, W, n_modes, m1, m2, eps, upsampfac, epi):
        nb, ntr, n_cap = W.shape
        _lib.check(_lib.lib().fv_nufft2d1(
            self._h, prec, bx.data_ptr(), by.data_ptr(), n_dev.data_ptr(), n_cap,
            _lib.doubles(scale), nb, ntr, W.data_ptr(), int(n_modes), m1.data_ptr(), m2.data_ptr(),
            m1.numel(), float(eps), float(upsampfac), ctypes.byref(epi)), "fv_nufft2d1")

    def type1_fused(self, prec, bx, by, n_dev, scale, W, modes: ModeSet, eps, upsampfac, epi):
        nb, ntr, n_cap = W.shape
        _lib.check(_lib.lib().fv_nufft2d1_fused(
            self._h, prec, bx.data_ptr(), by.data_ptr(), n_dev.data_ptr(), n_cap, _lib.doubles(scale), nb,
            ntr, W.data_ptr(), modes.handle, float(eps), float(upsampfac), ctypes.byref(epi)),
            "fv_nufft2d1_fused")

    def set_option(self, name: str, value: int):
        _lib.check(_lib.lib().fv_plan_set_option(self._h, name.encode(), int(value)), "fv_plan_set_option")

    def type3(self, prec, dim, xyz, n_dev, xlim, uvw, ulim, scale, W, eps, upsampfac, epi):
        nb, ntr, n_cap = W.shape
        z = xyz[2].data_ptr() if dim == 3 else None
        w = uvw[2].data_ptr() if dim == 3 else None
        _lib.check(_lib.lib().fv_nufft3(
            self._h, prec, dim, xyz[0].data_ptr(), xyz[1].data_ptr(), z, n_dev.data_ptr(), n_cap,
            _lib.doubles(xlim) if xlim is not None else None,
            uvw[0].data_ptr(), uvw[1].data_ptr(), w, uvw[0].numel(),
            _lib.doubles(ulim) if ulim is not None else None,
            _lib.doubles(scale), nb, ntr, W.data_ptr(), float(eps), float(upsampfac),
            ctypes.byref(epi)), "fv_nufft3")


_PLANS: dict = {}


def default_plan() -> NufftPlan:
    _lib.require_gpu()
    torch.cuda.init()
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    if key not in _PLANS:
        _PLANS[key] = NufftPlan()
    return _PLANS[key]


def _prec_of(weights) -> int:
    return 1 if np.asarray(weights).dtype in (np.complex64, np.float32) else 2


def _to_dev(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype, non_blocking=False)


def _one_shot(dim, pts, weights, targets, eps, upsample_factor):
    _lib.require_gpu()
    weights = np.asarray(weights)
    prec = _prec_of(weights)
    rdt, cdt = _RDT[prec], _CDT[prec]
    W = _to_dev(np.atleast_2d(weights), cdt).unsqueeze(0).contiguous()       # (1, ntr, n)
    n = W.shape[-1]
    xyz = [_to_dev(p, rdt) for p in pts]
    uvw = [_to_dev(t, rdt) for t in targets]
    nk = uvw[0].numel()
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    out = torch.empty((1, W.shape[1], nk), dtype=cdt, device="cuda")
    epi = _lib.make_epilogue(out.data_ptr(), out.stride(0), out.stride(1))
    plan = default_plan()
    if n == 0:
        out.zero_()
    else:
        plan.type3(prec, dim, xyz, n_dev, None, uvw, None, [1.0], W, eps, upsample_factor, epi)
    res = out[0].cpu().numpy()
    return res[0] if weights.ndim == 1 else res


def gpu_nufft2d(x, y, weights, u, v, eps, n_threads: int = 1, upsample_factor=2):
    """Type-3 2-D transform  out[k] = sum_s w_s exp(i (u_k x_s + v_k y_s))
    (stub: reference gpu/nufft.py:11-50; CPU: cpu/nufft.py:11-59).  ``n_threads`` is ignored."""
    return _one_shot(2, [x, y], weights, [u, v], eps, float(upsample_factor))


def gpu_nufft3d(x, y, z, weights, u, v, w, eps, n_threads: int = 1, upsample_factor=2):
    """Type-3 3-D transform (stub: reference gpu/nufft.py:53-98; CPU: cpu/nufft.py:62-118)."""
    return _one_shot(3, [x, y, z], weights, [u, v, w], eps, float(upsample_factor))


def gpu_nufft2d_type1(x, y, weights, n_modes, index, eps, upsample_factor=2, n_threads: int = 1,
                      method: str = "fused"):
    """Type-1 2-D transform onto ``n_modes`` x ``n_modes`` integer modes followed by the gather
    ``model[..., index[0], index[1]]`` (CPU: cpu/nufft.py:120-175).  ``index`` holds signed mode
    numbers; negative ones address the FFT-ordered negative frequencies, as in the reference.
    ``method``: "fused" (shared-memory spread + FFT, the default) or "cufft" (global fine grid)."""
    _lib.require_gpu()
    weights = np.asarray(weights)
    prec = _prec_of(weights)
    rdt, cdt = _RDT[prec], _CDT[prec]
    index = np.asarray(index)
    n_modes = int(n_modes)
    m = index[:2].astype(np.int64)
    if np.any(m >= n_modes) or np.any(m < -n_modes):     # numpy fancy-index semantics of cpu/nufft.py:175
        raise IndexError(f"index out of bounds for axis with size {n_modes}")
    # FFT-ordered array position (negative = from the end) -> signed mode number
    pos = np.where(m < 0, m + n_modes, m)
    m = np.where(pos > (n_modes - 1) // 2, pos - n_modes, pos)
    W = _to_dev(np.atleast_2d(weights), cdt).unsqueeze(0).contiguous()
    n = W.shape[-1]
    bx, by = _to_dev(x, rdt), _to_dev(y, rdt)
    m1 = torch.as_tensor(m[0].astype(np.int32)).cuda()
    m2 = torch.as_tensor(m[1].astype(np.int32)).cuda()
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    out = torch.zeros((1, W.shape[1], m1.numel()), dtype=cdt, device="cuda")
    epi = _lib.make_epilogue(out.data_ptr(), out.stride(0), out.stride(1))
    if n and m1.numel():
        if method == "fused":
            modes = ModeSet(m[0], m[1], n_modes)
            default_plan().type1_fused(prec, bx, by, n_dev, [1.0], W, modes, eps, float(upsample_factor), epi)
            torch.cuda.current_stream().synchronize()      # modes is released on return
        elif method == "cufft":
            default_plan().type1(prec, bx, by, n_dev, [1.0], W, n_modes, m1, m2, eps, float(upsample_factor), epi)
        else:
            raise ValueError("method must be 'fused' or 'cufft'")
    res = out[0].cpu().numpy()
    return res[0] if weights.ndim == 1 else res
