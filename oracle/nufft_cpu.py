"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the NUFFTs the reference calls.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product never does.

The reference delegates its transforms to the third-party package ``finufft``
(``/root/reference/src/fftvis/cpu/nufft.py:48`` nufft2d3, ``:105`` nufft3d3, ``:162`` nufft2d1;
dependency declared UNPINNED at ``/root/reference/pyproject.toml:35``; its source is not under
/root/reference and it cannot be installed offline).  This module restates finufft's published
algorithm (Barnett, Magland & af Klinteberg 2019, arXiv:1808.06736; parameter rules as recalled
in SURVEY.md Appendix B.1):

* kernel width  w = ceil(-log10(eps/10)) (sigma = 2) or ceil(-ln eps / (pi sqrt(1-1/sigma)))
* beta = 2.30 w (2.20, 2.26, 2.38 for w = 2, 3, 4) at sigma=2, 0.97 pi w (1 - 1/(2 sigma)) otherwise
* fine grid nf = next235even(max(sigma N, 2 w)); type 3 nf = next235even(2 sigma S X / pi + w + 1)
* spread -> FFT -> deconvolve by the kernel's Fourier transform (Gauss-Legendre quadrature)

PARITY UNPINNED: no golden vectors for this boundary exist in the reference and finufft is
absent, so this restatement is pinned only against the direct fp64 sum (``direct_sum``), to the
requested eps (tests/test_oracle.py).  Same public call shapes as the reference's wrappers:
``cpu_nufft2d``, ``cpu_nufft3d``, ``cpu_nufft2d_type1`` (cpu/nufft.py:11,62,120).
"""
from __future__ import annotations

import ctypes
import os
from functools import lru_cache

import numpy as np
import scipy.fft as sfft

from . import build as _build

_c_double_p = ctypes.POINTER(ctypes.c_double)


@lru_cache(maxsize=1)
def _lib():
    lib = ctypes.CDLL(str(_build.build()))
    lib.fvo_max_threads.restype = ctypes.c_int
    return lib


def max_threads() -> int:
    return int(_lib().fvo_max_threads())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------
# ground truth
# --------------------------------------------------------------------------------------
def direct_sum(x, y, z, weights, u, v, w, isign=+1, nthreads=None):
    """fp64 direct evaluation  V[t,k] = sum_s W[t,s] exp(i isign (u_k x_s + v_k y_s + w_k z_s)).

    ``z``/``w`` may be None.  ``weights`` (n,) or (ntr, n).  Returns complex128 like weights.
    """
    nthreads = nthreads or max_threads()
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    z = None if z is None else np.ascontiguousarray(z, np.float64)
    u = np.ascontiguousarray(u, np.float64)
    v = np.ascontiguousarray(v, np.float64)
    w = None if w is None else np.ascontiguousarray(w, np.float64)
    W = np.ascontiguousarray(np.atleast_2d(weights), np.complex128)
    ntr, n = W.shape
    out = np.empty((ntr, u.size), np.complex128)
    _lib().fvo_direct_sum(ctypes.c_int64(n), _ptr(x), _ptr(y), _ptr(z), ctypes.c_int(ntr),
                          _ptr(W), ctypes.c_int64(u.size), _ptr(u), _ptr(v), _ptr(w),
                          ctypes.c_int(isign), _ptr(out), ctypes.c_int(nthreads))
    return out[0] if np.ndim(weights) == 1 else out


# --------------------------------------------------------------------------------------
# kernel / grid parameter rules
# --------------------------------------------------------------------------------------
def kernel_params(eps: float, upsampfac: float = 2.0, precision: int = 2):
    """(w, beta) from the requested tolerance (finufft setup_spreader rules)."""
    mach = 1.1e-16 if precision == 2 else 6e-8
    eps = max(float(eps), mach)
    if upsampfac == 2.0:
        ns = int(np.ceil(-np.log10(eps / 10.0)))
    else:
        ns = int(np.ceil(-np.log(eps) / (np.pi * np.sqrt(1.0 - 1.0 / upsampfac))))
    ns = min(max(ns, 2), 16)
    if upsampfac == 2.0:
        bon = {2: 2.20, 3: 2.26, 4: 2.38}.get(ns, 2.30)
    else:
        bon = 0.97 * np.pi * (1.0 - 1.0 / (2.0 * upsampfac))
    return ns, bon * ns


def next235even(n: int) -> int:
    """Smallest even integer >= n whose only prime factors are 2, 3, 5."""
    n = int(n)
    if n <= 2:
        return 2
    if n % 2:
        n += 1
    while True:
        m = n
        for p in (2, 3, 5):
            while m % p == 0:
                m //= p
        if m == 1:
            return n
        n += 2


def type1_grid_size(n_modes: int, ns: int, upsampfac: float = 2.0) -> int:
    return next235even(max(int(upsampfac * n_modes), 2 * ns))


def _quad_nodes(ns: int):
    """Gauss-Legendre nodes/weights on (0, w/2) used for the kernel's Fourier transform."""
    q = int(2 + 3.0 * (ns / 2.0))
    zz, ww = np.polynomial.legendre.leggauss(2 * q)
    J2 = ns / 2.0
    z = zz[q:] * J2
    return z, ww[q:] * J2


def es_kernel(z, ns, beta):
    z = np.asarray(z, np.float64)
    arg = 1.0 - (2.0 * z / ns) ** 2
    out = np.zeros_like(z)
    m = arg > 0
    out[m] = np.exp(beta * (np.sqrt(arg[m]) - 1.0))
    return out


def kernel_ft_series(nf: int, ns: int, beta: float):
    """phihat(k) for k = 0 .. nf/2 on an nf grid, *including* the (-1)^k of the half-grid
    shift introduced by the fold (x = 0 lands on grid index nf/2)."""
    z, w = _quad_nodes(ns)
    f = w * es_kernel(z, ns, beta)
    k = np.arange(nf // 2 + 1)
    ph = 2.0 * (np.cos(2.0 * np.pi * np.outer(k, z) / nf) @ f)
    return ph * np.where(k % 2 == 0, 1.0, -1.0)


def kernel_ft_at(omega, ns: int, beta: float):
    """phihat(omega) for arbitrary real omega (radians per grid cell) -- type 3."""
    z, w = _quad_nodes(ns)
    f = w * es_kernel(z, ns, beta)
    return 2.0 * (np.cos(np.multiply.outer(np.asarray(omega, np.float64), z)) @ f)


# --------------------------------------------------------------------------------------
# spread / interp (C++, OpenMP)
# --------------------------------------------------------------------------------------
def _rdtype(cdtype):
    return np.float32 if np.dtype(cdtype) == np.complex64 else np.float64


def spread(pts, c, nfs, ns, beta, nthreads=None):
    """pts: list of d coordinate arrays (radians, periodic 2pi); c (ntr, n) complex.
    Returns fw (ntr, nf_d, ..., nf_1) -- x (pts[0]) is the fastest axis."""
    nthreads = nthreads or max_threads()
    cd = c.dtype
    rd = _rdtype(cd)
    d = len(pts)
    pts = [np.ascontiguousarray(p, rd) for p in pts]
    nf = list(nfs) + [1] * (3 - d)
    c = np.ascontiguousarray(c)
    ntr, n = c.shape
    fw = np.empty((ntr,) + tuple(nfs[::-1]), cd)
    fn = _lib().fvo_spread_f32 if rd == np.float32 else _lib().fvo_spread_f64
    fn(ctypes.c_int(d), ctypes.c_int64(nf[0]), ctypes.c_int64(nf[1]), ctypes.c_int64(nf[2]),
       ctypes.c_int64(n), _ptr(pts[0]), _ptr(pts[1]), _ptr(pts[2]) if d == 3 else None,
       ctypes.c_int(ntr), _ptr(c), _ptr(fw), ctypes.c_int(ns), ctypes.c_double(beta),
       ctypes.c_int(nthreads))
    return fw


def interp(pts, fw, ns, beta, nthreads=None):
    """fw (ntr, nf_d, ..., nf_1) -> values (ntr, n) at the NU points."""
    nthreads = nthreads or max_threads()
    cd = fw.dtype
    rd = _rdtype(cd)
    d = len(pts)
    pts = [np.ascontiguousarray(p, rd) for p in pts]
    nfs = fw.shape[1:][::-1]
    nf = list(nfs) + [1] * (3 - d)
    fw = np.ascontiguousarray(fw)
    ntr = fw.shape[0]
    n = pts[0].size
    c = np.empty((ntr, n), cd)
    fn = _lib().fvo_interp_f32 if rd == np.float32 else _lib().fvo_interp_f64
    fn(ctypes.c_int(d), ctypes.c_int64(nf[0]), ctypes.c_int64(nf[1]), ctypes.c_int64(nf[2]),
       ctypes.c_int64(n), _ptr(pts[0]), _ptr(pts[1]), _ptr(pts[2]) if d == 3 else None,
       ctypes.c_int(ntr), _ptr(fw), _ptr(c), ctypes.c_int(ns), ctypes.c_double(beta),
       ctypes.c_int(nthreads))
    return c


def _fft(fw, isign, nthreads):
    axes = tuple(range(1, fw.ndim))
    if isign > 0:  # e^{+i...}: unnormalised backward transform
        return sfft.ifftn(fw, axes=axes, norm="forward", workers=nthreads, overwrite_x=True)
    return sfft.fftn(fw, axes=axes, workers=nthreads, overwrite_x=True)


# --------------------------------------------------------------------------------------
# type 1 (2-D), type 2 (d-D, internal), type 3 (2-D / 3-D)
# --------------------------------------------------------------------------------------
def nufft2d1(x, y, c, n_modes, eps, isign=+1, upsampfac=2.0, nthreads=None):
    """F[t, k1, k2] = sum_j c[t,j] exp(i isign (k1 x_j + k2 y_j)), FFT-ordered output
    (``modeord=1``), k in [-(N-1)/2 .. (N-1)/2] for odd N.  c (n,) or (ntr, n)."""
    nthreads = nthreads or max_threads()
    c2 = np.atleast_2d(c)
    cd = c2.dtype
    prec = 1 if cd == np.complex64 else 2
    ns, beta = kernel_params(eps, upsampfac, prec)
    N = int(n_modes)
    nf = type1_grid_size(N, ns, upsampfac)
    fw = spread([x, y], c2, (nf, nf), ns, beta, nthreads)     # (ntr, nf_y, nf_x)
    fh = _fft(fw, isign, nthreads)
    ph = kernel_ft_series(nf, ns, beta)
    kmin = -(N // 2)
    k = np.arange(kmin, kmin + N)                              # CMCL order
    inv = (1.0 / ph[np.abs(k)]).astype(_rdtype(cd))
    sub = fh[:, (k % nf)[:, None], (k % nf)[None, :]]          # [t, k2, k1]
    out = sub * inv[None, :, None] * inv[None, None, :]
    out = np.transpose(out, (0, 2, 1))                         # [t, k1, k2]
    out = np.fft.ifftshift(out, axes=(1, 2))                   # CMCL -> FFT ordering
    out = np.ascontiguousarray(out.astype(cd, copy=False))
    return out[0] if np.ndim(c) == 1 else out


def _arraywidcen(a):
    lo, hi = float(np.min(a)), float(np.max(a))
    w, c = (hi - lo) / 2.0, (hi + lo) / 2.0
    if abs(c) < 0.1 * w:
        w += abs(c)
        c = 0.0
    return w, c


def type3_grid(S, X, ns, upsampfac=2.0):
    """finufft set_nhg_type3: (nf, h, gamma) for one dimension."""
    Xs, Ss = X, S
    if X == 0.0:
        if S == 0.0:
            Xs, Ss = 1.0, 1.0
        else:
            Xs = max(Xs, 1.0 / S)
    else:
        Ss = max(Ss, 1.0 / X)
    nfd = 2.0 * upsampfac * Ss * Xs / np.pi + (ns + 1)
    if not np.isfinite(nfd):
        nfd = 0.0
    nf = int(nfd)
    if nf < 2 * ns:
        nf = 2 * ns
    nf = next235even(nf)
    return nf, 2.0 * np.pi / nf, nf / (2.0 * upsampfac * Ss)


def _nufft_type2(pts, F, eps, isign, upsampfac, nthreads, prec):
    """values[t, k] = sum_m F[t, m_d.., m_1] exp(i isign m . pts_k); F in CMCL order
    (index i <-> mode i - N/2), x (pts[0]) fastest axis."""
    d = len(pts)
    cd = F.dtype
    ns, beta = kernel_params(eps, upsampfac, prec)
    Ns = F.shape[1:][::-1]
    nfs = [next235even(max(int(upsampfac * N), 2 * ns)) for N in Ns]
    fw = np.zeros((F.shape[0],) + tuple(nfs[::-1]), cd)
    idx, invs = [], []
    for N, nf in zip(Ns, nfs):
        k = np.arange(-(N // 2), -(N // 2) + N)
        ph = kernel_ft_series(nf, ns, beta)
        idx.append(k % nf)
        invs.append((1.0 / ph[np.abs(k)]).astype(_rdtype(cd)))
    if d == 2:
        fw[:, idx[1][:, None], idx[0][None, :]] = F * invs[1][None, :, None] * invs[0][None, None, :]
    else:
        fw[:, idx[2][:, None, None], idx[1][None, :, None], idx[0][None, None, :]] = (
            F * invs[2][None, :, None, None] * invs[1][None, None, :, None]
            * invs[0][None, None, None, :])
    fw = _fft(fw, isign, nthreads)
    return interp(pts, fw, ns, beta, nthreads)


def nufft_type3(xs, c, ss, eps, isign=+1, upsampfac=2.0, nthreads=None):
    """f[t,k] = sum_j c[t,j] exp(i isign s_k . x_j); xs / ss lists of d arrays (d = 2, 3)."""
    nthreads = nthreads or max_threads()
    c2 = np.atleast_2d(c)
    cd = c2.dtype
    rd = _rdtype(cd)
    prec = 1 if cd == np.complex64 else 2
    d = len(xs)
    ns, beta = kernel_params(eps, upsampfac, prec)
    xs = [np.asarray(a, rd) for a in xs]
    ss = [np.asarray(a, rd) for a in ss]
    nfs, hs, gams, Cs, Ds = [], [], [], [], []
    for a, s in zip(xs, ss):
        X, C = _arraywidcen(a)
        S, D = _arraywidcen(s)
        nf, h, gam = type3_grid(S, X, ns, upsampfac)
        nfs.append(nf), hs.append(h), gams.append(gam), Cs.append(C), Ds.append(D)
    xp = [((a - rd(C)) / rd(g)).astype(rd) for a, C, g in zip(xs, Cs, gams)]
    if any(D != 0.0 for D in Ds):
        ph = sum(rd(D) * a for D, a in zip(Ds, xs))
        cp = (c2 * np.exp(1j * isign * ph.astype(np.float64)).astype(cd)).astype(cd)
    else:
        cp = c2
    sp = [(rd(h * g) * (s - rd(D))).astype(rd) for s, h, g, D in zip(ss, hs, gams, Ds)]
    fw = spread(xp, cp, tuple(nfs), ns, beta, nthreads)        # step 1: spread, no FFT
    vals = _nufft_type2(sp, fw, eps, isign, upsampfac, nthreads, prec)   # step 2
    phihat = np.ones(ss[0].size)
    for s_ in sp:
        phihat = phihat * kernel_ft_at(s_.astype(np.float64), ns, beta)
    dec = 1.0 / phihat
    if any(C != 0.0 for C in Cs):
        phase = sum((s.astype(np.float64) - D) * C for s, D, C in zip(ss, Ds, Cs))
        dec = dec * np.exp(1j * isign * phase)
    out = (vals * dec.astype(cd if np.iscomplexobj(dec) else rd)[None, :]).astype(cd)
    return out[0] if np.ndim(c) == 1 else out


# --------------------------------------------------------------------------------------
# the reference's three entry points (same names / argument order, cpu/nufft.py:11,62,120)
# --------------------------------------------------------------------------------------
def cpu_nufft2d(x, y, weights, u, v, eps, n_threads=None, upsample_factor=2):
    return nufft_type3([x, y], weights, [u, v], eps, +1, float(upsample_factor), n_threads)


def cpu_nufft3d(x, y, z, weights, u, v, w, eps, upsample_factor=2, n_threads=None):
    return nufft_type3([x, y, z], weights, [u, v, w], eps, +1, float(upsample_factor), n_threads)


def cpu_nufft2d_type1(x, y, weights, n_modes, index, eps, upsample_factor=2, n_threads=None):
    model = nufft2d1(x, y, weights, n_modes, eps, +1, float(upsample_factor), n_threads)
    return model[..., index[0], index[1]]


# --------------------------------------------------------------------------------------
# restatements of the product's two algebraic shortcuts (checked on the CPU against the direct sum / numpy FFT)
# --------------------------------------------------------------------------------------
def nufft2d1_xdirect(x, y, c, m1, m2, eps, upsampfac=2.0):
    """Type-1 values at the integer modes (m1[k], m2[k]) by the hybrid the product's single-precision pass 1 uses
    (csrc/type1_xdirect.cuh): an exact Fourier sum along x over the distinct first mode numbers, the
    exponential-of-semicircle kernel + FFT + deconvolution along y only:

        T[col, row] = sum_j c_j phi((row - gy_j)) exp(+i k_col x_j);   V[k] = FFT_y(T[col(k)])[m2[k]] / phihat(m2[k])

    Plain numpy in fp64 (test infrastructure: small sizes)."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    c2 = np.atleast_2d(np.asarray(c, np.complex128))
    m1, m2 = np.asarray(m1), np.asarray(m2)
    ns, beta = kernel_params(eps, upsampfac, 2)
    n_modes = 2 * int(max(np.abs(m1).max(), np.abs(m2).max())) + 1
    nf = type1_grid_size(n_modes, ns, upsampfac)
    cols, col_of = np.unique(m1, return_inverse=True)
    # fold y onto the grid: y = -pi -> 0, 0 -> nf / 2 (the half-grid shift is inside kernel_ft_series' sign)
    gy = ((y / (2 * np.pi) + 0.5) % 1.0) * nf
    i0 = np.ceil(gy - ns / 2.0).astype(int)
    T = np.zeros((c2.shape[0], cols.size, nf), np.complex128)
    phase = np.exp(1j * np.outer(cols, x))                                  # (ncols, n): exact along x
    for j in range(ns):
        ker = es_kernel(i0 + j - gy, ns, beta)                               # (n,)
        rows = (i0 + j) % nf
        for t in range(c2.shape[0]):
            np.add.at(T[t], (slice(None), rows), phase * (c2[t] * ker)[None, :])
    Fy = sfft.ifft(T, axis=2, norm="forward")                               # e^{+i ...} along y
    ph = kernel_ft_series(nf, ns, beta)
    out = Fy[:, col_of, m2 % nf] / ph[np.abs(m2)][None, :]
    return out[0] if np.ndim(c) == 1 else out


def padded_fft_half_length(data):
    """The zero-padded, centred-mode transform of the type-3 inner FFT in the half-length form of
    csrc/type3_fft.cuh: ``data`` holds the nin (even) centred modes k - nin / 2; the result is the n = 2 nin point
    transform X[k] = sum_m data[m] exp(+2 pi i (m - nin / 2) k / n), from two nin-point transforms."""
    y = np.asarray(data, np.complex128)
    nin = y.shape[-1]
    n = 2 * nin
    j = np.arange(nin)
    even = sfft.ifft(y, axis=-1, norm="forward")
    odd = sfft.ifft(y * np.exp(2j * np.pi * j / n), axis=-1, norm="forward")
    sgn = np.where(j % 2 == 0, 1.0, -1.0)
    out = np.empty(y.shape[:-1] + (n,), np.complex128)
    out[..., 0::2] = sgn * even
    out[..., 1::2] = sgn * (-1j) * odd
    return out
