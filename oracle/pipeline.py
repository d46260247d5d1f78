"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's visibility pipeline.

Only tests/, __graft_entry__.smoke() and bench.py's ``cpu_baseline`` / ``--impl reference`` legs
may import this module; the product (fftvis_b200/) never does.

Two levels:

``simulate_direct``  the measurement equation evaluated term by term in fp64
    V[f,t,i,j,bl] = sum_{s above horizon} M_s[j,i] exp(+2 pi i f b.s_hat/c)
    (i.e. what /root/reference/src/fftvis/cpu/cpu_simulate.py:856-1071
    computes through the NUFFT).  No gridding, no plane rotation, no beam-pair flipping: it
    checks the engine's whole front half as well.

``simulate_cpu``  the reference's own loop structure (cpu_simulate.py:537-1071): dtype casts,
    default baselines from redundancy, griddability -> type 1 / plane rotation -> type 3,
    per-time rotate + horizon cut, per-frequency beam evaluation, apparent coherency per unique
    beam pair (with the flipped-baseline conjugation), NUFFT via oracle/nufft_cpu.py, basis
    contraction.  This is the arm timed as the CPU baseline ("port": finufft, matvis and
    pyuvdata are absent offline, so their arithmetic is restated -- PARITY UNPINNED at those
    third-party boundaries).
"""
from __future__ import annotations

import numpy as np

from . import beams as obeams
from . import coords as _coords
from . import host_ref as _utils        # reference modules by path where present, else the private restatement
from . import nufft_cpu

_grid = _utils
_catalog = _utils

C_LIGHT = 299792458.0
DEFAULT_EPS = {1: 6e-8, 2: 1e-13}     # reference core/simulate.py:16-19


def _dtypes(precision):
    return (np.float32, np.complex64) if precision == 1 else (np.float64, np.complex128)


def topocentric(ra, dec, times, telescope_loc, precision=2, coord_method="CoordinateRotationERFA",
                coord_method_params=None):
    """Per-time ENU unit vectors (nt, 3, Ns) in the working real dtype, computed in fp64 from the
    dtype-cast ra/dec (oracle/coords.py: the oracle's own chain)."""
    rd, _ = _dtypes(precision)
    return _coords.topocentric_enu(np.asarray(ra, rd), np.asarray(dec, rd), times, telescope_loc,
                                   coord_method, coord_method_params).astype(rd)


def prepare_beam_evaluation(antnums, baselines, beam_idx):
    """Unique (bi <= bj) beam pairs, baseline lists and flip flags -- cpu/beams.py:91-127."""
    if beam_idx is None:
        return [(0, 0)], {(0, 0): np.arange(len(baselines))}, {(0, 0): [False] * len(baselines)}
    ub = np.unique(beam_idx)
    pairs = [(ub[i], ub[j]) for i in range(len(ub)) for j in range(i, len(ub))]
    a2b = dict(zip(antnums, beam_idx))
    to_bls = {p: [] for p in pairs}
    to_flip = {p: [] for p in pairs}
    for k, (ai, aj) in enumerate(baselines):
        bi, bj = a2b[ai], a2b[aj]
        key, fl = ((bi, bj), False) if bi <= bj else ((bj, bi), True)
        to_bls[key].append(k)
        to_flip[key].append(fl)
    return pairs, to_bls, to_flip


def _evaluate_beams(beam_list, az, za, polarized, freq, fidx, spline_opts, cdtype):
    out = []
    for b in beam_list:
        ev = obeams.evaluate_beam(b, az, za, polarized, freq, fidx, spline_opts)
        out.append(np.asarray(ev).astype(cdtype))
    return out


# ----------------------------------------------------------------------------------------
def simulate_direct(ants, fluxes, ra, dec, freqs, times, beam_list, telescope_loc,
                    baselines=None, beam_idx=None, precision=2, polarized=False,
                    beam_spline_opts=None, beam_coefs=None, reference_flip_quirk=True,
                    nthreads=None, coord_method="CoordinateRotationERFA", coord_method_params=None):
    """Direct fp64 evaluation of the measurement equation on the dtype-cast inputs.
    Returns (nf, nt, nbls) or (nf, nt, 2, 2, nbls) complex128."""
    rd, _ = _dtypes(precision)
    freqs = np.asarray(freqs).astype(rd).astype(np.float64)
    nf, nt = freqs.size, len(_coords.jd_of(times))
    ants = {k: np.asarray(v, dtype=float) for k, v in ants.items()}
    antnums = list(ants.keys())
    if baselines is None:
        baselines = [r[0] for r in _utils.get_pos_reds(ants, include_autos=True)]
    baselines = [tuple(b) for b in baselines]
    nbls = len(baselines)
    beam_idx = _utils.validate_beam_idx(beam_idx, beam_coefs, len(beam_list), len(ants))
    coherency, pol_sky = _catalog.prepare_source_catalog(np.asarray(fluxes), polarized)
    coherency = coherency.astype(np.complex128)
    antvecs = np.array([ants[a] for a in antnums]).astype(rd).astype(np.float64)
    a_index = {a: i for i, a in enumerate(antnums)}
    blvec = np.array([antvecs[a_index[b[1]]] - antvecs[a_index[b[0]]] for b in baselines])  # (nbls,3)
    topo_all = topocentric(ra, dec, times, telescope_loc, precision, coord_method,
                           coord_method_params).astype(np.float64)
    nfeed = 2 if polarized else 1
    vis = np.zeros((nt, nbls, nfeed, nfeed, nf), np.complex128)

    if beam_coefs is None:
        pairs, to_bls, to_flip = prepare_beam_evaluation(antnums, baselines, beam_idx)
    for ti in range(nt):
        topo = topo_all[ti]
        up = topo[2] > 0
        if not up.any():
            continue
        tp = topo[:, up]
        flux = coherency[up]
        az, za = _coords.enu_to_az_za(tp[0], tp[1], "uvbeam")
        x, y, z = (2 * np.pi * tp[i] for i in range(3))
        for fi, f in enumerate(freqs):
            u, v, w = (blvec[:, i] * f / C_LIGHT for i in range(3))
            bev = _evaluate_beams(beam_list, az, za, polarized, f, fi, beam_spline_opts, np.complex128)
            if beam_coefs is not None:
                # per-antenna beams A_ant = sum_k c[ant,k,f] phi_k, then every baseline on its own
                for k, (a1, a2) in enumerate(baselines):
                    Ai = sum(beam_coefs[a_index[a1], q, fi] * bev[q] for q in range(len(bev)))
                    Aj = sum(beam_coefs[a_index[a2], q, fi] * bev[q] for q in range(len(bev)))
                    app = obeams.compute_apparent_coherency([Ai, Aj], 0, 1, flux[:, fi], polarized, pol_sky)
                    r = nufft_cpu.direct_sum(x, y, z, app, u[k:k + 1], v[k:k + 1], w[k:k + 1], nthreads=1)
                    vis[ti, k, :, :, fi] = r[:, 0].reshape(nfeed, nfeed).T
                continue
            for (bi, bj) in pairs:
                idx = np.asarray(to_bls[(bi, bj)], dtype=int)
                if idx.size == 0:
                    continue
                fl = np.asarray(to_flip[(bi, bj)], dtype=bool)
                app = obeams.compute_apparent_coherency(bev, bi, bj, flux[:, fi], polarized, pol_sky)
                sgn = np.where(fl, -1.0, 1.0)
                r = nufft_cpu.direct_sum(x, y, z, app, sgn * u[idx], sgn * v[idx], sgn * w[idx],
                                         nthreads=nthreads)          # (P, nb)
                if reference_flip_quirk:
                    r = np.where(fl[None, :], r.conj(), r)
                    vis[ti, idx, :, :, fi] = np.swapaxes(r.reshape(nfeed, nfeed, idx.size), 2, 0)
                else:
                    # physically exact: a flipped baseline is the conjugate *transpose*
                    blk = np.swapaxes(r.reshape(nfeed, nfeed, idx.size), 2, 0)
                    blk = np.where(fl[:, None, None], np.swapaxes(blk.conj(), 1, 2), blk)
                    vis[ti, idx, :, :, fi] = blk
    return np.transpose(vis, (4, 0, 2, 3, 1)) if polarized else np.moveaxis(vis[..., 0, 0, :], 2, 0)


# ----------------------------------------------------------------------------------------
def plan_array(ants, baselines, precision, flat_array_tol=1e-6, force_use_type3=False):
    """Front half of ``simulate`` (cpu_simulate.py:628-681): gridded? rotation, bls, n_modes."""
    rd, _ = _dtypes(precision)
    antnums = list(ants.keys())
    a_index = {a: i for i, a in enumerate(antnums)}
    antvecs = np.array([ants[a] for a in antnums], dtype=rd)
    basis_matrix, n_modes = None, None
    if np.abs(antvecs[:, -1]).max() > flat_array_tol or force_use_type3:
        is_gridded = False
    else:
        is_gridded, gridded, basis_matrix = _grid.check_antpos_griddability(ants)
    if not is_gridded:
        rot = np.ascontiguousarray(_utils.get_plane_to_xy_rotation_matrix(antvecs).T)
        rants = rot @ antvecs.T
        bls = np.array([rants[:, a_index[b[1]]] - rants[:, a_index[b[0]]] for b in baselines]).T
        is_coplanar = bool(np.all(np.abs(bls[2]) <= flat_array_tol))
        bls = (bls / C_LIGHT).astype(rd)
        rot = rot.astype(rd)
        basis_matrix = None
    else:
        bls = np.round(np.array([gridded[b[1]] - gridded[b[0]] for b in baselines]).T).astype(int)
        n_modes = 2 * int(np.round(np.max(np.abs(bls)))) + 1
        basis_matrix = (basis_matrix / C_LIGHT).astype(rd)
        is_coplanar = True
        rot = np.eye(3, dtype=rd)
    return dict(is_gridded=is_gridded, rotation_matrix=rot, bls=bls, is_coplanar=is_coplanar,
                basis_matrix=basis_matrix, n_modes=n_modes, antnums=antnums)


def _run_nufft(app, topo, uvw, bls, flipped, idx, use_type1, is_coplanar, tx, ty, n_modes, eps,
               nthreads, upsample_factor, nfeed, direct=False):
    """cpu_simulate.py:205-300."""
    flipped = np.asarray(flipped, dtype=bool)
    if use_type1:
        b = np.where(flipped, -bls[:, idx], bls[:, idx])
        if direct:
            r = nufft_cpu.direct_sum(tx, ty, None, app, b[0], b[1], None, nthreads=nthreads)
        else:
            r = nufft_cpu.cpu_nufft2d_type1(tx, ty, app, n_modes, b, eps, upsample_factor, nthreads)
    else:
        q = np.where(flipped, -uvw[:, idx], uvw[:, idx])
        if direct:
            r = nufft_cpu.direct_sum(topo[0], topo[1], None if is_coplanar else topo[2], app, q[0],
                                     q[1], None if is_coplanar else q[2], nthreads=nthreads)
        elif is_coplanar:
            r = nufft_cpu.cpu_nufft2d(topo[0], topo[1], app, q[0], q[1], eps, nthreads, upsample_factor)
        else:
            r = nufft_cpu.cpu_nufft3d(topo[0], topo[1], topo[2], app, q[0], q[1], q[2], eps,
                                      upsample_factor, nthreads)
    r = np.where(flipped, np.conj(r), r)
    return np.swapaxes(r.reshape(nfeed, nfeed, len(idx)), 2, 0)


def simulate_cpu(ants, fluxes, ra, dec, freqs, times, beam_list, telescope_loc, baselines=None,
                 beam_idx=None, precision=2, polarized=False, eps=None, upsample_factor=2,
                 beam_spline_opts=None, flat_array_tol=1e-6, force_use_type3=False,
                 beam_coefs=None, nthreads=None, direct=False, freq_slice=None, time_slice=None,
                 coord_method="CoordinateRotationERFA", coord_method_params=None):
    """Reference-structured CPU pipeline (see module docstring).  ``direct=True`` swaps the
    NUFFT for the fp64 direct sum while keeping every other step (casts, rotation, gridding)."""
    rd, cd = _dtypes(precision)
    eps = DEFAULT_EPS[precision] if eps is None else eps
    ants = {k: np.asarray(v, dtype=float) for k, v in ants.items()}
    ra, dec = np.asarray(ra).astype(rd), np.asarray(dec).astype(rd)
    freqs = np.asarray(freqs).astype(rd)
    beam_idx = _utils.validate_beam_idx(beam_idx, beam_coefs, len(beam_list), len(ants))
    if baselines is None:
        baselines = [r[0] for r in _utils.get_pos_reds(ants, include_autos=True)]
    baselines = [tuple(b) for b in baselines]
    nbls = len(baselines)
    coherency, pol_sky = _catalog.prepare_source_catalog(np.asarray(fluxes), polarized)
    coherency = coherency.astype(cd)
    plan = plan_array(ants, baselines, precision, flat_array_tol, force_use_type3)
    rot, bls, antnums = plan["rotation_matrix"], plan["bls"], plan["antnums"]
    use_type1, is_coplanar, basis = plan["is_gridded"], plan["is_coplanar"], plan["basis_matrix"]
    topo_all = topocentric(ra, dec, times, telescope_loc, precision, coord_method, coord_method_params)
    nt = topo_all.shape[0]
    t_idx = range(nt)[time_slice or slice(None)]
    f_idx = range(freqs.size)[freq_slice or slice(None)]
    nfeed = 2 if polarized else 1
    vis = np.zeros((len(t_idx), nbls, nfeed, nfeed, len(f_idx)), cd)
    use_basis = beam_coefs is not None
    if use_basis:
        a1 = np.array([antnums.index(b[0]) for b in baselines])
        a2 = np.array([antnums.index(b[1]) for b in baselines])
    else:
        pairs, to_bls, to_flip = prepare_beam_evaluation(antnums, baselines, beam_idx)
    rot_is_identity = np.allclose(rot, np.eye(3))
    for to, ti in enumerate(t_idx):
        topo = topo_all[ti]
        up = topo[2] > 0
        n = int(up.sum())
        if n == 0:
            continue
        topo = np.ascontiguousarray(topo[:, up])
        flux = coherency[up]
        az, za = _coords.enu_to_az_za(topo[0], topo[1], "uvbeam")
        if not rot_is_identity:
            topo = (rot @ topo).astype(rd)
        if basis is not None:
            topo = (basis.T @ topo).astype(rd)
        topo = topo * rd(2 * np.pi)
        for fo, fi in enumerate(f_idx):
            f = freqs[fi]
            uvw = None if use_type1 else bls * f
            bev = _evaluate_beams(beam_list, az, za, polarized, float(f), fi, beam_spline_opts, cd)
            tx, ty = (topo[0] * f, topo[1] * f) if use_type1 else (None, None)
            if use_basis:
                K = len(bev)
                c1 = beam_coefs[a1, :, fi].conj()
                c2 = beam_coefs[a2, :, fi]
                allb = np.arange(nbls)
                noflip = np.zeros(nbls, bool)
                acc = np.zeros((nbls, nfeed, nfeed), cd)
                for k in range(K):
                    for l in range(k, K):
                        app = obeams.compute_apparent_coherency(bev, k, l, flux[:, fi], polarized, pol_sky).astype(cd)
                        vkl = _run_nufft(app, topo, uvw, bls, noflip, allb, use_type1, is_coplanar,
                                         tx, ty, plan["n_modes"], eps, nthreads, upsample_factor, nfeed, direct)
                        acc += (c1[:, k] * c2[:, l])[:, None, None] * vkl
                        if l != k:
                            acc += (c1[:, l] * c2[:, k])[:, None, None] * vkl.swapaxes(1, 2)
                vis[to, :, :, :, fo] += acc
                continue
            for (bi, bj) in pairs:
                idx = np.asarray(to_bls[(bi, bj)], dtype=int)
                if idx.size == 0:
                    continue
                app = obeams.compute_apparent_coherency(bev, bi, bj, flux[:, fi], polarized, pol_sky).astype(cd)
                v = _run_nufft(app, topo, uvw, bls, to_flip[(bi, bj)], idx, use_type1, is_coplanar,
                               tx, ty, plan["n_modes"], eps, nthreads, upsample_factor, nfeed, direct)
                vis[to, idx, :, :, fo] += v
    return np.transpose(vis, (4, 0, 2, 3, 1)) if polarized else np.moveaxis(vis[..., 0, 0, :], 2, 0)
