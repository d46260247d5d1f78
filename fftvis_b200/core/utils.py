"""Host-side planners shared by every engine (numpy only, run once per ``simulate`` call).

Behavioural mirror of ``/root/reference/src/fftvis/core/utils.py`` -- same function names,
argument meaning, return values and error strings -- re-implemented with array operations
instead of the reference's Python double loops.  Results are
pinned against fixtures generated from the reference's own module
(``tests/golden/make_golden.py`` -> ``tests/golden/core_utils.npz``).
"""
from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger(__name__)

speed_of_light = 299792458.0  # m/s, reference core/utils.py:9


def memo_by_value(maxsize: int = 4):
    """Memoise a pure host planner on the VALUES of its arguments (antenna-position dicts and arrays are keyed
    by their bytes).  An array layout or a set of observation times is a property of the instrument, not of
    one ``simulate`` call: repeated calls with the same layout skip the O(N_ant^2) planning, a different
    layout recomputes.  Results are returned as deep copies, so callers may modify them."""
    import copy
    import functools

    def key_of(v):
        if isinstance(v, dict):
            return ("d", tuple(v.keys()), tuple(np.asarray(x, dtype=float).tobytes() for x in v.values()))
        if isinstance(v, np.ndarray):
            return ("a", v.dtype.str, v.shape, v.tobytes())
        if isinstance(v, (list, tuple)):
            return ("l", tuple(key_of(x) for x in v))
        return v

    def _fresh(r):
        # lists of (lists of) immutable tuples: new list objects; anything else: a deep copy
        if isinstance(r, list) and all(isinstance(g, (tuple, list)) for g in r):
            return [list(g) if isinstance(g, list) else g for g in r]
        return copy.deepcopy(r)

    def deco(fn):
        cache = {}

        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            try:
                key = (tuple(key_of(a) for a in args), tuple(sorted((k, key_of(v)) for k, v in kwargs.items())))
                hash(key)
            except TypeError:
                return fn(*args, **kwargs)
            if key not in cache:
                if len(cache) >= maxsize:
                    cache.pop(next(iter(cache)))
                cache[key] = fn(*args, **kwargs)
            return _fresh(cache[key])

        wrapped.cache_clear = cache.clear
        return wrapped
    return deco


def _pair_table(antpos: dict, include_autos: bool):
    """All (i<j [, i==j]) antenna pairs in the dict's iteration order (outer i, inner j)."""
    keys = list(antpos.keys())
    pos = np.array([np.asarray(antpos[k], dtype=float) for k in keys]).reshape(len(keys), -1)
    korder = np.array(keys)
    ii, jj = np.meshgrid(np.arange(len(keys)), np.arange(len(keys)), indexing="ij")
    sel = korder[ii] < korder[jj]
    if include_autos:
        sel |= ii == jj
    return keys, pos, ii[sel], jj[sel]


@memo_by_value()
def get_pos_reds(antpos, decimals=3, include_autos=True, representatives_only=False):
    """Redundant-baseline groups from antenna positions (reference core/utils.py:11-71).

    Groups pairs whose rounded (u, v) separation agrees up to sign (``w`` is ignored); groups
    appear in first-seen order, members in scan order, a member found with the opposite sign
    is stored reversed, and finally every group whose *first* member has ``b_y < 0`` is
    reversed as a whole.  ``representatives_only=True`` returns just the first member of every
    group (what the engines use as the default baseline list, reference cpu_simulate.py:614-616)
    without building the member lists.
    """
    keys, pos, i_idx, j_idx = _pair_table(antpos, include_autos)
    if i_idx.size == 0:
        return []
    uv = np.round(pos[j_idx] - pos[i_idx], decimals)[:, :2] + 0.0  # +0.0: fold -0.0 into 0.0
    # canonical orientation: first non-zero component positive
    neg = (uv[:, 0] < 0) | ((uv[:, 0] == 0) & (uv[:, 1] < 0))
    canon = np.where(neg[:, None], -uv, uv) + 0.0
    # group on one int64 key per pair (the separations are already rounded to `decimals`)
    ik = np.rint(canon * 10.0**decimals).astype(np.int64)
    key = ik[:, 0] * (np.int64(1) << 32) + (ik[:, 1] + (np.int64(1) << 31))
    if representatives_only:
        _, first = np.unique(key, return_index=True)
        first_sorted = np.sort(first)
        hi, hj = i_idx[first_sorted], j_idx[first_sorted]
        swap = (pos[hj, 1] - pos[hi, 1]) < 0
        a1 = np.where(swap, hj, hi).tolist()
        a2 = np.where(swap, hi, hj).tolist()
        return [(keys[p], keys[q]) for p, q in zip(a1, a2)]
    _, first, inverse = np.unique(key, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    rank = np.empty(first.size, dtype=int)
    rank[np.argsort(first, kind="stable")] = np.arange(first.size)
    group_of_pair = rank[inverse]
    first_sorted = np.sort(first)

    order = np.argsort(group_of_pair, kind="stable")
    bounds = np.searchsorted(group_of_pair[order], np.arange(first.size + 1))
    reds = []
    for g in range(first.size):
        members = order[bounds[g]:bounds[g + 1]]
        head = first_sorted[g]
        same = neg[members] == neg[head]
        # zero separation (autos): +d and -d coincide; the reference's scan stores them reversed,
        # which for an auto pair is the same tuple.
        grp = [
            (keys[i_idx[m]], keys[j_idx[m]]) if s else (keys[j_idx[m]], keys[i_idx[m]])
            for m, s in zip(members, same)
        ]
        a1, a2 = grp[0]
        if (np.asarray(antpos[a2], dtype=float) - np.asarray(antpos[a1], dtype=float))[1] < 0:
            grp = [(b, a) for (a, b) in grp]
        reds.append(grp)
    return reds


def get_plane_to_xy_rotation_matrix(antvecs):
    """Rotation taking the least-squares plane through the antennas onto z = const
    (reference core/utils.py:74-119).  Identity when the fitted slopes vanish."""
    antvecs = np.asarray(antvecs, dtype=float)
    design = np.column_stack([antvecs[:, 0], antvecs[:, 1], np.ones(len(antvecs))])
    from scipy import linalg  # same LAPACK driver (gelsd) as the reference

    (sx, sy, _), *_ = linalg.lstsq(design, antvecs[:, 2])
    if abs(sx) <= 1e-8 and abs(sy) <= 1e-8:
        return np.eye(3)
    nz = -1.0 / np.sqrt(sx * sx + sy * sy + 1.0)      # z-component of the unit normal
    k = np.array([sy, -sx, 0.0]) / np.hypot(sx, sy)   # rotation axis (in-plane)
    theta = np.arccos(-nz)
    ct, st = np.cos(theta), np.sin(theta)
    kx = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    # Rodrigues:  R = I + sin(t) [k]x + (1 - cos(t)) [k]x^2
    return np.eye(3) + st * kx + (1.0 - ct) * (kx @ kx)


def get_task_chunks(nprocesses: int, nfreqs: int, ntimes: int):
    """Split (frequency, time) into ``nprocesses`` rectangular tasks, frequencies kept whole
    when possible (reference core/utils.py:122-187).  Returns
    ``(nprocesses, freq_chunks, time_chunks, nf, nt)``."""
    ntasks = nfreqs * ntimes
    if ntasks < 2 * nprocesses:
        return 1, [slice(None)], [slice(None)], nfreqs, ntimes

    def _shape(nfc):
        return int(np.ceil(nfreqs / nfc)), int(np.ceil(ntimes / (nprocesses / nfc)))

    nfc, (nf, nt) = 1, _shape(1)
    sizes = [nf * nt]
    while nf > 1 and nprocesses * nf * nt > ntasks:
        nfc += 1
        nf, nt = _shape(nfc)
        sizes.append(nf * nt)
    nfc = 1 + int(np.argmin(sizes))
    nf, nt = _shape(nfc)
    ntc = int(np.ceil(nprocesses / nfc))
    freq_chunks = [slice(nf * i, min(nfreqs, nf * (i + 1))) for i in range(nfc)] * ntc
    time_chunks = []
    for i in range(ntc):
        time_chunks += [slice(nt * i, min(ntimes, nt * (i + 1)))] * nfc
    return nprocesses, freq_chunks, time_chunks, nf, nt


def inplace_rot_base(rot, b):
    """``b[:, s] <- rot @ b[:, s]`` in place (reference core/utils.py:190-211)."""
    b[...] = np.asarray(rot, dtype=b.dtype) @ b


def get_required_chunks(freemem, nax, nfeed, nant, nsrc, nbeam, nbeampix, precision,
                        source_buffer: float = 1.0, nprocesses: int = 1) -> int:
    """Number of source-axis chunks needed to fit ``freemem`` bytes
    (memory model of reference core/utils.py:213-285)."""
    rsize = 4 * precision
    csize = 2 * rsize
    ch = 0
    total = freemem
    while total >= freemem and ch < 100:
        ch += 1
        nchunk = int(nsrc // ch * source_buffer)
        total = (
            nant * 3 * rsize + nsrc * rsize + nbeampix * nfeed * nax * csize
            + 3 * nsrc * rsize + 3 * nsrc * rsize * nprocesses
            + 3 * nchunk * rsize * nprocesses + nchunk * rsize * nprocesses
            + nbeam * nfeed * nax * nchunk * csize * nprocesses
            + ch * nfeed * nant * nfeed * nant * csize
        )
    logger.info("free mem %.2f GB -> %d source chunks (estimate %.2f GB)",
                freemem / 1024**3, ch, total / 1024**3)
    return ch


def get_desired_chunks(freemem, min_chunks, beam_list, nax, nfeed, nant, nsrc, precision,
                       source_buffer: float = 1.0):
    """(nchunks, sources per chunk) -- reference core/utils.py:287-355."""
    nbeampix = sum(
        b.data_array.shape[-2] * b.data_array.shape[-1] for b in beam_list if hasattr(b, "data_array")
    )
    need = get_required_chunks(freemem, nax, nfeed, nant, nsrc, len(beam_list), nbeampix,
                               precision, source_buffer)
    nchunks = max(1, min(max(min_chunks, need), nsrc))     # (the reference divides by zero on an empty sky)
    return nchunks, int(np.ceil(nsrc / nchunks))


def validate_beam_idx(beam_idx, beam_coefs, nbeam: int, nant: int):
    """Validate / infer the antenna -> beam map (reference core/utils.py:358-430; the error
    strings are the ones the reference's tests match)."""
    if beam_coefs is not None:
        if beam_idx is not None:
            raise ValueError(
                "beam_idx should not be provided when beam_coefs is given. "
                "The mapping from antennas to beams is defined by beam_coefs."
            )
        return beam_idx
    if beam_idx is None:
        if nbeam == nant:
            beam_idx = np.arange(nant)
        elif nbeam != 1:
            raise ValueError(
                "If number of beams provided is not 1 or nant, beam_idx must be provided."
            )
    if beam_idx is not None:
        beam_idx = np.asarray(beam_idx)
        if beam_idx.shape != (nant,):
            raise ValueError("beam_idx must be length nant")
        if not np.all((beam_idx >= 0) & (beam_idx < nbeam)):
            raise ValueError("beam_idx contains indices greater than the number of beams")
    return beam_idx
