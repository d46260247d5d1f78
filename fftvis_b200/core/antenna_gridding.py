"""Lattice detection for the type-1 (gridded array) path -- host side, runs once per call.

Behavioural mirror of ``/root/reference/src/fftvis/core/antenna_gridding.py``: the same
decisions (shortest baseline + first non-collinear baseline as lattice basis, rational scaling
to integers, ``max_factor``) so that ``basis_matrix``, the integer antenna coordinates and hence
``n_modes`` agree with the reference.  Pinned by ``tests/golden/gridding.npz`` (generated from
the reference module itself) and the reference's own cases (tests/test_antenna_gridding.py:59-82).
"""
from __future__ import annotations

from fractions import Fraction
from math import lcm
from typing import Any, Dict, Tuple

import numpy as np

from .utils import memo_by_value as _memo


def find_integer_multiplier(arr: np.ndarray, max_denominator: int = 10**6) -> int:
    """Least common denominator of the (rationalised) non-zero entries of ``arr``
    (reference antenna_gridding.py:7-37)."""
    dens = [
        Fraction(float(v)).limit_denominator(max_denominator).denominator
        for v in np.ravel(arr) if v != 0
    ]
    return lcm(*dens) if dens else 1


def can_scale_to_int(arr: np.ndarray, tol: float = 1e-9, max_denominator: int = 10**6,
                     max_factor: int = None) -> Tuple[bool, int]:
    """(ok, f) with f*arr integral to ``tol`` (reference antenna_gridding.py:40-75)."""
    f = find_integer_multiplier(arr, max_denominator)
    if max_factor is not None and f > max_factor:
        return False, f
    scaled = f * np.asarray(arr, dtype=float)
    return bool(np.allclose(scaled, np.round(scaled), atol=tol)), f


def find_lattice_basis(antpos: Dict[Any, np.ndarray], tol: float = 1e-9):
    """2x2 matrix whose columns generate the antenna lattice, or None for autos only
    (reference antenna_gridding.py:77-137; including its collinear-array fallback, which
    stacks the vectors as *rows*)."""
    xy = np.array([np.asarray(antpos[a], dtype=float)[:2] for a in antpos])
    sep = (xy[:, None, :] - xy[None, :, :]).reshape(-1, 2)
    length = np.linalg.norm(sep, axis=1)
    keep = length > tol
    if not keep.any():
        return None
    sep = sep[keep][np.argsort(length[keep])]
    b1 = sep[0]
    cross = b1[0] * sep[1:, 1] - b1[1] * sep[1:, 0]
    hits = np.nonzero(np.abs(cross) > tol)[0]
    if hits.size == 0:
        return np.vstack([b1, np.array([0, 1])])
    return np.column_stack([b1, sep[1 + hits[0]]])


@_memo()
def check_antpos_griddability(antpos: Dict[Any, np.ndarray], tol: float = 1e-9,
                              max_denominator: int = 10**6, max_factor: int = 1000):
    """(is_griddable, integer antenna coords, 3x3 basis/factor) -- reference
    antenna_gridding.py:139-219.  Not griddable -> (False, antpos, eye(3))."""
    keys = list(antpos.keys())
    vecs = np.array([antpos[a] for a in keys], dtype=float)
    b2 = find_lattice_basis(antpos, tol=tol)
    if b2 is None:
        return False, antpos, np.eye(vecs.shape[-1])
    basis = np.zeros((3, 3))
    basis[:2, :2] = b2
    basis[2, 2] = 1.0
    coords = np.linalg.solve(basis, (vecs - vecs[0]).T).T
    ok, factor = can_scale_to_int(np.ravel(coords), tol=tol, max_denominator=max_denominator,
                                  max_factor=max_factor)
    if not ok:
        return False, antpos, np.eye(vecs.shape[-1])
    grid = np.round(factor * coords).astype(int)
    return True, {a: grid[i] for i, a in enumerate(keys)}, basis / factor
