"""TEST INFRASTRUCTURE ONLY -- the host planners the CPU pipeline restatement needs, standing on
their own (nothing here imports ``fftvis_b200``).

Where the reference tree is present (the authoring container: ``/root/reference``), the
reference's OWN modules are loaded by file path and used as they are:

    core/utils.py             get_pos_reds, get_plane_to_xy_rotation_matrix, validate_beam_idx
    core/antenna_gridding.py  check_antpos_griddability
    cpu/utils.py              prepare_source_catalog            (needs numba to import)

They depend on numpy / scipy / numba only, so they load without the package ``__init__``.  On the
GPU box the tree is absent; there the private restatements below are used.  They follow the same
reference lines (cited per function) and are pinned to the reference in two ways:
``tests/test_oracle.py::test_host_restatement_equals_reference_modules`` compares them with the
loaded modules whenever the tree exists, and ``tests/test_host_golden.py`` compares them with the
committed fixtures generated from those modules (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import importlib.util
import math
from fractions import Fraction
from pathlib import Path

import numpy as np

REF_SRC = Path("/root/reference/src/fftvis")
C_LIGHT = 299792458.0


def _load_by_path(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_modules():
    """(core_utils, antenna_gridding, cpu_utils) of the reference, or None where the tree (or a
    dependency of one of the three files) is missing."""
    if not REF_SRC.exists():
        return None
    try:
        return (_load_by_path("_fv_ref_core_utils", REF_SRC / "core" / "utils.py"),
                _load_by_path("_fv_ref_gridding", REF_SRC / "core" / "antenna_gridding.py"),
                _load_by_path("_fv_ref_cpu_utils", REF_SRC / "cpu" / "utils.py"))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# private restatements (used when the reference tree is absent)
# ---------------------------------------------------------------------------------------------
def _restated_get_pos_reds(antpos, decimals=3, include_autos=True):
    """core/utils.py:11-71.  Scan the pairs (ai, aj), ai < aj (or ai == aj with autos), outer ai;
    a pair joins the group of the first pair seen with the same rounded (u, v) -- reversed when it
    matches with the opposite sign; a group whose first member points to negative y is reversed."""
    keys = list(antpos)
    group_of = {}        # rounded (u, v) -> index into `groups`
    groups = []
    for ai in keys:
        for aj in keys:
            if not (ai < aj or (include_autos and ai == aj)):
                continue
            d = np.round(np.asarray(antpos[aj], float) - np.asarray(antpos[ai], float), decimals)
            u, v = float(d[0]), float(d[1])
            fwd, bwd = (u, v), (-u, -v)
            if fwd not in group_of and bwd not in group_of:
                group_of[fwd] = len(groups)
                groups.append([(ai, aj)])
            elif bwd in group_of:             # the reference tests the reversed key first
                groups[group_of[bwd]].append((aj, ai))
            else:
                groups[group_of[fwd]].append((ai, aj))
    out = []
    for g in groups:
        a, b = g[0]
        dy = float(np.asarray(antpos[b], float)[1] - np.asarray(antpos[a], float)[1])
        out.append([(q, p) for (p, q) in g] if dy < 0 else g)
    return out


def _restated_plane_rotation(antvecs):
    """core/utils.py:74-119: least-squares plane z = sx x + sy y + z0 (scipy.linalg.lstsq), then the
    Rodrigues rotation about (sy, -sx, 0) by the tilt angle; identity when both slopes are ~0."""
    from scipy import linalg
    a = np.asarray(antvecs, float)
    coef = linalg.lstsq(np.stack([a[:, 0], a[:, 1], np.ones(len(a))], axis=1), a[:, 2])[0]
    sx, sy = float(coef[0]), float(coef[1])
    if np.isclose(sx, 0) and np.isclose(sy, 0.0):
        return np.eye(3)
    nrm = np.array([sx, sy, -1.0])
    nrm /= np.linalg.norm(nrm)
    ax = np.array([sy, -sx, 0.0])
    ax /= np.linalg.norm(ax)
    th = math.acos(-nrm[2])
    K = np.zeros((3, 3))
    K[0, 1], K[0, 2], K[1, 0], K[1, 2], K[2, 0], K[2, 1] = -ax[2], ax[1], ax[2], -ax[0], -ax[1], ax[0]
    return np.eye(3) + math.sin(th) * K + (1.0 - math.cos(th)) * (K @ K)


def _restated_validate_beam_idx(beam_idx, beam_coefs, nbeam, nant):
    """core/utils.py:408-428 (same error strings)."""
    if beam_coefs is not None:
        if beam_idx is not None:
            raise ValueError("beam_idx should not be provided when beam_coefs is given. "
                             "The mapping from antennas to beams is defined by beam_coefs.")
        return None
    if beam_idx is None:
        if nbeam == nant:
            return np.arange(nant)
        if nbeam != 1:
            raise ValueError("If number of beams provided is not 1 or nant, beam_idx must be provided.")
        return None
    if beam_idx.shape != (nant,):
        raise ValueError("beam_idx must be length nant")
    if any((i < 0 or i >= nbeam) for i in beam_idx):
        raise ValueError("beam_idx contains indices greater than the number of beams")
    return beam_idx


def _restated_griddability(antpos, tol=1e-9, max_denominator=10**6, max_factor=1000):
    """core/antenna_gridding.py:72-219: basis = shortest separation + the shortest one not collinear
    with it; express the positions (relative to the first antenna) in that basis; griddable when the
    lcm of the rational approximations' denominators (<= max_factor) scales them all to integers."""
    keys = list(antpos)
    vec = np.array([np.asarray(antpos[k], float) for k in keys])
    xy = vec[:, :2]
    sep = (xy[:, None, :] - xy[None, :, :]).reshape(-1, 2)
    ln = np.linalg.norm(sep, axis=1)
    keep = ln > tol
    if not keep.any():
        return False, antpos, np.eye(vec.shape[-1])
    sep = sep[keep][np.argsort(ln[keep])]
    b1, b2 = sep[0], None
    for cand in sep[1:]:
        if abs(b1[0] * cand[1] - b1[1] * cand[0]) > tol:
            b2 = cand
            break
    B = np.zeros((3, 3))
    B[:2, :2] = np.column_stack([b1, b2]) if b2 is not None else np.vstack([b1, np.array([0, 1])])
    B[2, 2] = 1.0
    frac = np.linalg.solve(B, (vec - vec[0]).T).T
    dens = [Fraction(float(v)).limit_denominator(max_denominator).denominator for v in frac.ravel() if v != 0]
    factor = math.lcm(*dens) if dens else 1
    if factor > max_factor:
        return False, antpos, np.eye(vec.shape[-1])
    scaled = factor * frac
    if not np.allclose(scaled, np.round(scaled), atol=tol):
        return False, antpos, np.eye(vec.shape[-1])
    return True, {k: np.round(scaled[i]).astype(int) for i, k in enumerate(keys)}, B / factor


def _restated_source_catalog(sky_model, polarized_beam):
    """cpu/utils.py:26-80: 0.5 I, or the coherency 0.5 [[I+Q, U+iV], [U-iV, I-Q]] as (n, nf, 2, 2)."""
    sky_model = np.asarray(sky_model)
    if sky_model.ndim == 2:
        return 0.5 * sky_model, False
    if not (polarized_beam and sky_model.ndim == 3 and sky_model.shape[-1] == 4):
        if polarized_beam:
            raise ValueError("polarized_beam=True requires sky_model to be either:\n  2D unpolarized, or\n"
                             "  3D with last axis of length 4; "
                             f"got ndim={sky_model.ndim}, shape={sky_model.shape}")
        raise ValueError("polarized_beam=False requires sky_model to be 2D; "
                         f"got ndim={sky_model.ndim}, shape={sky_model.shape}")
    I, Q, U, V = np.moveaxis(sky_model, -1, 0)
    rows = np.stack([np.stack([I + Q, U + 1j * V], axis=-1), np.stack([U - 1j * V, I - Q], axis=-1)], axis=-2)
    return 0.5 * rows, True


RESTATED = dict(get_pos_reds=_restated_get_pos_reds, get_plane_to_xy_rotation_matrix=_restated_plane_rotation,
                validate_beam_idx=_restated_validate_beam_idx, check_antpos_griddability=_restated_griddability,
                prepare_source_catalog=_restated_source_catalog)

_REF = load_reference_modules()
SOURCE = "reference modules loaded from /root/reference" if _REF is not None else "private restatement"

if _REF is not None:
    get_pos_reds = _REF[0].get_pos_reds
    get_plane_to_xy_rotation_matrix = _REF[0].get_plane_to_xy_rotation_matrix
    validate_beam_idx = _REF[0].validate_beam_idx
    check_antpos_griddability = _REF[1].check_antpos_griddability
    prepare_source_catalog = _REF[2].prepare_source_catalog
else:
    get_pos_reds = _restated_get_pos_reds
    get_plane_to_xy_rotation_matrix = _restated_plane_rotation
    validate_beam_idx = _restated_validate_beam_idx
    check_antpos_griddability = _restated_griddability
    prepare_source_catalog = _restated_source_catalog
