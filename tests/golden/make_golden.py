"""Generate tests/golden/*.npz from the REFERENCE's own modules.

Run in the authoring container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference package cannot be imported as a whole (finufft / matvis / pyuvdata / astropy /
ray are absent), but three of its modules depend only on numpy / scipy / numba and load by file
path: core/utils.py, core/antenna_gridding.py, cpu/utils.py.  Their outputs on seeded inputs are
the golden vectors for the host-side planners.  The four numba coherency kernels live in
cpu/beams.py, which imports pyuvdata at module scope; their bodies are extracted *by line range*
from the reference source at generation time (nothing is copied into this repo) and compiled
with numba's decorators stripped to produce the golden coherency vectors.
"""
from __future__ import annotations

import importlib.util
import json
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src/fftvis")
OUT = Path(__file__).resolve().parent


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def hex_array(nside, spacing=14.6):
    """Hexagon with ``nside`` antennas per side (own generator; hera_sim is absent)."""
    pos = {}
    k = 0
    for row in range(-(nside - 1), nside):
        ncol = 2 * nside - 1 - abs(row)
        for col in range(ncol):
            x = (col - (ncol - 1) / 2.0) * spacing
            y = row * spacing * np.sqrt(3) / 2
            pos[k] = np.array([x, y, 0.0])
            k += 1
    return pos


def main():
    cu = _load("ref_core_utils", REF / "core/utils.py")
    ag = _load("ref_gridding", REF / "core/antenna_gridding.py")
    pu = _load("ref_cpu_utils", REF / "cpu/utils.py")

    rng = np.random.default_rng(42)
    arrays = {
        "hex2": hex_array(2), "hex3": hex_array(3), "hex4": hex_array(4),
        "line": {i: np.array([i * 10.0, 0.0, 0.0]) for i in range(5)},
        "square": {i * 4 + j: np.array([i * 7.0, j * 7.0, 0.0]) for i in range(4) for j in range(4)},
        "random": {i: np.append(rng.uniform(-50, 50, 2), 0.0) for i in range(12)},
        "tilted": {i: np.array([*p[:2], 0.05 * p[0] - 0.02 * p[1] + 1.0]) for i, p in hex_array(3).items()},
    }
    # holey hex: remove antennas (tests/test_cpu_simulate.py:199-271 does the same with rng 42)
    full = hex_array(4)
    drop = set(rng.choice(len(full), 7, replace=False).tolist())
    arrays["holey_hex4"] = {k: v for k, v in full.items() if k not in drop}
    shear = np.array([[1, 0.5], [0, 1]])
    arrays["sheared_square"] = {k: np.append(shear @ v[:2], 0.0) for k, v in arrays["square"].items()}

    gold = {}
    meta = {}
    for name, ants in arrays.items():
        reds = cu.get_pos_reds(ants, include_autos=True)
        gold[f"{name}/antpos"] = np.array([ants[k] for k in ants])
        gold[f"{name}/antkeys"] = np.array(list(ants.keys()))
        gold[f"{name}/red_first"] = np.array([r[0] for r in reds])
        gold[f"{name}/red_sizes"] = np.array([len(r) for r in reds])
        gold[f"{name}/red_flat"] = np.array([bl for r in reds for bl in r])
        reds_na = cu.get_pos_reds(ants, include_autos=False)
        gold[f"{name}/red_first_noautos"] = np.array([r[0] for r in reds_na])
        vec = np.array([ants[k] for k in ants])
        gold[f"{name}/plane_rot"] = cu.get_plane_to_xy_rotation_matrix(vec)
        ok, gridded, basis = ag.check_antpos_griddability(ants)
        meta[f"{name}/griddable"] = bool(ok)
        gold[f"{name}/basis"] = np.asarray(basis, dtype=float)
        if ok:
            gold[f"{name}/gridded"] = np.array([gridded[k] for k in ants])

    chunks = {}
    for args in [(3, 30, 1), (10, 5, 1), (8, 1024, 60), (2, 2, 3), (4, 16, 16), (6, 20, 30), (8, 7, 3)]:
        npz, fc, tc, nf, nt = cu.get_task_chunks(*args)
        chunks[str(args)] = dict(
            nproc=npz, nf=nf, nt=nt,
            fc=[[s.start, s.stop] for s in fc], tc=[[s.start, s.stop] for s in tc])
    meta["task_chunks"] = chunks

    # prepare_source_catalog
    sky_i = rng.uniform(0.5, 2, (6, 3))
    sky_iquv = rng.normal(size=(6, 3, 4))
    gold["catalog/sky_i"] = sky_i
    gold["catalog/coh_i"] = pu.prepare_source_catalog(sky_i, False)[0]
    gold["catalog/sky_iquv"] = sky_iquv
    gold["catalog/coh_iquv"] = pu.prepare_source_catalog(sky_iquv, True)[0]
    # inplace_rot
    rot = cu.get_plane_to_xy_rotation_matrix(np.array([arrays["tilted"][k] for k in arrays["tilted"]]))
    b = rng.normal(size=(3, 9))
    gold["rot/rot"], gold["rot/b_in"] = rot, b.copy()
    b2 = b.copy()
    cu.inplace_rot_base(rot, b2)
    gold["rot/b_out"] = b2

    # the four coherency kernels: compile the reference's function bodies (decorators stripped)
    src = (REF / "cpu/beams.py").read_text().splitlines()
    start = next(i for i, l in enumerate(src) if "def get_apparent_flux_polarized_beam(" in l)
    body = "\n".join(l[4:] if l.startswith("    ") else l for l in src[start - 2:])
    body = re.sub(r"@staticmethod\n", "", body)
    body = re.sub(r"@nb\.jit\([^)]*\)\n", "", body)
    ns = {"np": np}
    exec(compile(body, "ref_cpu_beams_kernels", "exec"), ns)
    n = 11
    bi = rng.normal(size=(2, 2, n)) + 1j * rng.normal(size=(2, 2, n))
    bj = rng.normal(size=(2, 2, n)) + 1j * rng.normal(size=(2, 2, n))
    flux = rng.uniform(0.1, 3, n)
    coh = rng.normal(size=(2, 2, n)) + 1j * rng.normal(size=(2, 2, n))
    gold["coh/beam_i"], gold["coh/beam_j"], gold["coh/flux"], gold["coh/coherency"] = bi, bj, flux, coh
    t = bi.copy(); ns["get_apparent_flux_polarized_beam"](t, flux); gold["coh/out_polarized_beam"] = t
    t = bi.copy(); ns["get_apparent_flux_polarized"](t, coh); gold["coh/out_polarized"] = t
    t = np.zeros_like(bi); ns["get_apparent_flux_polarized_beam_pair"](bi, bj, flux, t); gold["coh/out_beam_pair"] = t
    t = np.zeros_like(bi); ns["get_apparent_flux_polarized_pair"](bi, bj, coh, t); gold["coh/out_pair"] = t

    np.savez_compressed(OUT / "reference_host.npz", **gold)
    (OUT / "reference_host.json").write_text(json.dumps(meta, indent=1, sort_keys=True))
    print("wrote", OUT / "reference_host.npz", len(gold), "arrays")


if __name__ == "__main__":
    sys.exit(main())
