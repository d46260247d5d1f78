"""Synthetic arrays, skies and beams for tests and benchmarks (SURVEY.md section 8(d) "common
synthetic inputs"): there is no network for GLEAM / HERA layout files / CST beams, so every
workload is generated from a seed.  Host-side numpy only."""
from __future__ import annotations

import numpy as np

from .beam_models import UVBeamTable

HEX_SPACING = 14.6     # metres (HERA)


def hex_array(hex_num: int, spacing: float = HEX_SPACING, z: float = 0.0) -> dict:
    """Filled hexagon with ``hex_num`` antennas per side (3 hex_num (hex_num-1) + 1 antennas),
    numbered row by row."""
    ants, k = {}, 0
    n = hex_num - 1
    for r in range(-n, n + 1):
        for q in range(max(-n, -n - r), min(n, n - r) + 1):
            ants[k] = np.array([spacing * (q + 0.5 * r), spacing * np.sqrt(3) / 2 * r, z])
            k += 1
    return ants


def hex_rows(rows=(3, 4, 3), spacing: float = HEX_SPACING) -> dict:
    """Close-packed rows of antennas, each row centred (``(3, 4, 3)``: the 10-antenna hex of
    BASELINE configs[0])."""
    ants, k = {}, 0
    for r, cnt in enumerate(rows):
        for q in range(cnt):
            ants[k] = np.array([spacing * (q - 0.5 * (cnt - 1)), spacing * np.sqrt(3) / 2 * r, 0.0])
            k += 1
    return ants


def hera350_like(spacing: float = HEX_SPACING) -> dict:
    """A 350-element HERA-like layout that stays on one lattice (so the reference's gridding logic
    selects the type-1 path): hex core of 11 per side split into three sectors displaced by thirds
    of the lattice vectors (320 antennas after removing 11 on the sector seams), plus 30
    outriggers on the same third-spacing sub-lattice."""
    n = 10
    a1 = np.array([spacing, 0.0])
    a2 = np.array([0.5 * spacing, np.sqrt(3) / 2 * spacing])
    offs = [np.zeros(2), (a1 + a2) / 3.0, 2.0 * (a1 + a2) / 3.0 - a2]
    pos = []
    for r in range(-n, n + 1):
        for q in range(max(-n, -n - r), min(n, n - r) + 1):
            s = -q - r
            xy = q * a1 + r * a2
            # three 120-degree sectors of the hexagon
            sec = int(np.floor(np.mod(np.arctan2(xy[1], xy[0]) + 1e-9, 2 * np.pi) / (2 * np.pi / 3))) % 3
            pos.append((xy + offs[sec], sec, max(abs(q), abs(r), abs(s))))
    # drop 11 antennas from the outermost ring of sector 2 (deterministic) -> 320
    drop = [i for i, p in enumerate(pos) if p[1] == 2 and p[2] == n][:11]
    core = [p[0] for i, p in enumerate(pos) if i not in set(drop)]
    # 30 outriggers on the sub-lattice, two rings
    out = []
    sub1, sub2 = a1 / 3.0, a2 / 3.0
    for ring, cnt in ((45, 18), (60, 12)):
        for j in range(cnt):
            ang = 2 * np.pi * (j + 0.5 * (ring == 60)) / cnt
            target = ring * np.linalg.norm(sub1) * np.array([np.cos(ang), np.sin(ang)])
            c = np.linalg.solve(np.column_stack([sub1, sub2]), target)
            out.append(np.round(c[0]) * sub1 + np.round(c[1]) * sub2)
    allp = core + out
    return {i: np.array([p[0], p[1], 0.0]) for i, p in enumerate(allp)}


def random_array(nant: int, radius: float = 150.0, zspan: float = 2.0, seed: int = 42) -> dict:
    """Non-griddable layout: uniform in a disc, z ~ U(-zspan, zspan) (BASELINE config 4)."""
    rng = np.random.default_rng(seed)
    r = radius * np.sqrt(rng.uniform(0, 1, nant))
    th = rng.uniform(0, 2 * np.pi, nant)
    z = rng.uniform(-zspan, zspan, nant)
    return {i: np.array([r[i] * np.cos(th[i]), r[i] * np.sin(th[i]), z[i]]) for i in range(nant)}


def all_baselines(ants: dict, autos: bool = False) -> list:
    k = list(ants.keys())
    return [(k[i], k[j]) for i in range(len(k)) for j in range(i + (0 if autos else 1), len(k))]


def random_sky(nsrc: int, freqs, seed: int = 42, kind: str = "uniform", alpha: float = -0.8,
               polarized: bool = False):
    """(ra, dec, fluxes): sources uniform on the sphere; ``uniform``: S ~ U(0.5, 2) Jy;
    ``gleam``: power-law counts S = 0.05 u^(-1/1.5) clipped at 100 Jy; ``diffuse``: |N(0,1)|.
    flux(f) = S (f / 150 MHz)^alpha.  ``polarized`` adds Q, U, V at <= 10% of I."""
    rng = np.random.default_rng(seed)
    ra = rng.uniform(0, 2 * np.pi, nsrc)
    dec = np.arcsin(rng.uniform(-1, 1, nsrc))
    if kind == "uniform":
        s = rng.uniform(0.5, 2.0, nsrc)
    elif kind == "gleam":
        s = np.minimum(0.05 * rng.uniform(1e-12, 1.0, nsrc) ** (-1 / 1.5), 100.0)
    elif kind == "diffuse":
        s = np.abs(rng.normal(size=nsrc))
    else:
        raise ValueError(kind)
    freqs = np.atleast_1d(np.asarray(freqs, dtype=float))
    flux = s[:, None] * (freqs[None, :] / 150e6) ** alpha
    if polarized:
        frac = rng.uniform(-0.1, 0.1, (nsrc, 1, 3))
        flux = np.concatenate([flux[..., None], flux[..., None] * frac], axis=-1)
    return ra, dec, flux


def synthetic_uvbeam(freqs, naz: int = 360, nza: int = 181, diameter: float = 14.0, seed: int = 0,
                     perturb: float = 0.0, include_endpoint: bool = False) -> UVBeamTable:
    """E-field az/za table ``(2, 2, Nf, nza, naz)``: Gaussian envelope of a ``diameter`` dish times
    the dipole projection of two orthogonal feeds, with a small complex sidelobe ripple (and an
    optional per-beam perturbation so that per-antenna beams differ)."""
    freqs = np.atleast_1d(np.asarray(freqs, dtype=float))
    rng = np.random.default_rng(seed)
    if include_endpoint:
        az = np.linspace(0, 2 * np.pi, naz)
    else:
        az = np.arange(naz) * (2 * np.pi / naz)
    za = np.linspace(0, np.pi, nza)
    A, Z = np.meshgrid(az, za)                      # (nza, naz)
    data = np.empty((2, 2, freqs.size, nza, naz), np.complex128)
    p = rng.normal(size=4) * perturb
    for fi, f in enumerate(freqs):
        sig = np.arcsin(min(1.0, 2.2150894 * (299792458.0 / f) / (np.pi * diameter))) * 2.0 / 2.355
        env = np.exp(-(Z**2) / (2 * sig * sig * (1 + p[0]))) + 0.02 * np.cos(6 * Z) * np.exp(-Z)
        ripple = 1.0 + 0.05j * np.sin(3 * A + p[1]) * np.sin(Z) + p[2] * 0.1 * np.cos(A) * np.sin(Z)
        e = env * ripple
        # vec 0 = az-hat, vec 1 = za-hat; feed 0 = x (east), feed 1 = y (north)
        data[0, 0, fi] = -np.sin(A) * e
        data[1, 0, fi] = np.cos(A) * np.cos(Z) * e
        data[0, 1, fi] = np.cos(A) * e * (1 + 0.03 * p[3])
        data[1, 1, fi] = np.sin(A) * np.cos(Z) * e * (1 + 0.03 * p[3])
    return UVBeamTable(data, az, za, freqs, "efield")
