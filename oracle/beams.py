"""TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of beam evaluation and of the reference's
four coherency kernels.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import it.

Follows
  * /root/reference/src/fftvis/cpu/beams.py:12-89   (evaluate_beam: (2,2,n) E-field or (n,) power)
  * /root/reference/src/fftvis/cpu/beams.py:129-246 (the four numba products)
  * /root/reference/src/fftvis/cpu/cpu_simulate.py:90-202 (_compute_apparent_coherency branches)
The beam arithmetic itself lives in pyuvdata >= 3.1.2 (third party, absent): analytic formulas
and the az/za ``map_coordinates`` interpolation are restated from its documented behaviour
 -- PARITY UNPINNED for that boundary.  The coherency products are
pinned by the reference's own einsum identities (tests/test_cpu_beams.py:102,606,870,953).
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage, special

C_LIGHT = 299792458.0


def gaussian_sigma(diameter, freq):
    return np.arcsin(2.2150894 * (C_LIGHT / freq) / (np.pi * diameter)) * 2.0 / 2.355


def analytic_efield_amplitude(kind: int, diameter: float, za, freq: float):
    za = np.asarray(za, dtype=np.float64)
    if kind == 0:       # Gaussian
        s = gaussian_sigma(diameter, freq)
        return np.exp(-(za**2) / (2.0 * s * s))
    if kind == 1:       # Airy
        x = np.pi * diameter * np.sin(za) * freq / C_LIGHT
        out = np.ones_like(x)
        nz = x != 0
        out[nz] = 2.0 * special.j1(x[nz]) / x[nz]
        return out
    if kind == 2:
        return np.ones_like(za)
    raise ValueError(kind)


def table_fractional_index(table, az, za):
    """Fractional (za, az) pixel coordinates and whether azimuth wraps without an end point."""
    a1 = np.asarray(table.axis1_array, dtype=np.float64)
    a2 = np.asarray(table.axis2_array, dtype=np.float64)
    daz = a1[1] - a1[0]
    dza = a2[1] - a2[0]
    wraps = bool(np.isclose(a1.size * daz, 2 * np.pi, rtol=1e-6))
    az_i = (np.asarray(az, np.float64) - a1[0]) / daz
    za_i = (np.asarray(za, np.float64) - a2[0]) / dza
    return za_i, az_i, wraps


def interp_table(table, az, za, freq_index: int, order: int = 1):
    """(ncomp_vec, ncomp_feed, n) values of a tabulated beam, real and imaginary parts
    interpolated separately with ``scipy.ndimage.map_coordinates`` (mode 'nearest'), the azimuth
    axis extended by wrapping when the grid covers 2 pi without its end point."""
    za_i, az_i, wraps = table_fractional_index(table, az, za)
    data = table.data_array[:, :, freq_index]
    if wraps:
        npad = max(1, order + 1)
        data = np.concatenate([data[..., -npad:], data, data[..., :npad]], axis=-1)
        az_i = np.mod(az_i, table.axis1_array.size) + npad
    out = np.empty(data.shape[:2] + (az_i.size,), dtype=np.result_type(data.dtype, np.float64))
    for a in range(data.shape[0]):
        for f in range(data.shape[1]):
            d = data[a, f]
            if np.iscomplexobj(d):
                re = ndimage.map_coordinates(d.real, [za_i, az_i], order=order, mode="nearest")
                im = ndimage.map_coordinates(d.imag, [za_i, az_i], order=order, mode="nearest")
                out[a, f] = re + 1j * im
            else:
                out[a, f] = ndimage.map_coordinates(d, [za_i, az_i], order=order, mode="nearest")
    return out


def evaluate_beam(beam, az, za, polarized: bool, freq: float, freq_index: int = 0,
                  spline_opts=None, check: bool = False):
    """(2, 2, n) complex E-field [vector comp, feed, src] if polarized else (n,) power."""
    order = int((spline_opts or {}).get("order", 1))
    if beam.kind == 3:
        resp = interp_table(beam, az, za, freq_index, order)
        out = resp if polarized else resp[0, 0]
    else:
        e = analytic_efield_amplitude(beam.kind, beam.diameter, za, freq)
        if polarized:
            out = np.broadcast_to((e / np.sqrt(2.0))[None, None, :], (2, 2, e.size)).astype(complex)
        else:
            out = e * e
    if check:
        sm = np.sum(out)
        if np.isinf(sm) or np.isnan(sm):
            raise ValueError("Beam interpolation resulted in an invalid value")
    return out


# ---- the four products of cpu/beams.py:129-246 --------------------------------------------
def apparent_polarized_beam(beam, flux):
    """A^H diag(F) A, out[a,p,s] = sum_b conj(A[b,a,s]) A[b,p,s] F[s]   (cpu/beams.py:129-145)."""
    return np.einsum("bas,s,bps->aps", beam.conj(), flux, beam)


def apparent_polarized(beam, coherency):
    """A^H C A   (cpu/beams.py:147-180)."""
    return np.einsum("kin,kmn,mjn->ijn", beam.conj(), coherency, beam)


def apparent_polarized_beam_pair(beam_i, beam_j, flux):
    """A_i^H diag(F) A_j   (cpu/beams.py:182-212)."""
    return np.einsum("bas,s,bps->aps", beam_i.conj(), flux, beam_j)


def apparent_polarized_pair(beam_i, beam_j, coherency):
    """A_i^H C A_j   (cpu/beams.py:215-246)."""
    return np.einsum("bas,bks,kps->aps", beam_i.conj(), coherency, beam_j)


def compute_apparent_coherency(beam_evals, bi, bj, flux_f, polarized, polarized_sky_model):
    """(nfeeds^2, n) NUFFT strengths for beam pair (bi, bj) at one frequency;
    ``flux_f`` = flux[:, f] (n,) or (n, 2, 2).  Branches of cpu_simulate.py:138-191, including
    the axis-0 flip that is applied on the polarised-sky branch only."""
    if polarized and polarized_sky_model:
        coh = np.transpose(flux_f, (1, 2, 0))
        out = apparent_polarized_pair(np.flip(beam_evals[bi], 0), np.flip(beam_evals[bj], 0), coh)
        # (the same-beam branch writes through a flipped *view* of the work buffer and reshapes
        # that view, cpu_simulate.py:152-156,191: element order = view order, i.e. the same
        # A'^H C A' with A' the flipped beam.)
        return out.reshape(4, -1)
    if polarized:
        return apparent_polarized_beam_pair(beam_evals[bi], beam_evals[bj], flux_f).reshape(4, -1)
    # beams are cast to the complex dtype before the product (cpu_simulate.py:84-86), so the
    # square root is the complex principal root
    prod = np.asarray(beam_evals[bi], dtype=complex) * np.asarray(beam_evals[bj], dtype=complex)
    return (np.sqrt(prod) * flux_f)[None, :]
