"""Beam basis decomposition via SVD, on the GPU.

Mirror of ``compute_beam_basis`` (/root/reference/src/fftvis/core/beam_basis.py:17-154): every
input beam is evaluated on a common azimuth / zenith-angle grid at one frequency, the flattened
responses are decomposed by SVD, and the retained right-singular vectors come back as eigenbeams
with the per-input-beam coefficients ``U[:, :K] * s[:K]`` -- the ``beam`` list and ``beam_coefs``
of the basis-visibility path (``simulate_vis(..., beam_coefs=...)``, cpu_simulate.py:303-470).
Same arguments, error strings and return shapes.

B200 form: the beams are evaluated by ``fv_weights`` (one launch per beam, the responses never
leave the device) and the SVD of the wide ``(n_beams, n_pixels)`` matrix is done on the device as
a Householder QR of its tall transpose followed by the SVD of the small triangular factor --
backward stable, unlike the Gram-matrix shortcut, so singular values keep their full relative
accuracy down to the 1e-12 default threshold.  Eigenbeams are returned as ``UVBeamTable`` (the
subset of ``pyuvdata.UVBeam`` the hot path reads) instead of ``UVBeam`` copies.
"""
from __future__ import annotations

import logging

import numpy as np

from ..beam_models import UVBeamTable, as_beam_model, prepare_beam_unpolarized

logger = logging.getLogger(__name__)


def compute_beam_basis(beam_list, freq: float, polarized: bool, threshold: float = 1e-12,
                       axis1_array=None, axis2_array=None, n_axis1: int = 361, n_axis2: int = 181):
    """SVD beam basis of a collection of antenna beams (reference core/beam_basis.py:17-154).

    Returns ``(eigenbeams, beam_coefs)``: a list of K ``UVBeamTable`` on the common grid and an
    ``(n_beams, K)`` array with ``beam_i = sum_k beam_coefs[i, k] * eigenbeam_k``.
    """
    if len(beam_list) == 0:
        raise ValueError("beam_list must contain at least one beam.")
    if not (0.0 < threshold <= 1.0):
        raise ValueError("threshold must be in the interval (0, 1].")
    freq_grid = np.atleast_1d(freq).astype(float)
    if freq_grid.size != 1:
        raise ValueError("compute_beam_basis currently expects a scalar freq.")

    models = []
    for beam in beam_list:
        m = as_beam_model(beam)
        if polarized:
            if m.beam_type != "efield":
                raise ValueError("polarized=True requires efield beams.")
        else:
            m = prepare_beam_unpolarized(m)
        models.append(m)

    if (axis1_array is None) != (axis2_array is None):
        raise ValueError("axis1_array and axis2_array must be supplied together.")
    if axis1_array is None:
        for m in models:
            if isinstance(m, UVBeamTable):
                axis1_array, axis2_array = m.axis1_array, m.axis2_array
                break
        else:
            axis1_array = np.linspace(0.0, 2.0 * np.pi, n_axis1)
            axis2_array = np.linspace(0.0, np.pi, n_axis2)
    axis1_array = np.asarray(axis1_array, dtype=float)
    axis2_array = np.asarray(axis2_array, dtype=float)

    import torch

    from ..gpu.beams import evaluate_beam_device

    naz, nza = axis1_array.size, axis2_array.size
    az = np.tile(axis1_array, nza)                      # az fastest: data_array[..., za, az]
    za = np.repeat(axis2_array, naz)
    rows = []
    for m in models:
        if isinstance(m, UVBeamTable) and m.Nfreqs > 1:
            m = m.interp_freq(freq_grid)
        resp = evaluate_beam_device(m, az, za, polarized, float(freq_grid[0]), prec=2, order=1)
        rows.append(resp.reshape(-1) if polarized else resp[0].real.contiguous())
    flat = torch.stack(rows, dim=0)                     # (n_beams, n_comp * nza * naz)
    if flat.is_complex() and float(flat.imag.abs().max()) == 0.0:
        flat = flat.real.contiguous()                   # real responses: real singular vectors
    slice_shape = (2, 2, nza, naz) if polarized else (1, 1, nza, naz)

    # SVD of the wide matrix:  flat^H = Q R  (tall Householder QR),  R^H = U s W^H
    #   =>  flat = U s (Q W)^H
    q, r = torch.linalg.qr(flat.mH, mode="reduced")
    u, s, wh = torch.linalg.svd(r.mH, full_matrices=False)
    vh = wh @ q.mH
    s_norm = s / s[0]
    K = int((s_norm >= threshold).sum().item())
    beam_coefs = (u[:, :K] * s[:K][None, :]).cpu().numpy()
    vh_host = vh[:K].cpu().numpy()

    beam_type = "efield" if polarized else "power"
    eigenbeams = []
    for k in range(K):
        data = vh_host[k].reshape(slice_shape)[:, :, np.newaxis]
        eigenbeams.append(UVBeamTable(np.ascontiguousarray(data), axis1_array.copy(), axis2_array.copy(),
                                      freq_grid.copy(), beam_type))
    return eigenbeams, beam_coefs
