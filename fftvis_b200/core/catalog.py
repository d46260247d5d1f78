"""Sky-model preparation (host, once per call).

Mirror of ``prepare_source_catalog`` (/root/reference/src/fftvis/cpu/utils.py:26-80): Stokes I
becomes ``0.5 I``; Stokes (I, Q, U, V) becomes the 2x2 coherency
``0.5 [[I+Q, U+iV], [U-iV, I-Q]]`` laid out ``(Nsrc, Nfreq, 2, 2)``.  Same error strings.
"""
from __future__ import annotations

import numpy as np


def prepare_source_catalog(sky_model: np.ndarray, polarized_beam: bool):
    sky_model = np.asarray(sky_model)
    if sky_model.ndim == 2:
        return 0.5 * sky_model, False
    if polarized_beam and sky_model.ndim == 3 and sky_model.shape[-1] == 4:
        i, q, u, v = (sky_model[..., k] for k in range(4))
        coh = np.empty(sky_model.shape[:2] + (2, 2), dtype=np.result_type(sky_model.dtype, np.complex64))
        coh[..., 0, 0] = i + q
        coh[..., 0, 1] = u + 1j * v
        coh[..., 1, 0] = u - 1j * v
        coh[..., 1, 1] = i - q
        coh *= 0.5
        return coh, True
    if polarized_beam:
        raise ValueError(
            f"polarized_beam=True requires sky_model to be either:\n"
            f"  2D unpolarized, or\n"
            f"  3D with last axis of length 4; "
            f"got ndim={sky_model.ndim}, shape={sky_model.shape}"
        )
    raise ValueError(
        f"polarized_beam=False requires sky_model to be 2D; "
        f"got ndim={sky_model.ndim}, shape={sky_model.shape}"
    )
