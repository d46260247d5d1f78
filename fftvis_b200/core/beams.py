"""Abstract beam evaluator (mirror of /root/reference/src/fftvis/core/beams.py:10-139).

The reference derives from ``matvis.core.beams.BeamInterpolator``; matvis is absent, so the
attributes that base class carries are set here directly (asserted by the reference's
tests/test_gpu_beams.py:7-19 and tests/test_beam_evaluator.py).
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from .coords import enu_to_az_za


class BeamEvaluator(ABC):
    def __init__(self, **kwargs):
        self.beam_list = []
        self.beam_idx = None
        self.polarized = False
        self.nant = 0
        self.freq = 0.0
        self.nsrc = 0
        self.spline_opts = {}
        self.precision = 2
        for k, v in kwargs.items():
            setattr(self, k, v)

    @abstractmethod
    def evaluate_beam(self, beam, az, za, polarized, freq, check=False, spline_opts=None,
                      interpolation_function="az_za_map_coordinates"):  # pragma: no cover
        ...

    @abstractmethod
    def get_apparent_flux_polarized(self, beam, flux):  # pragma: no cover
        ...

    def interp(self, tx: np.ndarray, ty: np.ndarray, out: np.ndarray) -> np.ndarray:
        """matvis ``BeamInterpolator.interp`` bridge (reference core/beams.py:106-139)."""
        az, za = enu_to_az_za(enu_e=tx, enu_n=ty, orientation="uvbeam")
        self.nsrc = len(az)
        for i, bm in enumerate(self.beam_list):
            vals = self.evaluate_beam(bm, az, za, self.polarized, self.freq,
                                      spline_opts=self.spline_opts,
                                      interpolation_function="az_za_map_coordinates")
            out[i] = vals.transpose((1, 0, 2)) if (self.polarized and vals.ndim == 3) else vals
        return out
