"""Beam descriptions the GPU engine understands (parameter containers only -- the arithmetic is
in csrc/weights.cuh; the CPU restatement used by tests is oracle/beams.py).

The reference takes pyuvdata objects (``UVBeam``, ``AnalyticBeam``, ``BeamInterface``;
/root/reference/src/fftvis/wrapper.py:247-285) and evaluates them through
``BeamInterface.compute_response`` (/root/reference/src/fftvis/cpu/beams.py:69-81).  pyuvdata is
absent from this image, so the engine works on these light-weight equivalents;
``as_beam_model`` converts the pyuvdata objects by duck typing when they are present.

Analytic definitions (pyuvdata >= 3.1, recalled in SURVEY.md Appendix B.3):
  Airy      E = 2 J1(x)/x,  x = pi D sin(za) f / c,            power = E^2
  Gaussian  E = exp(-za^2 / (2 s^2)), s = asin(2.2150894 c/(f pi D)) * 2/2.355, power = E^2
  E-field response of an unpolarised analytic beam: E/sqrt(2) in every (vector, feed) slot.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

KIND_GAUSSIAN = 0
KIND_AIRY = 1
KIND_UNIFORM = 2
KIND_TABLE = 3


@dataclass
class AnalyticBeam:
    kind: int
    diameter: float = 14.0
    beam_type: str = "efield"     # "efield" (2x2 Jones) or "power" (scalar, one pol)

    def to_power(self) -> "AnalyticBeam":
        return type(self)(diameter=self.diameter, beam_type="power")


@dataclass
class GaussianBeam(AnalyticBeam):
    kind: int = KIND_GAUSSIAN


@dataclass
class AiryBeam(AnalyticBeam):
    kind: int = KIND_AIRY


@dataclass
class UniformBeam(AnalyticBeam):
    kind: int = KIND_UNIFORM


@dataclass
class UVBeamTable:
    """Tabulated az/za beam: the subset of ``pyuvdata.UVBeam`` the hot path reads.

    data_array : (Naxes_vec, Nfeeds, Nfreqs, Nza, Naz) complex E-field, or
                 (1, Npols, Nfreqs, Nza, Naz) real power.
    axis1_array: azimuth grid (radians, uniform);  axis2_array: zenith-angle grid.
    """

    data_array: np.ndarray
    axis1_array: np.ndarray
    axis2_array: np.ndarray
    freq_array: np.ndarray
    beam_type: str = "efield"
    kind: int = field(default=KIND_TABLE, init=False)

    @property
    def Nfreqs(self) -> int:
        return int(np.size(self.freq_array))

    def interp_freq(self, freqs: np.ndarray) -> "UVBeamTable":
        """Linear interpolation of the table onto ``freqs`` (what the reference's wrapper does
        once up front with ``UVBeam.interp(freq_array=freqs)``, wrapper.py:261-271)."""
        freqs = np.atleast_1d(np.asarray(freqs, dtype=float))
        f0 = np.asarray(self.freq_array, dtype=float)
        if f0.size == freqs.size and np.allclose(f0, freqs, rtol=1e-12, atol=0):
            return self
        if f0.size == 1:
            data = np.repeat(self.data_array, freqs.size, axis=2)
        else:
            hi = np.clip(np.searchsorted(f0, freqs), 1, f0.size - 1)
            lo = hi - 1
            t = ((freqs - f0[lo]) / (f0[hi] - f0[lo]))[None, None, :, None, None]
            data = (1 - t) * self.data_array[:, :, lo] + t * self.data_array[:, :, hi]
        return UVBeamTable(data, self.axis1_array, self.axis2_array, freqs, self.beam_type)

    def to_power(self, use_feed: str = "x") -> "UVBeamTable":
        """E-field -> single-pol power table (``matvis.prepare_beam_unpolarized``,
        reference wrapper.py:278-279): |E_0f|^2 + |E_1f|^2 for the chosen feed."""
        if self.beam_type == "power":
            if self.data_array.shape[1] == 1:
                return self
            ip = 0 if use_feed in ("x", "e", 0) else 1
            return UVBeamTable(self.data_array[:1, ip:ip + 1].real.copy(), self.axis1_array,
                               self.axis2_array, self.freq_array, "power")
        ifeed = 0 if use_feed in ("x", "e", 0) else 1
        p = (np.abs(self.data_array[:, ifeed]) ** 2).sum(axis=0)
        return UVBeamTable(p[None, None], self.axis1_array, self.axis2_array,
                           self.freq_array, "power")


def as_beam_model(beam):
    """Normalise user input (our models, or pyuvdata objects by duck typing)."""
    if isinstance(beam, (AnalyticBeam, UVBeamTable)):
        return beam
    inner = getattr(beam, "beam", None)
    if inner is not None and inner is not beam:          # pyuvdata BeamInterface
        model = as_beam_model(inner)
        bt = getattr(beam, "beam_type", None)
        if bt == "power" and model.beam_type != "power":
            model = model.to_power()
        return model
    if hasattr(beam, "data_array") and hasattr(beam, "axis1_array"):   # pyuvdata UVBeam
        return UVBeamTable(np.asarray(beam.data_array), np.asarray(beam.axis1_array),
                           np.asarray(beam.axis2_array), np.asarray(beam.freq_array).ravel(),
                           getattr(beam, "beam_type", "efield"))
    name = type(beam).__name__
    if name == "AiryBeam":
        return AiryBeam(diameter=float(beam.diameter))
    if name == "GaussianBeam" and getattr(beam, "diameter", None) is not None:
        return GaussianBeam(diameter=float(beam.diameter))
    if name == "UniformBeam":
        return UniformBeam()
    raise TypeError(f"fftvis_b200 cannot interpret beam object of type {name}")


def prepare_beam_unpolarized(beam, use_feed: str = "x"):
    """Power, single-polarisation version of ``beam`` (reference wrapper.py:278-279)."""
    beam = as_beam_model(beam)
    if isinstance(beam, UVBeamTable):
        return beam.to_power(use_feed)
    return beam.to_power()
