"""GPU beam evaluation: fills ``GPUBeamEvaluator`` (stub at
/root/reference/src/fftvis/gpu/beams.py:15-88) with the CPU evaluator's interface
(/root/reference/src/fftvis/cpu/beams.py:9-246): ``evaluate_beam``, ``prepare_beam_evaluation`` and
the four apparent-flux products.  The arithmetic is ``fv_weights`` (csrc/weights.cu).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ..beam_models import KIND_TABLE, UVBeamTable, as_beam_model
from ..core.beams import BeamEvaluator
from . import _lib

_RDT = {1: torch.float32, 2: torch.float64}
_CDT = {1: torch.complex64, 2: torch.complex128}


INTERPOLATION_FUNCTIONS = ("az_za_map_coordinates", "az_za_simple")


def resolve_interpolation(interpolation_function, spline_opts) -> int:
    """Spline order the device evaluates for the reference's ``interpolation_function`` +
    ``spline_opts`` (passed to pyuvdata at cpu/beams.py:69-74).

    ``az_za_map_coordinates``: ``spline_opts["order"]`` (``scipy.ndimage.map_coordinates``); 1 when
    absent.  ``az_za_simple``: ``RectBivariateSpline(kx, ky)``; ``kx = ky = 1`` is the same bilinear
    interpolant as order 1 (the identity the reference tests at tests/test_cpu_beams.py:15-87) and is
    evaluated as such, other degrees (FITPACK's not-a-knot splines) raise ``NotImplementedError``.
    Unknown names raise ``ValueError``; nothing is silently ignored."""
    opts = dict(spline_opts or {})
    if interpolation_function not in INTERPOLATION_FUNCTIONS:
        raise ValueError(f"unknown interpolation_function {interpolation_function!r}; "
                         f"expected one of {INTERPOLATION_FUNCTIONS}")
    if interpolation_function == "az_za_simple":
        kx, ky = int(opts.get("kx", 1)), int(opts.get("ky", 1))
        if (kx, ky) != (1, 1):
            raise NotImplementedError(
                "interpolation_function='az_za_simple' is evaluated on the GPU for kx = ky = 1 (bilinear) only; "
                f"got kx={kx}, ky={ky}")
        return 1
    return int(opts.get("order", 1))


class DeviceBeam:
    """A beam model uploaded to the GPU (owns the table tensor) + its ``fv_beam`` descriptor."""

    def __init__(self, beam, precision: int, order: int = 1, device="cuda"):
        beam = as_beam_model(beam)
        self.model = beam
        self.is_power = beam.beam_type == "power"
        self.desc = _lib.fv_beam()
        self.desc.kind = int(beam.kind)
        self.desc.is_power = int(self.is_power)
        self.desc.diameter = float(getattr(beam, "diameter", 0.0) or 0.0)
        self.desc.order = int(order)
        self.table = None
        if beam.kind == KIND_TABLE:
            self._upload_table(beam, precision, device)

    def _upload_table(self, beam: UVBeamTable, precision: int, device):
        if self.desc.order not in (0, 1, 3):
            raise NotImplementedError(
                "GPU beam interpolation supports spline order 0, 1 or 3 "
                f"(got order={self.desc.order}); pass beam_spline_opts={{'order': 1}} or {{'order': 3}}")
        az = np.asarray(beam.axis1_array, dtype=np.float64)
        za = np.asarray(beam.axis2_array, dtype=np.float64)
        daz, dza = az[1] - az[0], za[1] - za[0]
        data = np.asarray(beam.data_array)
        periodic = bool(np.isclose(az.size * daz, 2 * np.pi, rtol=1e-6))
        pad = 0
        order = int(self.desc.order)
        if periodic:   # grid covers 2 pi without its end point: extend by wrapping
            pad = max(2, order + 1)
            data = np.concatenate([data[..., -pad:], data, data[..., :pad]], axis=-1)
        spad = 0
        if order == 3:
            # scipy.ndimage.map_coordinates(order=3, mode="nearest") evaluates the cubic B-spline whose
            # coefficients are spline_filter(edge-pad(grid, 12)); the device evaluates the same
            # coefficients (4 x 4 taps), so the prefilter runs here once per table
            from scipy import ndimage
            spad = 12
            padw = [(0, 0)] * (data.ndim - 2) + [(spad, spad), (spad, spad)]
            padded = np.pad(data, padw, mode="edge")
            flat = padded.reshape((-1,) + padded.shape[-2:])
            coef = np.empty(flat.shape, dtype=np.complex128 if np.iscomplexobj(flat) else np.float64)
            for k in range(flat.shape[0]):
                if np.iscomplexobj(flat):
                    coef[k] = (ndimage.spline_filter(flat[k].real, order=3, output=np.float64, mode="nearest")
                               + 1j * ndimage.spline_filter(flat[k].imag, order=3, output=np.float64, mode="nearest"))
                else:
                    coef[k] = ndimage.spline_filter(flat[k], order=3, output=np.float64, mode="nearest")
            data = coef.reshape(padded.shape)
        # upload the table as it is and re-lay it out on the device (a host transpose of a 1024-frequency
        # E-field table is a multi-GB strided copy)
        dev_data = torch.as_tensor(data).to(device)
        if self.is_power:
            dt = _RDT[precision]
            tab = dev_data[0, 0]                                                   # (nf, nza, naz)
            tab = tab.real if tab.is_complex() else tab
        else:
            dt = _CDT[precision]
            nf = data.shape[2]
            tab = dev_data.permute(2, 0, 1, 3, 4).reshape(nf, 4, data.shape[3], data.shape[4])
        self.table = tab.to(dtype=dt).contiguous()
        del dev_data, tab
        host_shape = tuple(self.table.shape)
        self.nfreq_table = host_shape[0]
        d = self.desc
        d.table = self.table.data_ptr()
        d.nza, d.naz = int(host_shape[-2]), int(host_shape[-1])
        d.az_wrap_period = int(az.size) if periodic else 0
        d.az_pad = pad
        d.spline_pad = spad
        d.az0, d.daz, d.za0, d.dza = float(az[0]), float(daz), float(za[0]), float(dza)

    def descriptor(self, freq_offset: int) -> _lib.fv_beam:
        d = _lib.fv_beam()
        ctypes.memmove(ctypes.byref(d), ctypes.byref(self.desc), ctypes.sizeof(d))
        d.freq_offset = int(freq_offset) if (self.table is not None and self.nfreq_table > 1) else 0
        return d


def launch_weights(prec, mode, beam_i: DeviceBeam, beam_j: DeviceBeam, az, za, src_idx, n_dev, n_cap,
                   freqs_dev, f0, nf, flux, nsrc_total, out, out_beam=None, stream=None):
    """Thin wrapper over ``fv_weights`` on torch device tensors."""
    st = (stream or torch.cuda.current_stream()).cuda_stream
    bi, bj = beam_i.descriptor(f0), beam_j.descriptor(f0)
    _lib.check(_lib.lib().fv_weights(
        prec, mode, ctypes.byref(bi), ctypes.byref(bj), az.data_ptr(), za.data_ptr(),
        src_idx.data_ptr(), n_dev.data_ptr(), n_cap, freqs_dev.data_ptr(), nf, f0, flux.data_ptr(),
        nsrc_total, out.data_ptr(), out_beam.data_ptr() if out_beam is not None else None, st),
        "fv_weights")


class BeamTiles:
    """Sorted-by-beam-tile state of one time step's live set (``fv_tiles_*``, csrc/weights_tiled.cuh): owns the
    sort buffers; ``sort`` re-orders the live set in place, ``weights`` evaluates table beams from shared memory."""

    def __init__(self, stream=None):
        self._h = ctypes.c_void_p()
        st = (stream or torch.cuda.current_stream()).cuda_stream
        _lib.check(_lib.lib().fv_tiles_create(ctypes.byref(self._h), st), "fv_tiles_create")

    def __del__(self):
        try:
            if self._h:
                _lib.lib().fv_tiles_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def _array(beams, f0):
        arr = (_lib.fv_beam * len(beams))()
        for k, b in enumerate(beams):
            d = b.descriptor(f0)
            ctypes.memmove(ctypes.byref(arr[k]), ctypes.byref(d), ctypes.sizeof(d))
        return arr

    @classmethod
    def supported(cls, beams) -> bool:
        """All beams are az/za tables of order 0 / 1 on one grid (at most 6 staged together)."""
        if not beams or any(b.table is None for b in beams):
            return False
        return bool(_lib.lib().fv_tiles_supported(cls._array(beams, 0), len(beams)))

    def sort(self, prec, beam, xyz, az, za, src_idx, n_dev, n_cap):
        d = beam.descriptor(0)
        _lib.check(_lib.lib().fv_tiles_sort(self._h, prec, ctypes.byref(d), xyz.data_ptr(), az.data_ptr(),
                                            za.data_ptr(), src_idx.data_ptr(), n_dev.data_ptr(), n_cap), "fv_tiles_sort")

    def weights(self, prec, mode, beams, basis, az, za, src_idx, n_cap, freqs_dev, f0, nf, flux, nsrc_total, out):
        _lib.check(_lib.lib().fv_weights_tiled(
            self._h, prec, mode, self._array(beams, f0), len(beams), 1 if basis else 0, az.data_ptr(), za.data_ptr(),
            src_idx.data_ptr(), n_cap, freqs_dev.data_ptr(), nf, f0, flux.data_ptr(), nsrc_total, out.data_ptr()),
            "fv_weights_tiled")


def launch_weights_basis(prec, mode, beams, az, za, src_idx, n_dev, n_cap, freqs_dev, f0, nf, flux, nsrc_total,
                         out, stream=None):
    """``fv_weights_basis``: the K basis beams evaluated once per (source, frequency), all K (K + 1) / 2
    pair products written as the strengths ``out (nf, npairs * 4, n_cap)`` of one batched NUFFT."""
    st = (stream or torch.cuda.current_stream()).cuda_stream
    K = len(beams)
    arr = (_lib.fv_beam * K)()
    for k, b in enumerate(beams):
        d = b.descriptor(f0)
        ctypes.memmove(ctypes.byref(arr[k]), ctypes.byref(d), ctypes.sizeof(d))
    _lib.check(_lib.lib().fv_weights_basis(
        prec, mode, arr, K, az.data_ptr(), za.data_ptr(), src_idx.data_ptr(), n_dev.data_ptr(), n_cap,
        freqs_dev.data_ptr(), nf, f0, flux.data_ptr(), nsrc_total, out.data_ptr(), st), "fv_weights_basis")


def evaluate_beam_device(beam, az, za, polarized: bool, freq: float, prec: int = 2, order: int = 1) -> torch.Tensor:
    """One ``fv_weights`` launch: the response of ``beam`` at host directions (az, za) for one
    frequency, left on the device as ``(4, n)`` complex [vector component * 2 + feed] if
    ``polarized`` else ``(1, n)`` complex (power in the real part)."""
    _lib.require_gpu()
    model = as_beam_model(beam)
    if not polarized and model.beam_type != "power":
        model = model.to_power() if hasattr(model, "to_power") else model
    if isinstance(model, UVBeamTable) and model.Nfreqs > 1:
        fi = int(np.argmin(np.abs(np.asarray(model.freq_array) - freq)))
        model = UVBeamTable(model.data_array[:, :, fi:fi + 1], model.axis1_array,
                            model.axis2_array, np.atleast_1d(model.freq_array[fi]), model.beam_type)
    dbeam = DeviceBeam(model, prec, order)
    n = int(np.size(az))
    rdt, cdt = _RDT[prec], _CDT[prec]
    az_d = torch.as_tensor(np.ascontiguousarray(az)).to("cuda", rdt)
    za_d = torch.as_tensor(np.ascontiguousarray(za)).to("cuda", rdt)
    idx = torch.arange(n, dtype=torch.int32, device="cuda")
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    freqs = torch.tensor([float(freq)], dtype=torch.float64, device="cuda")
    P = 4 if polarized else 1
    flux = torch.ones((1, n), dtype=cdt, device="cuda")
    out = torch.empty((1, P, max(n, 1)), dtype=cdt, device="cuda")
    ob = torch.empty((1, P, max(n, 1)), dtype=cdt, device="cuda")
    if n:
        launch_weights(prec, 1 if polarized else 0, dbeam, dbeam, az_d, za_d, idx, n_dev, n, freqs, 0, 1,
                       flux, n, out, ob)
    return ob[0, :, :n]


class GPUBeamEvaluator(BeamEvaluator):
    """GPU implementation of the beam evaluator."""

    def evaluate_beam(self, beam, az, za, polarized, freq, check=False, spline_opts=None,
                      interpolation_function="az_za_map_coordinates"):
        """Beam response at (az, za) for one frequency: ``(2, 2, n)`` complex E-field
        [vector component, feed, source] if ``polarized`` else ``(n,)`` power.  Host arrays in and
        out like the CPU evaluator (cpu/beams.py:12-89); ``interpolation_function`` and
        ``spline_opts`` are resolved by ``resolve_interpolation`` (unsupported choices raise)."""
        _lib.require_gpu()
        self.polarized = polarized
        self.freq = freq
        self.spline_opts = spline_opts or {}
        order = resolve_interpolation(interpolation_function, self.spline_opts)
        az = np.asarray(az)
        prec = 1 if az.dtype == np.float32 else 2
        n = az.size
        ob = evaluate_beam_device(beam, az, za, polarized, freq, prec, order)
        res = ob.cpu().numpy()
        interp_beam = res.reshape(2, 2, n) if polarized else res[0].real.astype(az.dtype)
        if check:
            sm = np.sum(interp_beam)
            if np.isinf(sm) or np.isnan(sm):
                raise ValueError("Beam interpolation resulted in an invalid value")
        return interp_beam

    @staticmethod
    def prepare_beam_evaluation(antnums, baselines, beam_idx):
        """Unique (bi <= bj) beam pairs, per-pair baseline index lists and flip flags
        (cpu/beams.py:91-127), computed with array operations."""
        nbl = len(baselines)
        if beam_idx is None:
            return [(0, 0)], {(0, 0): np.arange(nbl)}, {(0, 0): [False] * nbl}
        beam_idx = np.asarray(beam_idx)
        ub = np.unique(beam_idx)
        pairs = [(ub[i], ub[j]) for i in range(len(ub)) for j in range(i, len(ub))]
        lookup = {a: b for a, b in zip(antnums, beam_idx)}
        bl = np.asarray(baselines).reshape(nbl, 2)
        try:
            b1 = np.array([lookup[a] for a in bl[:, 0]])
            b2 = np.array([lookup[a] for a in bl[:, 1]])
        except KeyError as e:  # pragma: no cover
            raise ValueError("Beam pair not in beam pair list") from e
        flipped = b1 > b2
        lo, hi = np.minimum(b1, b2), np.maximum(b1, b2)
        to_bls = {p: [] for p in pairs}
        to_flip = {p: [] for p in pairs}
        for p in pairs:
            sel = np.nonzero((lo == p[0]) & (hi == p[1]))[0]
            to_bls[p] = sel.tolist()
            to_flip[p] = flipped[sel].tolist()
        return pairs, to_bls, to_flip

    # ---- the four products (host arrays in/out; each is one fv_weights launch) -----------------
    @staticmethod
    def _product(mode, beam_i, beam_j, flux_or_coh):
        """One ``fv_coherency`` launch on host arrays: (2, 2, n) beams -> (2, 2, n) product."""
        _lib.require_gpu()
        beam_i = np.asarray(beam_i)
        prec = 1 if beam_i.dtype == np.complex64 else 2
        cdt = _CDT[prec]
        n = beam_i.shape[-1]
        bi = torch.as_tensor(np.ascontiguousarray(beam_i.reshape(4, n))).to("cuda", cdt)
        bj = torch.as_tensor(np.ascontiguousarray(np.asarray(beam_j).reshape(4, n))).to("cuda", cdt)
        fc = np.asarray(flux_or_coh)
        fc = torch.as_tensor(np.ascontiguousarray(fc.reshape(-1, n) if mode != 1 else fc)).to("cuda", cdt)
        out = torch.empty((4, n), dtype=cdt, device="cuda")
        _lib.check(_lib.lib().fv_coherency(prec, mode, bi.data_ptr(), bj.data_ptr(), fc.data_ptr(), n,
                                           out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "fv_coherency")
        return out.reshape(2, 2, n).cpu().numpy()

    def get_apparent_flux_polarized_beam(self, beam, flux):
        """In place ``beam <- A^H diag(F) A`` (cpu/beams.py:129-145)."""
        beam[...] = self._product(1, beam, beam, flux)

    def get_apparent_flux_polarized(self, beam, coherency):
        """In place ``beam <- A^H C A`` (cpu/beams.py:147-180)."""
        beam[...] = self._product(4, beam, beam, coherency)

    def get_apparent_flux_polarized_beam_pair(self, beam_i, beam_j, flux, out):
        """``out <- A_i^H diag(F) A_j`` (cpu/beams.py:182-212)."""
        out[...] = self._product(1, beam_i, beam_j, flux)

    def get_apparent_flux_polarized_pair(self, beam_i, beam_j, coherency, out):
        """``out <- A_i^H C A_j`` (cpu/beams.py:215-246)."""
        out[...] = self._product(4, beam_i, beam_j, coherency)
