"""Caller-facing plumbing around the engine: the UVData array layout, results
streamed to disk in time blocks, and a hera_sim-shaped simulator object.

The reference returns one in-memory array ``(nfreqs, ntimes[, nfeed, nfeed], nbls)`` (cpu_simulate.py:850-854)
and leaves packing to its callers: hera_sim's ``VisibilitySimulator`` plugin fills a pyuvdata ``UVData`` object
from it (docs/tutorials/fftvis_tutorial.ipynb cell 2).  pyuvdata and hera_sim are absent here, so this module
produces the ARRAYS a ``UVData`` holds, in pyuvdata's conventions, without importing it:

* baseline-time axis time-major (all baselines of the first time, then the second, ...);
* ``data_array (Nblts, Nfreqs, Npols)``, polarisations ``[-5, -6, -7, -8]`` = xx, yy, xy, yx taken from the
  2 x 2 feed matrix ``[[xx, xy], [yx, yy]]``; unpolarised results are stored once as pseudo-Stokes I (pol 1);
* ``uvw_array = enu(ant_2) - enu(ant_1)`` in metres, ``ant_1_array`` / ``ant_2_array`` / ``time_array`` per blt.

``simulate_to_npy`` runs an observation in blocks of time steps and writes every finished block into a ``.npy``
file through ``numpy.lib.format.open_memmap``: device and page-locked host memory hold two blocks, not the
observation (BASELINE configs[4] is 65 GB), and the file has the reference's own axis order.
"""
from __future__ import annotations

import json
import threading
from pathlib import Path

import numpy as np

POL_NUMS_LINEAR = np.array([-5, -6, -7, -8])       # xx, yy, xy, yx (pyuvdata polarisation integers)
_FEED_INDEX = ((0, 0), (1, 1), (0, 1), (1, 0))     # where each of them sits in the 2 x 2 feed matrix


def uvdata_arrays(vis: np.ndarray, ants: dict, baselines, times, freqs) -> dict:
    """The arrays of a pyuvdata ``UVData`` object for a result of ``simulate_vis``.

    ``vis``: ``(nfreqs, ntimes, nbls)`` or ``(nfreqs, ntimes, 2, 2, nbls)``; ``baselines``: the ``(ant1, ant2)``
    pairs of the last axis; ``times``: Julian dates.  Returns ``data_array (Nblts, Nfreqs, Npols)``,
    ``polarization_array``, ``ant_1_array``, ``ant_2_array``, ``baseline_array`` (pyuvdata's ``2048 ant1 + ant2 + 2^16``
    numbering of arrays with <= 2048 antennas), ``time_array``, ``uvw_array``, ``freq_array``, ``Nblts``,
    ``Nbls``, ``Ntimes``, ``Nfreqs``, ``Npols``."""
    vis = np.asarray(vis)
    baselines = [tuple(b) for b in baselines]
    times = np.atleast_1d(np.asarray(getattr(times, "jd", times), dtype=np.float64))
    freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    nf, nt, nbl = vis.shape[0], vis.shape[1], vis.shape[-1]
    if nf != freqs.size or nt != times.size or nbl != len(baselines):
        raise ValueError("vis axes do not match freqs / times / baselines")
    if vis.ndim == 5:
        pols = POL_NUMS_LINEAR
        stack = np.stack([vis[:, :, i, j, :] for i, j in _FEED_INDEX], axis=-1)      # (nf, nt, nbl, 4)
    elif vis.ndim == 3:
        pols = np.array([1])
        stack = vis[..., None]
    else:
        raise ValueError("vis must be (nfreqs, ntimes, nbls) or (nfreqs, ntimes, 2, 2, nbls)")
    data = np.ascontiguousarray(np.transpose(stack, (1, 2, 0, 3))).reshape(nt * nbl, nf, pols.size)
    a1 = np.array([b[0] for b in baselines], dtype=np.int64)
    a2 = np.array([b[1] for b in baselines], dtype=np.int64)
    pos = {k: np.asarray(v, dtype=np.float64) for k, v in ants.items()}
    uvw = np.array([pos[b[1]] - pos[b[0]] for b in baselines], dtype=np.float64).reshape(nbl, 3)
    return dict(
        data_array=data, polarization_array=pols, ant_1_array=np.tile(a1, nt), ant_2_array=np.tile(a2, nt),
        baseline_array=np.tile(2048 * a1 + a2 + 2 ** 16, nt), time_array=np.repeat(times, nbl),
        uvw_array=np.tile(uvw, (nt, 1)), freq_array=freqs, Nblts=nt * nbl, Nbls=nbl, Ntimes=nt, Nfreqs=nf,
        Npols=int(pols.size))


def fill_uvdata(uvd, arrays: dict):
    """Set the arrays of ``uvdata_arrays`` on a ``UVData``-like object (attribute assignment only; the caller
    owns telescope metadata, flags and nsamples, which pyuvdata requires and fftvis does not produce)."""
    for k, v in arrays.items():
        setattr(uvd, k, v)
    n = arrays["data_array"].shape
    if getattr(uvd, "flag_array", None) is None or np.shape(uvd.flag_array) != n:
        uvd.flag_array = np.zeros(n, dtype=bool)
    if getattr(uvd, "nsample_array", None) is None or np.shape(uvd.nsample_array) != n:
        uvd.nsample_array = np.ones(n, dtype=np.float32)
    return uvd


def simulate_to_npy(path, ants, fluxes, ra, dec, freqs, times, beam, telescope_loc, time_block: int = 8,
                    engine=None, **simulate_kwargs) -> dict:
    """Run an observation in blocks of ``time_block`` time steps and stream each finished block into ``path``
    (a ``.npy`` file of shape ``(nfreqs, ntimes[, 2, 2], nbls)``, the reference's return layout).  The next block
    computes while the previous one is copied to page-locked memory and written by a background thread.
    Returns ``{"path", "shape", "dtype", "baselines"}``; a ``<path>.json`` side-car holds the same plus times and
    frequencies."""
    import torch
    from .gpu.gpu_simulate import GPUSimulationEngine
    eng = engine or GPUSimulationEngine()
    beam_list = beam if isinstance(beam, (list, tuple)) else [beam]
    drop = ("nprocesses", "nthreads", "force_use_ray", "trace_mem", "enable_memory_monitor")
    kw = {k: v for k, v in simulate_kwargs.items() if k not in drop}
    plan = eng.prepare(ants, freqs, fluxes, list(beam_list), ra, dec, times, telescope_loc, **kw)
    nf, nt, nbl = plan.nf_local, plan.ntimes, plan.nbls
    P = 4 if plan.polarized else 1
    cdt = np.complex64 if plan.precision == 1 else np.complex128
    shape = (nf, nt, 2, 2, nbl) if plan.polarized else (nf, nt, nbl)
    path = Path(path)
    mm = np.lib.format.open_memmap(path, mode="w+", dtype=cdt, shape=shape)
    flat = mm.reshape(nf, nt, P * nbl)
    tb = max(1, min(int(time_block), nt))
    dev = plan.device
    tdt = torch.complex64 if plan.precision == 1 else torch.complex128
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream()
        dbuf = [torch.empty((nf, tb, P, nbl), dtype=tdt, device=dev) for _ in range(2)]
        hbuf = [torch.empty((nf, tb, P, nbl), dtype=tdt, pin_memory=True) for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        writers = [None, None]

        def write(slot, t0, n):
            done[slot].synchronize()
            flat[:, t0:t0 + n, :] = hbuf[slot][:, :n].numpy().reshape(nf, n, P * nbl)

        for i, t0 in enumerate(range(0, nt, tb)):
            slot, n = i & 1, min(tb, nt - t0)
            if writers[slot] is not None:
                writers[slot].join()                         # the slot's previous block is on disk
            out = dbuf[slot][:, :n] if n == tb else torch.empty((nf, n, P, nbl), dtype=tdt, device=dev)
            eng.run_plan(plan, time_range=(t0, t0 + n), out=out)
            hbuf[slot][:, :n].copy_(out, non_blocking=True)
            done[slot].record(st)
            writers[slot] = threading.Thread(target=write, args=(slot, t0, n))
            writers[slot].start()
        for wth in writers:
            if wth is not None:
                wth.join()
        eng.check_source_buffer(plan)
    mm.flush()
    del mm
    bls = _plan_baselines(ants, simulate_kwargs.get("baselines"))
    meta = dict(path=str(path), shape=list(shape), dtype=np.dtype(cdt).name, baselines=[list(map(int, b)) for b in bls])
    side = dict(meta, times=[float(t) for t in np.atleast_1d(np.asarray(getattr(times, "jd", times), dtype=float))],
                freqs=[float(f) for f in np.atleast_1d(np.asarray(freqs, dtype=float))])
    Path(str(path) + ".json").write_text(json.dumps(side))
    return meta


def _plan_baselines(ants, baselines):
    if baselines is not None:
        return [tuple(b) for b in baselines]
    from .core import utils
    return utils.get_pos_reds(ants, include_autos=True, representatives_only=True)


class FFTVisB200:
    """A visibility simulator shaped like hera_sim's ``VisibilitySimulator`` plugins (``simulate(data_model)``
    returning the visibilities in the data model's UVData layout): what a hera_sim maintainer registers next
    to the CPU ``FFTVis`` wrapper.  ``data_model`` is duck-typed: ``ants`` (dict of ENU positions), ``freqs``,
    ``times`` (JD), ``ra`` / ``dec`` (radians), ``fluxes``, ``beams`` (list), ``telescope_loc`` and optionally
    ``baselines`` / ``beam_idx``.  Keyword arguments of the constructor go to ``simulate_vis``."""

    def __init__(self, precision: int = 2, polarized: bool = False, **simulate_kwargs):
        self.precision, self.polarized, self.kw = precision, polarized, dict(simulate_kwargs)

    def simulate(self, data_model) -> dict:
        from .wrapper import simulate_vis
        dm = data_model
        baselines = getattr(dm, "baselines", None)
        beams = list(dm.beams)
        vis = simulate_vis(ants=dm.ants, fluxes=dm.fluxes, ra=dm.ra, dec=dm.dec, freqs=dm.freqs, times=dm.times,
                           beam=beams if len(beams) > 1 else beams[0], telescope_loc=dm.telescope_loc,
                           baselines=baselines, beam_idx=getattr(dm, "beam_idx", None), precision=self.precision,
                           polarized=self.polarized, **self.kw)
        return uvdata_arrays(vis, dm.ants, _plan_baselines(dm.ants, baselines), dm.times, dm.freqs)
