"""User-facing entry points: ``simulate_vis`` and the two factories, with the reference's
signatures (/root/reference/src/fftvis/wrapper.py:16-48, 51-82, 85-336).  Only ``backend="gpu"``
exists here -- this package *is* the GPU slot; ``backend="cpu"`` raises (no CPU fallback)."""
from __future__ import annotations

import numpy as np

from .beam_models import as_beam_model, prepare_beam_unpolarized
from .core.simulate import default_accuracy_dict
from .core.utils import get_desired_chunks, validate_beam_idx


def create_beam_evaluator(backend: str = "gpu", **kwargs):
    """reference wrapper.py:16-48 (where the gpu branch raises NotImplementedError)."""
    if backend == "gpu":
        from .gpu.beams import GPUBeamEvaluator
        return GPUBeamEvaluator(**kwargs)
    if backend == "cpu":
        raise NotImplementedError("fftvis_b200 provides the GPU backend only; use fftvis for backend='cpu'")
    raise ValueError(f"Unsupported backend: {backend}")


def create_simulation_engine(backend: str = "gpu", **kwargs):
    """reference wrapper.py:51-82."""
    if backend == "gpu":
        from .gpu.gpu_simulate import GPUSimulationEngine
        return GPUSimulationEngine(**kwargs)
    if backend == "cpu":
        raise NotImplementedError("fftvis_b200 provides the GPU backend only; use fftvis for backend='cpu'")
    raise ValueError(f"Unsupported backend: {backend}")


def _device_free_bytes() -> float:
    import torch
    if not torch.cuda.is_available():
        return np.inf
    # what this process can still take on an exclusively owned GPU: torch's cached blocks are reusable
    # (cudaMemGetInfo would count them as used, and costs ~15 ms per call on a 180 GB device)
    dev = torch.cuda.current_device()
    return float(torch.cuda.get_device_properties(dev).total_memory - torch.cuda.memory_allocated(dev))


def simulate_vis(ants, fluxes, ra, dec, freqs, times, beam, telescope_loc, beam_idx=None,
                 baselines=None, precision=2, polarized=False, eps=None, upsample_factor=2,
                 beam_spline_opts=None, use_feed="x", flat_array_tol=1e-6,
                 interpolation_function="az_za_map_coordinates", nprocesses=1, nthreads=None,
                 coord_method="CoordinateRotationERFA", coord_method_params=None,
                 force_use_type3=False, force_use_ray=False, trace_mem=False, backend="gpu",
                 max_memory=np.inf, min_chunks=1, source_buffer=1.0, beam_coefs=None):
    """Drop-in for ``fftvis.simulate_vis(..., backend="gpu")`` (reference wrapper.py:85-336):
    returns ``(nfreqs, ntimes, nbls)`` or ``(nfreqs, ntimes, 2, 2, nbls)`` complex64/128."""
    if eps is None:
        eps = default_accuracy_dict[precision]
    ants = {k: np.array(v) for k, v in ants.items()}
    _beam_list = beam if isinstance(beam, list) else [beam]
    nbeam, nant = len(_beam_list), len(ants)
    beam_idx = validate_beam_idx(beam_idx, beam_coefs, nbeam, nant)
    freqs = np.atleast_1d(np.asarray(freqs))
    beam_list = []
    for b in _beam_list:
        m = as_beam_model(b)
        if hasattr(m, "interp_freq") and m.Nfreqs > 1:
            m = m.interp_freq(freqs)                       # wrapper.py:261-271
        if not polarized and beam_coefs is None:
            m = prepare_beam_unpolarized(m, use_feed=use_feed)   # wrapper.py:278-279
        elif not polarized and beam_coefs is not None:
            raise ValueError(
                "Basis decomposition is not compatible with unpolarized simulations. "
                "Set polarized=True to use beam_coefs.")
        beam_list.append(m)
    nax = nfeed = 2 if polarized else 1
    # the memory that bounds the source-axis chunking is the GPU's, not the host's (wrapper.py:292-302)
    nchunks, _ = get_desired_chunks(min(max_memory, _device_free_bytes()), min_chunks, beam_list, nax,
                                    nfeed, nant, len(fluxes), precision, source_buffer=source_buffer)
    engine = create_simulation_engine(backend=backend)
    return engine.simulate(
        ants=ants, freqs=freqs, fluxes=fluxes, beam_list=beam_list, beam_idx=beam_idx, ra=ra, dec=dec,
        times=times, telescope_loc=telescope_loc, baselines=baselines, precision=precision,
        polarized=polarized, eps=eps, upsample_factor=upsample_factor, beam_spline_opts=beam_spline_opts,
        flat_array_tol=flat_array_tol, interpolation_function=interpolation_function,
        nprocesses=nprocesses, nthreads=nthreads, coord_method=coord_method,
        coord_method_params=coord_method_params, force_use_type3=force_use_type3,
        force_use_ray=force_use_ray, trace_mem=trace_mem, nchunks=nchunks, source_buffer=source_buffer,
        beam_coefs=beam_coefs)
