"""The reference-side shim files (shim/fftvis/gpu/*.py) import and export the
names the reference's wrapper looks up (/root/reference/src/fftvis/wrapper.py:45-46, 77-80)."""
import importlib
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_shim_modules_export_the_reference_names(monkeypatch):
    # a stand-in `fftvis` package whose `gpu` subpackage is the shim directory
    pkg = types.ModuleType("fftvis")
    pkg.__path__ = [str(ROOT / "shim" / "fftvis")]
    monkeypatch.setitem(sys.modules, "fftvis", pkg)
    for name in [m for m in sys.modules if m.startswith("fftvis.")]:
        monkeypatch.delitem(sys.modules, name)
    gpu = importlib.import_module("fftvis.gpu")
    sim = importlib.import_module("fftvis.gpu.gpu_simulate")
    beams = importlib.import_module("fftvis.gpu.beams")
    nufft = importlib.import_module("fftvis.gpu.nufft")
    utils = importlib.import_module("fftvis.gpu.utils")
    from fftvis_b200.core.beams import BeamEvaluator
    from fftvis_b200.core.simulate import SimulationEngine
    assert issubclass(sim.GPUSimulationEngine, SimulationEngine) and gpu.GPUSimulationEngine is sim.GPUSimulationEngine
    assert issubclass(beams.GPUBeamEvaluator, BeamEvaluator)
    for fn in ("gpu_nufft2d", "gpu_nufft3d", "gpu_nufft2d_type1"):
        assert callable(getattr(nufft, fn))
    assert callable(utils.inplace_rot)
    # the engine's simulate() takes the CPU engine's keyword set (cpu_simulate.py:537-569) and the chunk
    # evaluator the reference's positional/keyword names (cpu_simulate.py:856-884)
    import inspect
    sig = inspect.signature(sim.GPUSimulationEngine.simulate).parameters
    for kw in ("ants", "freqs", "fluxes", "beam_list", "ra", "dec", "times", "telescope_loc", "baselines", "beam_idx",
               "precision", "polarized", "eps", "upsample_factor", "beam_spline_opts", "flat_array_tol",
               "interpolation_function", "nprocesses", "nthreads", "coord_method", "coord_method_params",
               "force_use_ray", "force_use_type3", "trace_mem", "enable_memory_monitor", "nchunks", "source_buffer",
               "beam_coefs"):
        assert kw in sig, kw
    sig = inspect.signature(sim.GPUSimulationEngine._evaluate_vis_chunk).parameters
    for kw in ("time_idx", "freq_idx", "beam_list", "coord_mgr", "rotation_matrix", "antnums", "baselines", "bls",
               "freqs", "complex_dtype", "nfeeds", "beam_idx", "polarized", "polarized_sky_model", "eps",
               "upsample_factor", "beam_spline_opts", "interpolation_function", "n_threads", "is_coplanar",
               "use_type1", "basis_matrix", "type1_n_modes", "trace_mem", "nchunks", "beam_coefs"):
        assert kw in sig, kw
