"""GPU backend: the names the reference's stub package exports
(/root/reference/src/fftvis/gpu/__init__.py:3-5) plus ``gpu_nufft2d_type1`` and ``inplace_rot``."""
from .beams import GPUBeamEvaluator
from .gpu_simulate import GPUSimulationEngine, SimulationPlan
from .nufft import gpu_nufft2d, gpu_nufft2d_type1, gpu_nufft3d
from .utils import inplace_rot

__all__ = ["GPUBeamEvaluator", "GPUSimulationEngine", "SimulationPlan", "gpu_nufft2d",
           "gpu_nufft3d", "gpu_nufft2d_type1", "inplace_rot"]
