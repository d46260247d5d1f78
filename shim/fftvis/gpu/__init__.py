"""GPU backend of fftvis, provided by fftvis_b200 (replaces /root/reference/src/fftvis/gpu/__init__.py)."""
from .beams import GPUBeamEvaluator  # noqa: F401
from .gpu_simulate import GPUSimulationEngine  # noqa: F401
