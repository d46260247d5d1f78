"""Replaces the stub /root/reference/src/fftvis/gpu/gpu_simulate.py:20-91."""
from fftvis_b200.gpu.gpu_simulate import GPUSimulationEngine  # noqa: F401
