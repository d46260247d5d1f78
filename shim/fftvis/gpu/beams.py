"""Replaces the stub /root/reference/src/fftvis/gpu/beams.py:15-88."""
from fftvis_b200.gpu.beams import GPUBeamEvaluator  # noqa: F401
