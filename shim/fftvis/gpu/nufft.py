"""Replaces the stubs /root/reference/src/fftvis/gpu/nufft.py:11-98 (and adds the type-1 entry the CPU
backend has at cpu/nufft.py:120-175)."""
from fftvis_b200.gpu.nufft import gpu_nufft2d, gpu_nufft2d_type1, gpu_nufft3d  # noqa: F401
