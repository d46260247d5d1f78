"""Replaces the stub /root/reference/src/fftvis/gpu/utils.py:8-34."""
from fftvis_b200.gpu.utils import inplace_rot  # noqa: F401
