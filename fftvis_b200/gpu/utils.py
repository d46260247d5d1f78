"""GPU ``inplace_rot`` (stub at /root/reference/src/fftvis/gpu/utils.py:8-34; CPU version
/root/reference/src/fftvis/cpu/utils.py:5-24)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def inplace_rot(rot: np.ndarray, b):
    """``b[:, s] <- rot @ b[:, s]`` in place, on the GPU.

    ``b`` is either a CUDA torch tensor of shape (3, n) (rotated where it lives) or a host numpy
    array (copied to the device, rotated by the CUDA kernel and copied back into ``b``)."""
    _lib.require_gpu()
    rot = np.ascontiguousarray(rot, dtype=np.float64)
    if rot.shape != (3, 3):
        raise ValueError("rot must be 3x3")
    host = None
    if isinstance(b, np.ndarray):
        host = b
        b = torch.as_tensor(np.ascontiguousarray(b)).cuda()
    if b.dim() != 2 or b.shape[0] != 3:
        raise ValueError("b must have shape (3, n)")
    if not b.is_contiguous():
        raise ValueError("b must be contiguous")
    prec = 1 if b.dtype == torch.float32 else 2
    if b.dtype not in (torch.float32, torch.float64):
        raise TypeError("b must be float32 or float64")
    _lib.check(_lib.lib().fv_inplace_rot(prec, _lib.doubles(rot.ravel()), b.data_ptr(), b.shape[1],
                                         torch.cuda.current_stream().cuda_stream), "fv_inplace_rot")
    if host is not None:
        host[...] = b.cpu().numpy()
