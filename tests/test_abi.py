"""The C-ABI shared library loads and exports every symbol include/fftvis_b200.h declares (no
compute calls: this runs without a GPU), and the product never routes through the oracle."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "fftvis_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from fftvis_b200.gpu import _lib
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == names, (set(names) ^ set(_lib.EXPORTS))


def test_host_only_entry_points_work_without_gpu():
    from fftvis_b200.gpu import _lib
    lib = _lib.lib()
    w, beta = ctypes.c_int(0), ctypes.c_double(0)
    assert lib.fv_kernel_params(6e-8, 2.0, 1, ctypes.byref(w), ctypes.byref(beta)) == 0 and w.value == 9
    assert lib.fv_kernel_params(1e-13, 2.0, 2, ctypes.byref(w), ctypes.byref(beta)) == 0 and w.value == 14
    assert lib.fv_next235even(930) == 960 and lib.fv_next235even(82) == 90
    assert lib.fv_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fftvis_b200.gpu import _lib
    with pytest.raises(_lib.FVError, match="no CPU fallback"):
        _lib.require_gpu()
    import numpy as np
    from fftvis_b200.gpu import gpu_nufft2d
    with pytest.raises(_lib.FVError):
        gpu_nufft2d(np.zeros(3), np.zeros(3), np.ones(3, complex), np.zeros(2), np.zeros(2), 1e-9)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "fftvis_b200").rglob("*.py"):
        src = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
    for path in (ROOT / "fftvis_b200" / "csrc").glob("*.cu*"):
        assert "oracle/" not in path.read_text(), path


def test_oracle_never_imports_the_product():
    """The checker stands on its own: no module under oracle/ imports fftvis_b200 (beam containers and synthetic
    workloads reach it as arguments, duck-typed)."""
    for path in (ROOT / "oracle").glob("*.py"):
        src = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+fftvis_b200\b", src, flags=re.M), path
