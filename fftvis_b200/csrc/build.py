"""Build libfftvis_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m fftvis_b200.csrc.build [--force] [--ptxas-v]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
LIB = HERE.parent / "libfftvis_b200.so"
SOURCES = ["api.cu", "rotate_cut.cu", "weights.cu", "nufft.cu", "type1_fused.cu", "type1_small.cu", "type1_xdirect.cu", "type3.cu"]
# headers each translation unit includes (directly or through nufft_internal.cuh)
_NUFFT = ["common.cuh", "nufft_internal.cuh", "type1_fused.cuh"]
DEPS = {"api.cu": ["common.cuh"], "rotate_cut.cu": ["common.cuh"], "weights.cu": ["common.cuh"],
        "nufft.cu": _NUFFT, "type1_fused.cu": _NUFFT, "type1_small.cu": [*_NUFFT, "type1_small.cuh"],
        "type1_xdirect.cu": [*_NUFFT, "type1_xdirect.cuh"],
        "type3.cu": [*_NUFFT, "type3_tiles.cuh", "type3_fft.cuh"]}
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xcompiler", "-O3",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    objdir = HERE / "_obj"
    objdir.mkdir(exist_ok=True)
    jobs = []
    for src in SOURCES:
        obj = objdir / (src[:-3] + ".o")
        headers = [*(HERE / h for h in DEPS[src]), ROOT / "include" / "fftvis_b200.h"]
        if force or _stale(obj, [HERE / src, *headers]):
            cmd = [NVCC, *FLAGS, "-c", str(HERE / src), "-o", str(obj)]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append((src, cmd))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for (src, cmd), res in zip(jobs, ex.map(
                    lambda j: subprocess.run(j[1], capture_output=True, text=True), jobs)):
                if verbose or res.returncode:
                    sys.stderr.write(res.stderr)
                if res.returncode:
                    raise RuntimeError(f"nvcc failed on {src}:\n{res.stderr[-4000:]}")
    objs = [str(objdir / (s[:-3] + ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", str(LIB),
               "-lcufft", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            raise RuntimeError(f"link failed:\n{res.stderr[-4000:]}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--ptxas-v" in sys.argv))
