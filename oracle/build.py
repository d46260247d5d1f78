"""TEST INFRASTRUCTURE ONLY -- builds the CPU oracle library (oracle/_build/libfv_oracle.so).

The reference itself is pure Python over third-party natives (finufft, matvis, pyuvdata,
erfa) that are absent from this image, so there is nothing under /root/reference that could
be compiled into ``oracle/_ref``; the oracle is therefore the "port" kind.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_build" / "libfv_oracle.so"
SRC = HERE / "fv_oracle.cpp"


def build(force: bool = False) -> Path:
    OUT.parent.mkdir(exist_ok=True)
    if not force and OUT.exists() and OUT.stat().st_mtime >= SRC.stat().st_mtime:
        return OUT
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-std=c++17",
           str(SRC), "-o", str(OUT)]
    try:
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    except subprocess.CalledProcessError as e:  # older CPUs: retry without -march
        cmd.remove("-march=x86-64-v3")
        try:
            subprocess.run(cmd, check=True, capture_output=True, text=True)
        except subprocess.CalledProcessError as e2:
            sys.stderr.write(e.stderr + "\n" + e2.stderr)
            raise
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
