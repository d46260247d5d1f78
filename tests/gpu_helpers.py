"""Shared helpers of the GPU parity tests (oracle = checker only)."""
import numpy as np


def relerr(a, b):
    a, b = np.ravel(np.asarray(a)), np.ravel(np.asarray(b))
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def hex_ants(n):
    from fftvis_b200 import synth
    return synth.hex_array(n)


def small_sky(nsrc, freqs, seed=42, polarized=False):
    from fftvis_b200 import synth
    return synth.random_sky(nsrc, freqs, seed=seed, polarized=polarized)


TIMES = np.array([2459845.0, 2459845.0 + 600 / 86400.0])
