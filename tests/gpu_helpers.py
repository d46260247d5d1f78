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


F32_EPS = 6e-8


def f32_bar(cpu, want, eps=F32_EPS, slack=1.5):
    """fp32 parity bar: north_star's 10 x eps, or -- where the rounding of the fp32 *inputs* (phases of
    tens of radians) already exceeds that -- ``slack`` x the error the CPU restatement makes in the same
    precision on the same inputs.  No constant floor: a regression against the CPU path fails."""
    return max(10 * eps, slack * relerr(cpu, want))
