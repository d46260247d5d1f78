"""Abstract engine interface (mirror of /root/reference/src/fftvis/core/simulate.py:22-221).

The reference's ABC signature is stale (it still takes a single ``beam``); the wrapper calls
engines with the CPU engine's keyword set (wrapper.py:308-336, cpu_simulate.py:537-569), which is
what ``GPUSimulationEngine`` implements.  The ABC therefore only fixes the two method names.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

default_accuracy_dict = {1: 6e-8, 2: 1e-13}   # reference core/simulate.py:16-19


class SimulationEngine(ABC):
    @abstractmethod
    def simulate(self, *args, **kwargs):  # pragma: no cover
        ...

    @abstractmethod
    def _evaluate_vis_chunk(self, *args, **kwargs):  # pragma: no cover
        ...
