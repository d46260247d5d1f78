"""fftvis_b200 -- B200-native (sm_100a) GPU backend for the fftvis visibility simulator.

Public names mirror /root/reference/src/fftvis/__init__.py:4-31 for the GPU path."""
from . import core, gpu  # noqa: F401
from .beam_models import AiryBeam, GaussianBeam, UniformBeam, UVBeamTable  # noqa: F401
from .core.beam_basis import compute_beam_basis  # noqa: F401
from .core.coords import HERA_LOCATION, TelescopeLocation  # noqa: F401
from .wrapper import create_beam_evaluator, create_simulation_engine, simulate_vis  # noqa: F401

__version__ = "0.1.0"
