"""Host-side (numpy) planners of the fftvis hot path; see SURVEY.md section 8."""
from . import antenna_gridding, catalog, coords, utils  # noqa: F401
from .beam_basis import compute_beam_basis  # noqa: F401
