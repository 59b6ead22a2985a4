"""Mirror of pde_superresolution/analysis.py: the functions live in evaluation.py (torch reductions over
(sample, time, x) arrays instead of xarray objects)."""
from .evaluation import (calculate_survival, is_good, mostly_good, mostly_good_survival,  # noqa: F401
                         unify_x_coords)
from .duckarray import resample_mean  # noqa: F401
