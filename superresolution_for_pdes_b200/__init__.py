"""B200-native drop-in for the batched Poisson sample-generation path of
tahmidawal/Superresolution_for_PDEs.

Host side (this package): the reference's Python entry points with unchanged signatures, RNG
order and array layouts -- ``PoissonSolver`` (src/data_generation.py), ``EnhancedPoissonSolver``
(src/enhanced_data_generation.py), ``solve_multi_resolution`` (src/resolution_comparison.py),
``solve_large_scale`` / ``solve_multi_level`` (src/subdomain_upscaling.py), ``generate_test_data``
(src/compare_test_cases.py), ``generate_out_of_sample_data`` (src/test_out_of_sample.py).
Device side: ``libpsn_b200.so`` (hand-written sm_100a CUDA behind a C ABI, include/psn_b200.h).
There is no CPU fallback.
"""
from . import _lib
from ._lib import PsnError, describe, set_option
from .compare_test_cases import generate_out_of_sample_data, generate_test_data
from .data_generation import PoissonSolver
from .dataset_tensors import PDEDataset
from .enhanced_data_generation import EnhancedPoissonSolver
from .resolution_comparison import (bilinear_multi_level_upscale, cubic_multi_level_upscale, solve_multi_resolution,
                                    solve_multi_resolution_batch, split_into_subdomains, stitch_subdomains)
from .subdomain_upscaling import solve_large_scale, solve_multi_level

__all__ = ["PoissonSolver", "EnhancedPoissonSolver", "PDEDataset", "solve_multi_resolution",
           "solve_multi_resolution_batch", "solve_large_scale", "solve_multi_level", "generate_test_data",
           "generate_out_of_sample_data", "bilinear_multi_level_upscale", "cubic_multi_level_upscale",
           "split_into_subdomains", "stitch_subdomains", "set_option", "describe", "PsnError", "_lib"]
