"""Drop-in for the data-generating functions of the reference's comparison harnesses:
``generate_test_data`` (src/compare_test_cases.py:12-79) and ``generate_out_of_sample_data``
(src/test_out_of_sample.py:13-35).  Model evaluation, metrics and plots of those files are out of scope.
"""
from __future__ import annotations

from pathlib import Path
from typing import Tuple

import numpy as np

from . import _lib
from .data_generation import PoissonSolver


def generate_test_data(k_range: Tuple[float, float], n_samples: int = 10, label: str = "test",
                       constant_theta: bool = True, *, save: bool = True) -> dict:
    """src/compare_test_cases.py:12-79.  The standard pairs are drawn first (all k values), then, for
    ``constant_theta=False``, one theta block per sample on the fine grid (same stream order as the
    reference's per-sample draws); the coarse theta is its ``[::2, ::2]`` gather, f_coarse stays the forcing
    evaluated on the coarse grid.  The reference's per-sample re-solves become two batched calls; with
    constant theta the dataset's solutions already are the theta = 1 solutions.  ``save=False`` (new)
    skips the reference's ``data/<label>_dataset.npz`` write."""
    solver = PoissonSolver()
    dataset = solver.generate_dataset(n_samples=n_samples, k_range=k_range)
    if constant_theta:
        dataset['theta_fine'] = np.ones_like(dataset['theta_fine'])
        dataset['theta_coarse'] = np.ones_like(dataset['theta_coarse'])
    else:
        # one draw of n_samples * n_fine^2 doubles from the global legacy generator, replayed on the device
        theta_fine = _lib.mt19937_uniform((n_samples, solver.n_fine, solver.n_fine), 0.5, 2.0)
        theta_coarse = np.ascontiguousarray(theta_fine[:, ::2, ::2])
        dataset['theta_fine'] = theta_fine
        dataset['theta_coarse'] = theta_coarse
        if n_samples > 0:
            dataset['u_fine'] = solver.solve_poisson_batch(dataset['f_fine'], theta_fine, solver.n_fine)
            dataset['u_coarse'] = solver.solve_poisson_batch(dataset['f_coarse'], theta_coarse, solver.n_coarse)
    if save:
        save_path = Path(f'data/{label}_dataset.npz')
        save_path.parent.mkdir(exist_ok=True)
        np.savez(save_path, **dataset)
    return dataset


def generate_out_of_sample_data(n_samples: int = 10, *, save: bool = True) -> dict:
    """src/test_out_of_sample.py:13-35: the standard generator with k ~ U(5, 8), written to
    ``data/out_of_sample_dataset.npz``."""
    dataset = PoissonSolver().generate_dataset(n_samples=n_samples, k_range=(5.0, 8.0))
    if save:
        save_path = Path('data/out_of_sample_dataset.npz')
        save_path.parent.mkdir(exist_ok=True)
        np.savez(save_path, **dataset)
    return dataset
