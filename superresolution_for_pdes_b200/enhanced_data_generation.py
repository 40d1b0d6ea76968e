"""Drop-in for the reference's ``src/enhanced_data_generation.py`` (class ``EnhancedPoissonSolver``).

The 80x80 "super-fine" solve, the random 40x40 window and the stride-2 downsample of
``generate_subdomain_dataset`` (src/enhanced_data_generation.py:98-165) run as ONE fused kernel per
sample batch (``psn_subdomain_batch``).  ``plot_random_samples`` (plotting) is out of scope.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from . import _lib
from .data_generation import DEFAULT_MAX_ITER, DEFAULT_RTOL, PoissonSolver, sine_factors


class EnhancedPoissonSolver(PoissonSolver):
    def __init__(self, n_coarse: int = 20, n_fine: int = 40, n_superfine: int = 80, *,
                 rtol: float = DEFAULT_RTOL, max_iter: int = DEFAULT_MAX_ITER):
        """Reference: src/enhanced_data_generation.py:12-32."""
        super().__init__(n_coarse, n_fine, rtol=rtol, max_iter=max_iter)
        self.n_superfine = n_superfine
        self.x_superfine = np.linspace(0, 1, n_superfine)
        self.y_superfine = np.linspace(0, 1, n_superfine)
        self.X_superfine, self.Y_superfine = np.meshgrid(self.x_superfine, self.y_superfine)

    @property
    def L_superfine(self):
        return self._lazy_L(self.n_superfine)

    def generate_forcing_term_superfine(self, k1: float, k2: float) -> np.ndarray:
        """src/enhanced_data_generation.py:34-45"""
        return np.sin(2 * np.pi * k1 * self.X_superfine) * np.sin(2 * np.pi * k2 * self.Y_superfine)

    def solve_poisson_superfine(self, f: np.ndarray, theta: np.ndarray) -> np.ndarray:
        """src/enhanced_data_generation.py:47-68"""
        return self._solve_n(f, theta, self.n_superfine)

    def extract_subdomain(self, field: np.ndarray, start_x: int, start_y: int, size: int) -> np.ndarray:
        """rows = y, cols = x (src/enhanced_data_generation.py:70-83)"""
        return field[start_y:start_y + size, start_x:start_x + size]

    def downsample(self, field: np.ndarray, factor: int = 2) -> np.ndarray:
        """src/enhanced_data_generation.py:85-96"""
        return field[::factor, ::factor]

    def generate_subdomain_dataset(self, n_samples: int, k_range: Tuple[float, float] = (0.5, 12.0)) -> dict:
        """src/enhanced_data_generation.py:98-165.  The reference draws k1, k2, (solves), start_x,
        start_y per sample; the solve consumes no randomness, so all draws are made up front in that
        order and the whole batch is solved in one fused GPU call."""
        k1 = np.empty(n_samples)
        k2 = np.empty(n_samples)
        start_x = np.empty(n_samples, dtype=np.int32)
        start_y = np.empty(n_samples, dtype=np.int32)
        max_start = self.n_superfine - self.n_fine
        for i in range(n_samples):
            k1[i] = np.random.uniform(*k_range)
            k2[i] = np.random.uniform(*k_range)
            start_x[i] = np.random.randint(0, max_start)
            start_y[i] = np.random.randint(0, max_start)
        sx, sy = sine_factors(k1, k2, self.x_superfine)
        out, info = _lib.subdomain_batch(self.n_superfine, self.n_fine, n_samples, sx, sy, start_x, start_y,
                                         rtol=self.rtol, max_iter=self.max_iter)
        self.last_info = info
        info.raise_if_failed("generate_subdomain_dataset")
        dataset = {
            'u_coarse': out['u_coarse'],
            'u_fine': out['u_fine'],
            'f_coarse': out['f_coarse'],
            'f_fine': out['f_fine'],
            'theta_coarse': out['theta_coarse'],
            'theta_fine': out['theta_fine'],
            'k1': k1,
            'k2': k2,
            'is_subdomain': np.ones(n_samples, dtype=bool),
        }
        return dataset

    def combine_datasets(self, dataset1: Dict, dataset2: Dict) -> Dict:
        """src/enhanced_data_generation.py:167-191 (adds ``is_subdomain`` to dataset1 like the reference)."""
        combined_dataset = {}
        if 'is_subdomain' not in dataset1:
            dataset1['is_subdomain'] = np.zeros(len(dataset1['u_fine']), dtype=bool)
        for key in dataset1:
            if key in dataset2:
                combined_dataset[key] = np.concatenate([dataset1[key], dataset2[key]])
            else:
                combined_dataset[key] = dataset1[key]
        return combined_dataset
