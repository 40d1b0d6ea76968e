"""Drop-in for the solver part of the reference's ``src/subdomain_upscaling.py``: ``solve_large_scale``
(80/160) and ``solve_multi_level`` (40/80/160).  Same RNG order, dict keys and array views as the
reference; the two / three SuperLU factorisations become one ``psn_multires`` call (stride gathers of
theta and of the sine factors on the device).  The tiled U-Net upscaling and the plots of that file are
out of scope.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .data_generation import DEFAULT_MAX_ITER, DEFAULT_RTOL
from .resolution_comparison import _sines_kfirst, draw_examples


def _chain(sizes, rtol, max_iter, what):
    """k1, k2 ~ U(5, 8), f on the finest grid (``np.sin(k1 * 2 * np.pi * X)`` operand order), theta block,
    then every size in ``sizes`` (finest first, each half of the previous) solved on the [::2, ::2] gathers."""
    n_fine = sizes[0]
    for a, b in zip(sizes, sizes[1:]):
        if a != 2 * b:
            raise ValueError(f"{what}: grid sizes must halve ({a} -> {b}); the reference gathers [::2, ::2]")
    k1a, k2a, theta_solver, theta_host = draw_examples(1, n_fine, (5.0, 8.0))   # RNG replayed on the device
    k1, k2 = float(k1a[0]), float(k2a[0])
    x = np.linspace(0, 1, n_fine)
    sx, sy = _sines_kfirst(k1a, k2a, x)
    f = [sy[0][:, None] * sx[0][None, :]]
    theta = [theta_host[0]]
    for _ in sizes[1:]:
        f.append(f[-1][::2, ::2])
        theta.append(theta[-1][::2, ::2])
    us, info = _lib.multires(n_fine, list(sizes), 1, theta_solver, sx, sy, rtol=rtol, max_iter=max_iter)
    info.raise_if_failed(what)
    return k1, k2, f, theta, [us[n][0] for n in sizes]


def solve_large_scale(n_coarse: int = 80, n_fine: int = 160, *, rtol: float = DEFAULT_RTOL,
                      max_iter: int = DEFAULT_MAX_ITER, verbose: bool = False) -> dict:
    """src/subdomain_upscaling.py:11-60."""
    k1, k2, f, theta, u = _chain((n_fine, n_coarse), rtol, max_iter, "solve_large_scale")
    if verbose:
        print(f"u_fine - min: {u[0].min():.6f}, max: {u[0].max():.6f}")
        print(f"u_coarse - min: {u[1].min():.6f}, max: {u[1].max():.6f}")
    return {'u_fine': u[0], 'u_coarse': u[1], 'f_fine': f[0], 'f_coarse': f[1],
            'theta_fine': theta[0], 'theta_coarse': theta[1], 'k1': k1, 'k2': k2}


def solve_multi_level(n_coarse: int = 40, n_mid: int = 80, n_fine: int = 160, *, rtol: float = DEFAULT_RTOL,
                      max_iter: int = DEFAULT_MAX_ITER, verbose: bool = False) -> dict:
    """src/subdomain_upscaling.py:276-330."""
    k1, k2, f, theta, u = _chain((n_fine, n_mid, n_coarse), rtol, max_iter, "solve_multi_level")
    if verbose:
        for name, v in zip(("u_fine", "u_mid", "u_coarse"), u):
            print(f"{name} ({v.shape[0]}x{v.shape[1]}) - min: {v.min():.6f}, max: {v.max():.6f}")
    return {'u_fine': u[0], 'u_mid': u[1], 'u_coarse': u[2], 'f_fine': f[0], 'f_mid': f[1], 'f_coarse': f[2],
            'theta_fine': theta[0], 'theta_mid': theta[1], 'theta_coarse': theta[2], 'k1': k1, 'k2': k2}
