"""Drop-in for the solver part of the reference's ``src/resolution_comparison.py`` /
``src/resolution_comparison_statistical.py``: ``solve_multi_resolution`` (the 40/80/160/320/640
reference solutions).  The ML tiling / interpolation / plotting code of those files is out of scope.
"""
from __future__ import annotations

import time
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from .data_generation import DEFAULT_MAX_ITER, DEFAULT_RTOL


def _sines_kfirst(k1: np.ndarray, k2: np.ndarray, x: np.ndarray):
    """Operand order of the evaluation scripts: ``np.sin(k1 * 2 * np.pi * X)``
    (src/resolution_comparison.py:36), which rounds differently from ``2 * np.pi * k1 * X``."""
    k1 = np.asarray(k1, dtype=np.float64).reshape(-1, 1)
    k2 = np.asarray(k2, dtype=np.float64).reshape(-1, 1)
    return np.sin((k1 * 2 * np.pi) * x[None, :]), np.sin((k2 * 2 * np.pi) * x[None, :])


def draw_examples(n_examples: int, n_finest: int, k_range: Tuple[float, float], *, device_rng: bool = True, theta_on_host: bool = True,
                  defer_host: bool = False):
    """The RNG part of ``n_examples`` consecutive chains: per example k1, k2 ~ U(*k_range) and the master field
    theta ~ U(0.5, 2.0) on the finest grid, from numpy's GLOBAL legacy generator in the reference's order
    (src/resolution_comparison.py:27-28,42; src/subdomain_upscaling.py:19-20,31).  With ``device_rng`` the
    Mersenne-Twister stream is replayed on the GPU (``psn_mt19937_uniform_blocks``: bit-identical values, global state
    advanced as numpy would) and theta stays a CUDA tensor for the solve -- the 2 ms per 640 x 640 field of host RNG and
    its PCIe upload disappear.  Returns (k1 [E], k2 [E], theta for the solver, theta as a host array or None); with
    ``defer_host`` the last item is a zero-argument callable that waits for the (asynchronous) device-to-host copy and
    returns the array, so that the copy can run behind the caller's solves."""
    lo, hi = float(k_range[0]), float(k_range[1])
    nn = n_finest * n_finest
    if device_rng:
        theta_dev = _lib.device_empty_f64((n_examples, nn))
        _, raw = _lib.mt19937_uniform_blocks(n_examples, 2, nn, 0.5, 2.0, out=theta_dev)
        k1 = lo + (hi - lo) * raw[:, 0]        # numpy's random_uniform: lower + range * d (separate multiply and add)
        k2 = lo + (hi - lo) * raw[:, 1]
        theta_dev = _lib.reshaped(theta_dev, (n_examples, n_finest, n_finest))
        if theta_on_host and defer_host:
            return k1, k2, theta_dev, _lib.to_host_async(theta_dev)[1]
        theta_host = _lib.to_host(theta_dev) if theta_on_host else None
        return k1, k2, theta_dev, theta_host
    k1 = np.empty(n_examples)
    k2 = np.empty(n_examples)
    theta = np.empty((n_examples, n_finest, n_finest))
    for e in range(n_examples):
        k1[e] = np.random.uniform(lo, hi)
        k2[e] = np.random.uniform(lo, hi)
        theta[e] = np.random.uniform(0.5, 2.0, size=(n_finest, n_finest))
    return k1, k2, theta, ((lambda: theta) if (theta_on_host and defer_host) else theta)


def solve_multi_resolution(n_coarse: int = 40, resolutions: List[int] = [80, 160, 320, 640], *, device_rng: bool = True,
                           k_range: Tuple[float, float] = (10.0, 11.0), timed: bool = False,
                           rtol: float = DEFAULT_RTOL, max_iter: int = DEFAULT_MAX_ITER, verbose: bool = False):
    """src/resolution_comparison.py:13-78.  RNG order: k1, k2, then the theta block.
    ``timed=True`` adds ``solve_times`` (per-resolution wall seconds) like
    src/resolution_comparison_statistical.py:75-90, which uses ``k_range=(8.0, 12.0)``."""
    resolutions = list(resolutions)
    n_finest = max(resolutions)
    k1a, k2a, theta_solver, theta_host = draw_examples(1, n_finest, k_range, device_rng=device_rng)
    k1, k2 = float(k1a[0]), float(k2a[0])
    x = np.linspace(0, 1, n_finest)
    sx, sy = _sines_kfirst(k1a, k2a, x)
    f_finest = sy[0][:, None] * sx[0][None, :]
    theta_finest = theta_host[0]

    data = {'k1': k1, 'k2': k2, 'f': {}, 'theta': {}, 'u': {}}
    if timed:
        data['solve_times'] = {}
    all_res = [n_coarse] + resolutions
    for res in all_res:
        if n_finest % res != 0:
            raise ValueError(f"resolution {res} does not divide the finest grid {n_finest}")
        if res == n_finest:
            data['f'][res] = f_finest
            data['theta'][res] = theta_finest
        else:
            step = n_finest // res
            data['f'][res] = f_finest[::step, ::step]
            data['theta'][res] = theta_finest[::step, ::step]
    if timed:
        for res in all_res:
            t0 = time.time()
            us, info = _lib.multires(n_finest, [res], 1, theta_solver, sx, sy, rtol=rtol, max_iter=max_iter)
            data['solve_times'][res] = time.time() - t0
            info.raise_if_failed(f"solve_multi_resolution ({res}x{res})")
            data['u'][res] = us[res][0]
    else:
        us, info = _lib.multires(n_finest, all_res, 1, theta_solver, sx, sy, rtol=rtol, max_iter=max_iter)
        info.raise_if_failed("solve_multi_resolution")
        for res in all_res:
            data['u'][res] = us[res][0]
    if verbose:
        for res in all_res:
            print(f"u_{res} - min: {data['u'][res].min():.6f}, max: {data['u'][res].max():.6f}")
    return data


def solve_multi_resolution_batch(n_examples: int, n_coarse: int = 40, resolutions: Sequence[int] = (80, 160, 320, 640), *,
                                 k_range: Tuple[float, float] = (10.0, 11.0), rtol: float = DEFAULT_RTOL,
                                 max_iter: int = DEFAULT_MAX_ITER, device_rng: bool = True, theta_on_host: bool = True):
    """New: ``n_examples`` independent multi-resolution chains in one call per resolution.  The RNG is
    consumed exactly as ``n_examples`` consecutive calls of ``solve_multi_resolution`` would (see ``draw_examples``);
    ``theta_on_host=False`` leaves the master fields on the device (``theta_finest`` is then a CUDA tensor)."""
    resolutions = list(resolutions)
    n_finest = max(resolutions)
    k1, k2, theta_solver, theta_wait = draw_examples(n_examples, n_finest, k_range, device_rng=device_rng, theta_on_host=theta_on_host,
                                                     defer_host=True)
    x = np.linspace(0, 1, n_finest)
    sx, sy = _sines_kfirst(k1, k2, x)
    all_res = [n_coarse] + resolutions
    us, info = _lib.multires(n_finest, all_res, n_examples, theta_solver, sx, sy, rtol=rtol, max_iter=max_iter)
    theta_host = theta_wait() if theta_on_host else None   # device RNG: the copy of theta ran behind the solves
    info.raise_if_failed("solve_multi_resolution_batch")
    return {'k1': k1, 'k2': k2, 'theta_finest': theta_host if theta_host is not None else theta_solver, 'sin_x': sx, 'sin_y': sy,
            'u': us, 'info': info}


# --------------------------------------------------------------------------- interpolation baselines (row f4)
def _cascade(data: dict, target_resolution: int, mode: str) -> np.ndarray:
    current_res = 40
    levels = 0
    while current_res < target_resolution:
        current_res *= 2
        levels += 1
    u40 = np.asarray(data['u'][40])
    if levels == 0:
        return u40
    return _lib.upscale_cascade(u40[None], levels, mode)[0]


def bilinear_multi_level_upscale(data: dict, target_resolution: int) -> np.ndarray:
    """src/resolution_comparison_enhanced.py:19-41: 40 -> 80 -> ... -> target by x2 bilinear steps
    (align_corners=True, float32), all levels in one GPU call."""
    return _cascade(data, target_resolution, "bilinear")


def cubic_multi_level_upscale(data: dict, target_resolution: int) -> np.ndarray:
    """src/resolution_comparison_enhanced.py:43-65: the same cascade with bicubic steps."""
    return _cascade(data, target_resolution, "bicubic")


def split_into_subdomains(array: np.ndarray, subdomain_size: int) -> list:
    """src/resolution_comparison.py:123-139 (host-side views, as in the reference)."""
    n = array.shape[0] // subdomain_size
    return [[array[i * subdomain_size:(i + 1) * subdomain_size, j * subdomain_size:(j + 1) * subdomain_size]
             for j in range(n)] for i in range(n)]


def stitch_subdomains(subdomains: list) -> np.ndarray:
    """src/resolution_comparison.py:141-158"""
    n_rows, n_cols = len(subdomains), len(subdomains[0])
    size = subdomains[0][0].shape[0]
    result = np.zeros((n_rows * size, n_rows * size))
    for i in range(n_rows):
        for j in range(n_cols):
            result[i * size:(i + 1) * size, j * size:(j + 1) * size] = subdomains[i][j]
    return result
