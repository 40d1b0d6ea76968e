"""CPU restatement of the reference's Poisson sample-generation path (numpy + SciPy).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The product path never
imports this module; it must fail loudly when the CUDA library is missing.

What is restated, with the reference lines each function follows
(all paths relative to ``/root/reference``):

* ``create_laplacian``            src/data_generation.py:35-58
* ``forcing_term``                src/data_generation.py:60-77,
                                  src/enhanced_data_generation.py:34-45,
                                  src/resolution_comparison.py:32-36
* ``solve_poisson``               src/data_generation.py:79-104,
                                  src/enhanced_data_generation.py:47-68
* ``generate_dataset``            src/data_generation.py:106-159
* ``extract_subdomain``/``downsample``   src/enhanced_data_generation.py:70-96
* ``generate_subdomain_dataset``  src/enhanced_data_generation.py:98-165
* ``combine_datasets``            src/enhanced_data_generation.py:167-191
* ``solve_multi_resolution``      src/resolution_comparison.py:13-78 and the timed
                                  variant src/resolution_comparison_statistical.py:25-96
* ``solve_large_scale``           src/subdomain_upscaling.py:11-60
* ``solve_multi_level``           src/subdomain_upscaling.py:276-330
* ``generate_test_data``          src/compare_test_cases.py:12-79 (without the ``np.savez``)
* ``generate_out_of_sample_data`` src/test_out_of_sample.py:13-35 (without the ``np.savez``)

Third-party arithmetic: the reference's solve is ``scipy.sparse.linalg.spsolve``
(SciPy's vendored SuperLU ``gssv``, COLAMD ordering; UMFPACK branch inactive because
scikit-umfpack is absent).  The reference does not pin SciPy (README.md:183-190);
this image has SciPy 1.18.1 / numpy 2.3.5, and the oracle calls the very same SciPy
entry point on the very same matrix, so the arithmetic is the reference's own.

Pinning: the reference ships NO golden vectors or assertions for this path
(SURVEY.md section 4).  The oracle is therefore pinned against the *live reference*
imported in the build container: ``tests/golden/make_golden.py`` runs the unmodified
reference classes and this oracle on the same seeds, asserts bit-equality, and stores
the reference outputs as fixtures (``tests/golden/*.npz``) that travel to the GPU box.

A second, independent pin is ``solve_dst``: the closed-form discrete sine transform
(DST-I) solution of ``L u = f/theta``, which agrees with SuperLU to
<= 7e-13 relative L2 at every size and costs O(n^3) instead of a sparse LU.
"""
from __future__ import annotations

import time
import warnings
from typing import Dict, List, Sequence, Tuple

import numpy as np
from scipy.sparse import diags
from scipy.sparse.linalg import spsolve


# --------------------------------------------------------------------------- operator
def create_laplacian(n: int):
    """5-point Laplacian / h^2, h = 1/(n-1); all n*n nodes unknown, implicit zero ghosts.

    Follows src/data_generation.py:45-58 (diag -4, +-1 off-diagonals with the row-wrap
    entries zeroed, +-n off-diagonals, divided by h*h).
    """
    h = 1.0 / (n - 1)
    n2 = n * n
    main = -4 * np.ones(n2)
    off = np.ones(n2 - 1)
    off[np.arange(n - 1, n2 - 1, n)] = 0
    L = diags([main, off, off, np.ones(n * (n - 1)), np.ones(n * (n - 1))], [0, 1, -1, n, -n], shape=(n2, n2))
    return L / (h * h)


def grid(n: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """x = linspace(0,1,n) and the default 'xy' meshgrid (src/data_generation.py:22-29)."""
    x = np.linspace(0, 1, n)
    X, Y = np.meshgrid(x, x)
    return x, X, Y


def forcing_term(k1: float, k2: float, n: int) -> np.ndarray:
    """f[i,j] = sin(2 pi k1 x_j) sin(2 pi k2 y_i)   (src/data_generation.py:77)."""
    _, X, Y = grid(n)
    return np.sin(2 * np.pi * k1 * X) * np.sin(2 * np.pi * k2 * Y)


def forcing_term_kfirst(k1: float, k2: float, n: int) -> np.ndarray:
    """Same field with the operand order used by the evaluation scripts
    (``np.sin(k1 * 2 * np.pi * X)``, src/resolution_comparison.py:36)."""
    _, X, Y = grid(n)
    return np.sin(k1 * 2 * np.pi * X) * np.sin(k2 * 2 * np.pi * Y)


_L_CACHE: Dict[int, object] = {}


def solve_poisson(f: np.ndarray, theta: np.ndarray, n: int | None = None) -> np.ndarray:
    """Solve diag(theta) L u = f by sparse LU (src/data_generation.py:95-104)."""
    if n is None:
        n = int(round(np.sqrt(f.size)))
    L = _L_CACHE.get(n)
    if L is None:
        L = _L_CACHE[n] = create_laplacian(n)
    f_flat = f.reshape(-1)
    theta_flat = theta.reshape(-1)
    L_theta = diags(theta_flat) @ L
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # SparseEfficiencyWarning: DIA handed to spsolve, as in the reference
        u_flat = spsolve(L_theta, f_flat)
    return u_flat.reshape((n, n))


def solve_dst(f: np.ndarray, theta: np.ndarray) -> np.ndarray:
    """Closed-form solution of L u = f/theta by DST-I diagonalisation.

    Eigenpairs of the 1-D second difference with zero ghosts: s_p(i) = sin(pi (p+1)(i+1)/(n+1)),
    lambda_p = -(4/h^2) sin^2(pi (p+1) / (2 (n+1))).
    """
    from scipy.fft import dstn, idstn

    n = f.shape[0]
    h = 1.0 / (n - 1)
    lam = -(4.0 / (h * h)) * np.sin(np.pi * (np.arange(n) + 1) / (2.0 * (n + 1))) ** 2
    g = dstn(f / theta, type=1)
    g /= lam[:, None] + lam[None, :]
    return idstn(g, type=1)


def residual_rel(u: np.ndarray, f: np.ndarray, theta: np.ndarray) -> float:
    """||f - theta L u||_2 / ||f||_2, matrix-free (stencil of src/data_generation.py:49-58)."""
    n = u.shape[0]
    h = 1.0 / (n - 1)
    Lu = -4.0 * u
    Lu[1:, :] += u[:-1, :]
    Lu[:-1, :] += u[1:, :]
    Lu[:, 1:] += u[:, :-1]
    Lu[:, :-1] += u[:, 1:]
    Lu /= h * h
    return float(np.linalg.norm(f - theta * Lu) / np.linalg.norm(f))


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """||a-b||_2 / ||b||_2 (cf. src/utils.py:52-54, without its 1e-8 guard)."""
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


# --------------------------------------------------------------------------- datasets
def generate_dataset(n_samples: int, k_range: Tuple[float, float] = (1, 5), n_coarse: int = 20, n_fine: int = 40) -> dict:
    """Standard coarse/fine pairs; global np.random stream, k1 then k2 per sample
    (src/data_generation.py:128-157)."""
    keys = ["u_coarse", "u_fine", "f_coarse", "f_fine", "theta_coarse", "theta_fine", "k1", "k2"]
    ds = {k: [] for k in keys}
    for _ in range(n_samples):
        k1 = np.random.uniform(*k_range)
        k2 = np.random.uniform(*k_range)
        theta_fine = np.ones((n_fine, n_fine))
        theta_coarse = np.ones((n_coarse, n_coarse))
        f_fine = forcing_term(k1, k2, n_fine)
        f_coarse = forcing_term(k1, k2, n_coarse)
        u_fine = solve_poisson(f_fine, theta_fine, n_fine)
        u_coarse = solve_poisson(f_coarse, theta_coarse, n_coarse)
        for k, v in zip(keys, [u_coarse, u_fine, f_coarse, f_fine, theta_coarse, theta_fine, k1, k2]):
            ds[k].append(v)
    return {k: np.array(v) for k, v in ds.items()}


def extract_subdomain(field: np.ndarray, start_x: int, start_y: int, size: int) -> np.ndarray:
    """rows = y, cols = x (src/enhanced_data_generation.py:83)."""
    return field[start_y:start_y + size, start_x:start_x + size]


def downsample(field: np.ndarray, factor: int = 2) -> np.ndarray:
    """src/enhanced_data_generation.py:96"""
    return field[::factor, ::factor]


def generate_subdomain_dataset(n_samples: int, k_range: Tuple[float, float] = (0.5, 12.0),
                               n_coarse: int = 20, n_fine: int = 40, n_superfine: int = 80) -> dict:
    """One n_superfine^2 solve per sample, random window drawn AFTER the solve (x then y),
    stride-2 downsample (src/enhanced_data_generation.py:121-163)."""
    keys = ["u_coarse", "u_fine", "f_coarse", "f_fine", "theta_coarse", "theta_fine", "k1", "k2", "is_subdomain"]
    ds = {k: [] for k in keys}
    for _ in range(n_samples):
        k1 = np.random.uniform(*k_range)
        k2 = np.random.uniform(*k_range)
        theta_sf = np.ones((n_superfine, n_superfine))
        f_sf = forcing_term(k1, k2, n_superfine)
        u_sf = solve_poisson(f_sf, theta_sf, n_superfine)
        max_start = n_superfine - n_fine
        start_x = np.random.randint(0, max_start)
        start_y = np.random.randint(0, max_start)
        theta_fine = extract_subdomain(theta_sf, start_x, start_y, n_fine)
        f_fine = extract_subdomain(f_sf, start_x, start_y, n_fine)
        u_fine = extract_subdomain(u_sf, start_x, start_y, n_fine)
        vals = [downsample(u_fine), u_fine, downsample(f_fine), f_fine, downsample(theta_fine), theta_fine, k1, k2, True]
        for k, v in zip(keys, vals):
            ds[k].append(v)
    return {k: np.array(v) for k, v in ds.items()}


def combine_datasets(dataset1: Dict, dataset2: Dict) -> Dict:
    """src/enhanced_data_generation.py:167-191 (mutates dataset1 like the reference)."""
    out = {}
    if "is_subdomain" not in dataset1:
        dataset1["is_subdomain"] = np.zeros(len(dataset1["u_fine"]), dtype=bool)
    for key in dataset1:
        out[key] = np.concatenate([dataset1[key], dataset2[key]]) if key in dataset2 else dataset1[key]
    return out


def solve_multi_resolution(n_coarse: int = 40, resolutions: Sequence[int] = (80, 160, 320, 640),
                           k_lo: float = 10.0, k_hi: float = 11.0, timed: bool = False) -> dict:
    """src/resolution_comparison.py:13-78; ``timed=True`` adds ``solve_times`` like
    src/resolution_comparison_statistical.py:75-90 (which draws k from U(8,12))."""
    resolutions = list(resolutions)
    k1 = np.random.uniform(k_lo, k_hi)
    k2 = np.random.uniform(k_lo, k_hi)
    n_finest = max(resolutions)
    f_finest = forcing_term_kfirst(k1, k2, n_finest)
    theta_finest = np.random.uniform(0.5, 2.0, size=(n_finest, n_finest))
    data = {"k1": k1, "k2": k2, "f": {}, "theta": {}, "u": {}}
    if timed:
        data["solve_times"] = {}
    for res in [n_coarse] + resolutions:
        if res == n_finest:
            data["f"][res], data["theta"][res] = f_finest, theta_finest
        else:
            step = n_finest // res
            data["f"][res] = f_finest[::step, ::step]
            data["theta"][res] = theta_finest[::step, ::step]
        t0 = time.time()
        data["u"][res] = solve_poisson(data["f"][res], data["theta"][res], res)
        if timed:
            data["solve_times"][res] = time.time() - t0
    return data


# --------------------------------------------------------------------------- evaluation-script callers
def solve_large_scale(n_coarse: int = 80, n_fine: int = 160) -> dict:
    """src/subdomain_upscaling.py:11-60.  RNG order: k1, k2 ~ U(5, 8), then the theta block on the
    fine grid; f and theta of the coarse solve are the ``[::2, ::2]`` gathers of the fine fields."""
    k1 = np.random.uniform(5.0, 8.0)
    k2 = np.random.uniform(5.0, 8.0)
    f_fine = forcing_term_kfirst(k1, k2, n_fine)
    theta_fine = np.random.uniform(0.5, 2.0, size=(n_fine, n_fine))
    f_coarse, theta_coarse = f_fine[::2, ::2], theta_fine[::2, ::2]
    return {"u_fine": solve_poisson(f_fine, theta_fine, n_fine),
            "u_coarse": solve_poisson(f_coarse, theta_coarse, n_coarse),
            "f_fine": f_fine, "f_coarse": f_coarse, "theta_fine": theta_fine, "theta_coarse": theta_coarse,
            "k1": k1, "k2": k2}


def solve_multi_level(n_coarse: int = 40, n_mid: int = 80, n_fine: int = 160) -> dict:
    """src/subdomain_upscaling.py:276-330: three solves on f / theta gathered twice by ``[::2, ::2]``."""
    k1 = np.random.uniform(5.0, 8.0)
    k2 = np.random.uniform(5.0, 8.0)
    f_fine = forcing_term_kfirst(k1, k2, n_fine)
    f_mid = f_fine[::2, ::2]
    f_coarse = f_mid[::2, ::2]
    theta_fine = np.random.uniform(0.5, 2.0, size=(n_fine, n_fine))
    theta_mid = theta_fine[::2, ::2]
    theta_coarse = theta_mid[::2, ::2]
    return {"u_fine": solve_poisson(f_fine, theta_fine, n_fine), "u_mid": solve_poisson(f_mid, theta_mid, n_mid),
            "u_coarse": solve_poisson(f_coarse, theta_coarse, n_coarse),
            "f_fine": f_fine, "f_mid": f_mid, "f_coarse": f_coarse,
            "theta_fine": theta_fine, "theta_mid": theta_mid, "theta_coarse": theta_coarse, "k1": k1, "k2": k2}


def generate_test_data(k_range: Tuple[float, float], n_samples: int = 10, constant_theta: bool = True) -> dict:
    """src/compare_test_cases.py:12-79 minus the file write.  The standard dataset is drawn first (all k);
    with ``constant_theta=False`` one 40x40 theta block per sample follows in the RNG stream, the coarse
    theta is its ``[::2, ::2]`` gather, and f_coarse stays the forcing evaluated ON the coarse grid."""
    ds = generate_dataset(n_samples, k_range)
    n_c, n_f = ds["u_coarse"].shape[1], ds["u_fine"].shape[1]
    for i in range(n_samples):
        if constant_theta:
            th_f, th_c = np.ones((n_f, n_f)), np.ones((n_c, n_c))
        else:
            th_f = np.random.uniform(0.5, 2.0, size=(n_f, n_f))
            th_c = th_f[::2, ::2]
        ds["theta_fine"][i], ds["theta_coarse"][i] = th_f, th_c
        ds["u_fine"][i] = solve_poisson(ds["f_fine"][i], th_f, n_f)
        ds["u_coarse"][i] = solve_poisson(ds["f_coarse"][i], th_c, n_c)
    return ds


def generate_out_of_sample_data(n_samples: int = 10) -> dict:
    """src/test_out_of_sample.py:13-35 minus the file write: the standard generator with k ~ U(5, 8)."""
    return generate_dataset(n_samples, (5.0, 8.0))
