import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

KAT_K = [(0.5, 0.7), (2.3, 3.7), (11.9, 10.3)]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def kat_inputs(n, kidx, rough):
    """Same recipe as tests/golden/make_golden.py::kat_inputs."""
    k1, k2 = KAT_K[kidx]
    x = np.linspace(0, 1, n)
    X, Y = np.meshgrid(x, x)
    f = np.sin(2 * np.pi * k1 * X) * np.sin(2 * np.pi * k2 * Y)
    theta = np.random.RandomState(1000 + n + kidx).uniform(0.5, 2.0, size=(n, n)) if rough else np.ones((n, n))
    return f, theta


def kat_sub(n):
    return 1 if n <= 80 else n // 80


@pytest.fixture(scope="session")
def golden_kat():
    return dict(np.load(os.path.join(GOLDEN, "kat_solve_poisson.npz")))


@pytest.fixture(scope="session")
def gpu_lib():
    from superresolution_for_pdes_b200 import _lib
    if _lib.device_count() < 1:
        pytest.fail("gpu test selected but no CUDA device is visible")
    _lib.init(0)
    return _lib


def _rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def check_caller_dict(d, gold, prefix, tol_u):
    """Compare one caller's returned dict with the reference fixture (tests/golden/make_golden_callers.py)."""
    for k, v in d.items():
        v = np.asarray(v)
        key = f"{prefix}_{k}"
        is_u = k.startswith("u_")
        if key in gold.files:
            assert v.shape == gold[key].shape, key
            if is_u:
                assert _rel_l2(v, gold[key]) <= tol_u, key
            else:
                assert np.array_equal(v, gold[key]), key
        else:  # 160x160 field: norm / sum (+ [::2, ::2] samples of u)
            if is_u:
                assert _rel_l2(v[::2, ::2], gold[key + "_s2"]) <= tol_u, key
                assert abs(np.linalg.norm(v) - gold[key + "_norm"]) <= tol_u * gold[key + "_norm"], key
            else:
                assert v.sum() == gold[key + "_sum"] and np.linalg.norm(v) == gold[key + "_norm"], key
