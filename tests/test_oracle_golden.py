"""CPU tier: the oracle (oracle/poisson_oracle.py) against the golden fixtures that
tests/golden/make_golden.py generated from the LIVE, unmodified reference."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, check_caller_dict, kat_inputs, kat_sub
from oracle import poisson_oracle as O


@pytest.mark.parametrize("n", [20, 40, 80, 160, 320])
def test_oracle_reproduces_reference_kats(golden_kat, n):
    for k in range(3):
        for r in (0, 1):
            key = f"n{n}_k{k}_r{r}"
            f, theta = kat_inputs(n, k, r)
            u = O.solve_poisson(f, theta, n)
            s = kat_sub(n)
            # same SciPy entry point on the same matrix: equal to round-off of the LU (bit-equal in the
            # container that made the fixtures; a different BLAS/SciPy build may differ in the last bits)
            assert O.rel_l2(u[::s, ::s], golden_kat[key + "_u"]) <= 1e-12
            assert O.residual_rel(u, f, theta) <= 10 * golden_kat[key + "_res"] + 1e-15


def test_oracle_640_and_dst(golden_kat):
    n = 640
    f, theta = kat_inputs(n, 2, 1)
    ud = O.solve_dst(f, theta)
    s = kat_sub(n)
    assert O.rel_l2(ud[::s, ::s], golden_kat["n640_k2_r1_u"]) <= 1e-11     # independent closed form
    assert abs(np.linalg.norm(ud) - golden_kat["n640_k2_r1_norm"]) <= 1e-11 * golden_kat["n640_k2_r1_norm"]


def test_dst_agrees_with_lu_everywhere():
    for n in (5, 7, 20, 33, 80):
        f, theta = kat_inputs(n, 1, 1)
        assert O.rel_l2(O.solve_dst(f, theta), O.solve_poisson(f, theta, n)) <= 1e-12


def test_oracle_datasets_bit_exact():
    gold = np.load(os.path.join(GOLDEN, "dataset_standard_seed0.npz"))
    np.random.seed(0)
    ds = O.generate_dataset(6, k_range=(0.5, 5.0))
    for k in gold.files:
        if k.startswith("u_"):
            assert O.rel_l2(ds[k], gold[k]) <= 1e-13
        else:
            assert np.array_equal(ds[k], gold[k]), k
    gold = np.load(os.path.join(GOLDEN, "dataset_subdomain_seed1.npz"))
    np.random.seed(1)
    sd = O.generate_subdomain_dataset(4, k_range=(0.5, 12.0))
    for k in gold.files:
        if k.startswith("u_"):
            assert O.rel_l2(sd[k], gold[k]) <= 1e-13
        else:
            assert np.array_equal(sd[k], gold[k]), k


def test_structural_facts_of_the_reference_operator():
    """SURVEY.md F1-F3, pinned on the oracle."""
    n = 40
    f, theta = kat_inputs(n, 1, 1)
    one = np.ones((n, n))
    assert O.rel_l2(O.solve_poisson(f, theta, n), O.solve_poisson(f / theta, one, n)) <= 1e-13   # F1 row scaling
    u = O.solve_poisson(f, one, n)
    i, j = np.unravel_index(np.argmax(f), f.shape)
    assert u[i, j] < 0                                                                           # F2 sign
    assert np.abs(u[0, :]).max() > 1e-6                                                          # F3 free border
    # bit-exact forcing from 1-D sine factors
    x = np.linspace(0, 1, n)
    k1, k2 = 2.3, 3.7
    assert np.array_equal(np.outer(np.sin(2 * np.pi * k2 * x), np.sin(2 * np.pi * k1 * x)), O.forcing_term(k1, k2, n))


def test_dataset_tensor_oracle_against_golden():
    """oracle/dataset_oracle.py against outputs of the LIVE reference PDEDataset (fixture made by
    tests/golden/make_golden_tensors.py in the build container)."""
    from oracle.dataset_oracle import dataset_tensors
    gold = np.load(os.path.join(GOLDEN, "dataset_tensors_seed0.npz"))
    ds = np.load(os.path.join(GOLDEN, "dataset_standard_seed0.npz"))
    out = dataset_tensors({k: ds[k] for k in ds.files})
    assert np.abs(out["x"] - gold["x"]).max() <= 1e-6 and np.abs(out["y"] - gold["y"]).max() <= 1e-6
    assert np.allclose(out["stats"][:4], gold["stats"][:4], rtol=1e-6)


def test_upscale_oracle_against_golden():
    """oracle/upscale_oracle.py against outputs of the LIVE reference cascades (fixture made by
    tests/golden/make_golden_upscale.py in the build container).  float32 kernels of torch may differ in the
    last bits between CPUs (vectorised paths), hence 1e-6 absolute on O(0.01..1) values."""
    from oracle import upscale_oracle as uo
    gold = np.load(os.path.join(GOLDEN, "upscale_seed2.npz"))
    u40 = np.load(os.path.join(GOLDEN, "multires_seed2.npz"))["u40"]
    data = {"u": {40: u40}}
    for target in (80, 160):
        for mode in ("bilinear", "bicubic"):
            out = uo.multi_level_upscale(data, target, mode)
            assert out.dtype == np.float32 and out.shape == (target, target)
            assert np.abs(out - gold[f"{mode}_{target}"]).max() <= 1e-6 * max(1.0, np.abs(u40).max())
    assert uo.multi_level_upscale(data, 40, "bilinear").shape == (40, 40)
    parts = uo.split_into_subdomains(u40, 10)
    assert len(parts) == 4 and len(parts[0]) == 4 and np.array_equal(uo.stitch_subdomains(parts), u40)


def test_oracle_callers_against_golden():
    """Row a10: evaluation-script callers restated in the oracle reproduce the live reference's outputs."""
    gold = np.load(os.path.join(GOLDEN, "callers_seed5.npz"))
    np.random.seed(5)
    check_caller_dict(O.solve_large_scale(20, 40), gold, "large_small", 1e-13)
    np.random.seed(5)
    check_caller_dict(O.solve_large_scale(80, 160), gold, "large", 1e-13)
    np.random.seed(5)
    check_caller_dict(O.solve_multi_level(40, 80, 160), gold, "level", 1e-13)
    for name, const in (("test_const", True), ("test_rand", False)):
        np.random.seed(6)
        check_caller_dict(O.generate_test_data((1.0, 6.0), 3, const), gold, name, 1e-13)
    np.random.seed(7)
    check_caller_dict(O.generate_out_of_sample_data(3), gold, "oos", 1e-13)
