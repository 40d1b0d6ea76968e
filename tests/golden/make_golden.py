"""Generate the golden fixtures in this directory from the LIVE, UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own classes (matplotlib/seaborn are stubbed because the
reference imports them at module top and they are not installed), runs them on seeded
inputs, asserts that the oracle restatement (oracle/poisson_oracle.py) reproduces the
reference bit-for-bit, and stores the REFERENCE outputs.  Large fields are stored
sub-sampled (``u[::S, ::S]``) together with their full L2 norm so the fixtures stay small.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
for m in ("matplotlib", "matplotlib.pyplot", "seaborn"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, "/root/reference/src")
warnings.filterwarnings("ignore")

from data_generation import PoissonSolver  # noqa: E402  (reference)
from enhanced_data_generation import EnhancedPoissonSolver  # noqa: E402  (reference)
import resolution_comparison as ref_rc  # noqa: E402  (reference)

from oracle import poisson_oracle as O  # noqa: E402

KAT_SIZES = [20, 40, 80, 160, 320, 640]
KAT_K = [(0.5, 0.7), (2.3, 3.7), (11.9, 10.3)]


def kat_inputs(n, kidx, rough):
    """Deterministic KAT inputs; tests regenerate them with the same recipe."""
    k1, k2 = KAT_K[kidx]
    x = np.linspace(0, 1, n)
    X, Y = np.meshgrid(x, x)
    f = np.sin(2 * np.pi * k1 * X) * np.sin(2 * np.pi * k2 * Y)
    if rough:
        theta = np.random.RandomState(1000 + n + kidx).uniform(0.5, 2.0, size=(n, n))
    else:
        theta = np.ones((n, n))
    return f, theta


def sub(n):
    return 1 if n <= 80 else n // 80


def main():
    # ---- 1. solve_poisson known-answer tests
    kat = {}
    for n in KAT_SIZES:
        solver = PoissonSolver(n_coarse=n // 2, n_fine=n)
        for kidx in range(len(KAT_K)):
            for rough in (0, 1):
                if n == 640 and kidx != 2:
                    continue  # 7 s per reference solve; one smooth + one rough case at 640
                f, theta = kat_inputs(n, kidx, rough)
                u = solver.solve_poisson(f, theta, "fine")
                uo = O.solve_poisson(f, theta, n)
                assert np.array_equal(u, uo), "oracle != reference"
                ud = O.solve_dst(f, theta)
                assert O.rel_l2(ud, u) < 5e-12, (n, O.rel_l2(ud, u))
                s = sub(n)
                key = f"n{n}_k{kidx}_r{rough}"
                kat[key + "_u"] = u[::s, ::s].copy()
                kat[key + "_norm"] = np.linalg.norm(u)
                kat[key + "_res"] = O.residual_rel(u, f, theta)
                print(key, "ok  resid", kat[key + "_res"], "dst-vs-ref", O.rel_l2(ud, u))
    np.savez_compressed(os.path.join(HERE, "kat_solve_poisson.npz"), **kat)

    # ---- 2. standard dataset (seed 0) and subdomain dataset (seed 1)
    np.random.seed(0)
    ds = PoissonSolver().generate_dataset(6, k_range=(0.5, 5.0))
    np.random.seed(0)
    dso = O.generate_dataset(6, k_range=(0.5, 5.0))
    assert all(np.array_equal(ds[k], dso[k]) for k in ds)
    np.savez_compressed(os.path.join(HERE, "dataset_standard_seed0.npz"), **ds)

    np.random.seed(1)
    es = EnhancedPoissonSolver()
    sd = es.generate_subdomain_dataset(4, k_range=(0.5, 12.0))
    np.random.seed(1)
    sdo = O.generate_subdomain_dataset(4, k_range=(0.5, 12.0))
    assert all(np.array_equal(sd[k], sdo[k]) for k in sd)
    np.savez_compressed(os.path.join(HERE, "dataset_subdomain_seed1.npz"), **sd)

    comb = es.combine_datasets({k: v.copy() for k, v in ds.items()}, sd)
    np.savez_compressed(os.path.join(HERE, "dataset_combined_meta.npz"),
                        is_subdomain=comb["is_subdomain"], k1=comb["k1"], k2=comb["k2"],
                        u_fine_norms=np.linalg.norm(comb["u_fine"].reshape(len(comb["k1"]), -1), axis=1))

    # ---- 3. multi-resolution chain (seed 2): 40/80/160/320/640
    np.random.seed(2)
    mr = ref_rc.solve_multi_resolution(40, [80, 160, 320, 640])
    np.random.seed(2)
    mro = O.solve_multi_resolution(40, [80, 160, 320, 640])
    out = {"k1": mr["k1"], "k2": mr["k2"]}
    for res in (40, 80, 160, 320, 640):
        assert np.array_equal(mr["u"][res], mro["u"][res])
        assert np.array_equal(mr["f"][res], mro["f"][res])
        assert np.array_equal(mr["theta"][res], mro["theta"][res])
        s = sub(res)
        out[f"u{res}"] = mr["u"][res][::s, ::s].copy()
        out[f"u{res}_norm"] = np.linalg.norm(mr["u"][res])
        out[f"f{res}_norm"] = np.linalg.norm(mr["f"][res])
        out[f"theta{res}_sum"] = mr["theta"][res].sum()
    np.savez_compressed(os.path.join(HERE, "multires_seed2.npz"), **out)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
