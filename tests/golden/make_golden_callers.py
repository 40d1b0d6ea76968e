"""Golden fixture for the evaluation-script callers of the solve, from the
LIVE, UNMODIFIED reference:

    python tests/golden/make_golden_callers.py        (build container only: needs /root/reference)

Runs ``solve_large_scale`` / ``solve_multi_level`` (src/subdomain_upscaling.py), ``generate_test_data``
(src/compare_test_cases.py, both theta modes) and ``generate_out_of_sample_data`` (src/test_out_of_sample.py)
on fixed seeds, asserts that the oracle restatements reproduce every returned array bit for bit, and stores
the REFERENCE outputs in ``callers_seed5.npz`` (160x160 fields sub-sampled, see ``put``).  The reference functions write ``data/*.npz`` relative to
the working directory, so they run inside a scratch directory.
"""
import contextlib
import io
import os
import sys
import tempfile
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
for m in ("matplotlib", "matplotlib.pyplot", "seaborn"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, "/root/reference/src")
warnings.filterwarnings("ignore")

import subdomain_upscaling as ref_su  # noqa: E402  (reference)
import compare_test_cases as ref_ct  # noqa: E402  (reference)
import test_out_of_sample as ref_oos  # noqa: E402  (reference)

from oracle import poisson_oracle as O  # noqa: E402


def same(ref, ora, what):
    assert set(ref) == set(ora), (what, set(ref) ^ set(ora))
    for k in ref:
        assert np.array_equal(np.asarray(ref[k]), np.asarray(ora[k])), (what, k)


def put(out, key, v):
    """Fields larger than 80x80 are stored as [::2, ::2] samples plus their L2 norm and sum (the tests
    regenerate theta and f from the seed; the fixture pins them through norm and sum)."""
    if v.ndim == 2 and v.shape[0] > 80:
        out[key + "_s2"] = v[::2, ::2].copy() if key.split("_")[1] == "u" else np.zeros(0)
        out[key + "_norm"] = np.linalg.norm(v)
        out[key + "_sum"] = v.sum()
    else:
        out[key] = v


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def main():
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        os.mkdir("data")
        try:
            for name, ref_fn, ora_fn, args in (
                    ("large", ref_su.solve_large_scale, O.solve_large_scale, (80, 160)),
                    ("large_small", ref_su.solve_large_scale, O.solve_large_scale, (20, 40)),
                    ("level", ref_su.solve_multi_level, O.solve_multi_level, (40, 80, 160))):
                np.random.seed(5)
                r = quiet(ref_fn, *args)
                np.random.seed(5)
                o = ora_fn(*args)
                same(r, o, name)
                for k, v in r.items():
                    put(out, f"{name}_{k}", np.asarray(v))
            for name, const in (("test_const", True), ("test_rand", False)):
                np.random.seed(6)
                r = quiet(ref_ct.generate_test_data, (1.0, 6.0), 3, name, const)
                np.random.seed(6)
                o = O.generate_test_data((1.0, 6.0), 3, const)
                same(r, o, name)
                saved = np.load(os.path.join("data", f"{name}_dataset.npz"))
                assert set(saved.files) == set(r)
                for k, v in r.items():
                    out[f"{name}_{k}"] = v
            np.random.seed(7)
            r = quiet(ref_oos.generate_out_of_sample_data, 3)
            np.random.seed(7)
            same(r, O.generate_out_of_sample_data(3), "oos")
            for k, v in r.items():
                out[f"oos_{k}"] = v
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "callers_seed5.npz"), **out)
    print("callers_seed5.npz written:", len(out), "arrays")


if __name__ == "__main__":
    main()
