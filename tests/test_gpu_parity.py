"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) and the reference-signature
shims, against the oracle (live SciPy restatement) and the committed golden fixtures.

Tolerance (BASELINE.json north_star): relative L2 <= 1e-10 in fp64; forcing terms bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, KAT_K, check_caller_dict, kat_inputs, kat_sub
from oracle import poisson_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("n", [20, 40, 80, 160, 320, 640])
def test_kat_vs_golden(gpu_lib, golden_kat, n):
    """Known-answer tests against REFERENCE outputs (fixtures made by the unmodified reference)."""
    cases = [(k, r) for k in range(3) for r in (0, 1) if f"n{n}_k{k}_r{r}_u" in golden_kat]
    fs, ths = zip(*[kat_inputs(n, k, r) for k, r in cases])
    u, info = gpu_lib.solve_batch(n, len(cases), theta=np.stack(ths), f=np.stack(fs))
    s = kat_sub(n)
    for b, (k, r) in enumerate(cases):
        key = f"n{n}_k{k}_r{r}"
        err = O.rel_l2(u[b][::s, ::s], golden_kat[key + "_u"])
        nerr = abs(np.linalg.norm(u[b]) - golden_kat[key + "_norm"]) / golden_kat[key + "_norm"]
        res = O.residual_rel(u[b], fs[b], ths[b])
        print(key, "iters", info.iters[b], "status", info.status[b], "relres", info.relres[b], "err", err, "resid", res)
        assert info.accepted()[b]
        assert err <= TOL and nerr <= TOL
        # each solve reaches the reference's own residual level (or the 1e-12 target)
        assert res <= max(1e-12, 10 * golden_kat[key + "_res"])


@pytest.mark.parametrize("n", [5, 7, 13, 33, 64, 75, 86, 87, 100, 131, 200])
def test_odd_and_boundary_sizes_vs_oracle(gpu_lib, n):
    """Sizes around the resident/streamed switch, odd sizes, single-level grids."""
    rs = np.random.RandomState(n)
    B = 3
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    f = np.stack([O.forcing_term(*KAT_K[b], n) for b in range(B)])
    u, info = gpu_lib.solve_batch(n, B, theta=theta, f=f)
    for b in range(B):
        assert info.accepted()[b]
        assert O.rel_l2(u[b], O.solve_poisson(f[b], theta[b], n)) <= TOL


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_streamed_path_matches_resident(gpu_lib, mode):
    """Force the streamed kernels (mode 0: shared-memory tiles, mode 1: warp strips) on a grid the
    resident kernel also handles; both must agree."""
    n, B = 64, 5
    rs = np.random.RandomState(5)
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(0.5, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    u_res, _ = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    try:
        gpu_lib.set_option("smem_max_n", -1)
        gpu_lib.set_option("tile", 32)
        gpu_lib.set_option("vcycle_mode", mode)
        u_str, info = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    finally:
        gpu_lib.set_option("smem_max_n", 0)
        gpu_lib.set_option("tile", 64)
        gpu_lib.set_option("vcycle_mode", 2)
    assert info.accepted().all()
    for b in range(B):
        assert O.rel_l2(u_str[b], u_res[b]) <= 1e-11
        assert O.rel_l2(u_res[b], O.solve_poisson(np.outer(sy[b], sx[b]), theta[b], n)) <= TOL


@pytest.mark.parametrize("n,B", [(160, 40), (320, 3), (200, 7)])
def test_tile_and_strip_vcycles_agree(gpu_lib, n, B):
    """The two implementations of the streamed V-cycle legs run the same algorithm; small and large
    batches exercise different band splits of the strip kernels."""
    rs = np.random.RandomState(n)
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(0.5, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    try:
        gpu_lib.set_option("vcycle_mode", 0)
        u0, i0 = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
        gpu_lib.set_option("vcycle_mode", 1)
        u2, i2 = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
        gpu_lib.set_option("vcycle_mode", 2)
        gpu_lib.set_option("vcycle_f32", 0)
        u1, i1 = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    finally:
        gpu_lib.set_option("vcycle_mode", 2)
        gpu_lib.set_option("vcycle_f32", 1)
    assert i0.accepted().all() and i1.accepted().all() and i2.accepted().all()
    assert O.rel_l2(u1, u0) <= 1e-12
    assert O.rel_l2(u1, u2) <= 1e-13  # register-fed and TMA-fed strips do identical V-cycle arithmetic (the TMA mode's
    #                                   strip update kernel sums its partial dot products in another order)
    assert O.rel_l2(u1[0], O.solve_poisson(np.outer(sy[0], sx[0]), theta[0], n)) <= TOL
    # default configuration: V-cycle intermediates stored as float where every streamed level is a multiple of 4
    # (160, 320: yes; 200 -> 100 -> 50: no, stays fp64).  Same iteration counts, same accuracy bar.
    u3, i3 = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    assert i3.accepted().all() and (i3.iters <= i1.iters + 1).all() and i3.iters.sum() <= i1.iters.sum() + 1
    assert O.rel_l2(u3, u1) <= 1e-11
    assert O.rel_l2(u3[0], O.solve_poisson(np.outer(sy[0], sx[0]), theta[0], n)) <= TOL
    if n == 200:
        assert np.array_equal(u3, u1)


def test_float_stored_vcycle_is_scale_invariant(gpu_lib):
    """The float copies of the V-cycle intermediates are rescaled by a power of two of the residual norm, so data of
    any magnitude converges exactly like O(1) data (the reference's direct solve is scale-invariant too)."""
    n, B = 160, 6
    rs = np.random.RandomState(5)
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    f = rs.standard_normal((B, n, n))
    u, info = gpu_lib.solve_batch(n, B, theta=theta, f=f)
    for scale in (1e-40, 1e-25, 1e+25):
        us, infos = gpu_lib.solve_batch(n, B, theta=theta, f=f * scale)
        assert infos.accepted().all() and (infos.iters <= info.iters + 1).all()
        assert O.rel_l2(us / scale, u) <= 1e-10


@pytest.mark.parametrize("n,B", [(20, 700), (40, 650), (80, 310)])
def test_register_blocked_kernels_against_the_generic_one(gpu_lib, n, B):
    """20 x 20, 40 x 40 and 80 x 80 run in the register-blocked kernels (float V-cycle, levels n, n/2, n/4 in registers);
    `res_block = 0` sends them through the generic shared-memory kernel (fp64 V-cycle).  Same iteration counts to within one,
    same solutions to round-off, 1e-10 against the reference operator -- in every forcing form the kernel has a path for
    (sine factors, f array, wave numbers; theta or none).  (The fused window / downsample outputs of the 80 x 80 kernel are
    covered by the subdomain-dataset fixtures.)"""
    rs = np.random.RandomState(3 * n)
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(0.5, n / 4.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    f = sy[:, :, None] * sx[:, None, :]
    forms = [dict(theta=theta, sin_x=sx, sin_y=sy), dict(theta=theta, f=f), dict(theta=None, sin_x=sx, sin_y=sy),
             dict(theta=theta, k1=k[:, 0].copy(), k2=k[:, 1].copy())]
    assert gpu_lib.describe(n)["resident"]
    for kw in forms:
        u, info = gpu_lib.solve_batch(n, B, **kw)
        gpu_lib.set_option("res_block", 0)
        try:
            ug, infog = gpu_lib.solve_batch(n, B, **kw)
        finally:
            gpu_lib.set_option("res_block", 1)
        assert (info.status == 0).all() and (infog.status == 0).all()
        assert (np.abs(info.iters - infog.iters) <= 1).all()
        assert O.rel_l2(u, ug) <= 1e-11
    u, info = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    for b in (0, B // 2, B - 1):
        assert O.rel_l2(u[b], O.solve_poisson(f[b], theta[b], n)) <= TOL


@pytest.mark.parametrize("n,B", [(160, 150), (320, 7), (640, 3)])
def test_coarse_tail_in_the_register_blocked_vcycle_kernel(gpu_lib, n, B):
    """Hierarchies with an 80 x 80 level hand it (and everything below) to the register-blocked V-cycle kernel; with
    `coarse_block = 0` the strips go down to 40 x 40 and the shared-memory coarse kernel takes over.  Same iteration counts to
    within one, solutions equal to round-off, 1e-10 against the reference operator; the layout query reports the level."""
    rs = np.random.RandomState(n + 5)
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(0.5, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    lvl = gpu_lib.describe(n)["first_smem_level"]
    u, info = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    gpu_lib.set_option("coarse_block", 0)
    try:
        lvl0 = gpu_lib.describe(n)["first_smem_level"]
        u0, info0 = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    finally:
        gpu_lib.set_option("coarse_block", 1)
    assert lvl0 == lvl + 1 and n >> lvl == 80
    assert info.accepted().all() and info0.accepted().all()
    assert (np.abs(info.iters - info0.iters) <= 1).all()
    assert O.rel_l2(u, u0) <= 1e-11
    for b in (0, B - 1):
        assert O.rel_l2(u[b], O.solve_poisson(np.outer(sy[b], sx[b]), theta[b], n)) <= TOL


@pytest.mark.parametrize("nu", [1, 3])
@pytest.mark.parametrize("n,B", [(160, 300), (320, 5)])
def test_other_sweep_counts_on_the_streamed_path(gpu_lib, nu, n, B):
    """V(1,1) and V(3,3) cycles instantiate different strip geometries (halo widths, ring depths, float quads);
    large and small batches exercise different band splits.  Same 1e-10 bar against the reference operator."""
    rs = np.random.RandomState(10 * n + nu)
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(0.5, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    try:
        gpu_lib.set_option("nu", nu)
        u, info = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
        gpu_lib.set_option("vcycle_f32", 0)
        u64, info64 = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    finally:
        gpu_lib.set_option("nu", 2)
        gpu_lib.set_option("vcycle_f32", 1)
    assert info.accepted().all() and info64.accepted().all()
    assert (info.iters <= info64.iters + 1).all() and info.iters.max() <= (18 if nu == 1 else 9)
    assert O.rel_l2(u, u64) <= 1e-11
    for b in (0, B - 1):
        assert O.rel_l2(u[b], O.solve_poisson(np.outer(sy[b], sx[b]), theta[b], n)) <= TOL


def test_generate_dataset_matches_reference_fixture(gpu_lib):
    from superresolution_for_pdes_b200 import PoissonSolver
    gold = np.load(os.path.join(GOLDEN, "dataset_standard_seed0.npz"))
    np.random.seed(0)
    ds = PoissonSolver().generate_dataset(6, k_range=(0.5, 5.0))
    assert list(ds.keys()) == ["u_coarse", "u_fine", "f_coarse", "f_fine", "theta_coarse", "theta_fine", "k1", "k2"]
    np.random.seed(0)
    live = O.generate_dataset(6, k_range=(0.5, 5.0))
    for k in ("k1", "k2", "f_coarse", "f_fine", "theta_coarse", "theta_fine"):
        assert ds[k].dtype == np.float64 and ds[k].flags.c_contiguous
        assert np.array_equal(ds[k], live[k]), k          # bit-exact vs the oracle on THIS machine
        # fixtures were made on another CPU: numpy's SIMD sin may differ there in the last bit
        assert np.allclose(ds[k], gold[k], rtol=0, atol=1e-15), k
    for k in ("u_coarse", "u_fine"):
        assert ds[k].shape == gold[k].shape
        for b in range(6):
            assert O.rel_l2(ds[k][b], gold[k][b]) <= TOL, (k, b)


def test_generate_subdomain_dataset_matches_reference_fixture(gpu_lib):
    from superresolution_for_pdes_b200 import EnhancedPoissonSolver
    gold = np.load(os.path.join(GOLDEN, "dataset_subdomain_seed1.npz"))
    np.random.seed(1)
    es = EnhancedPoissonSolver()
    ds = es.generate_subdomain_dataset(4, k_range=(0.5, 12.0))
    np.random.seed(1)
    live = O.generate_subdomain_dataset(4, k_range=(0.5, 12.0))
    for k in ("k1", "k2", "f_coarse", "f_fine", "theta_coarse", "theta_fine", "is_subdomain"):
        assert np.array_equal(ds[k], live[k]), k           # bit-exact vs the oracle on THIS machine
        assert np.allclose(ds[k], gold[k], rtol=0, atol=1e-15), k
    assert ds["is_subdomain"].dtype == np.bool_
    for k in ("u_coarse", "u_fine"):
        for b in range(4):
            assert O.rel_l2(ds[k][b], gold[k][b]) <= TOL, (k, b)
    assert np.array_equal(ds["u_coarse"], ds["u_fine"][:, ::2, ::2])   # downsample identity
    # combine_datasets semantics (is_subdomain added to dataset1, concatenation order)
    meta = np.load(os.path.join(GOLDEN, "dataset_combined_meta.npz"))
    np.random.seed(0)
    d1 = es.generate_dataset(6, k_range=(0.5, 5.0))
    comb = es.combine_datasets(d1, ds)
    assert "is_subdomain" in d1
    assert np.array_equal(comb["is_subdomain"], meta["is_subdomain"])
    assert np.array_equal(comb["k1"], meta["k1"]) and np.array_equal(comb["k2"], meta["k2"])
    norms = np.linalg.norm(comb["u_fine"].reshape(10, -1), axis=1)
    assert np.allclose(norms, meta["u_fine_norms"], rtol=1e-10, atol=0)


def test_subdomain_dataset_on_a_streamed_superfine_grid(gpu_lib):
    """n_superfine = 120 is too large for the resident kernel: the window / downsample epilogue then
    runs as a separate gather kernel after the streamed solve.  Same results as the oracle."""
    from superresolution_for_pdes_b200 import EnhancedPoissonSolver
    np.random.seed(11)
    es = EnhancedPoissonSolver(n_coarse=30, n_fine=60, n_superfine=120)
    ds = es.generate_subdomain_dataset(3, k_range=(0.5, 9.0))
    np.random.seed(11)
    ref = O.generate_subdomain_dataset(3, k_range=(0.5, 9.0), n_coarse=30, n_fine=60, n_superfine=120)
    for k in ("k1", "k2", "f_coarse", "f_fine", "theta_coarse", "theta_fine", "is_subdomain"):
        assert np.array_equal(ds[k], ref[k]), k
    for k in ("u_coarse", "u_fine"):
        assert ds[k].shape == ref[k].shape
        assert O.rel_l2(ds[k], ref[k]) <= TOL


def test_multires_chain_matches_reference_fixture(gpu_lib):
    from superresolution_for_pdes_b200 import solve_multi_resolution
    gold = np.load(os.path.join(GOLDEN, "multires_seed2.npz"))
    np.random.seed(2)
    data = solve_multi_resolution(40, [80, 160, 320, 640])
    assert data["k1"] == gold["k1"] and data["k2"] == gold["k2"]
    f_live = O.forcing_term_kfirst(data["k1"], data["k2"], 640)
    for res in (40, 80, 160, 320, 640):
        s = kat_sub(res)
        assert data["u"][res].shape == (res, res)
        assert np.array_equal(data["f"][res], f_live[::640 // res, ::640 // res])   # bit-exact forcing (this machine)
        assert abs(np.linalg.norm(data["f"][res]) - gold[f"f{res}_norm"]) <= 1e-13 * gold[f"f{res}_norm"]
        assert data["theta"][res].sum() == gold[f"theta{res}_sum"]
        assert O.rel_l2(data["u"][res][::s, ::s], gold[f"u{res}"]) <= TOL
        assert abs(np.linalg.norm(data["u"][res]) - gold[f"u{res}_norm"]) <= TOL * gold[f"u{res}_norm"]
    np.random.seed(2)
    timed = solve_multi_resolution(40, [80, 160], timed=True)
    assert set(timed["solve_times"]) == {40, 80, 160}


def test_structural_identities(gpu_lib):
    """F1/F2/F3 of SURVEY.md: row scaling, sign, non-zero border; linearity of the solve."""
    n = 40
    f, theta = kat_inputs(n, 1, 1)
    one = np.ones((n, n))
    u, _ = gpu_lib.solve_batch(n, 3, theta=np.stack([theta, one, one]), f=np.stack([f, f / theta, 2.5 * f]))
    assert O.rel_l2(u[0], u[1]) <= 1e-11            # solve(f, theta) == solve(f/theta, 1)
    uo, _ = gpu_lib.solve_batch(n, 1, theta=one[None], f=f[None])
    assert O.rel_l2(u[2], 2.5 * uo[0]) <= 1e-11     # linearity
    i, j = np.unravel_index(np.argmax(f), f.shape)
    assert uo[0][i, j] < 0                          # L is negative definite: u has the opposite sign of f
    assert np.abs(uo[0][0, :]).max() > 1e-6         # border nodes are unknowns, not Dirichlet rows


def test_edge_cases(gpu_lib):
    n = 20
    # zero forcing -> exact zero, converged
    u, info = gpu_lib.solve_batch(n, 2, theta=None, f=np.zeros((2, n, n)))
    assert (u == 0).all() and (info.status == 0).all()
    # empty batch
    u, info = gpu_lib.solve_batch(n, 0, theta=None, f=np.zeros((0, n, n)))
    assert u.shape == (0, n, n)
    # singular theta (reference: MatrixRankWarning + non-finite output) -> reported, never silent
    th = np.ones((1, n, n))
    th[0, 3, 3] = 0.0
    f = O.forcing_term(2.0, 3.0, n)[None]
    u, info = gpu_lib.solve_batch(n, 1, theta=th, f=f)
    assert info.status[0] >= 2
    with pytest.raises(ValueError):
        gpu_lib.solve_batch(n, 1, theta=None, f=np.zeros((1, n + 1, n)))
    from superresolution_for_pdes_b200 import PoissonSolver
    with pytest.raises(ValueError):
        PoissonSolver().solve_poisson(np.zeros((39, 40)), np.ones((40, 40)))
    with pytest.raises(gpu_lib.PsnError):
        gpu_lib.solve_batch(n, 1, theta=None, f=None)   # no forcing given


def test_device_pointers_and_device_sin(gpu_lib):
    """Device-resident inputs/outputs (torch CUDA tensors) and the on-device forcing mode."""
    import torch
    n, B = 80, 16
    rs = np.random.RandomState(11)
    k = rs.uniform(0.5, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    u_host, _ = gpu_lib.solve_batch(n, B, sin_x=sx, sin_y=sy)
    dev = torch.device("cuda:0")
    out = torch.empty((B, n, n), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    gpu_lib.solve_batch(n, B, sin_x=torch.from_numpy(sx).to(dev), sin_y=torch.from_numpy(sy).to(dev), out=out, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), u_host)      # same kernel, same data -> bit-identical
    out2 = torch.empty_like(out)
    gpu_lib.solve_batch(n, B, k1=torch.from_numpy(k[:, 0].copy()).to(dev), k2=torch.from_numpy(k[:, 1].copy()).to(dev), out=out2, stream=st)
    torch.cuda.synchronize()
    assert O.rel_l2(out2.cpu().numpy(), u_host) <= 1e-9   # device sin differs from numpy's by ulps only


def test_large_batch_properties_640(gpu_lib):
    """BASELINE-size grid, properties that need no oracle: residual of every sample, agreement
    with the closed-form DST-I solution on a subset, batch-order independence."""
    n, B = 640, 24
    rs = np.random.RandomState(3)
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(8.0, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    u, info = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    assert (info.status == 0).all(), info.status
    print("640 iters", info.iters, "relres max", info.relres.max())
    for b in range(B):
        f = np.outer(sy[b], sx[b])
        assert O.residual_rel(u[b], f, theta[b]) <= 1e-11
    for b in range(0, B, 6):
        assert O.rel_l2(u[b], O.solve_dst(np.outer(sy[b], sx[b]), theta[b])) <= TOL
    perm = rs.permutation(B)
    u2, _ = gpu_lib.solve_batch(n, B, theta=theta[perm], sin_x=sx[perm], sin_y=sy[perm])
    assert np.array_equal(u2, u[perm])    # deterministic, batch-order independent


def test_repeated_runs_are_bit_identical(gpu_lib):
    """The kernels use fixed-order reductions and no atomics on floating-point data, so repeated runs of
    the same batch must agree bit for bit -- a cheap detector of data races (missing barriers show up as
    run-to-run differences) on the register-blocked 80x80 kernel, the generic resident kernel and the
    streamed path."""
    rs = np.random.RandomState(21)
    for n, B in ((80, 1500), (40, 3000), (36, 500), (160, 300)):
        theta = rs.uniform(0.5, 2.0, size=(B, n, n))
        k = rs.uniform(0.5, 12.0, size=(B, 2))
        x = np.linspace(0, 1, n)
        sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
        u0, i0 = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
        assert i0.accepted().all()
        for _ in range(3):
            u1, i1 = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
            assert np.array_equal(u1, u0) and np.array_equal(i1.iters, i0.iters)
    # generic resident kernel (register-blocked kernel switched off) gives the same answer at 80x80
    n, B = 80, 64
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    f = np.stack([O.forcing_term(*rs.uniform(0.5, 12.0, size=2), n) for _ in range(B)])
    ub, _ = gpu_lib.solve_batch(n, B, theta=theta, f=f)
    try:
        gpu_lib.set_option("res_block", 0)
        ug, _ = gpu_lib.solve_batch(n, B, theta=theta, f=f)
    finally:
        gpu_lib.set_option("res_block", 1)
    assert O.rel_l2(ub, ug) <= 1e-13
    assert O.rel_l2(ub[0], O.solve_poisson(f[0], theta[0], n)) <= TOL


@pytest.mark.parametrize("random_theta", [False, True])
def test_dataset_tensor_prologue_matches_reference_semantics(gpu_lib, random_theta):
    """Row f1: the consumer's float32 tensors (src/models.py:132-203) built by psn_dataset_tensors against the
    torch-CPU restatement (oracle/dataset_oracle.py, itself bit-equal to the live reference class).
    Tolerance: float32 arithmetic -- 2e-6 relative on the statistics, 2e-5 absolute on the z-scored tensors."""
    import torch
    from oracle.dataset_oracle import dataset_tensors as oracle_tensors
    from superresolution_for_pdes_b200 import PDEDataset, PoissonSolver
    np.random.seed(0)
    ds = PoissonSolver().generate_dataset(64, k_range=(0.5, 5.0))
    if random_theta:
        ds = dict(ds)
        ds["theta_fine"] = np.random.RandomState(1).uniform(0.5, 2.0, size=ds["theta_fine"].shape)
    ref = oracle_tensors(ds)
    # through the C ABI with host arrays ...
    x, y, stats = gpu_lib.dataset_tensors(ds["u_coarse"], ds["u_fine"], ds["f_fine"], ds["theta_fine"])
    assert x.dtype == np.float32 and x.shape == ref["x"].shape and y.shape == ref["y"].shape
    assert np.allclose(stats[:6], ref["stats"][:6], rtol=2e-6, atol=1e-12) and stats[6] == ref["stats"][6]
    assert np.abs(x - ref["x"]).max() <= 2e-5 and np.abs(y - ref["y"]).max() <= 2e-5
    # ... and through the reference-signature class with device-resident tensors
    d = PDEDataset(ds, device="cuda:0")
    assert len(d) == 64 and d.theta_is_constant == (not random_theta)
    xi, yi = d[5]
    assert tuple(xi.shape) == (3, 40, 40) and tuple(yi.shape) == (1, 40, 40) and xi.dtype == torch.float32
    assert np.abs(xi.cpu().numpy() - ref["x"][5]).max() <= 2e-5 and np.abs(yi.cpu().numpy() - ref["y"][5]).max() <= 2e-5
    back = d.denormalize(d.u_fine_norm[3]).cpu().numpy()
    assert np.abs(back - ds["u_fine"][3]).max() <= 1e-6 * np.abs(ds["u_fine"][3]).max() + 1e-9


@pytest.mark.parametrize("mode", ["bilinear", "bicubic"])
def test_upscale_cascades_match_reference_semantics(gpu_lib, mode):
    """Row f4: psn_upscale_cascade against the torch-CPU restatement (oracle/upscale_oracle.py, bit-equal to the
    live reference functions) and against the committed golden outputs of the reference.
    Tolerance: float32 arithmetic with a different summation order -- 4 ulp of the field's maximum per level."""
    from oracle import upscale_oracle as uo
    import superresolution_for_pdes_b200 as P
    gold = np.load(os.path.join(GOLDEN, "upscale_seed2.npz"))
    u40 = np.load(os.path.join(GOLDEN, "multires_seed2.npz"))["u40"]
    data = {"u": {40: u40}}
    fn = P.bilinear_multi_level_upscale if mode == "bilinear" else P.cubic_multi_level_upscale
    scale = float(np.abs(u40).max())
    for target, levels in ((80, 1), (160, 2), (320, 3), (640, 4)):
        out = fn(data, target)
        assert out.dtype == np.float32 and out.shape == (target, target)
        tol = 4 * levels * np.finfo(np.float32).eps * scale
        assert np.abs(out - uo.multi_level_upscale(data, target, mode)).max() <= tol
        if target <= 160:
            assert np.abs(out - gold[f"{mode}_{target}"]).max() <= tol
    assert fn(data, 40) is not None and fn(data, 40).shape == (40, 40)
    # batched, odd start size, through the C ABI
    rng = np.random.RandomState(3)
    u = rng.standard_normal((5, 23, 23))
    out = gpu_lib.upscale_cascade(u, 2, mode)
    ref = uo.upscale_batch(u, 2, mode)
    assert out.shape == (5, 92, 92) and np.abs(out - ref).max() <= 16 * np.finfo(np.float32).eps * np.abs(u).max()
    # the tiling helpers of the evaluation scripts round-trip
    parts = P.split_into_subdomains(u40, 10)
    assert np.array_equal(P.stitch_subdomains(parts), u40)


def test_streamed_edge_cases_and_speculative_pipeline(gpu_lib):
    """Large-grid (streamed, float-stored, deferred-x) path through HOST buffers in several pipeline chunks:
    (a) the first chunk is easy, later chunks need more outer iterations than the count the pipeline speculates with,
        so the deferred convergence check has to resume them;  (b) zero right-hand sides and a singular theta inside
        a batch;  (c) max_iter smaller than the iterations needed: status 2, finite partial solutions."""
    n, B = 176, 300
    rs = np.random.RandomState(11)
    x = np.linspace(0, 1, n)
    # high wave numbers converge in 10 outer iterations on this grid, smooth low ones need 11-12
    k = np.concatenate([rs.uniform(9.0, 12.0, size=(80, 2)), rs.uniform(0.5, 1.5, size=(B - 80, 2))])
    theta = np.concatenate([rs.uniform(0.5, 2.0, size=(80, n, n)), np.ones((B - 80, n, n))])
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    sy[7] = 0.0                                   # zero forcing inside the batch
    theta[150, 5, 9] = 0.0                        # singular sample
    u, info = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)
    assert (u[7] == 0).all() and info.status[7] == 0 and info.iters[7] == 0
    assert info.status[150] >= 2
    good = np.ones(B, bool); good[150] = False
    assert info.accepted()[good].all() and np.isfinite(u[good]).all()
    assert info.iters[good].max() > info.iters[:80].max()          # later chunks did need more iterations
    for b in (0, 79, 80, 200, B - 1):
        assert O.rel_l2(u[b], O.solve_poisson(np.outer(sy[b], sx[b]), theta[b], n)) <= TOL
    # the same batch through device buffers in one chunk: another band split of the strip kernels (it depends on the
    # chunk size), hence another summation order of the dot products -- same answers to round-off, iteration counts
    # of borderline samples may move by one
    import torch
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ud = torch.empty((B, n, n), dtype=torch.float64, device="cuda")
    _, info_d = gpu_lib.solve_batch(n, B, theta=d(theta), sin_x=d(sx), sin_y=d(sy), out=ud)
    torch.cuda.synchronize()
    info_d = info_d.to_host()
    assert np.abs(info_d.iters[good] - info.iters[good]).max() <= 1
    assert O.rel_l2(ud.cpu().numpy()[good], u[good]) <= 1e-11
    # (c)
    u3, info3 = gpu_lib.solve_batch(n, 90, theta=theta[80:170], sin_x=sx[80:170], sin_y=sy[80:170], max_iter=4)
    g3 = np.ones(90, bool); g3[70] = False
    assert (info3.status[g3] == 2).all() and (info3.iters[g3] == 4).all() and np.isfinite(u3[g3]).all()
    assert (info3.relres[g3] > 1e-9).all() and (info3.relres[g3] < 1e-2).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n,B", [(176, 300), (80, 5000)])
def test_page_locked_callers_take_the_deep_upload_pipeline(gpu_lib, n, B):
    """Host-buffer calls on resident grids with page-locked theta / sines / u use 5 staging slots with the uploads running
    3 chunks ahead on their own stream; pageable callers and streamed grids keep the 3-slot pipeline.  Same chunking, same
    kernels: the results must be bit-identical for every slot / prefetch / upload-stream setting and for page-locked
    vs pageable buffers, including chunks that the speculative enqueue has to resume (streamed case: later chunks need
    more outer iterations than the first)."""
    import torch
    rs = np.random.RandomState(21)
    x = np.linspace(0, 1, n)
    k = np.concatenate([rs.uniform(9.0, 12.0, size=(B // 4, 2)), rs.uniform(0.5, 1.5, size=(B - B // 4, 2))])
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    theta[B // 4:B // 2] = 1.0
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    u_ref, info_ref = gpu_lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy)           # pageable numpy buffers
    assert info_ref.accepted().all()
    for b in (0, B // 3, B - 1):
        assert O.rel_l2(u_ref[b], O.solve_poisson(np.outer(sy[b], sx[b]), theta[b], n)) <= TOL
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    th_p, sx_p, sy_p = pin(theta), pin(sx), pin(sy)
    try:
        for slots, ahead, own_stream in ((5, 3, 1), (6, 3, 1), (6, 2, 0), (4, 3, 1), (3, 1, 1)):
            gpu_lib.set_option("host_slots", slots)
            gpu_lib.set_option("host_prefetch", ahead)
            gpu_lib.set_option("host_h2d_stream", own_stream)
            u_p = torch.full((B, n, n), float("nan"), dtype=torch.float64).pin_memory()
            for _ in range(2):   # the second call reuses staging slots that the first one left behind
                _, info = gpu_lib.solve_batch(n, B, theta=th_p, sin_x=sx_p, sin_y=sy_p, out=u_p)
            assert np.array_equal(u_p.numpy(), u_ref), (slots, ahead, own_stream)
            assert np.array_equal(info.iters, info_ref.iters) and np.array_equal(info.status, info_ref.status)
    finally:
        gpu_lib.set_option("host_slots", 5)
        gpu_lib.set_option("host_prefetch", 3)
        gpu_lib.set_option("host_h2d_stream", 1)
    with pytest.raises(gpu_lib.PsnError):
        gpu_lib.set_option("host_slots", 7)


@pytest.mark.gpu
def test_evaluation_script_callers_match_reference_fixture(gpu_lib, tmp_path, monkeypatch):
    """Row a10: solve_large_scale / solve_multi_level (src/subdomain_upscaling.py), generate_test_data
    (src/compare_test_cases.py, both theta modes) and generate_out_of_sample_data (src/test_out_of_sample.py)
    through the shims, against the live reference's outputs on the same seeds (k, f, theta bit-exact, u <= 1e-10)."""
    import superresolution_for_pdes_b200 as pkg
    gold = np.load(os.path.join(GOLDEN, "callers_seed5.npz"))
    monkeypatch.chdir(tmp_path)
    np.random.seed(5)
    check_caller_dict(pkg.solve_large_scale(20, 40), gold, "large_small", TOL)
    np.random.seed(5)
    d = pkg.solve_large_scale()
    check_caller_dict(d, gold, "large", TOL)
    assert np.array_equal(d["f_coarse"], d["f_fine"][::2, ::2]) and np.array_equal(d["theta_coarse"], d["theta_fine"][::2, ::2])
    np.random.seed(5)
    check_caller_dict(pkg.solve_multi_level(), gold, "level", TOL)
    with pytest.raises(ValueError):
        pkg.solve_large_scale(60, 160)
    for name, const in (("test_const", True), ("test_rand", False)):
        np.random.seed(6)
        d = pkg.generate_test_data((1.0, 6.0), 3, name, const)
        check_caller_dict(d, gold, name, TOL)
        saved = np.load(tmp_path / "data" / f"{name}_dataset.npz")
        assert set(saved.files) == set(d) and all(np.array_equal(saved[k], d[k]) for k in d)
    np.random.seed(7)
    d = pkg.generate_out_of_sample_data(3)
    check_caller_dict(d, gold, "oos", TOL)
    assert (tmp_path / "data" / "out_of_sample_dataset.npz").exists()
    state = np.random.get_state()[1].copy()
    np.random.seed(7)
    O.generate_out_of_sample_data(3)
    assert np.array_equal(state, np.random.get_state()[1])      # same RNG consumption as the reference


def test_misaligned_device_arrays_are_rejected(gpu_lib):
    """Device arrays are accessed in place with 128-bit loads / stores and bulk copies: a view with an odd element
    offset (8-byte aligned only) must come back as PSN_ERR_INVALID, not as a sticky misaligned-address fault."""
    import torch
    n, B = 160, 4
    dev = torch.device("cuda:0")
    big = torch.rand(B * n * n + 1, dtype=torch.float64, device=dev) + 0.5
    theta_odd = big[1:].view(B, n, n)
    sx = torch.rand((B, n), dtype=torch.float64, device=dev)
    out = torch.empty((B, n, n), dtype=torch.float64, device=dev)
    assert theta_odd.data_ptr() % 16 == 8 and theta_odd.is_contiguous()
    with pytest.raises(gpu_lib.PsnError, match="16-byte aligned"):
        gpu_lib.solve_batch(n, B, theta=theta_odd, sin_x=sx, sin_y=sx, out=out)
    with pytest.raises(gpu_lib.PsnError, match="16-byte aligned"):
        gpu_lib.solve_batch(n, B, theta=big[:-1].view(B, n, n), sin_x=sx, sin_y=sx, out=theta_odd)
    # the context is intact: the aligned call works and the same (misaligned) data through HOST memory is fine
    u, info = gpu_lib.solve_batch(n, B, theta=big[:-1].view(B, n, n), sin_x=sx, sin_y=sx, out=out)
    torch.cuda.synchronize()
    assert info.accepted().all()
    uh, _ = gpu_lib.solve_batch(n, B, theta=theta_odd.cpu().numpy(), sin_x=sx.cpu().numpy(), sin_y=sx.cpu().numpy())
    assert np.isfinite(uh).all()


@pytest.mark.parametrize("n,B", [(80, 600), (160, 64)])
def test_calls_on_different_streams_are_ordered(gpu_lib, n, B):
    """Device-pointer calls only enqueue work and share one workspace / staging arena: a second call on ANOTHER
    stream (and a consumer-prologue call that recycles the staging arena) must wait for the first one's kernels."""
    import torch
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(n)
    th = [torch.rand((B, n, n), dtype=torch.float64, device=dev, generator=g) * 1.5 + 0.5 for _ in range(2)]
    xg = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    sx = torch.sin(2 * torch.pi * 3.3 * xg).repeat(B, 1).contiguous()
    sy = torch.sin(2 * torch.pi * 5.1 * xg).repeat(B, 1).contiguous()
    torch.cuda.synchronize()
    ref = []
    for t in th:   # serial reference, one stream, synchronised between calls
        o = torch.empty((B, n, n), dtype=torch.float64, device=dev)
        gpu_lib.solve_batch(n, B, theta=t, sin_x=sx, sin_y=sy, out=o, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref.append(o.clone())
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = [torch.empty((B, n, n), dtype=torch.float64, device=dev) for _ in range(2)]
    for rep in range(3):
        gpu_lib.solve_batch(n, B, theta=th[0], sin_x=sx, sin_y=sy, out=outs[0], stream=s1.cuda_stream)
        gpu_lib.solve_batch(n, B, theta=th[1], sin_x=sx, sin_y=sy, out=outs[1], stream=s2.cuda_stream)
        # recycles staging slot 0 of the library while the solves above may still be running
        up = gpu_lib.upscale_cascade(ref[0][:4, ::4, ::4].contiguous(), 1, stream=s1.cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(outs[0], ref[0]) and torch.equal(outs[1], ref[1]), rep
        assert np.isfinite(up).all()


def test_multires_uploads_the_master_field_once(gpu_lib):
    """psn_multires with HOST master fields stages theta / sine factors once per chunk of examples and solves every
    resolution from that device copy: results are bit-identical to per-resolution calls on the gathered inputs'
    device path, and to a call with device-resident master fields."""
    import torch
    N, B, res = 320, 5, [40, 80, 160, 320]
    rs = np.random.RandomState(21)
    theta = rs.uniform(0.5, 2.0, size=(B, N, N))
    k = rs.uniform(1.0, 6.0, size=(B, 2))
    x = np.linspace(0, 1, N)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    us, info = gpu_lib.multires(N, res, B, theta, sx, sy)
    assert info.accepted().all()
    for r in res:
        st = N // r
        u1, _ = gpu_lib.solve_batch(r, B, theta=theta, sin_x=sx, sin_y=sy, master_n=N, step=st)
        assert np.array_equal(us[r], u1), r
        for b in (0, B - 1):
            ref = O.solve_poisson(np.outer(sy[b], sx[b])[::st, ::st], theta[b, ::st, ::st], r)
            assert O.rel_l2(us[r][b], ref) <= TOL
    # host outputs are copied back while the next resolution is solved (largest grid first): same bits as the blocking
    # per-resolution schedule, for host and for device-resident master fields
    gpu_lib.set_option("multires_pipeline", 0)
    try:
        us0, info0 = gpu_lib.multires(N, res, B, theta, sx, sy)
    finally:
        gpu_lib.set_option("multires_pipeline", 1)
    usd, infod = gpu_lib.multires(N, res, B, torch.from_numpy(theta).cuda(), sx, sy)
    for r in res:
        assert np.array_equal(us[r], us0[r]) and np.array_equal(us[r], usd[r]), r
    for i_ in (info0, infod):
        assert np.array_equal(info.iters, i_.iters) and np.array_equal(info.status, i_.status) and np.array_equal(info.relres, i_.relres)
    # theta == 1 (NULL master field) with host sines
    us1, info1 = gpu_lib.multires(N, res[:2], B, None, sx, sy)
    for r in res[:2]:
        st = N // r
        assert O.rel_l2(us1[r][1], O.solve_poisson(np.outer(sy[1], sx[1])[::st, ::st], np.ones((r, r)), r)) <= TOL


def _host_blocks(rs, n_blocks, skip, count, low, high):
    out = np.empty((n_blocks, count))
    sk = np.empty((n_blocks, skip))
    for b in range(n_blocks):
        sk[b] = rs.random_sample(skip)
        out[b] = rs.uniform(low, high, count)
    return out, sk


@pytest.mark.parametrize("pre,n_blocks,skip,count", [(0, 1, 0, 10), (3, 5, 2, 1600), (311, 7, 1, 313), (312, 3, 0, 624),
                                                      (100, 64, 2, 25600), (5, 33, 3, 6401), (17, 4, 2, 409600), (9, 300, 0, 1)])
def test_mt19937_device_stream_is_bit_identical_to_numpy(gpu_lib, pre, n_blocks, skip, count):
    """psn_mt19937_uniform_blocks replays numpy's legacy generator on the device: every value, the raw skipped doubles
    and the generator state afterwards are identical to the host loop (all block / word alignments: odd offsets make
    every double straddle a regeneration boundary)."""
    import torch
    rs = np.random.RandomState(1234 + pre)
    rs.random_sample(pre)
    ref = np.random.RandomState()
    ref.set_state(rs.get_state())
    if pre % 2 == 1:          # odd word offset: consume a single 32-bit word on both
        rs.bytes(4)
        ref.bytes(4)
    want, want_sk = _host_blocks(ref, n_blocks, skip, count, 0.5, 2.0)
    got, got_sk = gpu_lib.mt19937_uniform_blocks(n_blocks, skip, count, 0.5, 2.0, random_state=rs)
    assert np.array_equal(got, want) and np.array_equal(got_sk, want_sk)
    assert np.array_equal(rs.random_sample(700), ref.random_sample(700))        # state handed back to numpy
    # device-resident output
    dev = torch.empty((n_blocks, count), dtype=torch.float64, device="cuda")
    rs2 = np.random.RandomState(77)
    ref2 = np.random.RandomState(77)
    gpu_lib.mt19937_uniform_blocks(n_blocks, skip, count, -1.0, 3.0, out=dev, random_state=rs2)
    want2, _ = _host_blocks(ref2, n_blocks, skip, count, -1.0, 3.0)
    assert np.array_equal(dev.cpu().numpy(), want2) and rs2.uniform() == ref2.uniform()


def test_mt19937_device_stream_from_pos_zero_and_long_draws(gpu_lib):
    rs = np.random.RandomState(8)
    st = rs.get_state()
    rs.set_state((st[0], st[1], 0, 0, 0.0))
    ref = np.random.RandomState()
    ref.set_state(rs.get_state())
    want, want_sk = _host_blocks(ref, 6, 2, 1000, 0.5, 2.0)
    got, got_sk = gpu_lib.mt19937_uniform_blocks(6, 2, 1000, 0.5, 2.0, random_state=rs)
    assert np.array_equal(got, want) and np.array_equal(got_sk, want_sk)
    assert np.array_equal(rs.random_sample(700), ref.random_sample(700))
    # more blocks than one grid dimension's 65,535: 70,000 blocks of 2 + 5 doubles
    rs = np.random.RandomState(5)
    ref = np.random.RandomState(5)
    got, got_sk = gpu_lib.mt19937_uniform_blocks(70000, 2, 5, 0.5, 2.0, random_state=rs)
    flat = ref.random_sample(70000 * 7).reshape(70000, 7)            # same stream: 2 raw + 5 mapped doubles per block
    assert np.array_equal(got_sk, flat[:, :2]) and np.array_equal(got, 0.5 + 1.5 * flat[:, 2:])
    assert rs.uniform() == ref.uniform()
    # one long draw, cut into 32,768-double blocks + remainder, on the GLOBAL generator
    np.random.seed(99)
    a = gpu_lib.mt19937_uniform((37, 80, 80), 0.5, 2.0)
    nxt = np.random.uniform()
    np.random.seed(99)
    b = np.random.uniform(0.5, 2.0, (37, 80, 80))
    assert np.array_equal(a, b) and nxt == np.random.uniform()


def test_multires_batch_with_device_rng_matches_host_rng(gpu_lib):
    """solve_multi_resolution_batch draws k1, k2 and the master theta fields on the device: same values, same
    solutions and same global generator state as the host-RNG path (the reference's own order of draws)."""
    from superresolution_for_pdes_b200 import solve_multi_resolution_batch
    np.random.seed(31)
    d = solve_multi_resolution_batch(5, 40, [80, 160, 320], k_range=(8.0, 12.0))
    nxt = np.random.uniform()
    np.random.seed(31)
    h = solve_multi_resolution_batch(5, 40, [80, 160, 320], k_range=(8.0, 12.0), device_rng=False)
    assert nxt == np.random.uniform()
    assert np.array_equal(d["k1"], h["k1"]) and np.array_equal(d["k2"], h["k2"])
    assert np.array_equal(d["theta_finest"], h["theta_finest"])
    for r in (40, 80, 160, 320):
        assert np.array_equal(d["u"][r], h["u"][r]), r
    np.random.seed(31)
    ref = O.solve_multi_resolution(40, [80, 160, 320], k_lo=8.0, k_hi=12.0)       # oracle = the reference's own loop body
    assert ref["k1"] == d["k1"][0] and ref["k2"] == d["k2"][0] and np.array_equal(ref["theta"][320], d["theta_finest"][0])
    for r in (40, 80, 160, 320):
        assert O.rel_l2(d["u"][r][0], ref["u"][r]) <= TOL
