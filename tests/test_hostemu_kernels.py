"""CPU tier: the CUDA kernel BODIES (psn_core.cuh) and the streamed-path driver compiled as host
code with one virtual thread per CTA (tests/hostemu -- test infrastructure, never loaded by the
product) against the oracle.  This checks indexing, halo widths, the temporal-blocking schedule,
the multigrid hierarchy for odd/even sizes and the per-sample state machine without a GPU;
data races are the one thing it cannot see (compute-sanitizer on the GPU box covers those)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import KAT_K, ROOT, kat_inputs
from oracle import poisson_oracle as O

HE_DIR = os.path.join(ROOT, "tests", "hostemu")
HE_LIB = os.path.join(HE_DIR, "libpsn_hostemu.so")
DP = ctypes.POINTER(ctypes.c_double)
IP = ctypes.POINTER(ctypes.c_int32)


@pytest.fixture(scope="module")
def he():
    srcs = [os.path.join(HE_DIR, "psn_hostemu.cpp")] + [
        os.path.join(ROOT, "superresolution_for_pdes_b200", "csrc", f)
        for f in ("psn_core.cuh", "psn_plan.h", "psn_stream_driver.h", "psn_warp.cuh", "psn_block.cuh")]
    if not os.path.exists(HE_LIB) or any(os.path.getmtime(s) > os.path.getmtime(HE_LIB) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", HE_LIB, srcs[0]], check=True)
    return ctypes.CDLL(HE_LIB)


def P(a):
    return None if a is None else a.ctypes.data_as(DP)


def PI(a):
    return None if a is None else a.ctypes.data_as(IP)


def run_smem(he, n, theta, f=None, sx=None, sy=None, rtol=1e-12, maxit=100, nu=2, sub=None):
    B = (f if f is not None else sx).shape[0]
    u = np.full((B, n, n), np.nan)
    it, rr, st = np.zeros(B, np.int32), np.zeros(B), np.zeros(B, np.int32)
    extra = [ctypes.c_int(0), None, None, None, None, None, None]
    outs = None
    if sub is not None:
        s, x0, y0 = sub
        sc = (s + 1) // 2
        outs = [np.full((B, s, s), np.nan), np.full((B, s, s), np.nan), np.full((B, sc, sc), np.nan), np.full((B, sc, sc), np.nan)]
        extra = [ctypes.c_int(s), PI(x0), PI(y0)] + [P(o) for o in outs]
    rc = he.hostemu_solve_smem(ctypes.c_int(n), ctypes.c_longlong(B), P(theta), P(f), P(sx), P(sy), None, None,
                               ctypes.c_int(n), ctypes.c_int(1), ctypes.c_double(rtol), ctypes.c_int(maxit), ctypes.c_int(nu),
                               P(u), PI(it), P(rr), PI(st), *extra)
    assert rc == 0
    return u, it, rr, st, outs


def run_stream(he, n, theta, f, TH, TW, ls, rtol=1e-12, maxit=100, nu=2, mn=None, step=1, mode=0, band=64, sx=None, sy=None):
    B = f.shape[0] if f is not None else sx.shape[0]
    u = np.full((B, n, n), np.nan)
    it, rr, st = np.zeros(B, np.int32), np.zeros(B), np.zeros(B, np.int32)
    rc = he.hostemu_solve_stream(ctypes.c_int(n), ctypes.c_longlong(B), P(theta), P(f), P(sx), P(sy), None, None,
                                 ctypes.c_int(mn or n), ctypes.c_int(step), ctypes.c_double(rtol), ctypes.c_int(maxit), ctypes.c_int(nu),
                                 ctypes.c_int(TH), ctypes.c_int(TW), ctypes.c_int(ls), ctypes.c_int(mode), ctypes.c_int(band), P(u), PI(it), P(rr), PI(st))
    assert rc > 0
    return u, it, rr, st


def batch_inputs(n, rough):
    fs, ths = zip(*[kat_inputs(n, k, rough) for k in range(3)])
    return np.stack(fs), np.stack(ths)


def test_plan_hierarchy():
    from superresolution_for_pdes_b200 import _lib
    he = ctypes.CDLL(HE_LIB) if os.path.exists(HE_LIB) else None
    if he is None:
        pytest.skip("host emulation library not built")
    ns = (ctypes.c_int * 12)()
    g = (ctypes.c_double * 12)()
    L = he.hostemu_plan(ctypes.c_int(640), ctypes.c_int(2), ns, g, None)
    assert list(ns)[:L] == [640, 320, 160, 80, 40, 20, 10]         # 10 x 10 = the dense-inverse level (PSN_COARSEST_MAX)
    assert list(g)[:L] == [0, 1, 3, 7, 15, 31, 63]               # gam' = 2 gam + 1 on even levels
    L = he.hostemu_plan(ctypes.c_int(37), ctypes.c_int(2), ns, g, None)
    assert list(ns)[:L] == [37, 18, 9] and list(g)[:L] == [0, 0, 1]
    L = he.hostemu_plan(ctypes.c_int(20), ctypes.c_int(2), ns, g, None)
    assert list(ns)[:L] == [20, 10, 5]                             # grids below 32 x 32 keep the 6 x 6 dense-inverse limit
    L = he.hostemu_plan(ctypes.c_int(45), ctypes.c_int(2), ns, g, None)
    assert list(ns)[:L] == [45, 22, 11, 5] and list(g)[:L] == [0, 0, 1, 0.5]   # gam' = gam / 2 on odd levels


@pytest.mark.parametrize("n", [5, 6, 7, 9, 13, 20, 37, 40, 75, 80, 86])
@pytest.mark.parametrize("rough", [0, 1])
def test_resident_body_matches_oracle(he, n, rough):
    f, theta = batch_inputs(n, rough)
    u, it, rr, st, _ = run_smem(he, n, theta, f=f)
    assert (st <= 1).all() and it.max() <= 12
    for b in range(3):
        assert O.rel_l2(u[b], O.solve_poisson(f[b], theta[b], n)) <= 1e-10


def test_resident_body_fused_subdomain_epilogue(he):
    n, s = 80, 40
    x = np.linspace(0, 1, n)
    k = np.array([[0.5, 11.0], [7.7, 3.1], [12.0, 12.0]])
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    x0 = np.array([0, 39, 17], dtype=np.int32)
    y0 = np.array([39, 0, 5], dtype=np.int32)
    u, it, rr, st, (uf, ff, uc, fc) = run_smem(he, n, None, sx=sx, sy=sy, sub=(s, x0, y0))
    for b in range(3):
        f = np.outer(sy[b], sx[b])
        win = (slice(y0[b], y0[b] + s), slice(x0[b], x0[b] + s))      # rows = y, cols = x
        assert np.array_equal(uf[b], u[b][win]) and np.array_equal(ff[b], f[win])
        assert np.array_equal(uc[b], u[b][win][::2, ::2]) and np.array_equal(fc[b], f[win][::2, ::2])
        assert O.rel_l2(u[b], O.solve_poisson(f, np.ones((n, n)), n)) <= 1e-10


@pytest.mark.parametrize("n,TH,TW,ls", [(96, 32, 32, 1), (100, 64, 64, 1), (131, 64, 64, 1), (160, 64, 64, 1),
                                        (160, 32, 64, 2), (64, 16, 16, 2), (45, 8, 24, 1)])
def test_streamed_bodies_match_oracle(he, n, TH, TW, ls):
    f, theta = batch_inputs(n, 1)
    u, it, rr, st = run_stream(he, n, theta, f, TH, TW, ls)
    assert (st <= 1).all() and it.max() <= 12
    for b in range(3):
        assert O.rel_l2(u[b], O.solve_poisson(f[b], theta[b], n)) <= 1e-10


@pytest.mark.parametrize("n,ls,band,nu", [(96, 1, 96, 2), (100, 1, 34, 2), (131, 1, 64, 2), (160, 2, 48, 2), (160, 1, 64, 1),
                                           (160, 2, 32, 3), (45, 1, 16, 2), (200, 2, 200, 2)])
def test_warp_strip_bodies_match_oracle(he, n, ls, band, nu):
    """Register-resident warp-strip V-cycle legs (mode 1), emulated with 32-lane vectors."""
    f, theta = batch_inputs(n, 1)
    u, it, rr, st = run_stream(he, n, theta, f, 64, 64, ls, mode=1, band=band, nu=nu)
    u0, it0, _, _ = run_stream(he, n, theta, f, 64, 64, ls, mode=0, nu=nu)
    u2, it2, _, _ = run_stream(he, n, theta, f, 64, 64, ls, mode=2, band=band, nu=nu)
    assert (st <= 1).all() and (it == it0).all()
    assert O.rel_l2(u, u0) <= 1e-13                      # same algorithm as the tile kernels
    # the ring-fed mode runs the same V-cycle arithmetic; its strip update kernel recomputes q = At p (bit-identical
    # r) but sums <p,q> and ||r||^2 in another order, so alpha differs in the last bits
    assert (it2 == it).all() and O.rel_l2(u2, u) <= 1e-13
    for b in range(3):
        assert O.rel_l2(u[b], O.solve_poisson(f[b], theta[b], n)) <= 1e-10


def test_warp_strip_vcycle_equals_resident_vcycle(he):
    """One V-cycle through the streamed strip kernels == the shared-memory V-cycle, node by node."""
    for n, nu, ls, band in ((96, 2, 1, 96), (160, 1, 1, 64), (132, 3, 1, 40), (256, 2, 2, 86), (46, 2, 1, 16), (640, 2, 3, 80)):
        r = np.random.RandomState(n).randn(n, n)
        zr = np.zeros((n, n))
        he.hostemu_vcycle(ctypes.c_int(n), ctypes.c_int(nu), P(r), P(zr))
        for mode in (0, 1, 2):
            z = np.full((n, n), np.nan)
            rz = np.zeros(1)
            rc = he.hostemu_stream_vcycle(ctypes.c_int(n), ctypes.c_int(nu), ctypes.c_int(64), ctypes.c_int(64), ctypes.c_int(ls),
                                          ctypes.c_int(mode), ctypes.c_int(band), P(r), P(z), P(rz))
            assert rc == 0
            assert np.abs(z - zr).max() <= 1e-12 * np.abs(zr).max()
            assert abs(rz[0] - (r * zr).sum()) <= 1e-12 * abs(rz[0])


def test_streamed_tiling_is_exact(he):
    """Temporal blocking must reproduce the untiled algorithm: different tile shapes give the same
    iterates up to summation order of the dot products."""
    n = 96
    f, theta = batch_inputs(n, 1)
    ua, ita, _, _ = run_stream(he, n, theta, f, 96, 96, 1)
    ub, itb, _, _ = run_stream(he, n, theta, f, 16, 32, 1)
    assert (ita == itb).all()
    assert O.rel_l2(ub, ua) <= 1e-13


def test_stride_gather_and_state_machine(he):
    n, mn, step = 100, 199, 2
    rs = np.random.RandomState(0)
    thm = rs.uniform(0.5, 2.0, size=(4, mn, mn))
    fm = np.stack([O.forcing_term(3.3, 4.4, mn), np.zeros((mn, mn)), O.forcing_term(0.5, 0.6, mn), O.forcing_term(9.0, 2.0, mn)])
    u, it, rr, st = run_stream(he, n, thm, fm, 64, 64, 1, mn=mn, step=step)
    assert st[1] == 0 and it[1] == 0 and (u[1] == 0).all()          # zero rhs finishes at once, others continue
    for b in (0, 2, 3):
        assert st[b] <= 1
        assert O.rel_l2(u[b], O.solve_poisson(fm[b][::2, ::2], thm[b][::2, ::2], n)) <= 1e-10
    # max_iter is honoured per sample and reported
    u, it, rr, st = run_stream(he, n, thm, fm, 64, 64, 1, mn=mn, step=step, maxit=3)
    assert (it[[0, 2, 3]] == 3).all() and (st[[0, 2, 3]] == 2).all()


def test_fp64_floor_status(he):
    """Very smooth forcing on a larger grid: rtol below the fp64 round-off floor ends with status 1
    (stagnation detected), with an error as small as the direct solver's."""
    n = 160
    f, theta = kat_inputs(n, 0, 0)
    u, it, rr, st = run_stream(he, n, theta[None], f[None], 64, 64, 1, rtol=1e-14)
    assert st[0] == 1 and it[0] < 30
    assert O.rel_l2(u[0], O.solve_poisson(f, theta, n)) <= 1e-10


@pytest.mark.parametrize("n,nu", [(80, 1), (80, 2), (80, 3), (40, 1), (40, 2), (40, 3), (20, 1), (20, 2), (20, 3)])
def test_register_blocked_resident_solver_matches_generic(he, n, nu):
    """psn_block.cuh (80x80 and 40x40: 4x4 / 2x2 / 1 node(s) per thread on levels 0 / 1 / 2 in registers, edge buffers in
    shared memory), emulated with the state of all 400 / 100 threads, against the generic shared-memory solver
    and the oracle."""
    f, theta = batch_inputs(n, 1)
    B = f.shape[0]
    u = np.full((B, n, n), np.nan)
    it, rr, st = np.zeros(B, np.int32), np.zeros(B), np.zeros(B, np.int32)
    w = n // 2
    x0 = np.array([0, w - 1, min(17, w - 3)], dtype=np.int32)
    y0 = np.array([w - 1, 0, 5], dtype=np.int32)
    uf, ff = np.full((B, w, w), np.nan), np.full((B, w, w), np.nan)
    uc, fc = np.full((B, w // 2, w // 2), np.nan), np.full((B, w // 2, w // 2), np.nan)
    rc = he.hostemu_solve_block(ctypes.c_int(n), ctypes.c_longlong(B), P(theta), P(f), None, None, ctypes.c_int(n), ctypes.c_int(1),
                                ctypes.c_double(1e-12), ctypes.c_int(100), ctypes.c_int(nu), P(u), PI(it), P(rr), PI(st),
                                ctypes.c_int(w), PI(x0), PI(y0), P(uf), P(ff), P(uc), P(fc))
    assert rc == 0
    ug, itg, _, _, _ = run_smem(he, n, theta, f=f, nu=nu)
    assert (it == itg).all() and (st == 0).all()
    assert O.rel_l2(u, ug) <= 1e-14
    for b in range(B):
        assert O.rel_l2(u[b], O.solve_poisson(f[b], theta[b], n)) <= 1e-10
        win = (slice(y0[b], y0[b] + w), slice(x0[b], x0[b] + w))
        assert np.array_equal(uf[b], u[b][win]) and np.array_equal(ff[b], f[b][win])
        assert np.array_equal(uc[b], u[b][win][::2, ::2]) and np.array_equal(fc[b], f[b][win][::2, ::2])


@pytest.mark.parametrize("n,nu", [(80, 1), (80, 2), (40, 2), (20, 2)])
def test_register_blocked_solver_phases_are_race_free(he, n, nu):
    """Every phase between two barriers of psn_block.cuh must not depend on the order of its threads: levels 1 and 2 exchange
    through planes of their own so that 'read the other level's edges' and 'publish mine' share a barrier interval, the float
    planes alias the fp64 ones, and the ghost rings are re-zeroed in passing.  The emulation runs the threads of a phase one
    after the other; forwards and backwards must give the same bits."""
    f, theta = batch_inputs(n, 1)
    B = f.shape[0]
    x0 = np.array([0, 39, 17], dtype=np.int32)
    y0 = np.array([39, 0, 5], dtype=np.int32)
    res = []
    for rev in (0, 1):
        he.hostemu_block_thread_order(ctypes.c_int(rev))
        u = np.full((B, n, n), np.nan)
        it, rr, st = np.zeros(B, np.int32), np.zeros(B), np.zeros(B, np.int32)
        try:
            rc = he.hostemu_solve_block(ctypes.c_int(n), ctypes.c_longlong(B), P(theta), P(f), None, None, ctypes.c_int(n), ctypes.c_int(1),
                                        ctypes.c_double(1e-12), ctypes.c_int(100), ctypes.c_int(nu), P(u), PI(it), P(rr), PI(st),
                                        ctypes.c_int(0), PI(x0), PI(y0), None, None, None, None)
        finally:
            he.hostemu_block_thread_order(ctypes.c_int(0))
        assert rc == 0 and (st == 0).all()
        res.append((u, it, rr))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])


@pytest.mark.parametrize("n,ls,band,nu", [(96, 1, 96, 2), (160, 2, 48, 2), (160, 1, 64, 1), (160, 2, 32, 3), (320, 3, 64, 2), (200, 1, 200, 2)])
def test_float_storage_vcycle_is_as_good_a_preconditioner(he, n, ls, band, nu):
    """Mode 3 = TMA-fed strips with the V-cycle intermediates STORED as float (arithmetic, CG vectors and dot
    products fp64).  One V-cycle agrees with the fp64 V-cycle to float rounding, CG needs the same number of
    iterations and the solution meets the same 1e-10 bar against the reference operator."""
    r = np.random.RandomState(n).randn(n, n) * 3.7e-9          # arbitrary magnitude: the kernels rescale by 2^k
    zr = np.zeros((n, n))
    he.hostemu_vcycle(ctypes.c_int(n), ctypes.c_int(nu), P(r), P(zr))
    z = np.full((n, n), np.nan)
    rz = np.zeros(1)
    rc = he.hostemu_stream_vcycle(ctypes.c_int(n), ctypes.c_int(nu), ctypes.c_int(64), ctypes.c_int(64), ctypes.c_int(ls),
                                  ctypes.c_int(3), ctypes.c_int(band), P(r), P(z), P(rz))
    assert rc == 0
    # no per-sample state in this unit call => scale 1: the float copy of z carries ~6e-8 relative rounding per level
    assert np.abs(z - zr).max() <= 2e-6 * np.abs(zr).max()
    assert abs(rz[0] - (r * zr).sum()) <= 2e-6 * abs(rz[0])
    f, theta = batch_inputs(n, 1)
    u, it, rr, st = run_stream(he, n, theta, f, 64, 64, ls, mode=3, band=band, nu=nu)
    u2, it2, _, _ = run_stream(he, n, theta, f, 64, 64, ls, mode=2, band=band, nu=nu)
    assert (st <= 1).all() and (it <= it2 + 1).all() and it.sum() <= it2.sum() + 1
    for b in range(3):
        assert O.rel_l2(u[b], O.solve_poisson(f[b], theta[b], n)) <= 1e-10
    # data of extreme magnitude: the power-of-two rescaling keeps the float copies in range
    for scale in (1e-30, 1e+30):
        us, its, _, sts = run_stream(he, n, theta[:1], f[:1] * scale, 64, 64, ls, mode=3, band=band, nu=nu)
        assert (sts <= 1).all() and its[0] <= it2[0] + 1
        assert O.rel_l2(us[0] / scale, u2[0]) <= 1e-10


def test_float_storage_needs_multiples_of_four(he):
    n = 100   # levels 100, 50: 50 % 4 != 0 on a streamed level
    r = np.zeros((n, n)); z = np.zeros((n, n)); rz = np.zeros(1)
    rc = he.hostemu_stream_vcycle(ctypes.c_int(n), ctypes.c_int(2), ctypes.c_int(64), ctypes.c_int(64), ctypes.c_int(2),
                                  ctypes.c_int(3), ctypes.c_int(64), P(r), P(z), P(rz))
    assert rc == -2


@pytest.mark.parametrize("n,ls,band", [(96, 1, 96), (160, 2, 48), (200, 2, 64), (256, 2, 128)])
def test_strip_update_and_strip_true_residual_kernels(he, n, ls, band):
    """Mode 2 with the product's call form (sine factors + theta): the CG update recomputes q = A p in strips (q is
    never stored) and the final true-residual pass runs in strips too.  Same iterates as mode 1 (2-D tile update that
    reads q, tile true residual) up to the summation order of the dot products; 1e-10 against the reference operator."""
    rs = np.random.RandomState(n)
    B = 3
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(0.5, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    u2, it2, rr2, st2 = run_stream(he, n, theta, None, 64, 64, ls, mode=2, band=band, sx=sx, sy=sy)
    u1, it1, rr1, st1 = run_stream(he, n, theta, None, 64, 64, ls, mode=1, band=band, sx=sx, sy=sy)
    assert (st2 <= 1).all() and (it2 == it1).all()
    assert O.rel_l2(u2, u1) <= 1e-13 and np.allclose(rr2, rr1, rtol=1e-6)
    for b in range(B):
        f = np.outer(sy[b], sx[b])
        assert O.rel_l2(u2[b], O.solve_poisson(f, theta[b], n)) <= 1e-10
        assert O.residual_rel(u2[b], f, theta[b]) <= 1e-11        # the reported convergence is the true residual's


def test_deferred_solution_update_restart_and_full_ring(he):
    """Mode 3 accumulates x from the ring of stored float directions only when a sample is checked.  (a) a failed
    check (rtol below the fp64 floor) restarts from the flushed x and ends with status 1; (b) V(1,1) needs more outer
    iterations than the ring has slots (12), so the ring-full flush runs mid-solve."""
    n = 160
    f, theta = kat_inputs(n, 0, 0)
    u, it, rr, st = run_stream(he, n, theta[None], f[None], 64, 64, 1, rtol=1e-14, mode=3, band=64)
    assert st[0] == 1 and it[0] < 30
    assert O.rel_l2(u[0], O.solve_poisson(f, theta, n)) <= 1e-10
    fb, thb = batch_inputs(n, 1)
    u1, it1, _, st1 = run_stream(he, n, thb, fb, 64, 64, 2, mode=3, band=48, nu=1, rtol=1e-13)
    assert (st1 <= 1).all() and it1.max() > 12
    for b in range(3):
        assert O.rel_l2(u1[b], O.solve_poisson(fb[b], thb[b], n)) <= 1e-10


@pytest.mark.parametrize("n,ls,band,nu", [(120, 2, 16, 2), (160, 2, 48, 2), (160, 1, 64, 2), (256, 2, 86, 2), (200, 1, 200, 2), (160, 2, 64, 1), (176, 1, 32, 3)])
def test_fused_update_and_down_leg(he, n, ls, band, nu):
    """Mode 4 = mode 3 with the CG update of iteration k and the level-0 down-leg of iteration k+1 in ONE strip kernel
    (double-buffered residual, p window one row ahead, V-cycle scale taken from the previous residual norm).  Same
    iteration counts as the unfused schedule, solutions equal to round-off, 1e-10 against the reference operator --
    for band splits with halo re-computation, several strips (n > 112) and V(1,1) / V(3,3) windows."""
    rs = np.random.RandomState(n + nu)
    B = 3
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(0.5, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    he.hostemu_fused_launches.restype = ctypes.c_longlong
    before = he.hostemu_fused_launches()
    u4, it4, rr4, st4 = run_stream(he, n, theta, None, 64, 64, ls, mode=4, band=band, nu=nu, sx=sx, sy=sy)
    assert he.hostemu_fused_launches() - before >= it4.max()        # one fused launch per outer iteration
    u3, it3, rr3, st3 = run_stream(he, n, theta, None, 64, 64, ls, mode=3, band=band, nu=nu, sx=sx, sy=sy)
    assert he.hostemu_fused_launches() - before < 2 * it4.max() + 8   # ... and none in the unfused mode
    assert (st4 <= 1).all() and (st3 <= 1).all()
    assert (np.abs(it4 - it3) <= 1).all() and it4.sum() <= it3.sum() + 1
    assert O.rel_l2(u4, u3) <= 1e-11
    # mode 5: the prologue (r_0 = bt from theta and the sine factors) runs inside the fused kernel too; its V-cycle scale is an
    # estimate (||sin_x|| ||sin_y|| h^2), a power of two like the exact one, so nothing changes in the iterates
    he.hostemu_fused_prologues.restype = ctypes.c_longlong
    pro = he.hostemu_fused_prologues()
    u5, it5, rr5, st5 = run_stream(he, n, theta, None, 64, 64, ls, mode=5, band=band, nu=nu, sx=sx, sy=sy)
    assert he.hostemu_fused_prologues() == pro + 1
    assert (st5 <= 1).all() and (it5 == it4).all() and O.rel_l2(u5, u4) <= 1e-12
    for b in range(B):
        f = np.outer(sy[b], sx[b])
        assert O.rel_l2(u4[b], O.solve_poisson(f, theta[b], n)) <= 1e-10
        assert O.residual_rel(u4[b], f, theta[b]) <= 1e-11


@pytest.mark.parametrize("n,ls,band,nu", [(160, 1, 64, 2), (320, 2, 64, 2), (160, 1, 48, 1)])
def test_coarse_tail_in_the_register_blocked_vcycle(he, n, ls, band, nu):
    """Mode 6 = mode 5 with the 80 x 80 level and everything below it in the register-blocked V-cycle of psn_block.cuh (float,
    level 0 of the block code carrying the boundary weight of a coarse level) instead of strips down to 40 x 40 plus the
    shared-memory coarse kernel: same iteration counts to within one, 1e-10 against the reference operator."""
    rs = np.random.RandomState(n + nu + 7)
    B = 2
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(0.5, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    u6, it6, rr6, st6 = run_stream(he, n, theta, None, 64, 64, ls, mode=6, band=band, nu=nu, sx=sx, sy=sy)
    u5, it5, rr5, st5 = run_stream(he, n, theta, None, 64, 64, ls + 1, mode=5, band=band, nu=nu, sx=sx, sy=sy)
    assert (st6 == 0).all() and (st5 == 0).all()
    assert (np.abs(it6 - it5) <= 1).all()
    assert O.rel_l2(u6, u5) <= 1e-11
    for b in range(B):
        f = np.outer(sy[b], sx[b])
        assert O.rel_l2(u6[b], O.solve_poisson(f, theta[b], n)) <= 1e-10
        assert O.residual_rel(u6[b], f, theta[b]) <= 1e-11


def test_fused_update_restart_full_ring_and_scales(he):
    """Fused schedule through its rare paths: (a) a failed check restarts from the recomputed residual (the plain
    down-leg then runs for that sample only); (b) more outer iterations than ring slots; (c) data of extreme magnitude
    (the V-cycle scale of iteration k+1 comes from ||r_k||)."""
    n = 160
    f, theta = kat_inputs(n, 0, 0)
    u, it, rr, st = run_stream(he, n, theta[None], f[None], 64, 64, 1, rtol=1e-14, mode=4, band=64)
    u3, it3, _, st3 = run_stream(he, n, theta[None], f[None], 64, 64, 1, rtol=1e-14, mode=3, band=64)
    assert st[0] == 1 and it[0] < 30 and abs(int(it[0]) - int(it3[0])) <= 2
    assert O.rel_l2(u[0], O.solve_poisson(f, theta, n)) <= 1e-10
    fb, thb = batch_inputs(n, 1)
    u1, it1, _, st1 = run_stream(he, n, thb, fb, 64, 64, 2, mode=4, band=48, nu=1, rtol=1e-13)
    assert (st1 <= 1).all() and it1.max() > 12
    for b in range(3):
        assert O.rel_l2(u1[b], O.solve_poisson(fb[b], thb[b], n)) <= 1e-10
    u2, it2, _, _ = run_stream(he, n, thb, fb, 64, 64, 2, mode=3, band=48)
    for scale in (1e-30, 1e+30):
        us, its, _, sts = run_stream(he, n, thb[:1], fb[:1] * scale, 64, 64, 2, mode=4, band=48)
        assert (sts <= 1).all() and its[0] <= it2[0] + 1
        assert O.rel_l2(us[0] / scale, u2[0]) <= 1e-10
    # fused prologue with sine factors of extreme magnitude, theta far from 1 (the scale estimate ignores theta) and a zero forcing
    rs = np.random.RandomState(3)
    x = np.linspace(0, 1, n)
    sx = np.stack([np.sin(2 * np.pi * 3.3 * x) * 1e-20, np.sin(2 * np.pi * 1.7 * x) * 1e18, np.zeros(n)])
    sy = np.stack([np.sin(2 * np.pi * 5.1 * x) * 1e-15, np.sin(2 * np.pi * 2.2 * x) * 1e10, np.sin(2 * np.pi * x)])
    th = np.stack([rs.uniform(1e-3, 2e-3, (n, n)), rs.uniform(50.0, 2000.0, (n, n)), rs.uniform(0.5, 2.0, (n, n))])
    u5, it5, _, st5 = run_stream(he, n, th, None, 64, 64, 2, mode=5, band=48, sx=sx, sy=sy)
    u3, it3, _, st3 = run_stream(he, n, th, None, 64, 64, 2, mode=3, band=48, sx=sx, sy=sy)
    assert (st5 <= 1).all() and (it5 == it3).all() and (u5[2] == 0).all()
    for b in range(2):
        assert O.rel_l2(u5[b], u3[b]) <= 1e-11
