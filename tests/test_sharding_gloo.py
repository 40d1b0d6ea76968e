"""CPU tier: the N>1 path (sample sharding + final gather) with world_size 2 over gloo.  The solve
itself is replaced by the oracle here (test-only injection) -- the sharding logic is what is tested."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT
from superresolution_for_pdes_b200.sharding import shard_bounds


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 7, 8, 9, 4096, 10000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(total, world, r)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))
    assert shard_bounds(4096, 8, 3) == (1536, 2048)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _oracle_solve(n, batch, *, theta=None, sin_x=None, sin_y=None, **kw):
    from oracle import poisson_oracle as O
    from superresolution_for_pdes_b200 import _lib
    u = np.empty((batch, n, n))
    for b in range(batch):
        f = np.outer(sin_y[b], sin_x[b])
        u[b] = O.solve_poisson(f, np.ones_like(f) if theta is None else theta[b], n)
    z = np.zeros(batch, dtype=np.int32)
    return u, _lib.SolveInfo(z + 9, np.zeros(batch), z)


def _oracle_sub(n_super, n_fine, batch, sx, sy, x0, y0):
    from oracle import poisson_oracle as O
    from superresolution_for_pdes_b200 import _lib
    out = {k: [] for k in ("u_fine", "f_fine", "theta_fine", "u_coarse", "f_coarse", "theta_coarse")}
    for b in range(batch):
        f = np.outer(sy[b], sx[b])
        th = np.ones_like(f)
        u = O.solve_poisson(f, th, n_super)
        for name, fld in (("u", u), ("f", f), ("theta", th)):
            w = fld[y0[b]:y0[b] + n_fine, x0[b]:x0[b] + n_fine]
            out[name + "_fine"].append(w)
            out[name + "_coarse"].append(w[::2, ::2])
    shapes = {"fine": (0, n_fine, n_fine), "coarse": (0, (n_fine + 1) // 2, (n_fine + 1) // 2)}
    res = {k: (np.array(v) if v else np.empty(shapes[k.split("_")[1]])) for k, v in out.items()}
    z = np.zeros(batch, dtype=np.int32)
    return res, _lib.SolveInfo(z, np.zeros(batch), z)


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from superresolution_for_pdes_b200 import EnhancedPoissonSolver
    from superresolution_for_pdes_b200 import sharding as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        es = EnhancedPoissonSolver()
        np.random.seed(0)
        ds = S.generate_dataset_sharded(es, 5, k_range=(0.5, 5.0), solve_fn=_oracle_solve)
        np.random.seed(1)
        sd = S.generate_subdomain_dataset_sharded(es, 3, k_range=(0.5, 12.0), sub_fn=_oracle_sub)
        x = np.linspace(0, 1, 24)
        k = np.random.RandomState(4).uniform(1, 5, size=(5, 2))
        sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
        u, iters, status = S.solve_batch_sharded(24, None, sx, sy, dst=0, solve_fn=_oracle_solve)
        if rank == 0:
            np.savez(os.path.join(tmp, "r0.npz"), u=u, iters=iters, **{"ds_" + k_: v for k_, v in ds.items()},
                     **{"sd_" + k_: v for k_, v in sd.items()})
        else:
            assert u is None
            np.savez(os.path.join(tmp, f"r{rank}.npz"), **{"ds_" + k_: v for k_, v in ds.items()})
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process(tmp_path):
    from oracle import poisson_oracle as O
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "r0.npz")
    r1 = np.load(tmp_path / "r1.npz")
    np.random.seed(0)
    ref = O.generate_dataset(5, k_range=(0.5, 5.0))
    for k in ref:
        assert np.array_equal(r0["ds_" + k], ref[k]), k        # same data as the unsharded reference loop
        assert np.array_equal(r1["ds_" + k], ref[k]), k        # ... on every rank
    np.random.seed(1)
    refs = O.generate_subdomain_dataset(3, k_range=(0.5, 12.0))
    for k in refs:
        assert np.array_equal(r0["sd_" + k], refs[k]), k
    assert r0["u"].shape == (5, 24, 24) and (r0["iters"] == 9).all()
