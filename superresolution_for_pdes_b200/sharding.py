"""Multi-GPU sharding of the sample batch: one process per GPU (torch.distributed), contiguous
index ranges per rank, NO collective on the solve path -- samples are independent
(reference loops: src/data_generation.py:128-153, src/enhanced_data_generation.py:121-159) --
and one final gather of the per-rank slices.

Every rank consumes the global ``np.random`` stream for the WHOLE batch exactly as the reference
would (so the stream ends in the same state on every rank and results do not depend on the world
size), then solves only its own slice.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np

from . import _lib
from .data_generation import PoissonSolver, sine_factors
from .enhanced_data_generation import EnhancedPoissonSolver


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`: ceil(total / world) samples per rank, the last
    ranks may be short or empty (NVSwitch is uniform, so no topology awareness is needed)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world / rank")
    per = -(-total // world)
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def bind_to_gpu_numa(device_index: int) -> Optional[list]:
    """Pin the calling process to the CPU cores that are local to GPU `device_index` (NVML's ideal affinity), so that
    the pinned staging buffers it allocates afterwards and its copy threads sit on that GPU's NUMA node.  With one
    process per GPU on a two-socket host this keeps the 3+ GB/step of host<->device traffic per GPU off the
    inter-socket link.  Returns the core list, or None when NVML / the cpuset does not allow it (never raises)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = [c for c in cores if c in allowed]
        if not cores or len(cores) == len(allowed):
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


def _dist():
    import torch.distributed as dist
    return dist


def world_info(group=None) -> Tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_slices(local: np.ndarray, total: int, group=None, dst: Optional[int] = None) -> Optional[np.ndarray]:
    """Assemble the full [total, ...] array from every rank's contiguous slice.  dst=None: every rank
    gets it (all_gather); otherwise only rank `dst` does.  NCCL groups move the data GPU-to-GPU over
    NVLink, gloo groups (CPU tests) over sockets."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return local
    per = -(-total // world)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad).to(dev)
    if dst is None:
        out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t, group=group) if hasattr(dist, "all_gather_into_tensor") and backend == "nccl" \
            else dist.all_gather(list(out.view((world, per) + tuple(local.shape[1:])).unbind(0)), t, group=group)
        return out.cpu().numpy()[:total]
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat(bufs, 0).cpu().numpy()[:total]


def generate_dataset_sharded(solver: PoissonSolver, n_samples: int, k_range=(1, 5), group=None,
                             solve_fn: Optional[Callable] = None) -> Dict[str, np.ndarray]:
    """``PoissonSolver.generate_dataset`` with the batch sharded over the ranks of `group`.
    Returns the full dataset on every rank.  `solve_fn(n, batch, sin_x=, sin_y=)` defaults to the
    CUDA library; tests inject a CPU stand-in to exercise the sharding logic without a GPU."""
    rank, world = world_info(group)
    solve = solve_fn or (lambda n, b, **kw: _lib.solve_batch(n, b, rtol=solver.rtol, max_iter=solver.max_iter, **kw))
    ks = np.random.uniform(*k_range, size=(n_samples, 2))
    k1, k2 = ks[:, 0].copy(), ks[:, 1].copy()
    lo, hi = shard_bounds(n_samples, world, rank)
    sxf, syf = sine_factors(k1, k2, solver.x_fine)
    sxc, syc = sine_factors(k1, k2, solver.x_coarse)
    u_f, info_f = solve(solver.n_fine, hi - lo, sin_x=sxf[lo:hi], sin_y=syf[lo:hi])
    u_c, info_c = solve(solver.n_coarse, hi - lo, sin_x=sxc[lo:hi], sin_y=syc[lo:hi])
    info_f.raise_if_failed("generate_dataset_sharded (fine)")
    info_c.raise_if_failed("generate_dataset_sharded (coarse)")
    return {
        'u_coarse': gather_slices(u_c, n_samples, group),
        'u_fine': gather_slices(u_f, n_samples, group),
        'f_coarse': syc[:, :, None] * sxc[:, None, :],
        'f_fine': syf[:, :, None] * sxf[:, None, :],
        'theta_coarse': np.ones((n_samples, solver.n_coarse, solver.n_coarse)),
        'theta_fine': np.ones((n_samples, solver.n_fine, solver.n_fine)),
        'k1': k1,
        'k2': k2,
    }


def generate_subdomain_dataset_sharded(solver: EnhancedPoissonSolver, n_samples: int, k_range=(0.5, 12.0), group=None,
                                       sub_fn: Optional[Callable] = None) -> Dict[str, np.ndarray]:
    """``EnhancedPoissonSolver.generate_subdomain_dataset`` sharded over the ranks of `group`."""
    rank, world = world_info(group)
    sub = sub_fn or (lambda *a: _lib.subdomain_batch(*a, rtol=solver.rtol, max_iter=solver.max_iter))
    k1 = np.empty(n_samples)
    k2 = np.empty(n_samples)
    start_x = np.empty(n_samples, dtype=np.int32)
    start_y = np.empty(n_samples, dtype=np.int32)
    max_start = solver.n_superfine - solver.n_fine
    for i in range(n_samples):
        k1[i] = np.random.uniform(*k_range)
        k2[i] = np.random.uniform(*k_range)
        start_x[i] = np.random.randint(0, max_start)
        start_y[i] = np.random.randint(0, max_start)
    lo, hi = shard_bounds(n_samples, world, rank)
    sx, sy = sine_factors(k1[lo:hi], k2[lo:hi], solver.x_superfine)
    out, info = sub(solver.n_superfine, solver.n_fine, hi - lo, sx, sy, start_x[lo:hi], start_y[lo:hi])
    info.raise_if_failed("generate_subdomain_dataset_sharded")
    ds = {k: gather_slices(out[k], n_samples, group) for k in
          ('u_coarse', 'u_fine', 'f_coarse', 'f_fine', 'theta_coarse', 'theta_fine')}
    ds['k1'], ds['k2'] = k1, k2
    ds['is_subdomain'] = np.ones(n_samples, dtype=bool)
    return ds


def solve_batch_sharded(n: int, theta: Optional[np.ndarray], sin_x: np.ndarray, sin_y: np.ndarray, group=None,
                        dst: Optional[int] = None, solve_fn: Optional[Callable] = None, **kw):
    """Shard a [B, ...] batch of large-grid problems (e.g. 4,096 x 640^2) across the ranks; every
    rank passes the FULL host arrays and solves its own contiguous slice.  Returns (u, iters, status)
    gathered on rank `dst` (or on every rank when dst is None)."""
    rank, world = world_info(group)
    B = sin_x.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    solve = solve_fn or _lib.solve_batch
    u, info = solve(n, hi - lo, theta=None if theta is None else theta[lo:hi], sin_x=sin_x[lo:hi], sin_y=sin_y[lo:hi], **kw)
    info = info.to_host()
    return (gather_slices(u, B, group, dst), gather_slices(info.iters, B, group, dst), gather_slices(info.status, B, group, dst))
