"""ctypes binding of libpsn_b200.so (the C ABI declared in include/psn_b200.h).

There is NO CPU fallback: if the CUDA library has not been built, or no CUDA device is
present, every compute call raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
-- the single nvcc command it runs is printed, so the library can also be rebuilt by hand).
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PSN_B200_LIB lets kernel experiments load an alternative build of the SAME CUDA library (never a CPU path)
LIB_PATH = os.environ.get("PSN_B200_LIB") or os.path.join(_HERE, "libpsn_b200.so")

PSN_OK = 0
STATUS_CONVERGED, STATUS_FP64_FLOOR, STATUS_MAXITER, STATUS_BREAKDOWN, STATUS_STAGNATED = 0, 1, 2, 3, 4
# a sample that ends at the fp64 round-off floor (status 1) is only accepted below this relative true residual
# (PSN_FLOOR_CEIL of the library; checked again here so that no entry point returns a stagnated sample silently)
FLOOR_RELRES_CEILING = 1e-10

EXPORTED_SYMBOLS = (
    "psn_version", "psn_get_device_count", "psn_init", "psn_shutdown", "psn_last_error",
    "psn_solve_batch", "psn_subdomain_batch", "psn_multires", "psn_set_option",
    "psn_profile_begin", "psn_profile_read", "psn_profile_end", "psn_describe", "psn_dataset_tensors", "psn_upscale_cascade",
    "psn_mt19937_uniform_blocks", "psn_mt19937_jump_host",
)


class PsnError(RuntimeError):
    """A call into libpsn_b200.so failed (the message is psn_last_error())."""


_lib: Optional[ctypes.CDLL] = None
_c_dp = ctypes.c_void_p


def load() -> ctypes.CDLL:
    """Load the shared library (once) and declare the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PsnError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built and this package has no "
            "CPU fallback. Run `python -c \"import __graft_entry__ as g; g.build()\"` at the repo root.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.psn_version.restype = ctypes.c_int
    lib.psn_get_device_count.restype = ctypes.c_int
    lib.psn_init.argtypes = [ctypes.c_int]
    lib.psn_init.restype = ctypes.c_int
    lib.psn_shutdown.restype = None
    lib.psn_last_error.restype = ctypes.c_char_p
    lib.psn_solve_batch.argtypes = [
        ctypes.c_int, ctypes.c_int64, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.c_int, ctypes.c_int,
        ctypes.c_int, ctypes.c_double, ctypes.c_int, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.c_void_p]
    lib.psn_solve_batch.restype = ctypes.c_int
    lib.psn_subdomain_batch.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int64, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.c_double, ctypes.c_int,
        _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.c_void_p]
    lib.psn_subdomain_batch.restype = ctypes.c_int
    lib.psn_multires.argtypes = [
        ctypes.c_int, _c_dp, ctypes.c_int, ctypes.c_int64, _c_dp, _c_dp, _c_dp, ctypes.c_double, ctypes.c_int,
        ctypes.POINTER(ctypes.c_void_p), _c_dp, _c_dp, _c_dp, ctypes.c_void_p]
    lib.psn_multires.restype = ctypes.c_int
    lib.psn_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int64]
    lib.psn_set_option.restype = ctypes.c_int
    lib.psn_profile_begin.restype = ctypes.c_int
    lib.psn_profile_read.argtypes = [ctypes.c_char_p, ctypes.c_int64]
    lib.psn_profile_read.restype = ctypes.c_int
    lib.psn_profile_end.restype = ctypes.c_int
    lib.psn_dataset_tensors.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp,
                                        ctypes.c_void_p]
    lib.psn_dataset_tensors.restype = ctypes.c_int
    lib.psn_upscale_cascade.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, _c_dp, _c_dp, ctypes.c_void_p]
    lib.psn_upscale_cascade.restype = ctypes.c_int
    lib.psn_mt19937_uniform_blocks.argtypes = [_c_dp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                               ctypes.c_double, _c_dp, _c_dp, _c_dp, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
    lib.psn_mt19937_uniform_blocks.restype = ctypes.c_int
    lib.psn_mt19937_jump_host.argtypes = [_c_dp, ctypes.c_int, ctypes.c_int64, _c_dp, ctypes.POINTER(ctypes.c_int)]
    lib.psn_mt19937_jump_host.restype = ctypes.c_int
    lib.psn_describe.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
    lib.psn_describe.restype = ctypes.c_int
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != PSN_OK:
        msg = load().psn_last_error().decode("utf-8", "replace")
        raise PsnError(f"libpsn_b200 error {rc}: {msg}")


def _ptr(a) -> Optional[int]:
    """Raw address of a numpy array (host) or a torch tensor (host or device); None stays NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(f"unsupported buffer type {type(a)!r}")


def _f64(a, shape=None, name="array"):
    """numpy inputs -> float64 C-contiguous (copy only if needed); torch tensors pass through."""
    if a is None:
        return None
    if hasattr(a, "data_ptr") and not isinstance(a, np.ndarray):
        if str(a.dtype) != "torch.float64" or not a.is_contiguous():
            raise ValueError(f"{name}: torch tensors must be contiguous float64")
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(a.shape)}")
        return a
    arr = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        if arr.size != int(np.prod(shape)):
            raise ValueError(f"{name}: expected {int(np.prod(shape))} elements (shape {tuple(shape)}), got shape {arr.shape}")
        arr = arr.reshape(shape)
    return arr


_TORCH_NAMES = {"float64": "torch.float64", "float32": "torch.float32", "int32": "torch.int32"}


def _out(a, shape, dtype: str, name: str):
    """Validate a caller-supplied OUTPUT buffer before its raw address crosses the C ABI: the library writes
    prod(shape) elements of `dtype` through it, so a wrong dtype, a non-contiguous view or a short array would
    corrupt host or device memory.  numpy arrays and torch tensors (host or CUDA) are accepted."""
    shape = tuple(int(v) for v in shape)
    if isinstance(a, np.ndarray):
        if a.dtype != np.dtype(dtype):
            raise ValueError(f"{name}: expected dtype {dtype}, got {a.dtype}")
        if not a.flags.c_contiguous:
            raise ValueError(f"{name}: output array must be C-contiguous")
        if not a.flags.writeable:
            raise ValueError(f"{name}: output array is read-only")
        if tuple(a.shape) != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(a.shape)}")
        return a
    if hasattr(a, "data_ptr"):
        if str(a.dtype) != _TORCH_NAMES[dtype]:
            raise ValueError(f"{name}: expected dtype {_TORCH_NAMES[dtype]}, got {a.dtype}")
        if not a.is_contiguous():
            raise ValueError(f"{name}: output tensor must be contiguous")
        if tuple(a.shape) != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(a.shape)}")
        return a
    raise TypeError(f"{name}: unsupported buffer type {type(a)!r}")


_PINNED_MIN = 1 << 20
_PINNED_MAX = int(os.environ.get("PSN_PINNED_MAX_BYTES", str(16 << 30)))
_TORCH_DTYPES = None


def host_empty(shape, dtype=np.float64) -> np.ndarray:
    """Result array for host callers.  Arrays of 1 MiB and more are backed by page-locked memory (a torch pinned tensor
    behind the numpy view, recycled by torch's caching host allocator): the library's copy pipeline then runs fully
    asynchronously and at link speed instead of through the driver's pageable staging (multi-resolution chain, 64
    examples: 165 -> 80 ms).  ``PSN_PINNED_MAX_BYTES=0`` turns it off; without torch / CUDA it is a plain np.empty."""
    global _TORCH_DTYPES
    shape = tuple(int(v) for v in (shape if hasattr(shape, "__len__") else (shape,)))
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dt.itemsize if len(shape) else dt.itemsize
    if _PINNED_MIN <= nbytes <= _PINNED_MAX:
        try:
            import torch
            if torch.cuda.is_available():
                if _TORCH_DTYPES is None:
                    _TORCH_DTYPES = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                                     np.dtype(np.int32): torch.int32}
                if dt in _TORCH_DTYPES:
                    return torch.empty(shape, dtype=_TORCH_DTYPES[dt], pin_memory=True).numpy()
        except Exception:
            pass
    return np.empty(shape, dtype=dt)


def device_count() -> int:
    return int(load().psn_get_device_count())


def init(device: int = -1) -> None:
    _check(load().psn_init(int(device)))


def shutdown() -> None:
    load().psn_shutdown()


def set_option(name: str, value: int) -> None:
    _check(load().psn_set_option(name.encode(), int(value)))


def describe(n: int) -> dict:
    info = (ctypes.c_int64 * 8)()
    _check(load().psn_describe(int(n), info))
    return {"resident": bool(info[0]), "levels": int(info[1]), "first_smem_level": int(info[2]),
            "threads": int(info[3]), "smem_bytes": int(info[4]), "tile": int(info[5]),
            "workspace_bytes_per_sample": int(info[6]), "vcycle_f32": bool(info[7] & 1),
            "update_strip": bool(info[7] & 2), "p_f32": bool(info[7] & 4),
            "x_defer": bool(info[7] & 8), "fused_update": bool(info[7] & 16)}


def profile_begin() -> None:
    _check(load().psn_profile_begin())


def profile_read() -> dict:
    lib = load()
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib.psn_profile_read(buf, len(buf))
    if n < 0:
        buf = ctypes.create_string_buffer(-n + 16)
        n = lib.psn_profile_read(buf, len(buf))
    return json.loads(buf.value.decode())


def profile_end() -> None:
    _check(load().psn_profile_end())


class SolveInfo:
    """Per-sample convergence record of one batched call."""

    def __init__(self, iters: np.ndarray, relres: np.ndarray, status: np.ndarray):
        self.iters, self.relres, self.status = iters, relres, status

    def to_host(self) -> "SolveInfo":
        cv = lambda a: a.cpu().numpy() if hasattr(a, "cpu") else a
        return SolveInfo(cv(self.iters), cv(self.relres), cv(self.status))

    def accepted(self) -> np.ndarray:
        """Per-sample bool: converged, or at the fp64 round-off floor with a relative true residual below
        FLOOR_RELRES_CEILING."""
        if hasattr(self.status, "cpu"):
            return self.to_host().accepted()
        status = np.asarray(self.status)
        return (status == STATUS_CONVERGED) | ((status == STATUS_FP64_FLOOR) & (np.asarray(self.relres) <= FLOOR_RELRES_CEILING))

    def raise_if_failed(self, what: str = "solve") -> None:
        if hasattr(self.status, "cpu"):
            return self.to_host().raise_if_failed(what)
        bad = np.flatnonzero(~self.accepted())
        if bad.size:
            i = int(bad[0])
            raise PsnError(
                f"{what}: {bad.size} of {self.status.size} samples did not converge "
                f"(first: sample {i}, status {int(self.status.flat[i])}, relres {float(self.relres.flat[i]):.3e}, "
                f"iters {int(self.iters.flat[i])}); refusing to return unconverged solutions")


def solve_batch(n: int, batch: int, *, theta=None, f=None, sin_x=None, sin_y=None, k1=None, k2=None,
                master_n: Optional[int] = None, step: int = 1, rtol: float = 2.5e-13, max_iter: int = 100,
                out=None, stream: int = 0, info_on_host: bool = True):
    """Batched solve of diag(theta) L u = f.  Buffers may be numpy arrays (host) or torch CUDA
    tensors (device, float64, contiguous).  Returns (u, SolveInfo)."""
    lib = load()
    mn = int(master_n) if master_n is not None else int(n)
    theta = _f64(theta, (batch, mn, mn), "theta")
    f = _f64(f, (batch, mn, mn), "f")
    sin_x = _f64(sin_x, (batch, mn), "sin_x")
    sin_y = _f64(sin_y, (batch, mn), "sin_y")
    k1 = _f64(k1, (batch,), "k1")
    k2 = _f64(k2, (batch,), "k2")
    if out is None:
        out = host_empty((batch, n, n))
    else:
        out = _out(out, (batch, n, n), "float64", "out")
    if hasattr(out, "is_cuda") and out.is_cuda:
        # device-resident call: keep the per-sample record on the device too, so that the call
        # only enqueues work on `stream` (no host pointer -> no staging, no synchronisation)
        import torch
        iters = torch.zeros(batch, dtype=torch.int32, device=out.device)
        relres = torch.zeros(batch, dtype=torch.float64, device=out.device)
        status = torch.zeros(batch, dtype=torch.int32, device=out.device)
    else:
        iters = np.zeros(batch, dtype=np.int32)
        relres = np.zeros(batch, dtype=np.float64)
        status = np.zeros(batch, dtype=np.int32)
    rc = lib.psn_solve_batch(int(n), int(batch), _ptr(theta), _ptr(f), _ptr(sin_x), _ptr(sin_y), _ptr(k1), _ptr(k2),
                             mn, int(step), 0, float(rtol), int(max_iter), _ptr(out), _ptr(iters), _ptr(relres),
                             _ptr(status), ctypes.c_void_p(stream))
    _check(rc)
    return out, SolveInfo(iters, relres, status)


def subdomain_batch(n_super: int, n_fine: int, batch: int, sin_x, sin_y, start_x, start_y, *,
                    rtol: float = 2.5e-13, max_iter: int = 100, stream: int = 0):
    """Fused super-fine solve + window + stride-2 downsample.  Returns (dict of arrays, SolveInfo)."""
    lib = load()
    nc = (n_fine + 1) // 2
    sin_x = _f64(sin_x, (batch, n_super), "sin_x")
    sin_y = _f64(sin_y, (batch, n_super), "sin_y")
    start_x = np.ascontiguousarray(start_x, dtype=np.int32).reshape(batch)
    start_y = np.ascontiguousarray(start_y, dtype=np.int32).reshape(batch)
    if batch and (start_x.min() < 0 or start_y.min() < 0 or start_x.max() + n_fine > n_super or start_y.max() + n_fine > n_super):
        raise ValueError("sub-domain window leaves the super-fine grid")
    out = {
        "u_fine": host_empty((batch, n_fine, n_fine)), "f_fine": host_empty((batch, n_fine, n_fine)),
        "theta_fine": host_empty((batch, n_fine, n_fine)),
        "u_coarse": host_empty((batch, nc, nc)), "f_coarse": host_empty((batch, nc, nc)),
        "theta_coarse": host_empty((batch, nc, nc)),
    }
    iters = np.zeros(batch, dtype=np.int32)
    relres = np.zeros(batch, dtype=np.float64)
    status = np.zeros(batch, dtype=np.int32)
    rc = lib.psn_subdomain_batch(int(n_super), int(n_fine), int(batch), _ptr(sin_x), _ptr(sin_y), _ptr(start_x), _ptr(start_y),
                                 float(rtol), int(max_iter), _ptr(out["u_fine"]), _ptr(out["f_fine"]), _ptr(out["theta_fine"]),
                                 _ptr(out["u_coarse"]), _ptr(out["f_coarse"]), _ptr(out["theta_coarse"]),
                                 _ptr(iters), _ptr(relres), _ptr(status), ctypes.c_void_p(stream))
    _check(rc)
    return out, SolveInfo(iters, relres, status)


def multires(n_finest: int, resolutions: Sequence[int], batch: int, theta_finest, sin_x, sin_y, *,
             rtol: float = 2.5e-13, max_iter: int = 100, stream: int = 0):
    """Solve at every resolution from one master theta / forcing per example (stride gathers).
    Returns ({res: u[batch, res, res]}, SolveInfo with [n_res, batch] arrays)."""
    lib = load()
    res = [int(r) for r in resolutions]
    theta_finest = _f64(theta_finest, (batch, n_finest, n_finest), "theta_finest")
    sin_x = _f64(sin_x, (batch, n_finest), "sin_x")
    sin_y = _f64(sin_y, (batch, n_finest), "sin_y")
    outs = [host_empty((batch, r, r)) for r in res]
    res_arr = np.asarray(res, dtype=np.int32)
    ptrs = (ctypes.c_void_p * len(res))(*[o.ctypes.data for o in outs])
    iters = np.zeros((len(res), batch), dtype=np.int32)
    relres = np.zeros((len(res), batch), dtype=np.float64)
    status = np.zeros((len(res), batch), dtype=np.int32)
    rc = lib.psn_multires(int(n_finest), _ptr(res_arr), len(res), int(batch), _ptr(theta_finest), _ptr(sin_x), _ptr(sin_y),
                          float(rtol), int(max_iter), ptrs, _ptr(iters), _ptr(relres), _ptr(status), ctypes.c_void_p(stream))
    _check(rc)
    return dict(zip(res, outs)), SolveInfo(iters, relres, status)


def solve_from_wavenumbers(n: int, k1, k2, *, theta=None, out=None, chunks: Optional[int] = None, kfirst: bool = False,
                           rtol: float = 2.5e-13, max_iter: int = 100, sines_out=None):
    """Host-side convenience used by the dataset shims and by bench.py's end-to-end leg: evaluates the 2n sine
    factors per sample with numpy (operand order of the reference, so f is bit-identical) for chunk c+1 on a
    helper thread WHILE chunk c is inside the C call (both release the GIL), then solves.  theta / out are host
    arrays [B, n, n] (theta may be None).  ``sines_out=(sin_x, sin_y)``: caller-owned float64 [B, n] arrays to
    evaluate the sines into -- a caller whose theta / out are page-locked passes page-locked ones here, so that
    every copy of the call is asynchronous and the library takes its deep upload pipeline.
    Returns (u, SolveInfo, sin_x, sin_y)."""
    from concurrent.futures import ThreadPoolExecutor

    k1 = np.ascontiguousarray(k1, dtype=np.float64).reshape(-1)
    k2 = np.ascontiguousarray(k2, dtype=np.float64).reshape(-1)
    B = k1.shape[0]
    x = np.linspace(0, 1, n)
    if out is None:
        out = host_empty((B, n, n))
    else:
        out = _out(out, (B, n, n), "float64", "out")
    if sines_out is not None:
        sx, sy = sines_out
        for a_ in (sx, sy):
            if a_.shape != (B, n) or a_.dtype != np.float64 or not a_.flags.c_contiguous:
                raise ValueError(f"sines_out arrays must be C-contiguous float64 of shape {(B, n)}")
    else:
        sx = host_empty((B, n))
        sy = host_empty((B, n))
    iters = np.zeros(B, dtype=np.int32)
    relres = np.zeros(B, dtype=np.float64)
    status = np.zeros(B, dtype=np.int32)
    if chunks is None:
        # small grids: the 2n sines per sample cost a fifth of the (PCIe-bound) call -> overlap them in 2 chunks (every extra
        # call pays the copy pipeline's fill and drain again: 80x80 x 18,944 measured 29.5 / 29.1 / 29.9 / 30.7 / 32.6 ms for
        # 1 / 2 / 3 / 4 / 8 chunks); large grids: sines are negligible and one call keeps the library's own pipeline intact
        chunks = 2 if n <= 100 else 1
        uneven = True
    else:
        uneven = False
    chunks = max(1, min(int(chunks), B)) if B else 1
    bounds = [(B * c) // chunks for c in range(chunks + 1)]
    if uneven and chunks == 2 and B >= 4096:
        # a short first chunk: the GPU starts after 1/8 of the sines, and the sines of the long second chunk are evaluated
        # while the first one is in the C call
        bounds = [0, B // 8, B]

    def sines_range(lo, hi):
        w1 = (k1[lo:hi] * 2 * np.pi) if kfirst else (2 * np.pi * k1[lo:hi])
        w2 = (k2[lo:hi] * 2 * np.pi) if kfirst else (2 * np.pi * k2[lo:hi])
        np.sin(w1[:, None] * x[None, :], out=sx[lo:hi])
        np.sin(w2[:, None] * x[None, :], out=sy[lo:hi])

    def sines(c):
        lo, hi = bounds[c], bounds[c + 1]
        work = (hi - lo) * n
        # up to 16 threads, but only this rank's share of the host cores (torchrun: one rank per GPU on the same host)
        cores = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)))
        nt = 1 if work < (1 << 16) else min(16, cores, max(1, (hi - lo) // 8))
        if nt <= 1:
            sines_range(lo, hi)
            return
        # numpy's sin releases the GIL: split the rows over a few threads (elementwise, so the values do not
        # depend on the split)
        cuts = [lo + ((hi - lo) * t) // nt for t in range(nt + 1)]
        with ThreadPoolExecutor(max_workers=nt) as tp:
            list(tp.map(lambda t: sines_range(cuts[t], cuts[t + 1]), range(nt)))

    def view(a, lo, hi):
        return None if a is None else a[lo:hi]

    with ThreadPoolExecutor(max_workers=1) as pool:
        fut = pool.submit(sines, 0) if B else None
        for c in range(chunks if B else 0):
            fut.result()
            if c + 1 < chunks:
                fut = pool.submit(sines, c + 1)
            lo, hi = bounds[c], bounds[c + 1]
            if hi == lo:
                continue
            dst = out[lo:hi]
            u_c, info = solve_batch(n, hi - lo, theta=view(theta, lo, hi), sin_x=sx[lo:hi], sin_y=sy[lo:hi],
                                    out=dst, rtol=rtol, max_iter=max_iter)
            if u_c is not dst:   # (a stand-in backend in the tests returns a fresh array)
                dst[...] = u_c
            iters[lo:hi], relres[lo:hi], status[lo:hi] = info.iters, info.relres, info.status
    return out, SolveInfo(iters, relres, status), sx, sy


def dataset_tensors(u_coarse, u_fine, f_fine, theta_fine, *, x=None, y=None, stream: int = 0):
    """float32 training tensors of the reference's PDEDataset (src/models.py:132-203) from float64 dataset arrays.
    Inputs: numpy arrays or torch CUDA float64 tensors; outputs follow `x` / `y` if given, else numpy float32.
    Returns (x [N,3,nf,nf], y [N,1,nf,nf], stats[8] as numpy)."""
    lib = load()
    N, nc = int(u_coarse.shape[0]), int(u_coarse.shape[-1])
    nf = int(u_fine.shape[-1])
    u_coarse = _f64(u_coarse, (N, nc, nc), "u_coarse")
    u_fine = _f64(u_fine, (N, nf, nf), "u_fine")
    f_fine = _f64(f_fine, (N, nf, nf), "f_fine")
    theta_fine = _f64(theta_fine, (N, nf, nf), "theta_fine")
    x = host_empty((N, 3, nf, nf), np.float32) if x is None else _out(x, (N, 3, nf, nf), "float32", "x")
    y = host_empty((N, 1, nf, nf), np.float32) if y is None else _out(y, (N, 1, nf, nf), "float32", "y")
    stats = np.zeros(8, dtype=np.float64)
    rc = lib.psn_dataset_tensors(N, nc, nf, _ptr(u_coarse), _ptr(u_fine), _ptr(f_fine), _ptr(theta_fine), _ptr(x), _ptr(y),
                                 _ptr(stats), ctypes.c_void_p(stream))
    _check(rc)
    return x, y, stats


def upscale_cascade(u, levels: int, mode: str = "bilinear", out=None, stream: int = 0):
    """`levels` successive x2 upsamplings (align_corners=True, float32) of u [B, n, n] (float64).
    mode: 'bilinear' | 'bicubic'.  Returns float32 [B, n*2^levels, n*2^levels]."""
    lib = load()
    B, n = int(u.shape[0]), int(u.shape[-1])
    u = _f64(u, (B, n, n), "u")
    m = n << levels
    out = host_empty((B, m, m), np.float32) if out is None else _out(out, (B, m, m), "float32", "out")
    rc = lib.psn_upscale_cascade({"bilinear": 0, "bicubic": 1}[mode], n, int(levels), B, _ptr(u), _ptr(out), ctypes.c_void_p(stream))
    _check(rc)
    return out


# --------------------------------------------------------------------------- numpy's legacy global RNG on the device
def _legacy_state(random_state=None):
    st = (random_state or np.random).get_state()
    if st[0] != "MT19937":
        raise ValueError(f"legacy MT19937 state expected, got {st[0]}")
    return st


def mt19937_jump(n_words: int, random_state=None) -> None:
    """Advance numpy's legacy generator (the global one, or a RandomState) by `n_words` 32-bit outputs -- one double is
    two words -- with the library's jump-ahead polynomials (host only)."""
    lib = load()
    st = _legacy_state(random_state)
    key = np.ascontiguousarray(st[1], dtype=np.uint32)
    key_out = np.empty(624, dtype=np.uint32)
    pos_out = ctypes.c_int(0)
    _check(lib.psn_mt19937_jump_host(_ptr(key), int(st[2]), int(n_words), _ptr(key_out), ctypes.byref(pos_out)))
    (random_state or np.random).set_state((st[0], key_out, int(pos_out.value), 0, 0.0))


def mt19937_uniform_blocks(n_blocks: int, skip: int, count: int, low: float, high: float, *, out=None, random_state=None,
                           stream: int = 0):
    """Device-side replay of ``for b in range(n_blocks): [np.random.random_sample() for _ in range(skip)];
    np.random.uniform(low, high, count)`` on numpy's legacy generator (global, or `random_state`), whose state is advanced
    exactly as the host loop would.  `out`: float64 [n_blocks, count] numpy array or torch CUDA tensor (default: a new
    numpy array).  Returns (out, skipped) with skipped [n_blocks, skip] the raw doubles in [0, 1) (numpy)."""
    lib = load()
    st = _legacy_state(random_state)
    if st[3] != 0 and skip + count > 0:
        pass   # a cached Gaussian is unaffected by uniform draws; it is carried over below
    key = np.ascontiguousarray(st[1], dtype=np.uint32)
    out = host_empty((n_blocks, count)) if out is None else _out(out, (n_blocks, count), "float64", "out")
    skipped = np.empty((n_blocks, skip), dtype=np.float64)
    key_out = np.empty(624, dtype=np.uint32)
    pos_out = ctypes.c_int(0)
    rc = lib.psn_mt19937_uniform_blocks(_ptr(key), int(st[2]), int(n_blocks), int(skip), int(count), float(low), float(high),
                                        _ptr(out), _ptr(skipped), _ptr(key_out), ctypes.byref(pos_out), ctypes.c_void_p(stream))
    _check(rc)
    (random_state or np.random).set_state((st[0], key_out, int(pos_out.value), st[3], st[4]))
    return out, skipped


_MT_SPLIT = 1 << 15   # doubles per block when one long draw is split over CTAs


def reshaped(a, shape):
    """C-order reshape of a numpy array or a contiguous torch tensor (no copy)."""
    return a.reshape(shape) if isinstance(a, np.ndarray) else a.view(*shape)


def to_host(a) -> np.ndarray:
    """Host copy of a CUDA tensor (through page-locked memory, see host_empty); numpy arrays pass through."""
    if isinstance(a, np.ndarray):
        return a
    import torch
    h = host_empty(tuple(a.shape), {"torch.float64": np.float64, "torch.float32": np.float32, "torch.int32": np.int32}[str(a.dtype)])
    torch.from_numpy(h).copy_(a)
    return h


_COPY_STREAM = None


def to_host_async(a):
    """Like to_host, but the copy runs on a side stream: returns (host array, wait) -- the array is valid once ``wait()`` has
    returned (it returns the array).  Used by the multi-resolution chain driver, whose master theta (produced on the device)
    goes back to the host while the solves are running."""
    global _COPY_STREAM
    if isinstance(a, np.ndarray):
        return a, (lambda: a)
    import torch
    h = host_empty(tuple(a.shape), {"torch.float64": np.float64, "torch.float32": np.float32, "torch.int32": np.int32}[str(a.dtype)])
    if _COPY_STREAM is None:
        _COPY_STREAM = torch.cuda.Stream()
    _COPY_STREAM.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(_COPY_STREAM):
        torch.from_numpy(h).copy_(a, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(_COPY_STREAM)

    def wait(_keep=a):   # (_keep: the device tensor must outlive the copy)
        ev.synchronize()
        return h

    return h, wait


def device_empty_f64(shape):
    """float64 CUDA tensor on the library's device (torch is the allocator here, nothing more)."""
    import torch
    return torch.empty(tuple(int(v) for v in shape), dtype=torch.float64, device="cuda")


def mt19937_uniform(shape, low: float, high: float, *, out=None, random_state=None, stream: int = 0):
    """``np.random.uniform(low, high, shape)`` of the legacy generator, produced on the device (bit-identical values,
    generator state advanced as numpy would).  One long draw is cut into blocks of 32,768 doubles, one CTA each.
    `out`: float64 array / CUDA tensor with prod(shape) elements, C-contiguous (default: new numpy array of `shape`)."""
    shape = tuple(int(v) for v in (shape if hasattr(shape, "__len__") else (shape,)))
    total = int(np.prod(shape)) if len(shape) else 1
    if out is None:
        out = host_empty(shape)
    flat = reshaped(out, (-1,))
    if flat.shape[0] != total:
        raise ValueError(f"out: expected {total} elements, got {flat.shape[0]}")
    nb, rem = divmod(total, _MT_SPLIT)
    if nb:
        head = flat[:nb * _MT_SPLIT]
        mt19937_uniform_blocks(nb, 0, _MT_SPLIT, low, high, out=reshaped(head, (nb, _MT_SPLIT)), random_state=random_state, stream=stream)
    if rem:
        tail = flat[nb * _MT_SPLIT:]
        mt19937_uniform_blocks(1, 0, rem, low, high, out=reshaped(tail, (1, rem)), random_state=random_state, stream=stream)
    return out
