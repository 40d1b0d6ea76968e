"""Numpy prototype of the MG-preconditioned CG used by the CUDA solver.

Development tool only (not imported by the package, tests or bench): it settles the
algorithmic choices (transfer operators, sweep counts, coarsest size, stopping rule)
on the CPU before they are frozen into the kernels, and it is compared here against
the reference's own SuperLU solve (imported from /root/reference when present).

Operator: A = -L*h^2-free form; we solve (-L) u = -f/theta with
(-L u)_ij = (4 u_ij - u_{i+-1,j} - u_{i,j+-1}) / h^2, zero ghost ring, h = 1/(n-1).
"""
import sys
import time
import numpy as np


def apply_A(u, inv_h2):
    q = 4.0 * u
    q[1:, :] -= u[:-1, :]
    q[:-1, :] -= u[1:, :]
    q[:, 1:] -= u[:, :-1]
    q[:, :-1] -= u[:, 1:]
    return q * inv_h2


def rbgs_half(u, b, h2, color):
    n, m = u.shape
    s = np.zeros_like(u)
    s[1:, :] += u[:-1, :]
    s[:-1, :] += u[1:, :]
    s[:, 1:] += u[:, :-1]
    s[:, :-1] += u[:, 1:]
    ii, jj = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    mask = ((ii + jj) & 1) == color
    u[mask] = 0.25 * (b[mask] * h2 + s[mask])


def restrict_sum(r):
    """sum over 2x2 aggregates (ragged last cell when n odd); caller scales."""
    n = r.shape[0]
    nc = (n + 1) // 2
    pad = np.zeros((2 * nc, 2 * nc))
    pad[:n, :n] = r
    return pad[0::2, 0::2] + pad[0::2, 1::2] + pad[1::2, 0::2] + pad[1::2, 1::2]


def prolong_const(e, n):
    return np.repeat(np.repeat(e, 2, axis=0), 2, axis=1)[:n, :n]


def prolong_bilinear(e, n):
    """cell-centred bilinear: fine node 2I+a gets 3/4 e_I + 1/4 e_{I-1+2a} (zero outside)."""
    nc = e.shape[0]
    p = np.zeros((nc + 2, nc + 2))
    p[1:-1, 1:-1] = e
    out = np.zeros((2 * nc, 2 * nc))
    for a in (0, 1):
        for c in (0, 1):
            di = -1 if a == 0 else 1
            dj = -1 if c == 0 else 1
            I = slice(1, nc + 1)
            out[a::2, c::2] = (9 * p[I, I] + 3 * p[1 + di:nc + 1 + di, I] + 3 * p[I, 1 + dj:nc + 1 + dj]
                               + p[1 + di:nc + 1 + di, 1 + dj:nc + 1 + dj]) / 16.0
    return out[:n, :n]


def restrict_bilinear(r):
    """transpose of prolong_bilinear (unscaled)"""
    n = r.shape[0]
    nc = (n + 1) // 2
    out = np.zeros((nc + 2, nc + 2))
    pad = np.zeros((2 * nc, 2 * nc))
    pad[:n, :n] = r
    for a in (0, 1):
        for c in (0, 1):
            di = -1 if a == 0 else 1
            dj = -1 if c == 0 else 1
            blk = pad[a::2, c::2]
            out[1:nc + 1, 1:nc + 1] += 9 * blk / 16
            out[1 + di:nc + 1 + di, 1:nc + 1] += 3 * blk / 16
            out[1:nc + 1, 1 + dj:nc + 1 + dj] += 3 * blk / 16
            out[1 + di:nc + 1 + di, 1 + dj:nc + 1 + dj] += blk / 16
    return out[1:-1, 1:-1]


class MG:
    def __init__(self, n, nu=2, transfer="const", coarsest=6, rscale=0.25):
        self.nu = nu
        self.transfer = transfer
        self.ns = [n]
        while self.ns[-1] > coarsest:
            self.ns.append((self.ns[-1] + 1) // 2)
        h = 1.0 / (n - 1)
        self.h2 = [(h * 2 ** l) ** 2 for l in range(len(self.ns))]
        self.rscale = rscale
        nc = self.ns[-1]
        # dense coarsest inverse
        N = nc * nc
        A = np.zeros((N, N))
        for k in range(N):
            e = np.zeros(N)
            e[k] = 1
            A[:, k] = apply_A(e.reshape(nc, nc), 1.0 / self.h2[-1]).ravel()
        self.Ainv = np.linalg.inv(A)

    def vcycle(self, b, l=0):
        n = self.ns[l]
        if l == len(self.ns) - 1:
            return (self.Ainv @ b.ravel()).reshape(n, n)
        h2 = self.h2[l]
        u = np.zeros_like(b)
        for _ in range(self.nu):
            rbgs_half(u, b, h2, 0)
            rbgs_half(u, b, h2, 1)
        r = b - apply_A(u, 1.0 / h2)
        if self.transfer == "const":
            rc = restrict_sum(r) * self.rscale
        else:
            rc = restrict_bilinear(r) * self.rscale
        ec = self.vcycle(rc, l + 1)
        if self.transfer == "const":
            u += prolong_const(ec, n)
        else:
            u += prolong_bilinear(ec, n)
        for _ in range(self.nu):
            rbgs_half(u, b, h2, 1)
            rbgs_half(u, b, h2, 0)
        return u


def mgpcg(b, mg, rtol=1e-12, maxit=200, verbose=False):
    n = b.shape[0]
    inv_h2 = 1.0 / mg.h2[0]
    x = np.zeros_like(b)
    r = b.copy()
    bnorm = np.linalg.norm(b)
    z = mg.vcycle(r)
    p = z.copy()
    rz = np.vdot(r, z)
    hist = []
    for it in range(1, maxit + 1):
        q = apply_A(p, inv_h2)
        alpha = rz / np.vdot(p, q)
        x += alpha * p
        r -= alpha * q
        rn = np.linalg.norm(r) / bnorm
        hist.append(rn)
        if rn <= rtol:
            rt = np.linalg.norm(b - apply_A(x, inv_h2)) / bnorm
            if rt <= rtol:
                break
            r = b - apply_A(x, inv_h2)
        z = mg.vcycle(r)
        rz_new = np.vdot(r, z)
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it, hist


def ref_solver(n):
    import types
    for m in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, "/root/reference/src")
    import warnings
    warnings.filterwarnings("ignore")
    from data_generation import PoissonSolver
    return PoissonSolver(n_coarse=max(n // 2, 2), n_fine=n)


if __name__ == "__main__":
    rng = np.random.RandomState(0)
    sizes = [int(s) for s in sys.argv[1].split(",")] if len(sys.argv) > 1 else [20, 40, 80, 160]
    for n in sizes:
        x = np.linspace(0, 1, n)
        X, Y = np.meshgrid(x, x)
        for case in ("smooth", "rough", "lowk"):
            if case == "lowk":
                k1, k2 = 0.5, 0.7
                theta = np.ones((n, n))
            elif case == "smooth":
                k1, k2 = 2.3, 3.7
                theta = np.ones((n, n))
            else:
                k1, k2 = 10.3, 11.9
                theta = rng.uniform(0.5, 2.0, (n, n))
            f = np.sin(2 * np.pi * k1 * X) * np.sin(2 * np.pi * k2 * Y)
            b = -f / theta
            uref = None
            if n <= 320:
                uref = ref_solver(n).solve_poisson(f, theta, "fine")
            for transfer, nu, rs in (("const", 2, 0.25), ("const", 1, 0.25), ("const", 2, 0.5), ("bilin", 2, 0.25), ("bilin", 1, 0.25)):
                mg = MG(n, nu=nu, transfer=transfer, rscale=rs)
                t = time.time()
                u, it, hist = mgpcg(b, mg)
                err = np.linalg.norm(u - uref) / np.linalg.norm(uref) if uref is not None else float("nan")
                print(f"n={n:4d} {case:6s} {transfer:5s} nu={nu} rs={rs} iters={it:3d} relres={hist[-1]:.2e} err_vs_ref={err:.2e} ({time.time()-t:.2f}s)")
