"""Prototype 2: vertex-style coarsening (coarse I <-> fine 2I+1), linear interpolation,
full weighting, rediscretised 5-point coarse operator whose last row/column carry the
1-D Galerkin boundary conductance (exact boundary distance). Development tool only."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/tools")
from prototype_mgpcg import ref_solver


class Level:
    def __init__(self, n, H2, gam):
        self.n, self.H2, self.gam = n, H2, gam   # diag = 4 + gam*(i==n-1) + gam*(j==n-1)
        d = np.full((n, n), 4.0)
        d[-1, :] += gam
        d[:, -1] += gam
        self.diag = d
        ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        self.red = ((ii + jj) & 1) == 0


def nbsum(u):
    s = np.zeros_like(u)
    s[1:, :] += u[:-1, :]; s[:-1, :] += u[1:, :]
    s[:, 1:] += u[:, :-1]; s[:, :-1] += u[:, 1:]
    return s


def apply(lv, u):
    return (lv.diag * u - nbsum(u)) / lv.H2


def half(lv, u, b, color):
    m = lv.red if color == 0 else ~lv.red
    s = nbsum(u)
    u[m] = ((b * lv.H2 + s) / lv.diag)[m]


def P1(e, n):
    nc = e.shape[0]
    pad = np.zeros((nc + 2,) + e.shape[1:])
    pad[1:-1] = e
    out = np.zeros((n,) + e.shape[1:])
    out[1::2] = e[: len(out[1::2])]
    ne = len(out[0::2])
    out[0::2] = 0.5 * (pad[0:ne] + pad[1:ne + 1])
    return out


def R1(r, nc):
    n = r.shape[0]
    pad = np.zeros((2 * nc + 2,) + r.shape[1:])
    pad[:n] = r
    return 0.5 * (0.5 * pad[0:2 * nc:2] + pad[1:2 * nc + 1:2] + 0.5 * pad[2:2 * nc + 2:2])


class MGV:
    def __init__(self, n, nu=2, coarsest=6):
        h = 1.0 / (n - 1)
        self.nu = nu
        H2 = h * h
        g = 1.0 / H2      # boundary conductance (regular)
        self.lv = [Level(n, H2, 0.0)]
        while self.lv[-1].n > coarsest:
            nf = self.lv[-1].n
            if nf % 2 == 0:
                g = g / 2
            else:
                g = 1.0 / (8 * H2) + g / 8
            H2 = 4 * H2
            nc = nf // 2
            self.lv.append(Level(nc, H2, g * H2 - 1.0))
        lc = self.lv[-1]
        N = lc.n ** 2
        A = np.zeros((N, N))
        for k in range(N):
            e = np.zeros(N); e[k] = 1
            A[:, k] = apply(lc, e.reshape(lc.n, lc.n)).ravel()
        self.Ainv = np.linalg.inv(A)
        self.gams = [l.gam for l in self.lv]

    def vcycle(self, b, l=0):
        lv = self.lv[l]
        if l == len(self.lv) - 1:
            return (self.Ainv @ b.ravel()).reshape(lv.n, lv.n)
        u = np.zeros_like(b)
        for _ in range(self.nu):
            half(lv, u, b, 0); half(lv, u, b, 1)
        r = b - apply(lv, u)
        nc = self.lv[l + 1].n
        rc = R1(R1(r, nc).T, nc).T
        ec = self.vcycle(rc, l + 1)
        u += P1(P1(ec, lv.n).T, lv.n).T
        for _ in range(self.nu):
            half(lv, u, b, 1); half(lv, u, b, 0)
        return u


def pcg(b, mg, rtol=1e-12, maxit=100):
    lv = mg.lv[0]
    x = np.zeros_like(b); r = b.copy(); bn = np.linalg.norm(b)
    z = mg.vcycle(r); p = z.copy(); rz = np.vdot(r, z)
    for it in range(1, maxit + 1):
        q = apply(lv, p)
        a = rz / np.vdot(p, q)
        x += a * p; r -= a * q
        rn = np.linalg.norm(r) / bn
        if rn <= rtol:
            r = b - apply(lv, x)
            if np.linalg.norm(r) / bn <= rtol:
                break
        z = mg.vcycle(r); rzn = np.vdot(r, z); p = z + (rzn / rz) * p; rz = rzn
    return x, it, np.linalg.norm(b - apply(lv, x)) / bn


if __name__ == "__main__":
    rng = np.random.RandomState(0)
    sizes = [int(s) for s in sys.argv[1].split(",")]
    for n in sizes:
        x = np.linspace(0, 1, n); X, Y = np.meshgrid(x, x)
        for case in ("smooth", "rough", "lowk"):
            if case == "lowk": k1, k2, theta = 0.5, 0.7, np.ones((n, n))
            elif case == "smooth": k1, k2, theta = 2.3, 3.7, np.ones((n, n))
            else: k1, k2, theta = 10.3, 11.9, rng.uniform(0.5, 2.0, (n, n))
            f = np.sin(2 * np.pi * k1 * X) * np.sin(2 * np.pi * k2 * Y)
            b = -f / theta
            uref = ref_solver(n).solve_poisson(f, theta, "fine") if n <= 320 else None
            for nu in (2, 1):
                mg = MGV(n, nu=nu)
                t = time.time()
                u, it, rr = pcg(b, mg)
                err = np.linalg.norm(u - uref) / np.linalg.norm(uref) if uref is not None else float("nan")
                print(f"n={n:4d} {case:6s} nu={nu} iters={it:3d} relres={rr:.2e} err={err:.2e} levels={[l.n for l in mg.lv]} ({time.time()-t:.2f}s)")
