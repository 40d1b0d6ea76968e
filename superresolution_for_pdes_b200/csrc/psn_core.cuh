// psn_core.cuh -- the numerical bodies of every kernel on the batched Poisson path.
//
// Each body is a template over an execution context `Ctx` that supplies
//   tid, nthreads, sync(), block_sum(v)
// so that the SAME source is compiled (a) by nvcc as __device__ code driven by one CTA
// (PsnDevCtx, psn_kernels.cu) and (b) by g++ as plain host code driven by a single virtual
// thread (PsnHostCtx, tests/hostemu) -- the latter is TEST INFRASTRUCTURE that lets the
// CPU-only test tier exercise the kernel logic (indexing, halos, state machine); it is never
// loaded by the product package.
//
// Operator / hierarchy conventions: see psn_plan.h.  Reference semantics being reproduced:
// src/data_generation.py:45-58 (operator), :77 (forcing), :95-104 (solve);
// src/enhanced_data_generation.py:83,96 (window extract, stride-2 downsample).
#pragma once
#include "psn_plan.h"

#if defined(__CUDACC__)
#define PSN_FN __device__ __forceinline__
#define PSN_HDFN __host__ __device__ inline
#else
#define PSN_FN inline
#define PSN_HDFN inline
#endif

#ifndef PSN_PI
#define PSN_PI 3.14159265358979323846
#endif

// ----------------------------------------------------------------------------- arguments
struct PsnRhs {
    const double* theta;  // [B, mn, mn] or null (theta == 1)
    const double* f;      // [B, mn, mn] or null
    const double* sx;     // [B, mn]: sin(2 pi k1 x_j) (host-evaluated, bit-exact forcing) or null
    const double* sy;     // [B, mn]: sin(2 pi k2 y_i) or null
    const double* k1;     // [B]: used when f, sx, sy are all null (device-side sin, synthetic mode)
    const double* k2;     // [B]
    int mn;               // master grid size: row length of theta / f / sx / sy
    int step;             // stride gather: grid node (i, j) reads master node (i*step, j*step)
};

struct PsnSolveArgs {
    PsnRhs rhs;
    double rtol;
    int max_iter;
    double* u;        // [B, n, n] or null
    int32_t* iters;   // [B]
    double* relres;   // [B]
    int32_t* status;  // [B]
    // fused sub-domain epilogue (src/enhanced_data_generation.py:136-148); sub == 0 disables it
    int sub;                 // window size (n_fine)
    const int32_t* start_x;  // [B] column offset of the window
    const int32_t* start_y;  // [B] row offset of the window
    double* u_fine;          // [B, sub, sub]
    double* f_fine;          // [B, sub, sub]
    double* u_coarse;        // [B, ceil(sub/2), ceil(sub/2)]
    double* f_coarse;        // [B, ceil(sub/2), ceil(sub/2)]
    double* th_fine;         // [B, sub, sub] window of theta (1.0 when theta is null)
    double* th_coarse;       // [B, ceil(sub/2), ceil(sub/2)]
};

// per-sample solver state of the streamed (large-grid) path
enum { PSN_PHASE_ITER = 0, PSN_PHASE_CHECK = 1, PSN_PHASE_DONE = 2 };
enum { PSN_ST_CONVERGED = 0, PSN_ST_FP64_FLOOR = 1, PSN_ST_MAXITER = 2, PSN_ST_BREAKDOWN = 3, PSN_ST_STAGNATED = 4 };

// A restart that fails to cut the true residual 4x ends the solve.  It is reported as the fp64 round-off floor
// (status 1, accepted by the callers) only when the relative true residual is below this ceiling -- the floor
// eps*||A||*||u||/||b|| is <= 0.4 n^2 eps = 3.6e-11 at n = 640; anything above it is a genuine stagnation (status 4,
// treated as a failure like max_iter).
#define PSN_FLOOR_CEIL 1e-10

struct PsnSampleState {
    double bb;         // ||bt||^2
    double rz_old;     // <r, z> of the previous iteration
    double rz;         // <r, z> of this iteration
    double pq;         // <p, A p>
    double rr;         // ||r||^2 (recursive, or true after a check)
    double prev_true;  // true ||r||^2 at the previous check
    double rr_prev;    // ||r||^2 one iteration earlier (scale of the float-stored previous search direction)
    int32_t iters;
    int32_t status;
    int32_t phase;
    int32_t restart;   // next direction is the preconditioned residual itself (beta = 0)
    int32_t nrestart;
    int32_t pad;
    double vs;         // power-of-two scale the float-stored V-cycle of the CURRENT outer iteration works with (psn_vscale of
                       // the residual norm it was started from; fused schedule: of the norm one iteration earlier)
};

// Deferred solution update (PsnCgArgs::xdefer): x is only needed when a sample is checked or finished, and the
// float-stored search directions of the last PSN_PEND_K outer iterations are still in their ring slots, so instead of
// x += alpha p every iteration (read + write of x: 16 of 83 bytes per node) the step lengths are recorded here and
// x (+)= sum_j a_j p_j is formed in one pass when the sample leaves the iteration (or the ring is full).
#define PSN_PEND_K 12
struct PsnPend {
    int32_t lo, hi;    // pending outer iterations [lo, hi): direction j lives in ring slot j % PSN_PEND_K
    int32_t xinit;     // x holds a value (0: the first flush writes instead of accumulating)
    int32_t pad;
    double a[PSN_PEND_K];   // alpha_j / s_j (the stored direction is s_j p_j)
};

struct PsnPatch {
    int oi, oj;  // global coordinates of patch element [0][0]
    int pitch;
};

// ----------------------------------------------------------------------------- contexts
PSN_FN unsigned psn_umulhi(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (unsigned)(((unsigned long long)a * (unsigned long long)b) >> 32);
#endif
}

// exact idx / d for idx * d < 2^32 (d >= 1); magic = ceil(2^32 / d)
PSN_FN unsigned psn_magic(unsigned d) { return d > 1 ? (0xFFFFFFFFu / d + 1u) : 0u; }
PSN_FN unsigned psn_fdiv(unsigned idx, unsigned magic) { return magic ? psn_umulhi(idx, magic) : idx; }

PSN_FN int psn_stagnation_status(double rt, double bb, double rtol) {
    const double c = rtol > PSN_FLOOR_CEIL ? rtol : PSN_FLOOR_CEIL;
    return (rt <= c * c * bb) ? PSN_ST_FP64_FLOOR : PSN_ST_STAGNATED;
}

// power-of-two scale that brings a residual of squared norm rr to O(1) before it is rounded to fp32
// (exact in fp64, makes the fp32-stored V-cycle independent of the magnitude of the data)
PSN_HDFN double psn_vscale(double rr) {
    if (!(rr > 0.0) || !(rr < 1e300)) return 1.0;
    int e = 0;
    frexp(rr, &e);
    return ldexp(1.0, -(e / 2));
}

PSN_FN int psn_imax(int a, int b) { return a > b ? a : b; }
PSN_FN int psn_imin(int a, int b) { return a < b ? a : b; }

PSN_FN double psn_ldg(const double* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

struct PsnHostCtx {
    int tid = 0, nthreads = 1;
    void sync() {}
    double block_sum(double v) { return v; }
};

#if defined(__CUDACC__)
// One CTA.  block_sum is deterministic: xor-butterfly inside each warp (every lane ends with
// the bit-identical value), one partial per warp in shared memory, fixed-order final sum.
struct PsnDevCtx {
    int tid, nthreads;
    double* red;  // shared scratch, 2 x 32 doubles
    int flip;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ double block_sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        double* buf = red + (flip << 5);
        flip ^= 1;
        if ((tid & 31) == 0) buf[tid >> 5] = v;
        __syncthreads();
        // fixed order, four interleaved chains: the dependent-add depth is nw / 4 + 2 instead of nw (every thread adds the same
        // values in the same order, so the result is identical across the CTA and from run to run)
        const int nw = (nthreads + 31) >> 5;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int w = 0;
        for (; w + 3 < nw; w += 4) {
            s0 += buf[w];
            s1 += buf[w + 1];
            s2 += buf[w + 2];
            s3 += buf[w + 3];
        }
        for (; w < nw; ++w) s0 += buf[w];
        return (s0 + s1) + (s2 + s3);
    }
};
#endif

#if defined(__CUDACC__)
// One warp acting as a (small) CTA: used for the tiny coarse levels, where a CTA-wide barrier per half
// sweep costs far more than the sweep itself.
struct PsnWarpCtxDev {
    int tid, nthreads;
    __device__ __forceinline__ void sync() { __syncwarp(); }
};
#endif

// ----------------------------------------------------------------------------- forcing / rhs
PSN_FN double psn_forcing(const PsnRhs& s, long long b, int i, int j) {
    const long long mi = (long long)i * s.step, mj = (long long)j * s.step;
    if (s.f) return s.f[(b * s.mn + mi) * s.mn + mj];
    if (s.sx) return s.sy[b * s.mn + mi] * s.sx[b * s.mn + mj];
    const double inv = 1.0 / (double)(s.mn - 1);
    return sin(2.0 * PSN_PI * s.k1[b] * ((double)mj * inv)) * sin(2.0 * PSN_PI * s.k2[b] * ((double)mi * inv));
}

// bt = -(f / theta) * h^2   (F1/F2 of SURVEY.md: reference solves diag(theta) L u = f, L negative definite)
PSN_FN double psn_rhs(const PsnRhs& s, long long b, int i, int j, double h2) {
    double f = psn_forcing(s, b, i, j);
    if (s.theta) {
        const long long mi = (long long)i * s.step, mj = (long long)j * s.step;
        // f * (1 / theta): a correctly rounded reciprocal + multiply is half the instructions of an fp64 division; the
        // result differs from f / theta by at most one ulp (far below the 1e-10 parity bar) and every kernel that
        // forms bt uses this same expression
        f = f * (1.0 / s.theta[(b * s.mn + mi) * s.mn + mj]);
    }
    return -(f * h2);
}

// ----------------------------------------------------------------------------- patch operations
// One red-black Gauss-Seidel half sweep over the window [wi0,wi1) x [wj0,wj1) (global coords,
// already clipped to the domain).  colour 0 = (i + j) even.  `first` = zero initial guess.
template <class Ctx>
PSN_FN void psn_half_sweep(Ctx& ctx, double* u, const double* b, PsnPatch g, const PsnLevel& lv,
                           int wi0, int wi1, int wj0, int wj1, int color, bool first) {
    const int rows = wi1 - wi0;
    const int w2 = (wj1 - wj0 + 1) >> 1;
    if (rows <= 0 || w2 <= 0) return;
    const unsigned cnt = (unsigned)rows * (unsigned)w2;
    const unsigned magic = psn_magic((unsigned)w2);
    const int last = lv.n - 1;
    const int pitch = g.pitch;
    for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
        const unsigned r = psn_fdiv(idx, magic);
        const int k = (int)(idx - r * (unsigned)w2);
        const int i = wi0 + (int)r;
        const int j = wj0 + ((i + wj0 + color) & 1) + 2 * k;
        if (j >= wj1) continue;
        const int c = (i - g.oi) * pitch + (j - g.oj);
        const double s = first ? 0.0 : (u[c - 1] + u[c + 1]) + (u[c - pitch] + u[c + pitch]);
        const double rd = (i == last) ? ((j == last) ? lv.rd2 : lv.rd1) : ((j == last) ? lv.rd1 : lv.rd0);
        u[c] = (b[c] + s) * rd;
    }
}


// Half sweep over a WHOLE level stored in the ghost-ringed shared layout (pitch n + 2).
// Even n: threads walk wrapped grid diagonals -- consecutive lanes are (i, j), (i+1, j+1), ... --
// all nodes of a diagonal have the same colour, and with the even pitch the lane stride (pitch + 1)
// is odd, so the node and its four neighbours are read bank-conflict free (a row-wise mapping of
// one colour strides by two doubles: 2-way conflicts).  Odd n falls back to the generic window sweep.
template <class Ctx>
PSN_FN void psn_half_sweep_full(Ctx& ctx, double* u, const double* b, const PsnLevel& lv, int color, bool first) {
    const int n = lv.n, P = lv.pitch;
    if (n & 1) {
        psn_half_sweep(ctx, u, b, PsnPatch{-1, -1, P}, lv, 0, n, 0, n, color, first);
        return;
    }
    const unsigned cnt = (unsigned)(n * (n >> 1));
    const unsigned magic = lv.nmagic;
    const int last = n - 1;
    // (a 4-way manually unrolled variant that issues all loads first was measured 13 % SLOWER on B200:
    //  the extra predication outweighs the added ILP)
    if (lv.gam == 0.0) {   // finest level: constant diagonal, no boundary selects
        for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
            const unsigned d2 = psn_fdiv(idx, magic);
            const int i = (int)(idx - d2 * (unsigned)n);
            int j = 2 * (int)d2 + color + i;
            j = (j >= n) ? j - n : j;
            const int c = (i + 1) * P + j + 1;
            const double s = first ? 0.0 : (u[c - 1] + u[c + 1]) + (u[c - P] + u[c + P]);
            u[c] = (b[c] + s) * 0.25;
        }
    } else {
        for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
            const unsigned d2 = psn_fdiv(idx, magic);
            const int i = (int)(idx - d2 * (unsigned)n);
            int j = 2 * (int)d2 + color + i;
            j = (j >= n) ? j - n : j;
            const int c = (i + 1) * P + j + 1;
            const double s = first ? 0.0 : (u[c - 1] + u[c + 1]) + (u[c - P] + u[c + P]);
            const double rd = (i == last) ? ((j == last) ? lv.rd2 : lv.rd1) : ((j == last) ? lv.rd1 : lv.rd0);
            u[c] = (b[c] + s) * rd;
        }
    }
}

PSN_FN double psn_resid_at(const double* u, const double* b, PsnPatch g, const PsnLevel& lv, int i, int j) {
    if (i < 0 || j < 0 || i >= lv.n || j >= lv.n) return 0.0;
    const int c = (i - g.oi) * g.pitch + (j - g.oj);
    const int last = lv.n - 1;
    const double d = 4.0 + (i == last ? lv.gam : 0.0) + (j == last ? lv.gam : 0.0);
    return b[c] - (d * u[c] - ((u[c - 1] + u[c + 1]) + (u[c - g.pitch] + u[c + g.pitch])));
}

// Coarse right-hand side for coarse nodes [I0,I1) x [J0,J1).  After a black half sweep the
// residual vanishes on black nodes, so only the centre and the four corner (red) nodes of the
// full-weighting stencil contribute:  bc = rho(c) + (rho(4 corners)) / 4   (scaled form).
template <class Ctx>
PSN_FN void psn_residual_restrict(Ctx& ctx, const double* u, const double* b, PsnPatch g, const PsnLevel& lv,
                                  double* out, PsnPatch go, int I0, int I1, int J0, int J1, unsigned cols_magic = 1u) {
    const int rows = I1 - I0, cols = J1 - J0;
    if (rows <= 0 || cols <= 0) return;
    const unsigned cnt = (unsigned)rows * (unsigned)cols;
    const unsigned magic = (cols_magic != 1u) ? cols_magic : psn_magic((unsigned)cols);   // 1 = "not precomputed"
    for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
        const unsigned r = psn_fdiv(idx, magic);
        const int I = I0 + (int)r;
        const int J = J0 + (int)(idx - r * (unsigned)cols);
        const int i = 2 * I + 1, j = 2 * J + 1;
        const double rc = psn_resid_at(u, b, g, lv, i, j);
        const double r00 = psn_resid_at(u, b, g, lv, i - 1, j - 1);
        const double r02 = psn_resid_at(u, b, g, lv, i - 1, j + 1);
        const double r20 = psn_resid_at(u, b, g, lv, i + 1, j - 1);
        const double r22 = psn_resid_at(u, b, g, lv, i + 1, j + 1);
        out[(long long)(I - go.oi) * go.pitch + (J - go.oj)] = rc + 0.25 * ((r00 + r02) + (r20 + r22));
    }
}

// u += P e on the RED nodes of the window only (the black nodes are overwritten by the black
// half sweep that follows and never read before).  e must read as zero outside the coarse grid.
template <class Ctx>
PSN_FN void psn_prolong_red(Ctx& ctx, double* u, PsnPatch g, const double* e, PsnPatch ge,
                            int wi0, int wi1, int wj0, int wj1) {
    const int rows = wi1 - wi0;
    const int w2 = (wj1 - wj0 + 1) >> 1;
    if (rows <= 0 || w2 <= 0) return;
    const unsigned cnt = (unsigned)rows * (unsigned)w2;
    const unsigned magic = psn_magic((unsigned)w2);
    for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
        const unsigned r = psn_fdiv(idx, magic);
        const int k = (int)(idx - r * (unsigned)w2);
        const int i = wi0 + (int)r;
        const int j = wj0 + ((i + wj0) & 1) + 2 * k;
        if (j >= wj1) continue;
        const int c = (i - g.oi) * g.pitch + (j - g.oj);
        double add;
        if (i & 1) {  // (odd, odd): coincides with coarse node ((i-1)/2, (j-1)/2)
            add = e[(((i - 1) >> 1) - ge.oi) * ge.pitch + (((j - 1) >> 1) - ge.oj)];
        } else {      // (even, even): centre of four coarse nodes
            const int ce = ((i >> 1) - ge.oi) * ge.pitch + ((j >> 1) - ge.oj);
            add = 0.25 * ((e[ce - ge.pitch - 1] + e[ce - ge.pitch]) + (e[ce - 1] + e[ce]));
        }
        u[c] += add;
    }
}


// u += P e on the red nodes of a WHOLE level in the ghost-ringed layout; same diagonal thread mapping
// as psn_half_sweep_full for even n.
template <class Ctx>
PSN_FN void psn_prolong_red_full(Ctx& ctx, double* u, const PsnLevel& lv, const double* e, const PsnLevel& lc) {
    const int n = lv.n, P = lv.pitch, Pc = lc.pitch;
    if (n & 1) {
        psn_prolong_red(ctx, u, PsnPatch{-1, -1, P}, e, PsnPatch{-1, -1, Pc}, 0, n, 0, n);
        return;
    }
    const unsigned cnt = (unsigned)(n * (n >> 1));
    const unsigned magic = lv.nmagic;
    for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
        const unsigned d2 = psn_fdiv(idx, magic);
        const int i = (int)(idx - d2 * (unsigned)n);
        int j = 2 * (int)d2 + i;
        j = (j >= n) ? j - n : j;
        const int c = (i + 1) * P + j + 1;
        double add;
        if (i & 1) {
            add = e[(((i - 1) >> 1) + 1) * Pc + ((j - 1) >> 1) + 1];
        } else {
            const int ce = ((i >> 1) + 1) * Pc + (j >> 1) + 1;
            add = 0.25 * ((e[ce - Pc - 1] + e[ce - Pc]) + (e[ce - 1] + e[ce]));
        }
        u[c] += add;
    }
}

// sum over groups of G adjacent lanes (G a power of two <= 32); every lane of a group ends with the group's sum
PSN_FN double psn_group_sum(PsnHostCtx&, double v, int) { return v; }
#if defined(__CUDACC__)
template <class Ctx>
__device__ __forceinline__ double psn_group_sum(Ctx&, double v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

// Coarsest level: u = inv * b (dense, N = nc*nc <= 100), both in shared level layout.  The N dot products are
// spread over the CTA: G = 2^k adjacent lanes share one output (interleaved k, fixed-order partial sums, one
// xor-butterfly), so a 100 x 100 solve is 25 independent loads + FMAs per thread instead of 100 dependent ones.
// inv_s: a copy of the inverse in shared memory (persistent resident kernel), or null: read it through L2.
template <class Ctx>
PSN_FN void psn_coarsest(Ctx& ctx, const PsnPlan& plan, const double* b, double* u, const double* inv_s = nullptr) {
    const PsnLevel& lv = plan.lv[plan.nlev - 1];
    const int nc = lv.n, N = nc * nc, pitch = lv.pitch;
    int G = 1;
    while (2 * G * N <= ctx.nthreads && G < 32) G *= 2;   // CTA sizes are multiples of 32 and G | 32: no group straddles a warp
    const int g = ctx.tid & (G - 1);
    const int per = ctx.nthreads >= G ? ctx.nthreads / G : 1;   // outputs per pass
    for (int base = 0; base < N; base += per) {                 // same trip count for every thread (the butterfly is warp-wide)
        const int t = base + ctx.tid / G;
        double acc = 0.0;
        if (t < N) {
            if (inv_s) {
                for (int k = g; k < N; k += G) {
                    const int ki = (int)psn_fdiv((unsigned)k, lv.nmagic);
                    acc += inv_s[k * N + t] * b[k + 2 * ki + pitch + 1];   // (ki+1)*pitch + kj+1
                }
            } else {
                for (int k = g; k < N; k += G) {
                    const int ki = (int)psn_fdiv((unsigned)k, lv.nmagic);
                    acc += psn_ldg(plan.coarse_inv + (size_t)k * N + t) * b[k + 2 * ki + pitch + 1];
                }
            }
        }
        acc = psn_group_sum(ctx, acc, G);
        if (t < N && g == 0) {
            const int ti = (int)psn_fdiv((unsigned)t, lv.nmagic);
            u[t + 2 * ti + pitch + 1] = acc;
        }
    }
}

// smem offset (in doubles, relative to the pool that starts at level 1) of level l's rhs array;
// its solution array follows at +size.
PSN_FN int psn_pool_off(const PsnPlan& plan, int l) { return (l >= 1 && l < plan.nlev) ? plan.lv[l].off : 0; }

// V(nu,nu) cycle entirely in shared memory for levels l0 .. nlev-1.
// b0 / u0: level-l0 arrays in shared level layout (ghost cells of u0 must be zero);
// pool: storage for levels l0+1.. (rhs then solution per level), zero-initialised once.
// Ends with a sync; u0 then holds M^{-1} b0.
template <class Ctx>
PSN_FN void psn_vcycle_smem(Ctx& ctx, const PsnPlan& plan, int l0, const double* b0, double* u0, double* pool,
                            const double* inv_s = nullptr) {
    const int L = plan.nlev;
    const int base = psn_pool_off(plan, l0 + 1);
    for (int l = l0; l < L - 1; ++l) {
        const PsnLevel& lv = plan.lv[l];
        const PsnLevel& lc = plan.lv[l + 1];
        const int nu = plan.nul[l];
        const int off = psn_pool_off(plan, l) - base;
        const double* bl = (l == l0) ? b0 : pool + off;
        double* ul = (l == l0) ? u0 : pool + off + lv.size;
        double* bc = pool + (psn_pool_off(plan, l + 1) - base);
        const PsnPatch g{-1, -1, lv.pitch};
        const PsnPatch gc{-1, -1, lc.pitch};
        psn_half_sweep_full(ctx, ul, bl, lv, 0, true);
        ctx.sync();
        psn_half_sweep_full(ctx, ul, bl, lv, 1, false);
        ctx.sync();
        for (int s = 1; s < nu; ++s) {
            psn_half_sweep_full(ctx, ul, bl, lv, 0, false);
            ctx.sync();
            psn_half_sweep_full(ctx, ul, bl, lv, 1, false);
            ctx.sync();
        }
        psn_residual_restrict(ctx, ul, bl, g, lv, bc, gc, 0, lc.n, 0, lc.n, lc.nmagic);
        ctx.sync();
    }
    {
        const int l = L - 1;
        const int off = psn_pool_off(plan, l) - base;
        const double* bl = (l == l0) ? b0 : pool + off;
        double* ul = (l == l0) ? u0 : pool + off + plan.lv[l].size;
        psn_coarsest(ctx, plan, bl, ul, inv_s);
        ctx.sync();
    }
    for (int l = L - 2; l >= l0; --l) {
        const PsnLevel& lv = plan.lv[l];
        const PsnLevel& lc = plan.lv[l + 1];
        const int nu = plan.nul[l];
        const int off = psn_pool_off(plan, l) - base;
        const double* bl = (l == l0) ? b0 : pool + off;
        double* ul = (l == l0) ? u0 : pool + off + lv.size;
        const double* uc = pool + (psn_pool_off(plan, l + 1) - base) + lc.size;
        const PsnPatch g{-1, -1, lv.pitch};
        const PsnPatch gc{-1, -1, lc.pitch};
        psn_prolong_red_full(ctx, ul, lv, uc, lc);
        ctx.sync();
        for (int s = 0; s < nu; ++s) {
            psn_half_sweep_full(ctx, ul, bl, lv, 1, false);
            ctx.sync();
            psn_half_sweep_full(ctx, ul, bl, lv, 0, false);
            ctx.sync();
        }
    }
}

// doubles of shared memory needed by psn_vcycle_smem's pool for levels l0+1..
PSN_HDFN int psn_pool_doubles(const PsnPlan& plan, int l0) {
    int s = 0;
    for (int l = l0 + 1; l < plan.nlev; ++l) s += 2 * plan.lv[l].size;
    return s;
}
// doubles of shared memory needed by the fully resident solver (r, p, z + pool)
PSN_HDFN int psn_smem_solver_doubles(const PsnPlan& plan) { return 3 * plan.lv[0].size + psn_pool_doubles(plan, 0); }

// ----------------------------------------------------------------------------- resident solver (n small)
// One CTA solves one sample completely in shared memory: MG-preconditioned CG with the
// iterate x held in registers (XPT values per thread).  Writes u (and / or the fused
// window-extract + stride-2 downsample outputs), iteration count, true relative residual, status.
template <int XPT, class Ctx>
PSN_FN void psn_smem_solve_body(Ctx& ctx, const PsnPlan& plan, const PsnSolveArgs& a, long long b, double* smem) {
    const PsnLevel& l0 = plan.lv[0];
    const int n = l0.n, P = l0.pitch, S0 = l0.size, nn = n * n;
    double* r = smem;
    double* p = smem + S0;
    double* z = smem + 2 * S0;
    double* pool = smem + 3 * S0;
    const int total = 3 * S0 + psn_pool_doubles(plan, 0);
    for (int idx = ctx.tid; idx < total; idx += ctx.nthreads) smem[idx] = 0.0;

    const unsigned nmagic = l0.nmagic;
    int cx[XPT];
    double x[XPT];
#pragma unroll
    for (int k = 0; k < XPT; ++k) {
        const int idx = ctx.tid + k * ctx.nthreads;
        x[k] = 0.0;
        cx[k] = -1;
        if (idx < nn) cx[k] = idx + 2 * (int)psn_fdiv((unsigned)idx, nmagic) + P + 1;   // (i+1)*P + j+1 with P = n+2
    }
    ctx.sync();

    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < XPT; ++k) {
        if (cx[k] >= 0) {
            const int idx = ctx.tid + k * ctx.nthreads;
            const int i = (int)psn_fdiv((unsigned)idx, nmagic), j = idx - i * n;
            const double v = psn_rhs(a.rhs, b, i, j, plan.h2);
            r[cx[k]] = v;
            acc += v * v;
        }
    }
    const double bb = ctx.block_sum(acc);
    const double tol2 = a.rtol * a.rtol * bb;

    int it = 0, status = PSN_ST_MAXITER, nrestart = 0;
    bool restart = true;
    double rz_old = 1.0, prev_true = 0.0, rr_final = bb;
    if (!(bb > 0.0)) {
        status = (bb == 0.0) ? PSN_ST_CONVERGED : PSN_ST_BREAKDOWN;  // zero rhs -> u = 0; NaN rhs -> breakdown
        rr_final = 0.0;
    } else {
        for (;;) {
            psn_vcycle_smem(ctx, plan, 0, r, z, pool);
            acc = 0.0;
#pragma unroll
            for (int k = 0; k < XPT; ++k)
                if (cx[k] >= 0) acc += r[cx[k]] * z[cx[k]];
            const double rz = ctx.block_sum(acc);
            const double beta = restart ? 0.0 : rz / rz_old;
            restart = false;
            rz_old = rz;
#pragma unroll
            for (int k = 0; k < XPT; ++k)
                if (cx[k] >= 0) p[cx[k]] = z[cx[k]] + beta * p[cx[k]];
            ctx.sync();
            acc = 0.0;
#pragma unroll
            for (int k = 0; k < XPT; ++k) {
                if (cx[k] >= 0) {
                    const int c = cx[k];
                    const double pc = p[c];
                    const double q = 4.0 * pc - ((p[c - 1] + p[c + 1]) + (p[c - P] + p[c + P]));
                    z[c] = q;
                    acc += pc * q;
                }
            }
            const double pq = ctx.block_sum(acc);
            if (!(pq > 0.0)) {
                status = PSN_ST_BREAKDOWN;
                break;
            }
            const double alpha = rz / pq;
            acc = 0.0;
#pragma unroll
            for (int k = 0; k < XPT; ++k) {
                if (cx[k] >= 0) {
                    const int c = cx[k];
                    x[k] += alpha * p[c];
                    const double rv = r[c] - alpha * z[c];
                    r[c] = rv;
                    acc += rv * rv;
                }
            }
            const double rr = ctx.block_sum(acc);
            rr_final = rr;
            ++it;
            if (rr <= tol2 || it >= a.max_iter) {
                // true residual r = bt - At x (x staged through p, which a restart rebuilds anyway)
#pragma unroll
                for (int k = 0; k < XPT; ++k)
                    if (cx[k] >= 0) p[cx[k]] = x[k];
                ctx.sync();
                acc = 0.0;
#pragma unroll
                for (int k = 0; k < XPT; ++k) {
                    if (cx[k] >= 0) {
                        const int c = cx[k];
                        const int idx = ctx.tid + k * ctx.nthreads;
                        const int i = (int)psn_fdiv((unsigned)idx, nmagic), j = idx - i * n;
                        const double bt = psn_rhs(a.rhs, b, i, j, plan.h2);
                        const double rv = bt - (4.0 * p[c] - ((p[c - 1] + p[c + 1]) + (p[c - P] + p[c + P])));
                        r[c] = rv;
                        acc += rv * rv;
                    }
                }
                const double rt = ctx.block_sum(acc);
                rr_final = rt;
                if (rt <= tol2) { status = PSN_ST_CONVERGED; break; }
                if (nrestart >= 1 && rt > prev_true * 0.0625) { status = psn_stagnation_status(rt, bb, a.rtol); break; }
                if (it >= a.max_iter) { status = PSN_ST_MAXITER; break; }
                prev_true = rt;
                ++nrestart;
                restart = true;
            }
        }
    }

    if (a.u) {
        double* ug = a.u + b * nn;
#pragma unroll
        for (int k = 0; k < XPT; ++k)
            if (cx[k] >= 0) ug[ctx.tid + k * ctx.nthreads] = x[k];
    }
    if (a.sub > 0) {
        const int s = a.sub, sc = (s + 1) >> 1;
        const int x0 = a.start_x[b], y0 = a.start_y[b];
#pragma unroll
        for (int k = 0; k < XPT; ++k) {
            if (cx[k] >= 0) {
                const int idx = ctx.tid + k * ctx.nthreads;
                const int i = (int)psn_fdiv((unsigned)idx, nmagic), j = idx - i * n;
                const int wi = i - y0, wj = j - x0;
                if (wi >= 0 && wi < s && wj >= 0 && wj < s) {
                    const double fv = psn_forcing(a.rhs, b, i, j);
                    const double tv = a.rhs.theta
                        ? a.rhs.theta[(b * a.rhs.mn + (long long)i * a.rhs.step) * a.rhs.mn + (long long)j * a.rhs.step] : 1.0;
                    const long long o = (b * s + wi) * s + wj;
                    if (a.u_fine) a.u_fine[o] = x[k];
                    if (a.f_fine) a.f_fine[o] = fv;
                    if (a.th_fine) a.th_fine[o] = tv;
                    if (((wi | wj) & 1) == 0) {
                        const long long oc = (b * sc + (wi >> 1)) * sc + (wj >> 1);
                        if (a.u_coarse) a.u_coarse[oc] = x[k];
                        if (a.f_coarse) a.f_coarse[oc] = fv;
                        if (a.th_coarse) a.th_coarse[oc] = tv;
                    }
                }
            }
        }
    }
    if (ctx.tid == 0) {
        if (a.iters) a.iters[b] = it;
        if (a.relres) a.relres[b] = (bb > 0.0) ? sqrt(rr_final / bb) : 0.0;
        if (a.status) a.status[b] = status;
    }
}

// ----------------------------------------------------------------------------- streamed path: tiles
struct PsnTileArgs {
    int level;         // level l of the plan this launch works on
    int TH, TW;        // interior tile size (even)
    int ntx, nty;      // tiles per row / column of one sample
    const double* b;   // [B, n_l, n_l] right-hand side of level l
    const double* uin; // post: [B, n_l, n_l] pre-smoothed iterate
    double* uout;      // [B, n_l, n_l]
    double* bc;        // pre:  [B, n_{l+1}, n_{l+1}] restricted residual (output)
    const double* ec;  // post: [B, n_{l+1}, n_{l+1}] coarse correction (input)
    double* part;      // post, level 0: [B, ntx*nty] partial <b, u>; null otherwise
    const PsnSampleState* state;  // skip samples whose phase != PSN_PHASE_ITER (null: none skipped)
};

PSN_HDFN int psn_even(int v) { return (v + 1) & ~1; }
// shared-memory doubles for the pre / post tile kernels
PSN_HDFN int psn_tile_pre_doubles(int TH, int TW, int nu) {
    const int PR = TH + 4 * nu + 1, PC = TW + 4 * nu + 1;
    return 2 * psn_even(PR * (PC | 1));
}
PSN_HDFN int psn_tile_post_doubles(int TH, int TW, int nu) {
    const int PR = TH + 4 * nu, PC = TW + 4 * nu;
    const int CPR = TH / 2 + 2 * nu + 1, CPC = TW / 2 + 2 * nu + 1;
    return 2 * psn_even(PR * (PC | 1)) + psn_even(CPR * CPC);
}

// Down-leg of the V-cycle on a streamed level, one tile: nu RB-GS sweeps from a zero guess
// (temporal blocking in shared memory: the valid window shrinks by one node per half sweep),
// then the restricted residual.  Reads b once (+halo), writes u and the coarse rhs.
template <class Ctx>
PSN_FN void psn_tile_pre_body(Ctx& ctx, const PsnPlan& plan, const PsnTileArgs& a, long long b, int ty, int tx, double* smem) {
    if (a.state && a.state[b].phase != PSN_PHASE_ITER) return;
    const PsnLevel& lv = plan.lv[a.level];
    const PsnLevel& lc = plan.lv[a.level + 1];
    const int n = lv.n, nc = lc.n, nu = plan.nul[a.level];
    const int Hlo = 2 * nu, Hhi = 2 * nu + 1;
    const int i0 = ty * a.TH, j0 = tx * a.TW;
    const int PR = a.TH + Hlo + Hhi, PC = a.TW + Hlo + Hhi, pitch = PC | 1;
    const PsnPatch g{i0 - Hlo, j0 - Hlo, pitch};
    double* sb = smem;
    double* su = smem + psn_even(PR * pitch);
    const double* bg = a.b + b * (long long)n * n;

    {
        const unsigned cnt = (unsigned)(PR * PC), magic = psn_magic((unsigned)PC);
        for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
            const unsigned pr = psn_fdiv(idx, magic);
            const int pc = (int)(idx - pr * (unsigned)PC);
            const int gi = g.oi + (int)pr, gj = g.oj + pc;
            const bool in = gi >= 0 && gi < n && gj >= 0 && gj < n;
            sb[pr * pitch + pc] = in ? bg[(long long)gi * n + gj] : 0.0;
            su[pr * pitch + pc] = 0.0;
        }
    }
    ctx.sync();
    for (int s = 0; s < 2 * nu; ++s) {
        const int wi0 = psn_imax(g.oi + s, 0), wi1 = psn_imin(g.oi + PR - s, n);
        const int wj0 = psn_imax(g.oj + s, 0), wj1 = psn_imin(g.oj + PC - s, n);
        psn_half_sweep(ctx, su, sb, g, lv, wi0, wi1, wj0, wj1, s & 1, s == 0);
        ctx.sync();
    }
    const PsnPatch go{0, 0, nc};
    psn_residual_restrict(ctx, su, sb, g, lv, a.bc + b * (long long)nc * nc, go,
                          i0 >> 1, psn_imin((i0 + a.TH) >> 1, nc), j0 >> 1, psn_imin((j0 + a.TW) >> 1, nc));
    {
        double* ug = a.uout + b * (long long)n * n;
        const unsigned cnt = (unsigned)(a.TH * a.TW), magic = psn_magic((unsigned)a.TW);
        for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
            const unsigned ii = psn_fdiv(idx, magic);
            const int gi = i0 + (int)ii, gj = j0 + (int)(idx - ii * (unsigned)a.TW);
            if (gi < n && gj < n) ug[(long long)gi * n + gj] = su[(gi - g.oi) * pitch + (gj - g.oj)];
        }
    }
}

// Up-leg on a streamed level, one tile: add the bilinear prolongation of the coarse correction
// (red nodes only), nu reverse-order (black, red) sweeps; on the finest level also the partial <b, u>.
template <class Ctx>
PSN_FN void psn_tile_post_body(Ctx& ctx, const PsnPlan& plan, const PsnTileArgs& a, long long b, int ty, int tx, double* smem) {
    if (a.state && a.state[b].phase != PSN_PHASE_ITER) return;
    const PsnLevel& lv = plan.lv[a.level];
    const PsnLevel& lc = plan.lv[a.level + 1];
    const int n = lv.n, nc = lc.n, nu = plan.nul[a.level];
    const int H = 2 * nu;
    const int i0 = ty * a.TH, j0 = tx * a.TW;
    const int PR = a.TH + 2 * H, PC = a.TW + 2 * H, pitch = PC | 1;
    const PsnPatch g{i0 - H, j0 - H, pitch};
    const int CPR = a.TH / 2 + H + 1, CPC = a.TW / 2 + H + 1;
    const PsnPatch ge{(i0 - H) / 2 - 1, (j0 - H) / 2 - 1, CPC};
    double* sb = smem;
    double* su = smem + psn_even(PR * pitch);
    double* se = smem + 2 * psn_even(PR * pitch);
    const double* bg = a.b + b * (long long)n * n;
    const double* ug_in = a.uin + b * (long long)n * n;
    const double* eg = a.ec + b * (long long)nc * nc;
    {
        const unsigned cnt = (unsigned)(PR * PC), magic = psn_magic((unsigned)PC);
        for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
            const unsigned pr = psn_fdiv(idx, magic);
            const int pc = (int)(idx - pr * (unsigned)PC);
            const int gi = g.oi + (int)pr, gj = g.oj + pc;
            const bool in = gi >= 0 && gi < n && gj >= 0 && gj < n;
            sb[pr * pitch + pc] = in ? bg[(long long)gi * n + gj] : 0.0;
            su[pr * pitch + pc] = in ? ug_in[(long long)gi * n + gj] : 0.0;
        }
        const unsigned ccnt = (unsigned)(CPR * CPC), cmagic = psn_magic((unsigned)CPC);
        for (unsigned idx = ctx.tid; idx < ccnt; idx += ctx.nthreads) {
            const unsigned pr = psn_fdiv(idx, cmagic);
            const int pc = (int)(idx - pr * (unsigned)CPC);
            const int gi = ge.oi + (int)pr, gj = ge.oj + pc;
            const bool in = gi >= 0 && gi < nc && gj >= 0 && gj < nc;
            se[pr * CPC + pc] = in ? eg[(long long)gi * nc + gj] : 0.0;
        }
    }
    ctx.sync();
    psn_prolong_red(ctx, su, g, se, ge, psn_imax(g.oi, 0), psn_imin(g.oi + PR, n), psn_imax(g.oj, 0), psn_imin(g.oj + PC, n));
    ctx.sync();
    for (int s = 0; s < 2 * nu; ++s) {
        const int sh = s + 1;
        const int wi0 = psn_imax(g.oi + sh, 0), wi1 = psn_imin(g.oi + PR - sh, n);
        const int wj0 = psn_imax(g.oj + sh, 0), wj1 = psn_imin(g.oj + PC - sh, n);
        psn_half_sweep(ctx, su, sb, g, lv, wi0, wi1, wj0, wj1, (s & 1) ^ 1, false);
        ctx.sync();
    }
    double acc = 0.0;
    {
        double* ug = a.uout + b * (long long)n * n;
        const unsigned cnt = (unsigned)(a.TH * a.TW), magic = psn_magic((unsigned)a.TW);
        for (unsigned idx = ctx.tid; idx < cnt; idx += ctx.nthreads) {
            const unsigned ii = psn_fdiv(idx, magic);
            const int gi = i0 + (int)ii, gj = j0 + (int)(idx - ii * (unsigned)a.TW);
            if (gi < n && gj < n) {
                const int c = (gi - g.oi) * pitch + (gj - g.oj);
                const double v = su[c];
                ug[(long long)gi * n + gj] = v;
                acc += v * sb[c];
            }
        }
    }
    if (a.part) {
        const double s = ctx.block_sum(acc);
        if (ctx.tid == 0) a.part[b * (long long)(a.ntx * a.nty) + (long long)ty * a.ntx + tx] = s;
    }
}

// Coarse part of a streamed V-cycle: levels l0.. of one sample entirely in shared memory.
template <class Ctx, class TV>
PSN_FN void psn_coarse_vcycle_body(Ctx& ctx, const PsnPlan& plan, int l0, const TV* bglob, TV* uglob,
                                   const PsnSampleState* state, long long b, double* smem) {
    if (state && state[b].phase != PSN_PHASE_ITER) return;
    const PsnLevel& lv = plan.lv[l0];
    const int n = lv.n, P = lv.pitch, nn = n * n;
    double* sb = smem;
    double* su = smem + lv.size;
    double* pool = smem + 2 * lv.size;
    const int total = 2 * lv.size + psn_pool_doubles(plan, l0);
    for (int idx = ctx.tid; idx < total; idx += ctx.nthreads) smem[idx] = 0.0;
    ctx.sync();
    const unsigned nmagic = lv.nmagic;
    const TV* bg = bglob + b * (long long)nn;
    for (int idx = ctx.tid; idx < nn; idx += ctx.nthreads) sb[idx + 2 * (int)psn_fdiv((unsigned)idx, nmagic) + P + 1] = (double)bg[idx];
    ctx.sync();
    psn_vcycle_smem(ctx, plan, l0, sb, su, pool);
    TV* ug = uglob + b * (long long)nn;
    for (int idx = ctx.tid; idx < nn; idx += ctx.nthreads) ug[idx] = (TV)su[idx + 2 * (int)psn_fdiv((unsigned)idx, nmagic) + P + 1];
}

// ----------------------------------------------------------------------------- streamed path: CG kernels (level 0, global memory)
// 2-D blocks of PSN_GT_ROWS x PSN_GT_COLS nodes; thread t handles column t % COLS and rows
// t / COLS + k * (nthreads / COLS).
#define PSN_GT_ROWS 32
#define PSN_GT_COLS 64

struct PsnCgArgs {
    int n;
    int ntx, nty;            // blocks per row / column
    int ntiles_rz;           // number of <r,z> partials per sample (post-kernel tiles)
    PsnRhs rhs;
    double h2;
    double* x;               // [B, n, n]
    double* r;
    double* z;
    double* p_old;
    double* p_new;
    double* q;
    double* part_rz;         // [B, ntiles_rz]
    double* part_a;          // [B, ntx*nty]   ||b||^2 partials (prologue)
    double* part_pq;         // [B, npq]       <p,q> partials
    int npq;                 // partials per sample written by the direction kernel
    int dir_mode;            // 0: 2-D tile kernel, 1: warp strips + register prefetch, 2: warp strips + TMA ring
    int dir_band, dir_nstrips, dir_nbands;   // strip decomposition of the direction kernel
    int vf32;                // z is stored as float, scaled by psn_vscale(||r||^2) (PsnStreamCfg::vf32)
    int upd_mode;            // 0: 2-D tile update that reads q; 2: strip update that recomputes q from p (q is never stored)
    int upd_band, upd_nstrips, upd_nbands;   // strip decomposition of the update kernel
    int pf32;                // the search direction is stored as float, scaled by psn_vscale(||r||^2) (needs vf32 and upd_mode 2)
    int xdefer;              // needs pf32: x is not touched by the update kernel; see PsnPend
    int it;                  // outer iteration index of this launch (ring slot of p_new = it % PSN_PEND_K)
    const float* pring;      // [PSN_PEND_K][B, n, n] ring of float search directions (xdefer)
    long long ring_stride;   // floats between two ring slots
    PsnPend* pend;           // [B]
    int fused;               // the CG update runs inside the level-0 down-leg of the next V-cycle (psn_strip_upre_body):
                             // r -> r_out double buffer, ||r||^2 partials indexed by the down-leg's tasks (fu_nparts)
    double* r_out;           // fused: [B, n, n] residual written by this iteration's update
    int fu_nparts;           // fused: partials per sample in part_u
    int pro_fused;           // the prologue ran inside the fused kernel too (mode 1): ||bt||^2 partials are in part_u (fu_nparts of
                             // them) and the first V-cycle was scaled with vs0[b]
    double* vs0;             // [B] scale estimate of the fused prologue: psn_vscale(||sin_x||^2 ||sin_y||^2 h^4) (= ||bt||^2 for theta == 1)
    int chk_mode;            // 0: 2-D tile true-residual kernel; 2: strip form (sine-factor forcing + theta on the solve grid)
    double* part_u;          // [B, upd_nstrips*upd_nbands] ||r||^2 partials of the strip update / strip true-residual kernels
    double* part_b;          // [B, ntx*nty]   ||r||^2 partials
    PsnSampleState* state;   // [B]
    double rtol;
    int max_iter;
};

template <class Ctx>
PSN_FN double psn_sum_partials(Ctx& ctx, const double* part, int count) {
    double acc = 0.0;
    for (int k = ctx.tid; k < count; k += ctx.nthreads) acc += part[k];
    return ctx.block_sum(acc);
}

#define PSN_GT_LOOP(ctx, a, ty, tx, ...)                                                      \
    {                                                                                         \
        const int col_ = (tx) * PSN_GT_COLS + (ctx.tid % PSN_GT_COLS);                         \
        const int rstep_ = ctx.nthreads >= PSN_GT_COLS ? ctx.nthreads / PSN_GT_COLS : 1;       \
        if (ctx.nthreads >= PSN_GT_COLS) {                                                     \
            for (int rr_ = ctx.tid / PSN_GT_COLS; rr_ < PSN_GT_ROWS; rr_ += rstep_) {          \
                const int i = (ty) * PSN_GT_ROWS + rr_, j = col_;                              \
                if (i < (a).n && j < (a).n) { __VA_ARGS__ }                                           \
            }                                                                                  \
        } else {                                                                               \
            for (int t_ = ctx.tid; t_ < PSN_GT_ROWS * PSN_GT_COLS; t_ += ctx.nthreads) {       \
                const int i = (ty) * PSN_GT_ROWS + t_ / PSN_GT_COLS, j = (tx) * PSN_GT_COLS + t_ % PSN_GT_COLS; \
                if (i < (a).n && j < (a).n) { __VA_ARGS__ }                                           \
            }                                                                                  \
        }                                                                                      \
    }

// r = bt, x = 0, partial ||bt||^2
template <class Ctx>
PSN_FN void psn_prologue_body(Ctx& ctx, const PsnCgArgs& a, long long b, int ty, int tx) {
    const long long base = b * (long long)a.n * a.n;
    double acc = 0.0;
    PSN_GT_LOOP(ctx, a, ty, tx, {
        const double v = psn_rhs(a.rhs, b, i, j, a.h2);
        a.r[base + (long long)i * a.n + j] = v;
        if (!a.xdefer) a.x[base + (long long)i * a.n + j] = 0.0;   // xdefer: the first flush writes x
        acc += v * v;
    })
    const double s = ctx.block_sum(acc);
    if (ctx.tid == 0) a.part_a[b * (long long)(a.ntx * a.nty) + (long long)ty * a.ntx + tx] = s;
}

// p_new = z + beta p_old;  q = At p_new;  partial <p_new, q>
template <class Ctx>
PSN_FN void psn_dir_apply_body(Ctx& ctx, const PsnCgArgs& a, long long b, int ty, int tx) {
    const PsnSampleState st = a.state[b];
    if (st.phase != PSN_PHASE_ITER) return;
    const double rz = psn_sum_partials(ctx, a.part_rz + b * (long long)a.ntiles_rz, a.ntiles_rz);
    const bool fresh = st.restart != 0;
    const double beta = fresh ? 0.0 : rz / st.rz_old;
    const int n = a.n;
    const long long base = b * (long long)n * n;
    const double* z = a.z + base;
    const double* po = a.p_old + base;
    double acc = 0.0;
    PSN_GT_LOOP(ctx, a, ty, tx, {
        const long long c = (long long)i * n + j;
        double pc, pl = 0.0, pr = 0.0, pu = 0.0, pd = 0.0;
        if (fresh) {  // restart: p = z, the old direction is never read (it may be uninitialised)
            pc = z[c];
            if (j > 0) pl = z[c - 1];
            if (j < n - 1) pr = z[c + 1];
            if (i > 0) pu = z[c - n];
            if (i < n - 1) pd = z[c + n];
        } else {
            pc = z[c] + beta * po[c];
            if (j > 0) pl = z[c - 1] + beta * po[c - 1];
            if (j < n - 1) pr = z[c + 1] + beta * po[c + 1];
            if (i > 0) pu = z[c - n] + beta * po[c - n];
            if (i < n - 1) pd = z[c + n] + beta * po[c + n];
        }
        const double q = 4.0 * pc - ((pl + pr) + (pu + pd));
        a.p_new[base + c] = pc;
        a.q[base + c] = q;
        acc += pc * q;
    })
    const double s = ctx.block_sum(acc);
    if (ctx.tid == 0) {
        a.part_pq[b * (long long)a.npq + (long long)ty * a.ntx + tx] = s;
        if (ty == 0 && tx == 0) a.state[b].rz = rz;
    }
}

// x += alpha p;  r -= alpha q;  partial ||r||^2
template <class Ctx>
PSN_FN void psn_update_body(Ctx& ctx, const PsnCgArgs& a, long long b, int ty, int tx) {
    const PsnSampleState st = a.state[b];
    if (st.phase != PSN_PHASE_ITER) return;
    const double pq = psn_sum_partials(ctx, a.part_pq + b * (long long)a.npq, a.npq);
    const double alpha = (pq > 0.0) ? st.rz / pq : 0.0;
    const int n = a.n;
    const long long base = b * (long long)n * n;
    double acc = 0.0;
    PSN_GT_LOOP(ctx, a, ty, tx, {
        const long long c = base + (long long)i * n + j;
        a.x[c] += alpha * a.p_new[c];
        const double rv = a.r[c] - alpha * a.q[c];
        a.r[c] = rv;
        acc += rv * rv;
    })
    const double s = ctx.block_sum(acc);
    if (ctx.tid == 0) {
        a.part_b[b * (long long)(a.ntx * a.nty) + (long long)ty * a.ntx + tx] = s;
        if (ty == 0 && tx == 0) a.state[b].pq = pq;
    }
}

// samples in PSN_PHASE_CHECK: r = bt - At x, partial ||r||^2
template <class Ctx>
PSN_FN void psn_true_residual_body(Ctx& ctx, const PsnCgArgs& a, long long b, int ty, int tx) {
    if (a.state[b].phase != PSN_PHASE_CHECK) return;
    const int n = a.n;
    const long long base = b * (long long)n * n;
    const double* x = a.x + base;
    double acc = 0.0;
    PSN_GT_LOOP(ctx, a, ty, tx, {
        const long long c = (long long)i * n + j;
        const double xl = (j > 0) ? x[c - 1] : 0.0;
        const double xr = (j < n - 1) ? x[c + 1] : 0.0;
        const double xu = (i > 0) ? x[c - n] : 0.0;
        const double xd = (i < n - 1) ? x[c + n] : 0.0;
        const double rv = psn_rhs(a.rhs, b, i, j, a.h2) - (4.0 * x[c] - ((xl + xr) + (xu + xd)));
        a.r[base + c] = rv;
        acc += rv * rv;
    })
    const double s = ctx.block_sum(acc);
    if (ctx.tid == 0) a.part_b[b * (long long)(a.ntx * a.nty) + (long long)ty * a.ntx + tx] = s;
}

// --- per-sample state machine (one thread per sample)
PSN_FN void psn_state_init(const PsnCgArgs& a, long long b) {
    PsnSampleState st;
    double bb = 0.0;
    const int npa = a.pro_fused ? a.fu_nparts : a.ntx * a.nty;
    const double* pa = a.pro_fused ? a.part_u + b * (long long)npa : a.part_a + b * (long long)npa;
    for (int k = 0; k < npa; ++k) bb += pa[k];
    st.bb = bb;
    st.vs = a.pro_fused ? a.vs0[b] : psn_vscale(bb);
    st.rz_old = 1.0;
    st.rz = 0.0;
    st.pq = 0.0;
    st.rr = bb;
    st.rr_prev = bb;
    st.prev_true = 0.0;
    st.iters = 0;
    st.status = PSN_ST_MAXITER;
    st.phase = PSN_PHASE_ITER;
    st.restart = 1;
    st.nrestart = 0;
    st.pad = 0;
    if (!(bb > 0.0)) {
        st.phase = PSN_PHASE_DONE;
        st.status = (bb == 0.0) ? PSN_ST_CONVERGED : PSN_ST_BREAKDOWN;
        st.rr = 0.0;
    }
    a.state[b] = st;
    if (a.xdefer) {
        PsnPend pd;
        pd.lo = pd.hi = 0;
        pd.xinit = 0;
        pd.pad = 0;
        for (int k = 0; k < PSN_PEND_K; ++k) pd.a[k] = 0.0;
        a.pend[b] = pd;
    }
}

// 4 consecutive elements (16-byte aligned address): vector accesses on the device
PSN_FN void psn_ld4(const float* p, double (&o)[4]) {
#if defined(__CUDA_ARCH__)
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
#else
    for (int e = 0; e < 4; ++e) o[e] = p[e];
#endif
}
PSN_FN void psn_ld4(const double* p, double (&o)[4]) {
#if defined(__CUDA_ARCH__)
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *reinterpret_cast<const double2*>(p + 2);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
#else
    for (int e = 0; e < 4; ++e) o[e] = p[e];
#endif
}
PSN_FN void psn_st4(double* p, const double (&v)[4]) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
#else
    for (int e = 0; e < 4; ++e) p[e] = v[e];
#endif
}

// does the flush pass act on this sample now?  (evaluated identically by the flush kernel and, afterwards, by
// psn_state_after_check, which retires the pending range)
PSN_FN bool psn_flush_due(const PsnSampleState& st, const PsnPend& pd) {
    const bool leaving = st.phase != PSN_PHASE_ITER;
    const bool full = pd.hi - pd.lo >= PSN_PEND_K;
    return (leaving || full) && (pd.hi > pd.lo || !pd.xinit);
}

// x (+)= sum over the pending directions; block `blk` of `nblk` per sample, grid-stride over the sample's nodes
#define PSN_FLUSH_BLOCKS 20
template <class Ctx>
PSN_FN void psn_flush_body(Ctx& ctx, const PsnCgArgs& a, long long b, int blk, int nblk) {
    const PsnSampleState st = a.state[b];
    const PsnPend pd = a.pend[b];
    if (!psn_flush_due(st, pd)) return;
    const long long nn = (long long)a.n * a.n;
    const float* ring = a.pring + b * nn;
    const double* av = a.pend[b].a;   // read through the cache (uniform per sample), not a register-indexed copy
    double* x = a.x + b * nn;
    // four consecutive nodes per step (n % 4 == 0 on this path): 16-byte loads of the float slots, 2 x 16 bytes of x
    const long long nq = nn >> 2;
    for (long long q = (long long)blk * ctx.nthreads + ctx.tid; q < nq; q += (long long)nblk * ctx.nthreads) {
        const long long c = q << 2;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        if (pd.xinit) psn_ld4(x + c, acc);
        for (int k = pd.lo; k < pd.hi; ++k) {
            const int slot = k % PSN_PEND_K;
            double v[4];
            psn_ld4(ring + (long long)slot * a.ring_stride + c, v);
            const double al = av[slot];
            acc[0] += al * v[0];
            acc[1] += al * v[1];
            acc[2] += al * v[2];
            acc[3] += al * v[3];
        }
        psn_st4(x + c, acc);
    }
}

// after update: bookkeeping + convergence test on the recursive residual
PSN_FN void psn_state_after_update(const PsnCgArgs& a, long long b) {
    PsnSampleState st = a.state[b];
    if (st.phase != PSN_PHASE_ITER) return;
    double rr = 0.0;
    const int cnt = a.fused ? a.fu_nparts : (a.upd_mode ? a.upd_nstrips * a.upd_nbands : a.ntx * a.nty);
    const double* pb = (a.upd_mode || a.fused) ? a.part_u + b * (long long)cnt : a.part_b + b * (long long)cnt;
    for (int k = 0; k < cnt; ++k) rr += pb[k];
    st.vs = psn_vscale(a.fused ? st.rr : rr);   // fused: the next down-leg has already run, scaled with the norm before this update
    st.rr_prev = st.rr;
    st.rr = rr;
    st.rz_old = st.rz;
    st.restart = 0;
    st.iters += 1;
    if (!(st.pq > 0.0)) {
        st.phase = PSN_PHASE_DONE;
        st.status = PSN_ST_BREAKDOWN;
    } else if (rr <= a.rtol * a.rtol * st.bb || st.iters >= a.max_iter) {
        st.phase = PSN_PHASE_CHECK;
    }
    a.state[b] = st;
}

// after the true-residual pass; returns 1 if the sample keeps iterating
PSN_FN int psn_state_after_check(const PsnCgArgs& a, long long b) {
    PsnSampleState st = a.state[b];
    if (a.xdefer) {
        PsnPend pd = a.pend[b];
        if (psn_flush_due(st, pd)) {   // the flush kernel that ran just before has put these directions into x
            a.pend[b].lo = pd.hi;
            a.pend[b].xinit = 1;
        }
    }
    if (st.phase == PSN_PHASE_CHECK) {
        double rt = 0.0;
        const int cnt = a.chk_mode ? a.upd_nstrips * a.upd_nbands : a.ntx * a.nty;
        const double* pb = a.chk_mode ? a.part_u + b * (long long)cnt : a.part_b + b * (long long)cnt;
        for (int k = 0; k < cnt; ++k) rt += pb[k];
        st.rr = rt;
        const double tol2 = a.rtol * a.rtol * st.bb;
        if (rt <= tol2) {
            st.phase = PSN_PHASE_DONE;
            st.status = PSN_ST_CONVERGED;
        } else if (st.nrestart >= 1 && rt > st.prev_true * 0.0625) {
            st.phase = PSN_PHASE_DONE;
            st.status = psn_stagnation_status(rt, st.bb, a.rtol);
        } else if (st.iters >= a.max_iter) {
            st.phase = PSN_PHASE_DONE;
            st.status = PSN_ST_MAXITER;
        } else {
            st.prev_true = rt;
            st.nrestart += 1;
            st.restart = 1;
            st.vs = psn_vscale(rt);
            st.phase = PSN_PHASE_ITER;
        }
        a.state[b] = st;
    }
    return st.phase == PSN_PHASE_ITER ? 1 : 0;
}
