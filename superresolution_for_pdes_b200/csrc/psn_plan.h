// psn_plan.h -- multigrid hierarchy ("plan") for one grid size n.  Plain C++, shared by the
// CUDA library (psn_api.cu) and by the host-emulation test harness (tests/hostemu).
//
// Reference operator (src/data_generation.py:45-58,95-102 in the reference repo):
//     diag(theta) * L * u = f,  L = 5-point Laplacian / h^2, h = 1/(n-1), all n*n nodes
//     unknown, out-of-array neighbours are implicit zeros.
// We solve the equivalent SPD system in h^2-scaled form
//     At x = bt,   (At x)_ij = 4 x_ij - x_{i-1,j} - x_{i+1,j} - x_{i,j-1} - x_{i,j+1},
//     bt = -(f/theta) * h^2,   x = u.
//
// Hierarchy: level l+1 has n_{l+1} = floor(n_l / 2) nodes per side, coarse node I sits on
// fine node 2I+1, so the left/top ghost (index -1, the true zero boundary) nests exactly.
// On the right/bottom the true boundary stays one FINE spacing behind the last node; the
// 1-D Galerkin product with linear interpolation keeps the coarse operator tridiagonal and
// only changes the last diagonal entry.  In H^2-scaled form the level operator is
//     (At_l u)_ij = d_ij u_ij - (4 neighbours),  d_ij = 4 + gam_l*[i==n_l-1] + gam_l*[j==n_l-1],
// with gam_0 = 0 and  gam_{l+1} = 2*gam_l + 1 (n_l even),  gam_{l+1} = gam_l / 2 (n_l odd).
// Transfers (scaled form): restriction r_{l+1} = 1/4 [1 2 1; 2 4 2; 1 2 1] applied to the fine
// residual centred on fine (2I+1, 2J+1); prolongation = bilinear.  The coarsest level
// (n <= coarsest_max: PSN_COARSEST_MAX = 10, at most 100 unknowns, for grids of 32 x 32 and more; 6 below) is solved
// with a precomputed dense inverse -- one small mat-vec spread over the CTA instead of two more smoothing levels of
// barrier-bound sweeps (80^2: +23 %, 40^2: +13 % solves/s).  Tiny grids keep the 6 x 6 limit: at 8 M solves/s the
// 80 KB inverse of a 10 x 10 level would be re-read from L2 at > 4 TB/s (20^2 measured 20 % slower with it).
#pragma once
#include <stdint.h>
#include <vector>
#include <cmath>

#define PSN_MAX_LEVELS 12
#define PSN_COARSEST_MAX 10

struct PsnLevel {
    int n;       // nodes per side
    int pitch;   // shared-memory pitch n + 2 (zero ghost column on both sides).  For even n the pitch is even,
                 // so a step along a grid diagonal is an ODD number of doubles: diagonal thread mappings of the
                 // red-black sweeps are bank-conflict free
    int size;    // doubles in one shared-memory level array: (n + 2)^2, rounded up to even
    int half;    // ceil(n / 2): slots per row and colour
    int off;     // offset (doubles) of this level's rhs array in a shared pool that starts at level 1
    unsigned nmagic;  // ceil(2^32 / n): exact idx / n by multiply-high (0 for n == 1)
    unsigned hmagic;  // same for `half`
    double gam;  // extra diagonal weight on the last row / last column
    double rd0;  // 1 / 4
    double rd1;  // 1 / (4 + gam)
    double rd2;  // 1 / (4 + 2 gam)
};

struct PsnPlan {
    int n;
    int nlev;
    int nu;                   // RB-GS sweeps before and after the coarse correction on level 0
    int nul[PSN_MAX_LEVELS];  // ... on level l (nul[0] == nu).  Default: 2 on level 0, 3 below -- a third sweep on the coarse
                              // levels (1/3 of the work of level 0) saves one PCG iteration in nine on every size tried
    double h2;                // h*h of the finest level, h = 1/(n-1)
    PsnLevel lv[PSN_MAX_LEVELS];
    const double* coarse_inv; // [nc*nc x nc*nc] inverse of the coarsest scaled operator (device ptr on GPU)
};

static inline int psn_level_size(int n) {
    // full zero ghost ring: (n + 2) x (n + 2), element (i, j) at (i + 1) * (n + 2) + (j + 1)
    int s = (n + 2) * (n + 2);
    return (s + 1) & ~1;
}

// Fill every field except coarse_inv.
#define PSN_COARSEST_SMALL 6
// (n = 40 runs in the register-blocked kernel, which keeps three levels in registers above the dense one: 40, 20, 10 | 5)
static inline int psn_default_coarsest(int n) { return n == 40 ? 5 : n >= 32 ? PSN_COARSEST_MAX : PSN_COARSEST_SMALL; }

// coarsest_max <= 0: the default for this n; nu_coarse <= 0: the same number of sweeps on every level
static inline void psn_plan_levels(PsnPlan* p, int n, int nu, int coarsest_max = 0, int nu_coarse = 0) {
    if (coarsest_max <= 0) coarsest_max = psn_default_coarsest(n);
    p->n = n;
    p->nu = nu;
    for (int l = 0; l < PSN_MAX_LEVELS; ++l) p->nul[l] = (l == 0 || nu_coarse <= 0) ? nu : nu_coarse;
    double h = (n > 1) ? 1.0 / (double)(n - 1) : 1.0;
    p->h2 = h * h;
    int nl = 0;
    int cur = n;
    double gam = 0.0;
    for (;;) {
        PsnLevel& L = p->lv[nl];
        L.n = cur;
        L.pitch = cur + 2;
        L.size = psn_level_size(cur);
        L.half = (cur + 1) / 2;
        L.off = 0;
        L.nmagic = cur > 1 ? (0xFFFFFFFFu / (unsigned)cur + 1u) : 0u;
        L.hmagic = L.half > 1 ? (0xFFFFFFFFu / (unsigned)L.half + 1u) : 0u;
        L.gam = gam;
        L.rd0 = 0.25;
        L.rd1 = 1.0 / (4.0 + gam);
        L.rd2 = 1.0 / (4.0 + 2.0 * gam);
        ++nl;
        if (cur <= coarsest_max || nl == PSN_MAX_LEVELS) break;
        gam = (cur % 2 == 0) ? 2.0 * gam + 1.0 : 0.5 * gam;
        cur = cur / 2;
    }
    p->nlev = nl;
    for (int l = 2; l < nl; ++l) p->lv[l].off = p->lv[l - 1].off + 2 * p->lv[l - 1].size;
    p->coarse_inv = nullptr;
}

// Dense inverse of the coarsest scaled operator (row-major N x N, N = nc*nc); symmetric.
static inline std::vector<double> psn_coarse_inverse(const PsnPlan& p) {
    const PsnLevel& L = p.lv[p.nlev - 1];
    const int nc = L.n, N = nc * nc;
    std::vector<long double> a((size_t)N * 2 * N, 0.0L);
    for (int i = 0; i < nc; ++i)
        for (int j = 0; j < nc; ++j) {
            int r = i * nc + j;
            long double d = 4.0L + (i == nc - 1 ? (long double)L.gam : 0.0L) + (j == nc - 1 ? (long double)L.gam : 0.0L);
            a[(size_t)r * 2 * N + r] = d;
            if (i > 0) a[(size_t)r * 2 * N + (r - nc)] = -1.0L;
            if (i < nc - 1) a[(size_t)r * 2 * N + (r + nc)] = -1.0L;
            if (j > 0) a[(size_t)r * 2 * N + (r - 1)] = -1.0L;
            if (j < nc - 1) a[(size_t)r * 2 * N + (r + 1)] = -1.0L;
            a[(size_t)r * 2 * N + N + r] = 1.0L;
        }
    // Gauss-Jordan with partial pivoting (N <= 100)
    for (int c = 0; c < N; ++c) {
        int piv = c;
        for (int r = c + 1; r < N; ++r)
            if (fabsl(a[(size_t)r * 2 * N + c]) > fabsl(a[(size_t)piv * 2 * N + c])) piv = r;
        if (piv != c)
            for (int k = 0; k < 2 * N; ++k) std::swap(a[(size_t)c * 2 * N + k], a[(size_t)piv * 2 * N + k]);
        long double inv = 1.0L / a[(size_t)c * 2 * N + c];
        for (int k = 0; k < 2 * N; ++k) a[(size_t)c * 2 * N + k] *= inv;
        for (int r = 0; r < N; ++r) {
            if (r == c) continue;
            long double f = a[(size_t)r * 2 * N + c];
            if (f == 0.0L) continue;
            for (int k = 0; k < 2 * N; ++k) a[(size_t)r * 2 * N + k] -= f * a[(size_t)c * 2 * N + k];
        }
    }
    std::vector<double> out((size_t)N * N);
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
            // symmetrise so that the preconditioner is exactly symmetric
            long double v = 0.5L * (a[(size_t)r * 2 * N + N + c] + a[(size_t)c * 2 * N + N + r]);
            out[(size_t)r * N + c] = (double)v;
        }
    return out;
}
