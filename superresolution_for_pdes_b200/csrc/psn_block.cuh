// psn_block.cuh -- register-blocked resident solver for n = 4*BX (the 80x80 "super-fine" grid: BX = 20).
//
// One CTA solves one sample.  Thread (bx, by) owns the 4x4 block of finest-level nodes
// rows 4by..4by+3, cols 4bx..4bx+3 and keeps r (residual), p (search direction), q = A p (fp64) and the
// V-cycle's rhs and iterate (float, see PSN_BLOCK_VF32) of that block in REGISTERS; x sits in shared memory
// in a [16][thread] layout.
// A red-black half sweep updates 8 nodes per thread with neighbours taken from registers, except the
// 8 block-edge neighbours, which come from four shared-memory edge buffers ([slot][thread] layout:
// lanes are contiguous, so every access is bank-conflict free) that the owning thread re-publishes
// after each half sweep.  Red and black edge slots alternate, so a half sweep reads only slots of the
// other colour while it writes its own: one __syncthreads per half sweep, no double buffering.
// Compared with the generic shared-memory kernel (psn_smem_solve_body) a node update drops from
// 5 LDS + 1 STS + ~12 integer instructions to 1 LDS + 1 STS amortised and no index arithmetic.
// Levels 1 and 2 (40^2, 20^2) are register-blocked the same way (2x2 and 1 node per thread); the 10^2 level is the dense inverse.
//
// Dual build like the rest of the bodies: nvcc compiles one thread's view (state arrays of extent 1);
// the host emulation (tests/hostemu) keeps the state of all threads and runs every phase as a loop
// over the virtual thread ids.
#pragma once
#include "psn_core.cuh"

#ifndef PSN_BLOCK_DEEP
#define PSN_BLOCK_DEEP 1
#endif

// The edge buffers are indexed by the thread's position in a thread grid with a ZERO GHOST RING, (BX + 2) x (BX + 2): the
// neighbour in every direction exists for every thread and reads as zero outside the grid, so no halo load needs a
// boundary predicate.  On the device the shared address of the thread's own edge cell (and of its x column) sits in one
// register each, produced by a volatile asm so that the compiler cannot re-derive it: every halo access is then one LDS / STS
// with an immediate offset.  (Before: S2R tid + ISETP boundary tests + FSEL and the S2UR / ULEA re-derivation of the shared
// window base in front of the accesses were 21 % of the kernel's instructions and 16 % of its stall samples.)
#if defined(__CUDACC__)
#define PSN_BK_TS 1
#define PSN_BK_FOR(ctx, NT) for (int t_ = (ctx).tid, once_ = 1; once_; once_ = 0)
#define PSN_BK_TI 0
#define PSN_BK_XA(S, t) ((S).xa_dev)
#else
// Host emulation: a phase is a loop over the virtual threads.  psn_bk_reverse_order() = 1 walks them backwards: a phase in which
// one thread reads what another writes (a missing barrier) then gives different bits, which tests/test_hostemu_kernels.py checks.
inline int& psn_bk_reverse_order() {
    static int rev = 0;
    return rev;
}
#define PSN_BK_TS NTPAD
#define PSN_BK_FOR(ctx, NT) \
    for (int i_ = 0, t_ = psn_bk_reverse_order() ? (NT) - 1 : 0; i_ < (NT); ++i_, t_ += psn_bk_reverse_order() ? -1 : 1)
#define PSN_BK_TI t_
#define PSN_BK_XA(S, t) ((S).xs.at(t))
#endif

// Shared-memory array of doubles addressed by a 32-bit shared-space address (device): every access is
// one ld.shared / st.shared with an immediate offset.  (With generic double* members the compiler
// re-derived the cluster-window base -- S2UR SR_CgaCtaId + ULEA -- in front of almost every access:
// 18 % of all executed instructions and a long-latency link in every dependency chain.)
#if defined(__CUDACC__)
struct PsnShF64 {
    unsigned base;
    __device__ __forceinline__ explicit PsnShF64(double* p) : base((unsigned)__cvta_generic_to_shared(p)) {}
    __device__ __forceinline__ PsnShF64 at(int off) const { PsnShF64 r(*this); r.base = base + 8u * (unsigned)off; return r; }
    __device__ __forceinline__ double ld(int i) const {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + 8u * (unsigned)i) : "memory");
        return v;
    }
    __device__ __forceinline__ void st(int i, double v) const {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + 8u * (unsigned)i), "d"(v) : "memory");
    }
    __device__ __forceinline__ void add(int i, double v) const { st(i, ld(i) + v); }
};
#else
struct PsnShF64 {
    double* base;
    explicit PsnShF64(double* p) : base(p) {}
    PsnShF64 at(int off) const { return PsnShF64(base + off); }
    double ld(int i) const { return base[i]; }
    void st(int i, double v) const { base[i] = v; }
    void add(int i, double v) const { base[i] += v; }
};
#endif

#if defined(__CUDACC__)
struct PsnShF32 {
    unsigned base;
    __device__ __forceinline__ explicit PsnShF32(double* p) : base((unsigned)__cvta_generic_to_shared(p)) {}
    __device__ __forceinline__ PsnShF32 at(int off) const { PsnShF32 r(*this); r.base = base + 4u * (unsigned)off; return r; }
    __device__ __forceinline__ float ld(int i) const {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u * (unsigned)i) : "memory");
        return v;
    }
    __device__ __forceinline__ void st(int i, float v) const {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 4u * (unsigned)i), "f"(v) : "memory");
    }
};
#else
struct PsnShF32 {
    float* base;
    explicit PsnShF32(double* p) : base(reinterpret_cast<float*>(p)) {}
    PsnShF32 at(int off) const { PsnShF32 r(*this); r.base = base + off; return r; }
    float ld(int i) const { return base[i]; }
    void st(int i, float v) const { base[i] = v; }
};
#endif

// V-cycle inside the register-blocked solver in fp32 (1, default) or fp64 (0): like the float-stored preconditioner of the
// streamed path, everything between "r in" and "z out" -- iterates, restricted residuals, edge exchange -- is float, scaled by
// a power of two near 1 / ||r||; r, p, q, x, every inner product and the stopping rule stay fp64.
#ifndef PSN_BLOCK_VF32
#define PSN_BLOCK_VF32 1
#endif

template <int BX>
struct PsnBlockGeom {
    static constexpr int N = 4 * BX;
    static constexpr int NT = BX * BX;
    static constexpr int NTPAD = (NT + 31) & ~31;
    // Thread grid with a ghost ring: BX + 2 rows of pitch BP cells.  Threads are numbered tile by tile -- 16 consecutive
    // threads (one half warp = one 64-bit shared-memory transaction) form a 4 x 4 patch of the thread grid -- and the pitch is
    // 12 mod 16, so that the four rows of a patch fall on four disjoint groups of banks: every edge access is conflict free.
    // (Row-major numbering with pitch BX + 2: 3 of 5 half warps straddle a row end, 1.6 wavefronts per access.)
    // Threads are numbered row by row (always for BX = 10 / 5, which the 4 x 4 tiles do not divide).  With ghost COLUMNS
    // (pitch BX + 2) 32 consecutive threads spread over 32 + 3 * 2 positions and every edge access takes two wavefronts --
    // the 40 x 40 kernel ran at 76 % of the shared-memory pipe, 41 % of it conflict wavefronts.  So these grids have ghost rows
    // only, pitch BX: thread t sits at cell t + BX, consecutive threads on consecutive cells, no conflict at any alignment; the
    // left / right neighbour of a thread in the first / last column is masked with the thread's flags instead (EL below), and
    // the idle threads of the last warp sit 32 cells behind their row-major position (same banks, behind the bottom ghost row).
// (PSN_BLOCK_TILED = 1: the layout of the first half of round 2 -- ghost columns, threads numbered in 4 x 4 tiles, pitch 12 mod
// 16 -- which left 13 % conflict wavefronts on the float planes, in the warps whose two tiles straddle a tile row.  The
// ghost-row-only layout below is conflict free for every warp: 80 x 80 1.553M -> 1.620M solves/s.)
#ifndef PSN_BLOCK_TILED
#define PSN_BLOCK_TILED 0
#endif
    static constexpr bool TILED = BX % 4 == 0 && PSN_BLOCK_TILED;
    static constexpr bool GHOSTCOLS = TILED;
    static constexpr int BROWS = BX + 2;
    static constexpr int BP = TILED ? ((BX + 2 + 3) / 16) * 16 + 12 : BX;
    static_assert(!TILED || (BP >= BX + 2 && BP % 16 == 12), "tile numbering: pitch 12 mod 16");
    // cells per edge-buffer plane: the padded grid, then one private cell per idle thread of the last warp (their accesses
    // need no predicate either), then slack for those threads' neighbour reads
    static constexpr int NTP = (BROWS * BP + (TILED ? 0 : 32) + (NTPAD - NT) + BP + 1 + 7) & ~7;
    // float planes of the V-cycle exchange (they alias the fp64 planes): a warp is two tiles = 4 rows x 8 cells per 32-bit
    // transaction, so the pitch is 8 mod 32.  (Two tiles are disjoint in banks iff their tile columns differ by an odd number;
    // the 2 of 13 warps that hold the last tile of a tile row and the first of the next conflict 2-way: with 25 tiles no pitch
    // pairs them all.)
    static constexpr int BPF = TILED ? ((BX + 2 - 8 + 31) / 32) * 32 + 8 : BX;
    static constexpr int NTPF = (BROWS * BPF + (TILED ? 0 : 32) + (NTPAD - NT) + BPF + 1 + 7) & ~7;
    static_assert((!TILED || (BPF >= BX + 2 && BPF % 32 == 8)) && 16 * NTPF * 4 <= 16 * NTP * 8, "float planes must fit into the fp64 planes");
    // shared memory (doubles): x[16][NTPAD], edges[4 sides][4 slots][NTP]
    static constexpr int XS = 16 * NTPAD;
    static constexpr int ES = 16 * NTP;
};

// doubles of dynamic shared memory the block kernel needs (x, edge buffers, coarse levels >= 1)
// Levels 0 and 1 always live in registers; level 2 too when there is a level below it (n, n/2, n/4 | n/8 ...), otherwise
// level 2 is the dense one (n = 20: 20, 10 | 5) and its right-hand side / solution pass through shared memory.
// (L0: the level of the plan that the block code treats as its level 0 -- 0 for the stand-alone solver, the first coarse
// level handed over by the streamed solver for the V-cycle-only kernel)
PSN_HDFN int psn_block_pool_doubles(const PsnPlan& plan, int L0 = 0) {
    return psn_pool_doubles(plan, L0 + (plan.nlev >= L0 + 4 ? 2 : 1));
}
template <int BX>
PSN_HDFN int psn_block_smem_doubles(const PsnPlan& plan) {
    return PsnBlockGeom<BX>::XS + PsnBlockGeom<BX>::ES + psn_block_pool_doubles(plan);
}
// the register-blocked code needs the hierarchy n, n/2, n/4, ... with at least 3 levels from level L0 on
template <int BX>
PSN_HDFN bool psn_block_ok_at(const PsnPlan& plan, int L0) {
    return L0 >= 0 && plan.nlev >= L0 + 3 && plan.lv[L0].n == 4 * BX && plan.lv[L0 + 1].n == 2 * BX && plan.lv[L0 + 2].n == BX;
}
template <int BX>
PSN_HDFN bool psn_block_ok(const PsnPlan& plan) {
    return plan.n == 4 * BX && psn_block_ok_at<BX>(plan, 0);
}
// shared memory of the V-cycle-only kernel: edge planes and the levels below the register levels (no x)
template <int BX>
PSN_HDFN int psn_block_vcycle_smem_doubles(const PsnPlan& plan, int L0) {
    return PsnBlockGeom<BX>::ES + psn_block_pool_doubles(plan, L0);
}
// doubles of the dense coarsest-level inverse, kept behind the arrays above by the persistent kernel when it fits
PSN_HDFN int psn_coarse_inv_doubles(const PsnPlan& plan) {
    const int N = plan.lv[plan.nlev - 1].n * plan.lv[plan.nlev - 1].n;
    return N * N;
}

template <int BX, class Ctx>
struct PsnBlockSolver {
    typedef PsnBlockGeom<BX> G;
    static constexpr int N = G::N, NT = G::NT, NTPAD = G::NTPAD;
    static constexpr int TILES = G::TILED ? BX / 4 : 1;   // 4 x 4 patches of threads per row of the thread grid
    // thread t -> its block (bx, by) of the thread grid: tile t / 16 (row-major over TILES x TILES), lane t % 16 inside it;
    // row by row when BX is not a multiple of 4
    PSN_FN static int bx_of(int t) { return G::TILED ? 4 * ((t >> 4) % TILES) + (t & 3) : t % BX; }
    PSN_FN static int by_of(int t) { return G::TILED ? 4 * ((t >> 4) / TILES) + ((t >> 2) & 3) : t / BX; }
    enum { TOP = 0, BOT = 1, LEFT = 2, RIGHT = 3 };

#if PSN_BLOCK_VF32
    typedef float VT;                      // element type inside the V-cycle
    typedef PsnShF32 ShV;
    static constexpr int BPV = G::BPF, NTPV = G::NTPF;
#else
    typedef double VT;
    typedef PsnShF64 ShV;
    static constexpr int BPV = G::BP, NTPV = G::NTP;
#endif
    // cursor = shared address of the thread's own edge cell, with the geometry of its planes
    struct CurD {                          // fp64 planes: CG exchange of p and x
        PsnShF64 a;
        int fl;                            // the thread's flags (first / last column masks of the layouts without ghost columns)
        typedef double T;
        static constexpr int P = G::BP, PL = G::NTP;
    };
    struct CurV {                          // V-cycle planes
        ShV a;
        int fl;
        typedef VT T;
        static constexpr int P = BPV, PL = NTPV;
    };

    Ctx& ctx;
    PsnShF64 xs;     // [16][NTPAD]
    PsnShF64 edge;   // [4][4][NTP] doubles / [4][4][NTPV] VT, cell index = position in the ghost-ringed thread grid
    ShV edgev;
    double r[PSN_BK_TS][16], p[PSN_BK_TS][16];   // CG residual and search direction
    double q[PSN_BK_TS][16];                     // A p; block of x during the true-residual check
    VT z[PSN_BK_TS][16], bv[PSN_BK_TS][16];      // V-cycle: iterate and (scaled) right-hand side on level 0
#if defined(__CUDACC__)
    PsnShF64 ea_dev, xa_dev;   // this thread's own edge cell / x column (shared addresses pinned in registers)
    ShV ev_dev;
    int fl_dev;                // bit 0: real thread (t < NT), bit 1: (bx + by) odd, bit 2: bx and by both odd,
                               // bits 3 / 4: last thread row / column, bit 5: first thread column
#endif

    const double* inv_s = nullptr;   // shared-memory copy of the coarsest-level inverse (or null)
    int L0 = 0;                      // level of the plan that is this solver's level 0
    // GAM0 (V-cycle-only kernel: level 0 is a coarse level of a larger hierarchy, whose last row / column carry the extra
    // diagonal weight gam): reciprocal diagonals 1/4, 1/(4 + gam), 1/(4 + 2 gam) and gam itself
    VT R0[3] = {(VT)0.25, (VT)0.25, (VT)0.25};
    VT gam0 = (VT)0;
    template <bool GAM0>
    PSN_FN VT rdiag0(int a, int cc, int fl) const {
        if (!GAM0 || (a < 3 && cc < 3)) return (VT)0.25;
        const bool lr = a == 3 && (fl & 8), lc = cc == 3 && (fl & 16);
        return lr ? (lc ? R0[2] : R0[1]) : (lc ? R0[1] : R0[0]);
    }
    template <bool GAM0>
    PSN_FN VT diag0(int a, int cc, int fl) const {
        if (!GAM0 || (a < 3 && cc < 3)) return (VT)4;
        return (VT)4 + ((a == 3 && (fl & 8)) ? gam0 : (VT)0) + ((cc == 3 && (fl & 16)) ? gam0 : (VT)0);
    }

    // position of thread t in the ghost-ringed thread grid of pitch P; the idle threads of the last warp get private cells behind it
    PSN_FN static int tpad(int t, int P) {
        if (!G::GHOSTCOLS) return t < NT ? t + BX : t + BX + 32;   // (P == BX: cell (by + 1) * BX + bx)
        return t < NT ? (by_of(t) + 1) * P + bx_of(t) + 1 : G::BROWS * P + (t - NT);
    }
    PSN_FN static int flags_of(int t) {
        const int bx = bx_of(t), by = by_of(t);
        return (t < NT ? 1 : 0) | ((((bx + by) & 1) && t < NT) ? 2 : 0) | (((bx & 1) && (by & 1) && t < NT) ? 4 : 0) |
               ((by == BX - 1 && t < NT) ? 8 : 0) | ((bx == BX - 1 && t < NT) ? 16 : 0) | ((bx == 0 && t < NT) ? 32 : 0);
    }
    PSN_FN int flags(int t) const {
#if defined(__CUDACC__)
        return fl_dev;
#else
        return flags_of(t);
#endif
    }
    PSN_FN CurD curd(int t) const {
#if defined(__CUDACC__)
        return CurD{ea_dev, fl_dev};
#else
        return CurD{edge.at(tpad(t, G::BP)), flags_of(t)};
#endif
    }
    PSN_FN CurV curv(int t) const {
#if defined(__CUDACC__)
        return CurV{ev_dev, fl_dev};
#else
        return CurV{edgev.at(tpad(t, BPV)), flags_of(t)};
#endif
    }

    PSN_FN PsnBlockSolver(Ctx& c, double* smem)
        : ctx(c), xs(smem), edge(smem + G::XS), edgev(smem + G::XS)
#if defined(__CUDACC__)
          , ea_dev(smem), xa_dev(smem), ev_dev(smem)
#endif
    {
#if defined(__CUDACC__)
        unsigned e = edge.base + 8u * (unsigned)tpad(c.tid, G::BP), x = xs.base + 8u * (unsigned)c.tid;
        unsigned v = edgev.base + (unsigned)sizeof(VT) * (unsigned)tpad(c.tid, BPV);
        int f = flags_of(c.tid);
        asm volatile("mov.u32 %0, %0;" : "+r"(e));   // opaque to the compiler: kept in a register (or spilled), never re-derived
        asm volatile("mov.u32 %0, %0;" : "+r"(x));
        asm volatile("mov.u32 %0, %0;" : "+r"(v));
        asm volatile("mov.s32 %0, %0;" : "+r"(f));
        ea_dev.base = e;
        xa_dev.base = x;
        ev_dev.base = v;
        fl_dev = f;
#endif
    }

    // Planes.  Level 0 (and the fp64 exchange of the CG part) uses plane 4 side + slot.  With the float V-cycle there is room
    // for 22 float planes in the space of the 16 fp64 ones, and levels 1 and 2 get planes that the neighbouring level does not
    // touch at the moment of a transfer, so that "read the other level's edges" and "publish my own" share one barrier interval:
    //   level 1 (slots 0, 1): TOP -> 16 + slot, LEFT -> 18 + slot, BOT / RIGHT -> slots 0 and 2 of the level-0 side.  Going down,
    //     the restriction reads the level-0 planes TOP / LEFT 0, 2 while the first level-1 half sweep writes 16, 18, BOT 2, RIGHT 2;
    //     going up, the level-0 correction reads BOT / RIGHT 0, 2 while the red edges of z go to TOP / LEFT 0, 2 and BOT / RIGHT 1, 3.
    //   level 2: iterate -> 20, residual -> 21.
    static constexpr bool SPLIT = PSN_BLOCK_VF32 != 0;
    static constexpr int NPLV = SPLIT ? 22 : 16;
    static_assert((long long)NPLV * NTPV * (long long)sizeof(VT) <= 16ll * G::NTP * 8, "V-cycle planes must fit into the fp64 planes");
    template <int LV>
    PSN_FN static constexpr int plane(int side, int slot) {
        return (!SPLIT || LV == 0) ? side * 4 + slot
               : (LV == 1)         ? (side == TOP ? 16 + slot : side == LEFT ? 18 + slot : side * 4 + 2 * slot)
                                   : (side == TOP ? 20 : 21);
    }
    // edge buffer (side, slot) of level LV, cell (dy, dx) relative to the thread's own cell
    template <int LV = 0, class C>
    PSN_FN static typename C::T EL(const C& c, int side, int slot, int dy, int dx) {
        typename C::T v = c.a.ld(plane<LV>(side, slot) * C::PL + dy * C::P + dx);
        if (!G::GHOSTCOLS) {   // no ghost columns: the cell left of the first / right of the last column belongs to another row
            if (dx < 0 && (c.fl & 32)) v = (typename C::T)0;
            if (dx > 0 && (c.fl & 16)) v = (typename C::T)0;
        }
        return v;
    }
    template <int LV = 0, class C>
    PSN_FN static void ES(const C& c, int side, int slot, typename C::T v) { c.a.st(plane<LV>(side, slot) * C::PL, v); }

    // halo values of field v needed by a half sweep of colour `col` (slots of the other colour); ghost cells read as zero
    template <class C, class T>
    PSN_FN static void load_halo(const C& c, int col, T (&hT)[4], T (&hB)[4], T (&hL)[4], T (&hR)[4]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hT[k] = hB[k] = hL[k] = hR[k] = (T)0;
            if ((k & 1) == col) {           // (0,k) and (k,0) have colour col: they need the node above / left
                hT[k] = EL(c, BOT, k, -1, 0);
                hL[k] = EL(c, RIGHT, k, 0, -1);
            } else {                        // (3,k) and (k,3) have colour col: they need the node below / right
                hB[k] = EL(c, TOP, k, 1, 0);
                hR[k] = EL(c, LEFT, k, 0, 1);
            }
        }
    }
    // all 16 halo values (both colours)
    template <class C, class T>
    PSN_FN static void load_halo_all(const C& c, T (&hT)[4], T (&hB)[4], T (&hL)[4], T (&hR)[4]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hT[k] = EL(c, BOT, k, -1, 0);
            hB[k] = EL(c, TOP, k, 1, 0);
            hL[k] = EL(c, RIGHT, k, 0, -1);
            hR[k] = EL(c, LEFT, k, 0, 1);
        }
    }
    // publish the edge values of colour `col` (col < 0: both colours) of field v
    template <class C, class T>
    PSN_FN static void publish(const C& c, const T (&v)[16], int col) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (col < 0 || (k & 1) == col) {        // (0,k), (k,0) have colour k&1
                ES(c, TOP, k, v[k]);
                ES(c, LEFT, k, v[4 * k]);
            }
            if (col < 0 || (k & 1) != col) {        // (3,k), (k,3) have colour (k+1)&1
                ES(c, BOT, k, v[12 + k]);
                ES(c, RIGHT, k, v[4 * k + 3]);
            }
        }
    }
    // The float planes of the V-cycle alias the fp64 planes of the CG exchange, so each side overwrites the other's ghost ring:
    // thread t re-zeroes its share of the ring cells of all planes of layout C (called once per PCG iteration and layout,
    // in a phase where nobody reads the planes; the next barrier orders it before the first halo load).
    static constexpr int RING = G::GHOSTCOLS ? 2 * (BX + 2) + 2 * BX : 2 * BX;   // (ghost rows only: top and bottom row)
    // Two schedules.  One warp per sample (BX = 5: 10 ring cells -- the two ghost rows -- for 32 threads): thread t owns ring cell t in every plane, so the
    // cell's offset is worked out once per call and the stores need no index arithmetic -- with a division and a branch per
    // cell the re-zeroing was 15 % of that kernel's instructions (23.0M -> 25.5M solves/s).  Larger CTAs: the cells of all
    // planes are dealt round-robin to all threads (3 - 4 each); the per-cell-owner form leaves most warps idle while a few
    // store, and measured 1 % (40 x 40) and 5 % (80 x 80) slower.
    static constexpr bool ZOWNER = NTPAD <= 32;
    static_assert(NTPAD >= RING, "every ring cell needs a thread");
    PSN_FN static int ring_cell(int c, int P) {
        if (!G::GHOSTCOLS) return c < BX ? c : (BX + 1) * P + (c - BX);
        if (c < BX + 2) return c;
        if (c < 2 * (BX + 2)) return (BX + 1) * P + (c - (BX + 2));
        return (1 + ((c - 2 * (BX + 2)) >> 1)) * P + (((c - 2 * (BX + 2)) & 1) ? BX + 1 : 0);
    }
    template <class C, class Sh>
    PSN_FN static void zero_ghosts(const Sh& planes, int t, int nplanes) {
        if (ZOWNER) {
            if (t >= RING) return;
            const Sh own = planes.at(ring_cell(t, C::P));
            for (int pl = 0; pl < nplanes; ++pl) own.st(pl * C::PL, (typename C::T)0);
        } else {
            for (int g = t; g < nplanes * RING; g += NTPAD) {
                const int pl = g / RING, c = g - pl * RING;
                planes.st(pl * C::PL + ring_cell(c, C::P), (typename C::T)0);
            }
        }
    }
    // sum of the four neighbours of local node (a, c)
    template <class T>
    PSN_FN static T nbsum(const T (&v)[16], int a, int c, const T (&hT)[4], const T (&hB)[4], const T (&hL)[4], const T (&hR)[4]) {
        const T up = (a > 0) ? v[4 * (a - 1) + c] : hT[c];
        const T dn = (a < 3) ? v[4 * (a + 1) + c] : hB[c];
        const T lf = (c > 0) ? v[4 * a + c - 1] : hL[a];
        const T rt = (c < 3) ? v[4 * a + c + 1] : hR[a];
        return (up + dn) + (lf + rt);
    }

    // one RB-GS half sweep on z with rhs bv; ends with the edges of colour `col` published and a sync.  Every thread of the CTA
    // takes part: the idle threads of the last warp work on zeros in their private cells.
    // (last: the final half sweep of the V-cycle -- nobody reads z's edges after it, and the next phase has a barrier of its own)
    template <bool GAM0 = false>
    PSN_FN void half_sweep(int col, bool first, bool last = false) {
        PSN_BK_FOR(ctx, NTPAD) {
            const CurV c = curv(t_);
            const int fl = GAM0 ? flags(t_) : 0;
            VT (&zz)[16] = z[PSN_BK_TI];
            const VT (&rr)[16] = bv[PSN_BK_TI];
            VT hT[4], hB[4], hL[4], hR[4];
            if (!first) load_halo(c, col, hT, hB, hL, hR);
            VT nv[16];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    if (((a + cc) & 1) == col)
                        nv[4 * a + cc] = (rr[4 * a + cc] + (first ? (VT)0 : nbsum(zz, a, cc, hT, hB, hL, hR))) * rdiag0<GAM0>(a, cc, fl);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if ((((k >> 2) + (k & 3)) & 1) == col) zz[k] = nv[k];
            if (!last) publish(c, zz, col);
        }
        if (!last) ctx.sync();
    }

    VT rho[PSN_BK_TS][16];

    // ------------------------------------------------------------------ three register-blocked levels
    // Level 1 (2 BX)^2: each thread owns the 2x2 block under its fine block; level 2 (BX)^2: one node per
    // thread.  Same edge-buffer exchange (slots k < block size); levels >= 3 stay in the shared-memory code.
    VT z1[PSN_BK_TS][4], r1[PSN_BK_TS][4], z2[PSN_BK_TS][1], r2[PSN_BK_TS][1];

    // (the BS-blocked helpers serve level 1: BS = 2)
    template <int BS>
    PSN_FN static void load_halo_bs(const CurV& c, int col, VT (&hT)[BS], VT (&hB)[BS], VT (&hL)[BS], VT (&hR)[BS]) {
#pragma unroll
        for (int k = 0; k < BS; ++k) {
            hT[k] = hB[k] = hL[k] = hR[k] = (VT)0;
            if (col < 0 || (k & 1) == col) {
                hT[k] = EL<1>(c, BOT, k, -1, 0);
                hL[k] = EL<1>(c, RIGHT, k, 0, -1);
            }
            if (col < 0 || (k & 1) != col) {
                hB[k] = EL<1>(c, TOP, k, 1, 0);
                hR[k] = EL<1>(c, LEFT, k, 0, 1);
            }
        }
    }
    template <int BS>
    PSN_FN static void publish_bs(const CurV& c, const VT (&v)[BS * BS], int col) {
#pragma unroll
        for (int k = 0; k < BS; ++k) {
            if (col < 0 || (k & 1) == col) {
                ES<1>(c, TOP, k, v[k]);
                ES<1>(c, LEFT, k, v[BS * k]);
            }
            if (col < 0 || (k & 1) != col) {
                ES<1>(c, BOT, k, v[BS * (BS - 1) + k]);
                ES<1>(c, RIGHT, k, v[BS * k + BS - 1]);
            }
        }
    }
    template <int BS>
    PSN_FN static VT nbsum_bs(const VT (&v)[BS * BS], int a, int c, const VT (&hT)[BS], const VT (&hB)[BS],
                              const VT (&hL)[BS], const VT (&hR)[BS]) {
        const VT up = (a > 0) ? v[BS * (a - 1) + c] : hT[c];
        const VT dn = (a < BS - 1) ? v[BS * (a + 1) + c] : hB[c];
        const VT lf = (c > 0) ? v[BS * a + c - 1] : hL[a];
        const VT rt = (c < BS - 1) ? v[BS * a + c + 1] : hR[a];
        return (up + dn) + (lf + rt);
    }
    // Diagonal weights of the thread's block on a BS-blocked level: only the threads of the last thread row / column differ
    // from 4 (extra weight gam on the last grid row / column).  Kept as two per-thread values (last row, last column).
    struct Diag {
        VT gr, gc;   // gam if the thread's block touches the last grid row / column, else 0
    };
    PSN_FN static Diag diag_of(const PsnLevel& lv, int t) {
        const int bx = bx_of(t), by = by_of(t);
        Diag d;
        d.gr = (t < NT && by == BX - 1) ? (VT)lv.gam : (VT)0;
        d.gc = (t < NT && bx == BX - 1) ? (VT)lv.gam : (VT)0;
        return d;
    }
    template <int BS>
    PSN_FN static VT diag_bs(const Diag& d, int a, int c) {
        return (VT)4 + ((a == BS - 1) ? d.gr : (VT)0) + ((c == BS - 1) ? d.gc : (VT)0);
    }
    // half sweep of a BS-blocked level held in zz / rr; publishes the updated colour.  rdv: reciprocal diagonals of the block
    template <int BS>
    PSN_FN static void half_sweep_bs(const CurV& c, const VT (&rdv)[BS * BS], VT (&zz)[BS * BS], const VT (&rr)[BS * BS], int col, bool first) {
        VT hT[BS], hB[BS], hL[BS], hR[BS];
        if (!first) load_halo_bs<BS>(c, col, hT, hB, hL, hR);
        VT nv[BS * BS];
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int cc = 0; cc < BS; ++cc)
                if (((a + cc) & 1) == col)
                    nv[BS * a + cc] = (rr[BS * a + cc] + (first ? (VT)0 : nbsum_bs<BS>(zz, a, cc, hT, hB, hL, hR))) * rdv[BS * a + cc];
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int cc = 0; cc < BS; ++cc)
                if (((a + cc) & 1) == col) zz[BS * a + cc] = nv[BS * a + cc];
        publish_bs<BS>(c, zz, col);
    }
    // level 2: one node per thread, colour (bx + by) & 1; all four sides read the single slot TOP[0]
    PSN_FN static VT nb2(const CurV& c) {
        return (EL<2>(c, TOP, 0, -1, 0) + EL<2>(c, TOP, 0, 1, 0)) + (EL<2>(c, TOP, 0, 0, -1) + EL<2>(c, TOP, 0, 0, 1));
    }
    PSN_FN static void half_sweep_2_body(const CurV& c, int fl, VT rd, VT& zz, VT rr, int col, bool first) {
        if ((fl & 1) && ((fl >> 1) & 1) == col) {
            const VT v = (rr + (first ? (VT)0 : nb2(c))) * rd;
            zz = v;
            ES<2>(c, TOP, 0, v);
        }
    }
    PSN_FN void half_sweep_2(const VT (&rd2)[PSN_BK_TS], int col, bool first) {
        PSN_BK_FOR(ctx, NTPAD) { half_sweep_2_body(curv(t_), flags(t_), rd2[PSN_BK_TI], z2[PSN_BK_TI][0], r2[PSN_BK_TI][0], col, first); }
        ctx.sync();
    }

    // z = M^{-1} bv with levels 0, 1, 2 in registers (n = 4 BX, 2 BX, BX) and levels >= 3 in shared memory (fp64)
    template <bool GAM0>
    PSN_FN void vcycle3(const PsnPlan& plan, double* pool) {
        const int nu = plan.nul[L0], nu1 = plan.nul[L0 + 1], nu2 = plan.nul[L0 + 2];
        const PsnLevel& l1 = plan.lv[L0 + 1];
        const PsnLevel& l2 = plan.lv[L0 + 2];
        const bool dense2 = plan.nlev == L0 + 3;              // level 2 is the dense level (n = 20); CTA-uniform
        const PsnLevel& l3 = plan.lv[L0 + (dense2 ? 2 : 3)];  // first level kept in shared memory
        double* b3 = pool;
        double* u3 = pool + l3.size;
        double* pool4 = pool + 2 * l3.size;
        // per-thread diagonals of levels 1 and 2 (4 is the interior value; the last thread row / column carries gam)
        VT rd1[PSN_BK_TS][4], rd2[PSN_BK_TS];
        Diag d1[PSN_BK_TS], d2[PSN_BK_TS];
        PSN_BK_FOR(ctx, NTPAD) {
            d1[PSN_BK_TI] = diag_of(l1, t_);
            d2[PSN_BK_TI] = diag_of(l2, t_);
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const bool lr = a == 1 && d1[PSN_BK_TI].gr != (VT)0, lc = c == 1 && d1[PSN_BK_TI].gc != (VT)0;
                    rd1[PSN_BK_TI][2 * a + c] = (VT)(lr ? (lc ? l1.rd2 : l1.rd1) : (lc ? l1.rd1 : l1.rd0));
                }
            {
                const bool lr = d2[PSN_BK_TI].gr != (VT)0, lc = d2[PSN_BK_TI].gc != (VT)0;
                rd2[PSN_BK_TI] = (VT)(lr ? (lc ? l2.rd2 : l2.rd1) : (lc ? l2.rd1 : l2.rd0));
            }
        }
        // ---------------- level 0 down
        half_sweep<GAM0>(0, true);
        half_sweep<GAM0>(1, false);
        for (int s = 1; s < nu; ++s) {
            half_sweep<GAM0>(0, false);
            half_sweep<GAM0>(1, false);
        }
        PSN_BK_FOR(ctx, NTPAD) {
            const CurV c = curv(t_);
            const VT (&zz)[16] = z[PSN_BK_TI];
            const VT (&rr)[16] = bv[PSN_BK_TI];
            VT (&pp)[16] = rho[PSN_BK_TI];
            VT hT[4], hB[4], hL[4], hR[4];
            load_halo(c, 0, hT, hB, hL, hR);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    if (((a + cc) & 1) == 0)
                        pp[4 * a + cc] = rr[4 * a + cc] - (diag0<GAM0>(a, cc, GAM0 ? flags(t_) : 0) * zz[4 * a + cc] - nbsum(zz, a, cc, hT, hB, hL, hR));
            // (reads above: black slots; writes below: red slots of TOP / LEFT -> no barrier needed in between)
            ES(c, TOP, 0, pp[0]);
            ES(c, TOP, 2, pp[2]);
            ES(c, LEFT, 0, pp[0]);
            ES(c, LEFT, 2, pp[8]);
        }
        ctx.sync();
        PSN_BK_FOR(ctx, NTPAD) {
            const CurV c = curv(t_);
            const VT (&pp)[16] = rho[PSN_BK_TI];
            const VT r04 = EL(c, LEFT, 0, 0, 1), r24 = EL(c, LEFT, 2, 0, 1);
            const VT r40 = EL(c, TOP, 0, 1, 0), r42 = EL(c, TOP, 2, 1, 0);
            const VT r44 = EL(c, TOP, 0, 1, 1);
            VT (&qq)[4] = r1[PSN_BK_TI];
            qq[0] = pp[5] + (VT)0.25 * ((pp[0] + pp[2]) + (pp[8] + pp[10]));
            qq[1] = pp[7] + (VT)0.25 * ((pp[2] + r04) + (pp[10] + r24));
            qq[2] = pp[13] + (VT)0.25 * ((pp[8] + pp[10]) + (r40 + r42));
            qq[3] = pp[15] + (VT)0.25 * ((pp[10] + r24) + (r42 + r44));
            // own planes: the first half sweep of level 1 (no halo reads) writes none of the planes read above
            if (SPLIT) half_sweep_bs<2>(c, rd1[PSN_BK_TI], z1[PSN_BK_TI], r1[PSN_BK_TI], 0, true);
        }
        ctx.sync();
        // ---------------- level 1 down (2x2 per thread)
        for (int s = 0; s < nu1; ++s) {
            if (!(SPLIT && s == 0)) {
                PSN_BK_FOR(ctx, NTPAD) { half_sweep_bs<2>(curv(t_), rd1[PSN_BK_TI], z1[PSN_BK_TI], r1[PSN_BK_TI], 0, s == 0); }
                ctx.sync();
            }
            PSN_BK_FOR(ctx, NTPAD) { half_sweep_bs<2>(curv(t_), rd1[PSN_BK_TI], z1[PSN_BK_TI], r1[PSN_BK_TI], 1, false); }
            ctx.sync();
        }
        PSN_BK_FOR(ctx, NTPAD) {
            const CurV c = curv(t_);
            const VT (&zz)[4] = z1[PSN_BK_TI];
            const VT (&rr)[4] = r1[PSN_BK_TI];
            VT hT[2], hB[2], hL[2], hR[2];
            load_halo_bs<2>(c, 0, hT, hB, hL, hR);
            const VT rho00 = rr[0] - (diag_bs<2>(d1[PSN_BK_TI], 0, 0) * zz[0] - nbsum_bs<2>(zz, 0, 0, hT, hB, hL, hR));
            const VT rho11 = rr[3] - (diag_bs<2>(d1[PSN_BK_TI], 1, 1) * zz[3] - nbsum_bs<2>(zz, 1, 1, hT, hB, hL, hR));
            rho[PSN_BK_TI][0] = rho00;
            rho[PSN_BK_TI][1] = rho11;
            ES<1>(c, TOP, 0, rho00);   // red slot of TOP: not read by this phase
        }
        ctx.sync();
        PSN_BK_FOR(ctx, NTPAD) {
            const CurV c = curv(t_);
            const VT c02 = EL<1>(c, TOP, 0, 0, 1), c20 = EL<1>(c, TOP, 0, 1, 0);
            const VT c22 = EL<1>(c, TOP, 0, 1, 1);
            r2[PSN_BK_TI][0] = rho[PSN_BK_TI][1] + (VT)0.25 * ((rho[PSN_BK_TI][0] + c02) + (c20 + c22));
            if (dense2) {
                if (flags(t_) & 1) b3[(by_of(t_) + 1) * l3.pitch + bx_of(t_) + 1] = (double)r2[PSN_BK_TI][0];
            } else if (SPLIT) {
                half_sweep_2_body(c, flags(t_), rd2[PSN_BK_TI], z2[PSN_BK_TI][0], r2[PSN_BK_TI][0], 0, true);
            }
        }
        ctx.sync();
        if (dense2) {
            // ---------------- level 2 is the dense level: solve it, every thread fetches its node and publishes it
            psn_vcycle_smem(ctx, plan, L0 + 2, b3, u3, pool4, inv_s);
            PSN_BK_FOR(ctx, NTPAD) {
                if (flags(t_) & 1) {
                    const VT v = (VT)u3[(by_of(t_) + 1) * l3.pitch + bx_of(t_) + 1];
                    z2[PSN_BK_TI][0] = v;
                    ES<2>(curv(t_), TOP, 0, v);
                }
            }
            ctx.sync();
        } else {
        // ---------------- level 2 down (one node per thread)
        for (int s = 0; s < nu2; ++s) {
            if (!(SPLIT && s == 0)) half_sweep_2(rd2, 0, s == 0);
            half_sweep_2(rd2, 1, false);
        }
        PSN_BK_FOR(ctx, NTPAD) {   // residual on red nodes -> second plane of level 2 (the first is being read by the neighbours)
            const CurV c = curv(t_);
            const int fl = flags(t_);
            VT rh = (VT)0;
            if ((fl & 3) == 1) rh = r2[PSN_BK_TI][0] - (diag_bs<1>(d2[PSN_BK_TI], 0, 0) * z2[PSN_BK_TI][0] - nb2(c));
            rho[PSN_BK_TI][0] = rh;
            ES<2>(c, BOT, 0, rh);
        }
        ctx.sync();
        PSN_BK_FOR(ctx, NT) {
            if (flags(t_) & 4) {   // bx, by both odd: centre of coarse node ((by-1)/2, (bx-1)/2) of level 3
                const CurV c = curv(t_);
                const int bx = bx_of(t_), by = by_of(t_);
                const VT c00 = EL<2>(c, BOT, 0, -1, -1), c02 = EL<2>(c, BOT, 0, -1, 1);
                const VT c20 = EL<2>(c, BOT, 0, 1, -1), c22 = EL<2>(c, BOT, 0, 1, 1);
                b3[(((by - 1) >> 1) + 1) * l3.pitch + ((bx - 1) >> 1) + 1] = (double)(rho[PSN_BK_TI][0] + (VT)0.25 * ((c00 + c02) + (c20 + c22)));
            }
        }
        ctx.sync();
        // ---------------- levels >= 3 in shared memory, fp64 (all warps; running them on a single warp with
        // __syncwarp instead of CTA barriers was measured 7 % slower)
        psn_vcycle_smem(ctx, plan, L0 + 3, b3, u3, pool4, inv_s);
        // ---------------- level 2 up
        PSN_BK_FOR(ctx, NT) {
            if ((flags(t_) & 3) == 1) {   // real thread, (bx + by) even
                const CurV c = curv(t_);
                const int bx = bx_of(t_), by = by_of(t_);
                const int P3 = l3.pitch;
                double add;
                if (by & 1) {
                    add = u3[(((by - 1) >> 1) + 1) * P3 + ((bx - 1) >> 1) + 1];
                } else {
                    const double* e = u3 + ((by >> 1) + 1) * P3 + (bx >> 1) + 1;
                    add = 0.25 * ((e[-P3 - 1] + e[-P3]) + (e[-1] + e[0]));
                }
                z2[PSN_BK_TI][0] += (VT)add;
                ES<2>(c, TOP, 0, z2[PSN_BK_TI][0]);
            }
        }
        ctx.sync();
        for (int s = 0; s < nu2; ++s) {
            half_sweep_2(rd2, 1, false);
            half_sweep_2(rd2, 0, false);
        }
        }   // !dense2
        // ---------------- level 1 up: corrections from z2 (its plane holds every thread's current z2)
        PSN_BK_FOR(ctx, NTPAD) {
            const CurV c = curv(t_);
            const VT e00 = z2[PSN_BK_TI][0];
            const VT em0 = EL<2>(c, TOP, 0, -1, 0), e0m = EL<2>(c, TOP, 0, 0, -1);
            const VT emm = EL<2>(c, TOP, 0, -1, -1);
            z1[PSN_BK_TI][3] += e00;
            z1[PSN_BK_TI][0] += (VT)0.25 * ((emm + em0) + (e0m + e00));
            if (SPLIT) publish_bs<2>(c, z1[PSN_BK_TI], 0);
        }
        ctx.sync();
        if (!SPLIT) {   // shared planes: the plane just read takes the edges of z1 only now
            PSN_BK_FOR(ctx, NTPAD) { publish_bs<2>(curv(t_), z1[PSN_BK_TI], 0); }
            ctx.sync();
        }
        for (int s = 0; s < nu1; ++s) {
            PSN_BK_FOR(ctx, NTPAD) { half_sweep_bs<2>(curv(t_), rd1[PSN_BK_TI], z1[PSN_BK_TI], r1[PSN_BK_TI], 1, false); }
            ctx.sync();
            PSN_BK_FOR(ctx, NTPAD) { half_sweep_bs<2>(curv(t_), rd1[PSN_BK_TI], z1[PSN_BK_TI], r1[PSN_BK_TI], 0, false); }
            ctx.sync();
        }
        // ---------------- level 0 up: corrections from z1 (every edge of z1 is current in the buffers)
        PSN_BK_FOR(ctx, NTPAD) {
            const CurV c = curv(t_);
            VT (&zz)[16] = z[PSN_BK_TI];
            const VT (&ee)[4] = z1[PSN_BK_TI];
            const VT e00 = ee[0], e01 = ee[1], e10 = ee[2], e11 = ee[3];
            const VT em0 = EL<1>(c, BOT, 0, -1, 0), em1 = EL<1>(c, BOT, 1, -1, 0);
            const VT e0m = EL<1>(c, RIGHT, 0, 0, -1), e1m = EL<1>(c, RIGHT, 1, 0, -1);
            const VT emm = EL<1>(c, BOT, 1, -1, -1);
            zz[5] += e00;
            zz[7] += e01;
            zz[13] += e10;
            zz[15] += e11;
            zz[0] += (VT)0.25 * ((emm + em0) + (e0m + e00));
            zz[2] += (VT)0.25 * ((em0 + em1) + (e00 + e01));
            zz[8] += (VT)0.25 * ((e0m + e00) + (e1m + e10));
            zz[10] += (VT)0.25 * ((e00 + e01) + (e10 + e11));
            if (SPLIT) publish(c, zz, 0);
        }
        ctx.sync();
        if (!SPLIT) {   // shared planes: the z1 edges have been read; the buffers now take the red edges of z
            PSN_BK_FOR(ctx, NTPAD) { publish(curv(t_), z[PSN_BK_TI], 0); }
            ctx.sync();
        }
        for (int s = 0; s < nu; ++s) {
            half_sweep<GAM0>(1, false);
            half_sweep<GAM0>(0, false, s == nu - 1);
        }
    }

    // bt of the thread's 16 nodes.  Standard call form (sine factors, theta on the solve grid itself): row / column factors and
    // theta come through 8 + 4 + 4 loads from per-thread base pointers -- the generic psn_rhs per node (forcing-mode branches,
    // 64-bit index products, once in the prologue and once per true-residual check) was 20 % of the kernel's instructions.
    // Same operations in the same order as psn_rhs, so the values are bit-identical.
    PSN_FN void rhs16(const PsnSolveArgs& a, long long b, int t, double h2, double (&out)[16]) const {
        const int bx = bx_of(t), by = by_of(t);
        const PsnRhs& s = a.rhs;
        if (s.sx && !s.f && s.step == 1 && s.mn == N) {
            const double* sxp = s.sx + b * (long long)N + 4 * bx;
            const double* syp = s.sy + b * (long long)N + 4 * by;
            double sxv[4], syv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                sxv[k] = psn_ldg(sxp + k);
                syv[k] = psn_ldg(syp + k);
            }
            if (s.theta) {
                const double* thp = s.theta + (b * (long long)N + 4 * by) * N + 4 * bx;
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) {
                    double th[4];
                    psn_ld4(thp + aa * N, th);      // 2 x 128-bit loads (device arrays are 16-byte aligned, N % 4 == 0)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        double f = syv[aa] * sxv[c];
                        f = f * (1.0 / th[c]);
                        out[4 * aa + c] = -(f * h2);
                    }
                }
            } else {
#pragma unroll
                for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                    for (int c = 0; c < 4; ++c) out[4 * aa + c] = -((syv[aa] * sxv[c]) * h2);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) out[k] = psn_rhs(s, b, 4 * by + (k >> 2), 4 * bx + (k & 3), h2);
        }
    }

    // block-wide sum of per-thread partials (device: acc[0] of every thread; host: all entries)
    PSN_FN double tsum(const double (&acc)[PSN_BK_TS]) {
#if defined(__CUDACC__)
        return ctx.block_sum(acc[0]);
#else
        double s = 0.0;
        for (int t = 0; t < NT; ++t) s += acc[t];
        return s;
#endif
    }
};

// Whole solve of sample b.  smem: dynamic shared memory (psn_block_smem_doubles doubles).
template <int BX, class Ctx>
PSN_FN void psn_block_solve_body(Ctx& ctx, const PsnPlan& plan, const PsnSolveArgs& a, long long b, double* smem,
                                 const double* inv_s = nullptr) {
    typedef PsnBlockGeom<BX> G;
    typedef PsnBlockSolver<BX, Ctx> SV;
    typedef typename SV::VT VT;
    constexpr int N = G::N, NT = G::NT, NTPAD = G::NTPAD;
    SV S(ctx, smem);
    S.inv_s = inv_s;
    double* pool = smem + G::XS + G::ES;          // levels >= 3 (levels 1 and 2 live in registers)
    {
        // zero the ghost rings: of the coarse levels, and of the thread grid the edge buffers are indexed by
        const int zn = G::ES + psn_block_pool_doubles(plan);
        double* z0 = smem + G::XS;
        for (int idx = ctx.tid; idx < zn; idx += ctx.nthreads) z0[idx] = 0.0;
    }
    double acc[PSN_BK_TS];
#if defined(__CUDACC__)
    const PsnShF64 xa_ = S.xa_dev;
#define PSN_BK_XA2(t) xa_
#else
#define PSN_BK_XA2(t) S.xs.at(t)
#endif

    // ---- prologue: r = bt, x = 0
    PSN_BK_FOR(ctx, NTPAD) {
        const PsnShF64 xa = PSN_BK_XA2(t_);
        double bt16[16];
        if (t_ < NT) {
            S.rhs16(a, b, t_, plan.h2, bt16);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) bt16[k] = 0.0;     // idle threads of the last warp carry zeros through every phase
        }
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double v = bt16[k];
            S.r[PSN_BK_TI][k] = v;
            S.p[PSN_BK_TI][k] = 0.0;
            S.q[PSN_BK_TI][k] = 0.0;
            S.z[PSN_BK_TI][k] = (VT)0;
            S.bv[PSN_BK_TI][k] = (VT)0;
            xa.st(k * NTPAD, 0.0);
            s += v * v;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) S.z1[PSN_BK_TI][k] = S.r1[PSN_BK_TI][k] = (VT)0;
        S.z2[PSN_BK_TI][0] = S.r2[PSN_BK_TI][0] = (VT)0;
        acc[PSN_BK_TI] = s;
    }
    ctx.sync();
    const double bb = S.tsum(acc);
    const double tol2 = a.rtol * a.rtol * bb;

    int it = 0, status = PSN_ST_MAXITER, nrestart = 0;
    bool restart = true;
    double rz_old = 1.0, prev_true = 0.0, rr_final = bb;
    if (!(bb > 0.0)) {
        status = (bb == 0.0) ? PSN_ST_CONVERGED : PSN_ST_BREAKDOWN;
        rr_final = 0.0;
    } else {
        for (;;) {
            // right-hand side of the V-cycle: the residual, scaled by a power of two near 1 / ||r|| when the cycle runs in float
            // (exact in fp64; undone in <r,z> and in the direction update), so that the preconditioner does not depend on the
            // magnitude of the data
            const double vs = (sizeof(VT) == 4) ? psn_vscale(rr_final) : 1.0;
            const double vinv = 1.0 / vs;
            PSN_BK_FOR(ctx, NTPAD) {
#pragma unroll
                for (int k = 0; k < 16; ++k) S.bv[PSN_BK_TI][k] = (VT)(S.r[PSN_BK_TI][k] * vs);
            }
            S.template vcycle3<false>(plan, pool);
            PSN_BK_FOR(ctx, NTPAD) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 16; ++k) s += S.r[PSN_BK_TI][k] * (double)S.z[PSN_BK_TI][k];
                acc[PSN_BK_TI] = (S.flags(t_) & 1) ? s : 0.0;   // (float cycle: the idle lanes may hold garbage from aliased cells)
            }
            const double rz = S.tsum(acc) * vinv;   // (device: block_sum syncs, so every thread is past the V-cycle's last edge read)
            const double beta = restart ? 0.0 : rz / rz_old;
            restart = false;
            rz_old = rz;
            // p = z + beta p, publish every edge of p
            PSN_BK_FOR(ctx, NTPAD) {
#pragma unroll
                for (int k = 0; k < 16; ++k) S.p[PSN_BK_TI][k] = (double)S.z[PSN_BK_TI][k] * vinv + beta * S.p[PSN_BK_TI][k];
                SV::publish(S.curd(t_), S.p[PSN_BK_TI], -1);
                if (PSN_BLOCK_VF32) SV::template zero_ghosts<typename SV::CurD>(S.edge, t_, 16);   // the V-cycle's float planes overwrote them
            }
            ctx.sync();
            // q = At p, <p, q>
            PSN_BK_FOR(ctx, NTPAD) {
                double hT[4], hB[4], hL[4], hR[4];
                SV::load_halo_all(S.curd(t_), hT, hB, hL, hR);
                double s = 0.0;
#pragma unroll
                for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double pc = S.p[PSN_BK_TI][4 * aa + c];
                        const double qv = 4.0 * pc - SV::nbsum(S.p[PSN_BK_TI], aa, c, hT, hB, hL, hR);
                        S.q[PSN_BK_TI][4 * aa + c] = qv;
                        s += pc * qv;
                    }
                acc[PSN_BK_TI] = (S.flags(t_) & 1) ? s : 0.0;
            }
            const double pq = S.tsum(acc);
            if (!(pq > 0.0)) {
                status = PSN_ST_BREAKDOWN;
                break;
            }
            const double alpha = rz / pq;
            PSN_BK_FOR(ctx, NTPAD) {
                const PsnShF64 xa = PSN_BK_XA2(t_);
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    xa.add(k * NTPAD, alpha * S.p[PSN_BK_TI][k]);
                    const double rv = S.r[PSN_BK_TI][k] - alpha * S.q[PSN_BK_TI][k];
                    S.r[PSN_BK_TI][k] = rv;
                    s += rv * rv;
                }
                acc[PSN_BK_TI] = (S.flags(t_) & 1) ? s : 0.0;
                if (PSN_BLOCK_VF32) SV::template zero_ghosts<typename SV::CurV>(S.edgev, t_, SV::NPLV);   // every halo of p has been read
            }
            const double rr = S.tsum(acc);
            rr_final = rr;
            ++it;
            if (rr <= tol2 || it >= a.max_iter) {
                // true residual r = bt - At x: x block into q, publish, stencil
                PSN_BK_FOR(ctx, NTPAD) {
                    const PsnShF64 xa = PSN_BK_XA2(t_);
#pragma unroll
                    for (int k = 0; k < 16; ++k) S.q[PSN_BK_TI][k] = xa.ld(k * NTPAD);
                    SV::publish(S.curd(t_), S.q[PSN_BK_TI], -1);
                }
                ctx.sync();
                PSN_BK_FOR(ctx, NTPAD) {
                    double hT[4], hB[4], hL[4], hR[4];
                    SV::load_halo_all(S.curd(t_), hT, hB, hL, hR);
                    double s = 0.0;
                    double bt16[16];
                    if (t_ < NT) {
                        S.rhs16(a, b, t_, plan.h2, bt16);
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; ++k) bt16[k] = 0.0;
                    }
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const double bt = bt16[4 * aa + c];
                            const double rv = bt - (4.0 * S.q[PSN_BK_TI][4 * aa + c] - SV::nbsum(S.q[PSN_BK_TI], aa, c, hT, hB, hL, hR));
                            S.r[PSN_BK_TI][4 * aa + c] = rv;
                            s += rv * rv;
                        }
                    acc[PSN_BK_TI] = (S.flags(t_) & 1) ? s : 0.0;
                }
                const double rt = S.tsum(acc);
                rr_final = rt;
                if (rt <= tol2) { status = PSN_ST_CONVERGED; break; }
                if (nrestart >= 1 && rt > prev_true * 0.0625) { status = psn_stagnation_status(rt, bb, a.rtol); break; }
                if (it >= a.max_iter) { status = PSN_ST_MAXITER; break; }
                prev_true = rt;
                ++nrestart;
                restart = true;
                // the fp64 planes have just carried x: the float ghost ring must read as zero again (tsum above was the barrier
                // behind the last halo load of x; the first half sweep of the V-cycle has one before its first float load)
                if (PSN_BLOCK_VF32) {
                    PSN_BK_FOR(ctx, NTPAD) { SV::template zero_ghosts<typename SV::CurV>(S.edgev, t_, SV::NPLV); }
                }
            }
        }
    }

    // ---- epilogue: u and / or the fused window + downsample outputs
    PSN_BK_FOR(ctx, NT) {
        if (t_ < NT) {
            const PsnShF64 xa = PSN_BK_XA2(t_);
            const int bx = SV::bx_of(t_), by = SV::by_of(t_);
            double xv[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) xv[k] = xa.ld(k * NTPAD);
            if (a.u) {
                double* ug = a.u + b * (long long)(N * N) + (4 * by) * N + 4 * bx;
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) {
                    const double row[4] = {xv[4 * aa], xv[4 * aa + 1], xv[4 * aa + 2], xv[4 * aa + 3]};
                    psn_st4(ug + aa * N, row);      // 2 x 128-bit stores
                }
            }
            if (a.sub > 0) {
                const int s = a.sub, sc = (s + 1) >> 1;
                const int x0 = a.start_x[b], y0 = a.start_y[b];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int i = 4 * by + (k >> 2), j = 4 * bx + (k & 3);
                    const int wi = i - y0, wj = j - x0;
                    if (wi >= 0 && wi < s && wj >= 0 && wj < s) {
                        const double fv = psn_forcing(a.rhs, b, i, j);
                        const double tv = a.rhs.theta
                            ? a.rhs.theta[(b * a.rhs.mn + (long long)i * a.rhs.step) * a.rhs.mn + (long long)j * a.rhs.step] : 1.0;
                        const long long o = (b * s + wi) * s + wj;
                        if (a.u_fine) a.u_fine[o] = xv[k];
                        if (a.f_fine) a.f_fine[o] = fv;
                        if (a.th_fine) a.th_fine[o] = tv;
                        if (((wi | wj) & 1) == 0) {
                            const long long oc = (b * sc + (wi >> 1)) * sc + (wj >> 1);
                            if (a.u_coarse) a.u_coarse[oc] = xv[k];
                            if (a.f_coarse) a.f_coarse[oc] = fv;
                            if (a.th_coarse) a.th_coarse[oc] = tv;
                        }
                    }
                }
            }
        }
    }
#undef PSN_BK_XA2
    if (ctx.tid == 0) {
        if (a.iters) a.iters[b] = it;
        if (a.relres) a.relres[b] = (bb > 0.0) ? sqrt(rr_final / bb) : 0.0;
        if (a.status) a.status[b] = status;
    }
}

// ----------------------------------------------------------------------------- V-cycle only (coarse tail of the streamed solver)
// u = M^{-1} b on level L0 of the plan (n_L0 = 4 BX) and everything below it, for sample b of [B, n, n] arrays of TV (the float
// storage of the streamed V-cycle, or double): the register-blocked V-cycle above with level L0 as its level 0 -- which, being a
// coarse level, carries the boundary weight gam in its diagonal (GAM0).  Replaces the last strip levels and the shared-memory
// coarse kernel of the streamed path when an 80 x 80 level exists (n = 160, 320, 640).  `zeroed`: the caller has zeroed the
// edge planes and the pool once for this CTA -- the V-cycle never writes a ghost cell, so that holds for every sample.
template <int BX, class TV, class Ctx>
PSN_FN void psn_block_vcycle_body(Ctx& ctx, const PsnPlan& plan, int L0, const TV* bglob, TV* uglob, const PsnSampleState* state,
                                  long long b, double* smem, const double* inv_s, bool zeroed) {
    if (state && state[b].phase != PSN_PHASE_ITER) return;   // (uniform over the CTA)
    typedef PsnBlockGeom<BX> G;
    typedef PsnBlockSolver<BX, Ctx> SV;
    typedef typename SV::VT VT;
    constexpr int N = G::N, NT = G::NT, NTPAD = G::NTPAD;
    (void)NTPAD;   // (only the host emulation's phase loops use it)
    // the solver addresses "x, then the planes": hand it a base such that the planes start at smem (x is never touched here)
    SV S(ctx, smem - G::XS);
    S.inv_s = inv_s;
    S.L0 = L0;
    const PsnLevel& lv = plan.lv[L0];
    S.R0[0] = (VT)lv.rd0; S.R0[1] = (VT)lv.rd1; S.R0[2] = (VT)lv.rd2;
    S.gam0 = (VT)lv.gam;
    double* pool = smem + G::ES;
    if (!zeroed) {
        const int zn = G::ES + psn_block_pool_doubles(plan, L0);
        for (int idx = ctx.tid; idx < zn; idx += ctx.nthreads) smem[idx] = 0.0;
        ctx.sync();
    }
    const TV* bg = bglob + b * (long long)(N * N);
    TV* ug = uglob + b * (long long)(N * N);
    PSN_BK_FOR(ctx, NTPAD) {
        const int bx = SV::bx_of(t_), by = SV::by_of(t_);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                S.bv[PSN_BK_TI][4 * a + c] = (t_ < NT) ? (VT)bg[(4 * by + a) * N + 4 * bx + c] : (VT)0;
                S.z[PSN_BK_TI][4 * a + c] = (VT)0;
            }
#pragma unroll
        for (int k = 0; k < 4; ++k) S.z1[PSN_BK_TI][k] = S.r1[PSN_BK_TI][k] = (VT)0;
        S.z2[PSN_BK_TI][0] = S.r2[PSN_BK_TI][0] = (VT)0;
    }
    S.template vcycle3<true>(plan, pool);
    PSN_BK_FOR(ctx, NT) {
        if (t_ < NT) {
            const int bx = SV::bx_of(t_), by = SV::by_of(t_);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) ug[(4 * by + a) * N + 4 * bx + c] = (TV)S.z[PSN_BK_TI][4 * a + c];
        }
    }
    ctx.sync();   // the next sample's first half sweep publishes into planes that the last up-sweep of this one still reads
}
