// psn_warp.cuh -- "warp vector" types for kernels written at warp granularity.
//
// Device build (nvcc): wd / wi / wb are one lane's double / int / bool and the helpers are
// warp shuffles and predicated global accesses.  Host build (g++, tests/hostemu only): the same
// names are 32-lane arrays with element-wise operators, so a warp-synchronous kernel body runs
// unchanged on the CPU test tier.  Control flow in such bodies must be warp-uniform; everything
// lane-dependent goes through w_sel().
#pragma once
#include "psn_core.cuh"
#include <string.h>

#if defined(__CUDACC__)

typedef double wd;
typedef float wf;
typedef int wi;
typedef bool wb;
struct PsnWarp {
    int lane;
};

PSN_FN wd w_bcast(double s) { return s; }
PSN_FN wi w_lane(const PsnWarp& w) { return w.lane; }
PSN_FN wd w_sel(wb c, wd a, wd b) { return c ? a : b; }
PSN_FN wd w_rcp(wd a) { return 1.0 / a; }
// value of the previous / next lane; the edge lane receives 0
PSN_FN wd w_up(const PsnWarp& w, wd v) {
    const wd r = __shfl_up_sync(0xffffffffu, v, 1);
    return w.lane == 0 ? 0.0 : r;
}
PSN_FN wd w_down(const PsnWarp& w, wd v) {
    const wd r = __shfl_down_sync(0xffffffffu, v, 1);
    return w.lane == 31 ? 0.0 : r;
}
PSN_FN double w_sum(wd v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
PSN_FN bool w_lane0(const PsnWarp& w) { return w.lane == 0; }
// (An explicit elect.sync around the bulk-copy issue of the row feeders removes the compiler's per-instruction ELECT / PLOP3
//  election sequences -- 138 -> 19 of them in the fused kernel, +2.6 % at 640^2 -- but the fp64-stored V(1,1) strips then
//  converged in a varying number of iterations, i.e. raced, with or without a memory clobber or a re-sync after the issue;
//  the cause was not found, so the lane-0 form stays.)
// two adjacent doubles at row[col], row[col+1] (col even); masked elements read as 0.
// vec: the row base and col are 16-byte aligned, so one 128-bit access is legal.
PSN_FN void w_ld2(const double* row, wi col, wb m0, wb m1, bool vec, wd& a, wd& b) {
    if (vec && m0 && m1) {
        const double2 v = *reinterpret_cast<const double2*>(row + col);
        a = v.x;
        b = v.y;
    } else {
        a = m0 ? row[col] : 0.0;
        b = m1 ? row[col + 1] : 0.0;
    }
}
PSN_FN void w_st2(double* row, wi col, wb m0, wb m1, bool vec, wd a, wd b) {
    if (vec && m0 && m1) {
        *reinterpret_cast<double2*>(row + col) = make_double2(a, b);
    } else {
        if (m0) row[col] = a;
        if (m1) row[col + 1] = b;
    }
}
// raw neighbour-lane values (edge lanes receive their own value; callers only use these where the
// edge lane's result is halo / out-of-domain data)
PSN_FN wd w_up_raw(wd v) { return __shfl_up_sync(0xffffffffu, v, 1); }
PSN_FN wd w_down_raw(wd v) { return __shfl_down_sync(0xffffffffu, v, 1); }
// predicated 128-bit pair access: both elements valid or none (even n, even column)
PSN_FN void w_ldp(const double* row, wi col, wb m, wd& a, wd& b) {
    double2 v = make_double2(0.0, 0.0);
    if (m) v = *reinterpret_cast<const double2*>(row + col);
    a = v.x;
    b = v.y;
}
PSN_FN void w_stp(double* row, wi col, wb m, wd a, wd b) {
    if (m) *reinterpret_cast<double2*>(row + col) = make_double2(a, b);
}
// hint: pull the lane's 32 bytes of a future row into L2 (no register cost)
PSN_FN void w_prefetch_l2(const double* row, wi col, wb m) {
    if (m) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + col));
}
PSN_FN wd w_ld1(const double* row, wi idx, wb m) { return m ? row[idx] : 0.0; }
PSN_FN void w_st1(double* row, wi idx, wb m, wd v) {
    if (m) row[idx] = v;
}

// ---- per-warp asynchronous row staging: bulk global->shared copies (TMA engine, 1-D) completing on
// an mbarrier.  One lane issues, every lane waits.  A protocol error traps instead of hanging.
PSN_FN unsigned psn_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
PSN_FN void w_bar_init(const PsnWarp& w, unsigned long long* bars, int count) {
    if (w.lane == 0) {
        for (int i = 0; i < count; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(psn_smem_addr(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // generic-proxy initialisation of the ring (zero fill) must be visible to the async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
}
PSN_FN void w_sync(const PsnWarp&) { __syncwarp(); }
// lane 0: arm the barrier with the byte count of the copies that follow
PSN_FN void w_bar_expect(const PsnWarp& w, unsigned long long* bar, unsigned bytes) {
    if (w.lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(psn_smem_addr(bar)), "r"(bytes) : "memory");
}
PSN_FN void w_bulk_g2s(const PsnWarp& w, void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    if (w.lane == 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(psn_smem_addr(dst)), "l"(src), "r"(bytes), "r"(psn_smem_addr(bar)) : "memory");
}
// the same two operations for a caller that is already inside `if (w_lane0(w))` (one divergent region per row
// instead of one per instruction)
PSN_FN void w_bar_expect_l0(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(psn_smem_addr(bar)), "r"(bytes) : "memory");
}
PSN_FN void w_bulk_g2s_l0(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(psn_smem_addr(dst)), "l"(src), "r"(bytes), "r"(psn_smem_addr(bar)) : "memory");
}
PSN_FN void w_bar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = psn_smem_addr(bar);
    unsigned done = 0;
    long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) break;
        if (spin == 64) t0 = clock64();
        if (spin > 64 && (spin & 1023) == 0 && clock64() - t0 > 4000000000ll) __trap();   // ~2 s: protocol error
    }
}
// the lane's 4 doubles (columns 4*lane .. 4*lane+3) of a staged row
PSN_FN void w_lds4(const PsnWarp& w, const double* row, wd (&out)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(row + 4 * w.lane);
    const double2 b = *reinterpret_cast<const double2*>(row + 4 * w.lane + 2);
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}
PSN_FN void w_zero_smem(const PsnWarp& w, double* p, int count) {
    for (int i = w.lane; i < count; i += 32) p[i] = 0.0;
}
// float overloads: V-cycle intermediates are kept in fp32 in HBM and the strip kernels that work on them compute in
// fp32 as well (wf); the CG vectors, the level-0 right-hand side and every inner product stay fp64 (wd)
PSN_FN wf w_sel(wb c, wf a, wf b) { return c ? a : b; }
PSN_FN wf w_up_raw(wf v) { return __shfl_up_sync(0xffffffffu, v, 1); }
PSN_FN wf w_down_raw(wf v) { return __shfl_down_sync(0xffffffffu, v, 1); }
PSN_FN void w_prefetch_l2(const float* row, wi col, wb m) {
    if (m) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + col));
}
PSN_FN wf w_ld1(const float* row, wi idx, wb m) { return m ? row[idx] : 0.f; }
PSN_FN void w_st1(float* row, wi idx, wb m, wf v) {
    if (m) row[idx] = v;
}
PSN_FN void w_lds4(const PsnWarp& w, const float* row, wf (&out)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(row + 4 * w.lane);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
PSN_FN void w_zero_smem(const PsnWarp& w, float* p, int count) {
    for (int i = w.lane; i < count; i += 32) p[i] = 0.f;
}
// the lane's four outputs of a row: two pair stores (double) or one 128-bit store (float; strip geometry
// guarantees that both pairs of a lane are inside or outside the owned range together)
PSN_FN void w_stq(double* row, wi col0, wb m0, wb m1, const wd (&v)[4]) {
    w_stp(row, col0, m0, v[0], v[1]);
    w_stp(row, col0 + 2, m1, v[2], v[3]);
}
PSN_FN void w_stq(float* row, wi col0, wb m0, wb, const wf (&v)[4]) {
    if (m0) *reinterpret_cast<float4*>(row + col0) = make_float4(v[0], v[1], v[2], v[3]);
}
template <class S> struct PsnWV;
template <> struct PsnWV<double> { typedef wd type; };
template <> struct PsnWV<float> { typedef wf type; };
template <class S> PSN_FN typename PsnWV<S>::type w_bcastT(double s) { return (S)s; }
template <class S> PSN_FN typename PsnWV<S>::type w_cvt(wd v) { return (S)v; }
template <class S> PSN_FN typename PsnWV<S>::type w_cvt(wf v) { return (S)v; }

#else  // ------------------------------------------------------------------ host emulation

template <class S>
struct wvt {
    S v[32];
};
typedef wvt<double> wd;
typedef wvt<float> wf;
struct wi {
    int v[32];
};
struct wb {
    bool v[32];
};
struct PsnWarp {
    int lane = 0;
};

#define PSN_W_FOR for (int l_ = 0; l_ < 32; ++l_)
inline wd w_bcast(double s) { wd r; PSN_W_FOR r.v[l_] = s; return r; }
inline wi w_lane(const PsnWarp&) { wi r; PSN_W_FOR r.v[l_] = l_; return r; }
template <class S> inline wvt<S> operator+(const wvt<S>& a, const wvt<S>& b) { wvt<S> r; PSN_W_FOR r.v[l_] = a.v[l_] + b.v[l_]; return r; }
template <class S> inline wvt<S> operator-(const wvt<S>& a, const wvt<S>& b) { wvt<S> r; PSN_W_FOR r.v[l_] = a.v[l_] - b.v[l_]; return r; }
template <class S> inline wvt<S> operator*(const wvt<S>& a, const wvt<S>& b) { wvt<S> r; PSN_W_FOR r.v[l_] = a.v[l_] * b.v[l_]; return r; }
inline wd operator*(double a, const wd& b) { wd r; PSN_W_FOR r.v[l_] = a * b.v[l_]; return r; }
inline wi operator*(const wi& a, int b) { wi r; PSN_W_FOR r.v[l_] = a.v[l_] * b; return r; }
inline wi operator+(const wi& a, int b) { wi r; PSN_W_FOR r.v[l_] = a.v[l_] + b; return r; }
inline wb operator>=(const wi& a, int b) { wb r; PSN_W_FOR r.v[l_] = a.v[l_] >= b; return r; }
inline wb operator<(const wi& a, int b) { wb r; PSN_W_FOR r.v[l_] = a.v[l_] < b; return r; }
inline wb operator==(const wi& a, int b) { wb r; PSN_W_FOR r.v[l_] = a.v[l_] == b; return r; }
inline wb operator&&(const wb& a, const wb& b) { wb r; PSN_W_FOR r.v[l_] = a.v[l_] && b.v[l_]; return r; }
template <class S> inline wvt<S> w_rcp(const wvt<S>& a) { wvt<S> r; PSN_W_FOR r.v[l_] = (S)1 / a.v[l_]; return r; }
template <class S> inline wvt<S> w_sel(const wb& c, const wvt<S>& a, const wvt<S>& b) { wvt<S> r; PSN_W_FOR r.v[l_] = c.v[l_] ? a.v[l_] : b.v[l_]; return r; }
inline wd w_up(const PsnWarp&, const wd& v) { wd r; PSN_W_FOR r.v[l_] = l_ == 0 ? 0.0 : v.v[l_ - 1]; return r; }
inline wd w_down(const PsnWarp&, const wd& v) { wd r; PSN_W_FOR r.v[l_] = l_ == 31 ? 0.0 : v.v[l_ + 1]; return r; }
inline double w_sum(const wd& v) { double s = 0.0; PSN_W_FOR s += v.v[l_]; return s; }
inline bool w_lane0(const PsnWarp&) { return true; }
inline void w_ld2(const double* row, const wi& col, const wb& m0, const wb& m1, bool, wd& a, wd& b) {
    PSN_W_FOR {
        a.v[l_] = m0.v[l_] ? row[col.v[l_]] : 0.0;
        b.v[l_] = m1.v[l_] ? row[col.v[l_] + 1] : 0.0;
    }
}
inline void w_st2(double* row, const wi& col, const wb& m0, const wb& m1, bool, const wd& a, const wd& b) {
    PSN_W_FOR {
        if (m0.v[l_]) row[col.v[l_]] = a.v[l_];
        if (m1.v[l_]) row[col.v[l_] + 1] = b.v[l_];
    }
}
template <class S> inline wvt<S> w_up_raw(const wvt<S>& v) { wvt<S> r; PSN_W_FOR r.v[l_] = l_ == 0 ? v.v[0] : v.v[l_ - 1]; return r; }
template <class S> inline wvt<S> w_down_raw(const wvt<S>& v) { wvt<S> r; PSN_W_FOR r.v[l_] = l_ == 31 ? v.v[31] : v.v[l_ + 1]; return r; }
inline void w_ldp(const double* row, const wi& col, const wb& m, wd& a, wd& b) {
    PSN_W_FOR {
        a.v[l_] = m.v[l_] ? row[col.v[l_]] : 0.0;
        b.v[l_] = m.v[l_] ? row[col.v[l_] + 1] : 0.0;
    }
}
inline void w_stp(double* row, const wi& col, const wb& m, const wd& a, const wd& b) {
    PSN_W_FOR if (m.v[l_]) { row[col.v[l_]] = a.v[l_]; row[col.v[l_] + 1] = b.v[l_]; }
}
inline void w_prefetch_l2(const double*, const wi&, const wb&) {}
inline wd w_ld1(const double* row, const wi& idx, const wb& m) { wd r; PSN_W_FOR r.v[l_] = m.v[l_] ? row[idx.v[l_]] : 0.0; return r; }
inline void w_st1(double* row, const wi& idx, const wb& m, const wd& v) { PSN_W_FOR if (m.v[l_]) row[idx.v[l_]] = v.v[l_]; }

inline void w_bar_init(const PsnWarp&, unsigned long long*, int) {}
inline void w_sync(const PsnWarp&) {}
inline void w_bar_expect(const PsnWarp&, unsigned long long*, unsigned) {}
inline void w_bulk_g2s(const PsnWarp&, void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
inline void w_bar_wait(unsigned long long*, unsigned) {}
inline void w_bar_expect_l0(unsigned long long*, unsigned) {}
inline void w_bulk_g2s_l0(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
inline void w_lds4(const PsnWarp&, const double* row, wd (&out)[4]) {
    PSN_W_FOR for (int e = 0; e < 4; ++e) out[e].v[l_] = row[4 * l_ + e];
}
inline void w_zero_smem(const PsnWarp&, double* p, int count) { for (int i = 0; i < count; ++i) p[i] = 0.0; }
inline void w_prefetch_l2(const float*, const wi&, const wb&) {}
inline wf w_ld1(const float* row, const wi& idx, const wb& m) { wf r; PSN_W_FOR r.v[l_] = m.v[l_] ? row[idx.v[l_]] : 0.f; return r; }
inline void w_st1(float* row, const wi& idx, const wb& m, const wf& v) { PSN_W_FOR if (m.v[l_]) row[idx.v[l_]] = v.v[l_]; }
inline void w_lds4(const PsnWarp&, const float* row, wf (&out)[4]) {
    PSN_W_FOR for (int e = 0; e < 4; ++e) out[e].v[l_] = row[4 * l_ + e];
}
inline void w_zero_smem(const PsnWarp&, float* p, int count) { for (int i = 0; i < count; ++i) p[i] = 0.f; }
inline void w_stq(double* row, const wi& col0, const wb& m0, const wb& m1, const wd (&v)[4]) {
    w_stp(row, col0, m0, v[0], v[1]);
    w_stp(row, col0 + 2, m1, v[2], v[3]);
}
inline void w_stq(float* row, const wi& col0, const wb& m0, const wb& m1, const wf (&v)[4]) {
    PSN_W_FOR {
        // mirrors the device's single 128-bit store: the two pair masks must agree
        if (m0.v[l_] != m1.v[l_]) abort();
        if (m0.v[l_]) for (int e = 0; e < 4; ++e) row[col0.v[l_] + e] = v[e].v[l_];
    }
}
template <class S> struct PsnWV { typedef wvt<S> type; };
template <class S> inline wvt<S> w_bcastT(double s) { wvt<S> r; PSN_W_FOR r.v[l_] = (S)s; return r; }
template <class S, class F> inline wvt<S> w_cvt(const wvt<F>& v) { wvt<S> r; PSN_W_FOR r.v[l_] = (S)v.v[l_]; return r; }

#endif

// ----------------------------------------------------------------------------- strip kernels
// One warp owns a strip of PSN_STRIP_COLS = 128 columns (lane l: the 4 columns c0+4l .. c0+4l+3,
// c0 even) and marches down the rows of its band.  A rolling window of W = 2*NU+2 rows lives in
// registers; at time t the pipeline applies
//     pre : R(t) [zero guess], B(t-1), R(t-2), ..., B(t-2NU+1), residual(t-2NU) -> restriction
//     post: prolong(t), B(t-1), R(t-2), ..., R(t-2NU)                           -> <b,u> partial
// so every sweep of a V-cycle leg is done in ONE pass over the rows, with vertical neighbours in
// registers and horizontal neighbours in registers or through ONE shuffle per stage.  Row parity
// works out so that at time t every stage touches the SAME register offsets {P, P+2}, P = t & 1.
// The validity window shrinks by one column per stage from the strip edges (halo) exactly as in
// the tiled version, so the result is identical to the untiled algorithm.  Out-of-domain columns
// are kept at zero by a zero reciprocal diagonal (no per-stage masks); rows are handled by a
// checked chunk at the band ends and a branch-free chunk in the steady state.
// Requires even n (every 2-column pair is entirely inside or outside the domain, 128-bit accesses).
#define PSN_STRIP_COLS 128
#define PSN_STRIP_CPL 4
#define PSN_STRIP_L2_AHEAD 16   // rows ahead of the march that are pulled into L2

// Element types: b is double on level 0 (the CG residual) and TV on coarse levels; every other grid array is TV,
// the storage type of the V-cycle intermediates (double, or float when PsnStreamCfg::vf32 is set).
struct PsnStripArgs {
    int level;
    int band;            // output rows per task (even)
    int nstrips, nbands; // tasks per sample = nstrips * nbands
    const void* b;       // [B, n, n]
    const void* uin;     // post
    void* uout;
    void* bc;            // pre:  [B, nc, nc] out
    const void* ec;      // post: [B, nc, nc] in
    double* part;        // post, level 0: [B, nstrips*nbands] partial <b,u>
    const PsnSampleState* state;
    int fused;           // the CG update is fused with the level-0 down-leg (psn_strip_upre_body): see psn_vcycle_scale
    int only_restart;    // pre, level 0, fused schedule: only samples whose residual was rebuilt (prologue / true residual)
};

// useful (output) columns per strip: pre needs 2NU halo columns left and 2NU+1 right; post needs
// 2NU+2 left (the prolongation onto an even column reads the coarse value one lane to the left)
// and 2NU right
// With fp32 storage (f32 != 0) every strip origin and owned range is a multiple of 4 columns, so that float rows
// are accessed in whole 16-byte quads (bulk copies, 128-bit stores): the left halo is rounded up to a multiple of
// 4 and the useful width down.
PSN_HDFN int psn_strip_hl_pre(int nu, int f32) { return f32 ? ((2 * nu + 3) & ~3) : 2 * nu; }
PSN_HDFN int psn_strip_hl_post(int nu, int f32) { return f32 ? ((2 * nu + 2 + 3) & ~3) : 2 * nu + 2; }
PSN_HDFN int psn_strip_use_pre(int nu, int f32) {
    const int u = PSN_STRIP_COLS - psn_strip_hl_pre(nu, f32) - (2 * nu + 1);
    return f32 ? (u & ~3) : (u & ~1);
}
PSN_HDFN int psn_strip_use_post(int nu, int f32) {
    const int u = PSN_STRIP_COLS - psn_strip_hl_post(nu, f32) - 2 * nu;
    return f32 ? (u & ~3) : (u & ~1);
}
// Scale of the float-stored V-cycle of a sample's CURRENT outer iteration (a power of two near 1 / ||r||), kept in the sample
// state by the state kernels.  Unfused: from the residual norm itself.  Fused schedule: the down-leg of iteration k+1 runs
// inside the update kernel of iteration k, before ||r_{k+1}|| exists, so it scales with ||r_k||.  (The scale only positions the
// data in the float exponent range; a residual 30x smaller than its scale assumes loses nothing.)
PSN_FN double psn_vcycle_scale(const PsnSampleState& st, int /*fused*/) { return st.vs; }

// horizontal neighbour sum of offset e (static) in a row held as v[0..3]
#define PSN_LR(v, e, upv, dnv) \
    (((e) == 0 ? (upv) : (v)[((e) + PSN_STRIP_CPL - 1) % PSN_STRIP_CPL]) + ((e) == PSN_STRIP_CPL - 1 ? (dnv) : (v)[((e) + 1) % PSN_STRIP_CPL]))

template <int NU>
struct PsnStripGeom {
    static constexpr int NS = 2 * NU;
    static constexpr int W = NS + 2;
    static constexpr int PF = W / 2;
};


// ---- row feeders: how the next rows of the streamed arrays reach the warp
// PsnRegFeeder: prefetch ring of PF rows in registers + L2 prefetch hints further ahead.
// PsnTmaFeeder: ring of W rows per array in shared memory, filled W-1 rows ahead by 1-D bulk
//               copies (TMA engine) that complete on one mbarrier per ring slot; no registers are
//               tied up by loads in flight, so the prefetch distance is 2-3x deeper.
template <int NU, int NARR>
struct PsnRegFeeder {
    typedef wd V0;
    typedef wd V1;
    static constexpr int W = 2 * NU + 2, PF = W / 2;
    wd pf[NARR][PF][4];
    const double* base[NARR];
    int n, tload;
    wi col0;
    wb mp[2];
    PSN_FN void set_base0(const double* p) { base[0] = p; }
    PSN_FN void set_base1(const double* p) { base[NARR - 1] = p; }
    PSN_FN void prime(const PsnWarp&, int ts) {
#pragma unroll
        for (int a = 0; a < NARR; ++a)
#pragma unroll
            for (int k = 0; k < PF; ++k) {
#pragma unroll
                for (int e = 0; e < 4; ++e) pf[a][k][e] = w_bcast(0.0);
                if (ts + k <= tload) {
                    w_ldp(base[a] + (long long)(ts + k) * n, col0, mp[0], pf[a][k][0], pf[a][k][1]);
                    w_ldp(base[a] + (long long)(ts + k) * n, col0 + 2, mp[1], pf[a][k][2], pf[a][k][3]);
                }
            }
    }
    // values of row tc (zero beyond tload) of array 0 (and 1); starts the load of row tc + PF
    template <bool FAST>
    PSN_FN void fetch(const PsnWarp&, int tt, int tc, int /*par*/, wd (&o0)[4], wd (&o1)[4]) {
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (a == 0) o0[e] = pf[a][tt % PF][e];
                else o1[e] = pf[a][tt % PF][e];
            }
            if (FAST || tc + PF <= tload) {
                const double* rowp = base[a] + (long long)(tc + PF) * n;
                w_ldp(rowp, col0, mp[0], pf[a][tt % PF][0], pf[a][tt % PF][1]);
                w_ldp(rowp, col0 + 2, mp[1], pf[a][tt % PF][2], pf[a][tt % PF][3]);
                if (tc + PSN_STRIP_L2_AHEAD <= tload) w_prefetch_l2(base[a] + (long long)(tc + PSN_STRIP_L2_AHEAD) * n, col0, mp[0]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) pf[a][tt % PF][e] = w_bcast(0.0);
            }
        }
    }
};

template <class T>
struct PsnElemOr {
    typedef T type;
    static constexpr int bytes = (int)sizeof(T);
    static constexpr int narr = 1;
};
template <>
struct PsnElemOr<void> {
    typedef double type;
    static constexpr int bytes = 0;
    static constexpr int narr = 0;
};

// T0 / T1 / T2: element types of the one to three streamed arrays (void: absent)
template <int NU, class T0, class T1 = void, class T2 = void>
struct PsnTmaFeeder {
    typedef typename PsnElemOr<T1>::type U1;
    typedef typename PsnElemOr<T2>::type U2;
    typedef typename PsnWV<T0>::type V0;   // register type the rows are delivered in: the array's own precision
    typedef typename PsnWV<U1>::type V1;
    typedef typename PsnWV<U2>::type V2;
    static constexpr int NARR = 1 + PsnElemOr<T1>::narr + PsnElemOr<T2>::narr;
    static constexpr int W = 2 * NU + 2, D = W - 1;
    static constexpr int ROW_BYTES = (int)sizeof(T0) + PsnElemOr<T1>::bytes + PsnElemOr<T2>::bytes;
    static constexpr int RING_BYTES = W * PSN_STRIP_COLS * ROW_BYTES;
    T0* ring0;                   // [W][PSN_STRIP_COLS], shared memory of this warp
    U1* ring1;
    U2* ring2;
    unsigned long long* bar;     // [W]
    const T0* base0;
    const U1* base1;
    const U2* base2;
    const T0* nx0;               // source of the next row to be issued (rows are issued strictly in order, so the
    const U1* nx1;               // address advances by one row pitch per issue instead of being recomputed)
    const U2* nx2;
    int n, tload, ts;
    int rows1_max;               // array 1 is only copied for rows <= rows1_max (fused update + down-leg: its rows run one ahead)
    int cA, cB, c0;              // in-domain column range [cA, cB) of the strip, strip origin
#if defined(__CUDACC__)
    __host__ __device__
#endif
    static constexpr int smem_bytes() { return RING_BYTES + W * 8 + (128 - (W * 8) % 128) % 128; }
    PSN_FN void attach(unsigned char* mem) {
        ring0 = reinterpret_cast<T0*>(mem);
        ring1 = reinterpret_cast<U1*>(mem + (size_t)W * PSN_STRIP_COLS * sizeof(T0));
        ring2 = reinterpret_cast<U2*>(mem + (size_t)W * PSN_STRIP_COLS * (sizeof(T0) + PsnElemOr<T1>::bytes));
        bar = reinterpret_cast<unsigned long long*>(mem + RING_BYTES);
        base1 = nullptr;
        base2 = nullptr;
        rows1_max = 0x7fffffff;
    }
    PSN_FN void set_base0(const T0* p) { base0 = p; }
    PSN_FN void set_base1(const U1* p) { base1 = p; }
    PSN_FN void set_base2(const U2* p) { base2 = p; }
    PSN_FN void issue(const PsnWarp& w, int row /*always the successor of the previously issued row*/, int slot) {
        if (w_lane0(w)) {
            const unsigned cnt = (unsigned)(cB - cA);
            const int so = slot * PSN_STRIP_COLS + (cA - c0);
            const bool a1 = NARR >= 2 && row <= rows1_max;
            w_bar_expect_l0(bar + slot, cnt * (unsigned)(ROW_BYTES - (a1 ? 0 : PsnElemOr<T1>::bytes)));
            w_bulk_g2s_l0(ring0 + so, nx0, cnt * (unsigned)sizeof(T0), bar + slot);
            if (a1) w_bulk_g2s_l0(ring1 + so, nx1, cnt * (unsigned)sizeof(U1), bar + slot);
            if (NARR >= 3) w_bulk_g2s_l0(ring2 + so, nx2, cnt * (unsigned)sizeof(U2), bar + slot);
        }
        nx0 += n;
        if (NARR >= 2) nx1 += n;
        if (NARR >= 3) nx2 += n;
    }
    PSN_FN void prime(const PsnWarp& w, int ts_) {
        ts = ts_;
        nx0 = base0 + (long long)ts * n + cA;
        nx1 = (NARR >= 2) ? base1 + (long long)ts * n + cA : nullptr;
        nx2 = (NARR >= 3) ? base2 + (long long)ts * n + cA : nullptr;
        w_zero_smem(w, ring0, W * PSN_STRIP_COLS);   // out-of-domain columns stay zero
        if (NARR >= 2) w_zero_smem(w, ring1, W * PSN_STRIP_COLS);
        if (NARR >= 3) w_zero_smem(w, ring2, W * PSN_STRIP_COLS);
        w_bar_init(w, bar, W);
        for (int k = 0; k < D; ++k)
            if (ts + k <= tload) issue(w, ts + k, k);
    }
    // refill the slot of the previous row, then wait for row tc; returns false (rows read as zero) beyond tload
    template <bool FAST>
    PSN_FN bool advance(const PsnWarp& w, int tt, int tc, int par) {
        // every lane has finished reading the slot of row tc-1 (previous time step): refill it with row tc+D
        w_sync(w);
        if (FAST || tc + D <= tload) issue(w, tc + D, (tt + W - 1) % W);
        if (FAST || tc <= tload) {
            w_bar_wait(bar + tt, (unsigned)par);
            return true;
        }
        return false;
    }
    template <bool FAST>
    PSN_FN void fetch(const PsnWarp& w, int tt, int tc, int par, V0 (&o0)[4], V1 (&o1)[4]) {
        if (advance<FAST>(w, tt, tc, par)) {
            w_lds4(w, ring0 + tt * PSN_STRIP_COLS, o0);
            if (NARR >= 2) w_lds4(w, ring1 + tt * PSN_STRIP_COLS, o1);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                o0[e] = w_bcastT<T0>(0.0);
                if (NARR >= 2) o1[e] = w_bcastT<U1>(0.0);
            }
        }
    }
    template <bool FAST>
    PSN_FN void fetch3(const PsnWarp& w, int tt, int tc, int par, V0 (&o0)[4], V1 (&o1)[4], V2 (&o2)[4]) {
        if (advance<FAST>(w, tt, tc, par)) {
            w_lds4(w, ring0 + tt * PSN_STRIP_COLS, o0);
            w_lds4(w, ring1 + tt * PSN_STRIP_COLS, o1);
            w_lds4(w, ring2 + tt * PSN_STRIP_COLS, o2);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                o0[e] = w_bcastT<T0>(0.0);
                o1[e] = w_bcastT<U1>(0.0);
                o2[e] = w_bcastT<U2>(0.0);
            }
        }
    }
};

// TMA feeder of a V-cycle leg: pre streams the rhs (TB); post streams the rhs (TB) and the pre-smoothed iterate (TV)
template <int NU, bool POST, class TB, class TV>
struct PsnStripFeederSel {
    typedef PsnTmaFeeder<NU, TB, void> type;
};
template <int NU, class TB, class TV>
struct PsnStripFeederSel<NU, true, TB, TV> {
    typedef PsnTmaFeeder<NU, TB, TV> type;
};

template <int NU, int NARR>
PSN_FN void psn_feeder_geom(PsnRegFeeder<NU, NARR>& fd, const wi& col0, const wb (&mp)[2], int, int) {
    fd.col0 = col0;
    fd.mp[0] = mp[0];
    fd.mp[1] = mp[1];
}
template <int NU, class T0, class T1, class T2>
PSN_FN void psn_feeder_geom(PsnTmaFeeder<NU, T0, T1, T2>& fd, const wi&, const wb (&)[2], int c0, int n) {
    fd.c0 = c0;
    fd.cA = psn_imax(c0, 0);
    fd.cB = psn_imin(c0 + PSN_STRIP_COLS, n);
}

template <int NU, bool FAST, bool SCALE, class TV, class Feeder>
PSN_FN void psn_strip_pre_chunk(const PsnWarp& w, int t, int ts, int tload, int tend, int last, int n, int nc,
                                int i0, int i1, int I0, int I1,
                                typename PsnWV<TV>::type (&val)[2 * NU + 2][4], typename PsnWV<TV>::type (&bb)[2 * NU + 2][4],
                                Feeder& fd, int par, typename PsnWV<TV>::type (&acc)[2],
                                const typename PsnWV<TV>::type (&rdA)[4], const typename PsnWV<TV>::type (&dA)[4],
                                const typename PsnWV<TV>::type (&mk)[4], const PsnLevel& lv,
                                const wd& bsc, TV* ug, TV* cg, const wi& col0, const wb (&mp)[2], const wb (&so)[2],
                                const wi& cj, const wb (&sc)[2]) {
    constexpr int NS = 2 * NU, W = NS + 2;
    typedef typename PsnWV<TV>::type WV;   // compute type of the sweeps = storage type of the iterate
    const WV zero = w_bcastT<TV>(0.0);
    const WV q25 = w_bcastT<TV>(0.25);
    // last-row diagonals are only needed by the checked chunk at the bottom of the grid: build them on demand
    WV rdB[4], dB[4];
    if (!FAST) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const wb lc_ = (col0 + e) == last;
            rdB[e] = mk[e] * w_sel(lc_, w_bcastT<TV>(lv.rd2), w_bcastT<TV>(lv.rd1));
            dB[e] = w_sel(lc_, w_bcastT<TV>(4.0 + 2.0 * lv.gam), w_bcastT<TV>(4.0 + lv.gam));
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) rdB[e] = dB[e] = zero;
    }
#pragma unroll
    for (int tt = 0; tt < W; ++tt) {
        const int tc = t + tt;
        if (!FAST && tc > tend) break;
        const int P = tt & 1;
        const int k0 = tt;
        // ---- new row: rhs from the feeder; first red half sweep from the zero guess
        {
            typename Feeder::V0 r0[4];
            typename Feeder::V1 r1[4];
            fd.template fetch<FAST>(w, tt, tc, par, r0, r1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if constexpr (SCALE) bb[k0][e] = w_cvt<TV>(r0[e] * bsc);
                else bb[k0][e] = w_cvt<TV>(r0[e]);
            }
        }
        if (FAST || tc <= last) {
            const bool lastrow = !FAST && tc == last;
            val[k0][P] = bb[k0][P] * (lastrow ? rdB[P] : rdA[P]);
            val[k0][P + 2] = bb[k0][P + 2] * (lastrow ? rdB[P + 2] : rdA[P + 2]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) bb[k0][e] = zero;
            val[k0][P] = val[k0][P + 2] = zero;
        }
        val[k0][1 - P] = val[k0][3 - P] = zero;
        // ---- smoothing stages 1 .. NS-1 on rows tc-1 .. tc-NS+1
#pragma unroll
        for (int s = 1; s < NS; ++s) {
            const int r = tc - s;
            if (FAST || (r >= ts && r <= last)) {
                const int ks = (tt - s + 2 * W) % W, ku = (tt - s - 1 + 2 * W) % W, kd = (tt - s + 1 + 2 * W) % W;
                const bool lastrow = !FAST && r == last;
                const WV upv = (P == 0) ? w_up_raw(val[ks][3]) : zero;
                const WV dnv = (P == 1) ? w_down_raw(val[ks][0]) : zero;
                const WV sa = (val[ku][P] + val[kd][P]) + PSN_LR(val[ks], P, upv, dnv);
                const WV sb = (val[ku][P + 2] + val[kd][P + 2]) + PSN_LR(val[ks], P + 2, upv, dnv);
                val[ks][P] = (bb[ks][P] + sa) * (lastrow ? rdB[P] : rdA[P]);
                val[ks][P + 2] = (bb[ks][P + 2] + sb) * (lastrow ? rdB[P + 2] : rdA[P + 2]);
            }
        }
        // ---- residual on the red nodes of row tc-NS, accumulate the restriction
        {
            const int r = tc - NS;
            WV ra = zero, rb = zero;
            if (FAST || (r >= ts && r <= last)) {
                const int ks = (tt - NS + 2 * W) % W, ku = (tt - NS - 1 + 2 * W) % W, kd = (tt - NS + 1 + 2 * W) % W;
                const bool lastrow = !FAST && r == last;
                const WV upv = (P == 0) ? w_up_raw(val[ks][3]) : zero;
                const WV dnv = (P == 1) ? w_down_raw(val[ks][0]) : zero;
                const WV sa = (val[ku][P] + val[kd][P]) + PSN_LR(val[ks], P, upv, dnv);
                const WV sb = (val[ku][P + 2] + val[kd][P + 2]) + PSN_LR(val[ks], P + 2, upv, dnv);
                ra = mk[P] * (bb[ks][P] - ((lastrow ? dB[P] : dA[P]) * val[ks][P] - sa));
                rb = mk[P + 2] * (bb[ks][P + 2] - ((lastrow ? dB[P + 2] : dA[P + 2]) * val[ks][P + 2] - sb));
            }
            if (P == 0) {  // even row 2I: corners of coarse rows I-1 (bottom) and I (top); lane holds J = cj, cj+1
                const WV c0v = q25 * (ra + rb);
                const WV c1v = q25 * (rb + w_down_raw(ra));
                const int I = (r >> 1) - 1;
                if (FAST || (r >= 2 && I >= I0 && I < I1)) {
                    TV* crow = cg + (long long)I * nc;
                    w_st1(crow, cj, sc[0], acc[0] + c0v);
                    w_st1(crow, cj + 1, sc[1], acc[1] + c1v);
                }
                acc[0] = c0v;
                acc[1] = c1v;
            } else {       // odd row 2I+1: centres
                acc[0] = acc[0] + ra;
                acc[1] = acc[1] + rb;
            }
        }
        // ---- row tc-NS+1 is final after its last black half sweep
        {
            const int ro = tc - (NS - 1);
            if (FAST || (ro >= i0 && ro < i1)) {
                const int ko = (tt - (NS - 1) + 2 * W) % W;
                w_stq(ug + (long long)ro * n, col0, so[0], so[1], val[ko]);
            }
        }
    }
}

// TB: element type of the right-hand side (double on level 0, TV below); TV: storage type of the iterate and of the
// restricted residual.  With TV = float the level-0 rhs is scaled by psn_vscale(||r||^2) on load.
template <int NU, class TB, class TV, class Feeder>
PSN_FN void psn_strip_pre_body(const PsnWarp& w, const PsnPlan& plan, const PsnStripArgs& a, long long b, int band, int strip, Feeder& fd) {
    if (a.state && a.state[b].phase != PSN_PHASE_ITER) return;
    if (a.state && a.only_restart && !a.state[b].restart) return;
    constexpr int NS = 2 * NU, W = NS + 2;
    constexpr int F32 = sizeof(TV) == 4 ? 1 : 0;
    constexpr bool SCALE = F32 && sizeof(TB) == 8;
    const int USE = psn_strip_use_pre(NU, F32);
    const int HL = psn_strip_hl_pre(NU, F32);
    const PsnLevel& lv = plan.lv[a.level];
    const PsnLevel& lc = plan.lv[a.level + 1];
    const int n = lv.n, nc = lc.n, last = n - 1;
    const int a0 = strip * USE, a1 = psn_imin(a0 + USE, n);
    const int c0 = a0 - HL;
    const int i0 = band * a.band, i1 = psn_imin(i0 + a.band, n);
    const int I0 = i0 >> 1, I1 = psn_imin(i1 >> 1, nc);
    const int J0 = a0 >> 1, J1 = psn_imin(a1 >> 1, nc);

    const wi col0 = w_lane(w) * 4 + c0;
    const wi col2 = col0 + 2;
    const wb mp[2] = {(col0 >= 0) && (col0 < n), (col2 >= 0) && (col2 < n)};
    const wb so[2] = {mp[0] && (col0 >= a0) && (col0 < a1), mp[1] && (col2 >= a0) && (col2 < a1)};
    const wi cj = w_lane(w) * 2 + (c0 >> 1);
    const wb sc[2] = {(cj >= J0) && (cj < J1), (cj + 1 >= J0) && (cj + 1 < J1)};
    typedef typename PsnWV<TV>::type WV;
    const WV zero = w_bcastT<TV>(0.0);
    WV rdA[4], dA[4], mk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const wi ce = col0 + e;
        const wb in = (e < 2) ? mp[0] : mp[1];
        const wb lc_ = ce == last;
        rdA[e] = w_sel(in, w_sel(lc_, w_bcastT<TV>(lv.rd1), w_bcastT<TV>(lv.rd0)), zero);
        dA[e] = w_sel(lc_, w_bcastT<TV>(4.0 + lv.gam), w_bcastT<TV>(4.0));
        mk[e] = w_sel(in, w_bcastT<TV>(1.0), zero);
    }

    const int ts = psn_imax(i0 - NS, 0);
    const int tload = psn_imin(i1 + NS, last);
    const int tend = i1 + NS;
    const TB* bg = reinterpret_cast<const TB*>(a.b) + b * (long long)n * n;
    TV* ug = reinterpret_cast<TV*>(a.uout) + b * (long long)n * n;
    TV* cg = reinterpret_cast<TV*>(a.bc) + b * (long long)nc * nc;
    const wd bsc = w_bcast((SCALE && a.state) ? a.state[b].vs : 1.0);

    WV val[W][4], bb[W][4], acc[2];
#pragma unroll
    for (int k = 0; k < W; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) val[k][e] = bb[k][e] = zero;
    fd.set_base0(bg);
    fd.n = n;
    fd.tload = tload;
    psn_feeder_geom(fd, col0, mp, c0, n);
    fd.prime(w, ts);
    acc[0] = acc[1] = zero;

    for (int t = ts; t <= tend; t += W) {
        // steady state: every row touched by this chunk is an interior row of the band
        const bool fast = (t - NS - 1 >= ts) && (t - NS >= 2) && (t + W - 1 + (W - 1) <= tload) && (t + W - 1 < last) &&
                          (t - (NS - 1) >= i0) && (t + W - 1 - (NS - 1) < i1) &&
                          (((t - NS) >> 1) - 1 >= I0) && (((t + W - 1 - NS) >> 1) - 1 < I1);
        const int par = ((t - ts) / W) & 1;
        if (fast)
            psn_strip_pre_chunk<NU, true, SCALE, TV>(w, t, ts, tload, tend, last, n, nc, i0, i1, I0, I1, val, bb, fd, par, acc, rdA, dA, mk, lv,
                                                 bsc, ug, cg, col0, mp, so, cj, sc);
        else
            psn_strip_pre_chunk<NU, false, SCALE, TV>(w, t, ts, tload, tend, last, n, nc, i0, i1, I0, I1, val, bb, fd, par, acc, rdA, dA, mk, lv,
                                                  bsc, ug, cg, col0, mp, so, cj, sc);
    }
}

template <int NU, bool FAST, bool SCALE, class TV, class Feeder>
PSN_FN void psn_strip_post_chunk(const PsnWarp& w, int t, int ts, int tload, int tend, int last, int n, int nc, int i0, int i1,
                                 typename PsnWV<TV>::type (&val)[2 * NU + 2][4], typename PsnWV<TV>::type (&bb)[2 * NU + 2][4],
                                 Feeder& fd, int par, typename PsnWV<TV>::type (&eprev)[2], typename PsnWV<TV>::type (&enext)[4], wd& dot,
                                 const typename PsnWV<TV>::type (&rdA)[4], const typename PsnWV<TV>::type (&mk)[4], const PsnLevel& lv,
                                 const wd& bsc, const TV* eg, TV* ug, const wi& col0,
                                 const wb (&mp)[2], const wb (&so)[2], const wi& cj, const wb (&mc)[2], bool want_dot) {
    constexpr int NS = 2 * NU, W = NS + 2;
    typedef typename PsnWV<TV>::type WV;
    const WV zero = w_bcastT<TV>(0.0);
    const WV q25 = w_bcastT<TV>(0.25);
    const wd zerod = w_bcast(0.0);
    WV rdB[4];
    if (!FAST) {
#pragma unroll
        for (int e = 0; e < 4; ++e) rdB[e] = mk[e] * w_sel((col0 + e) == last, w_bcastT<TV>(lv.rd2), w_bcastT<TV>(lv.rd1));
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) rdB[e] = zero;
    }
#pragma unroll
    for (int tt = 0; tt < W; ++tt) {
        const int tc = t + tt;
        if (!FAST && tc > tend) break;
        const int P = tt & 1;
        const int k0 = tt;
        // ---- new row: rhs and pre-smoothed iterate, then the prolongated correction on the red nodes
        {
            typename Feeder::V0 r0[4];
            typename Feeder::V1 r1[4];
            fd.template fetch<FAST>(w, tt, tc, par, r0, r1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if constexpr (SCALE) bb[k0][e] = w_cvt<TV>(r0[e] * bsc);
                else bb[k0][e] = w_cvt<TV>(r0[e]);
                val[k0][e] = w_cvt<TV>(r1[e]);
            }
        }
        if (FAST || tc <= last) {
            if (P == 0) {  // even row 2I, red = even columns: mean of coarse (I-1,J-1), (I-1,J), (I,J-1), (I,J)
                const int I = tc >> 1;
                // coarse row I was requested four fine rows ago (two-deep queue enext[0..1] = row I, enext[2..3] =
                // row I+1); request row I+2 now
                const WV enew[2] = {enext[0], enext[1]};
                enext[0] = enext[2];
                enext[1] = enext[3];
                enext[2] = enext[3] = zero;
                if (FAST || I + 2 < nc) {
                    const TV* erow = eg + (long long)(I + 2) * nc;
                    enext[2] = w_ld1(erow, cj, mc[0]);
                    enext[3] = w_ld1(erow, cj + 1, mc[1]);
                    if (I + 10 < nc) w_prefetch_l2(eg + (long long)(I + 10) * nc, cj, mc[0]);
                }
                const WV s0 = eprev[0] + enew[0], s1 = eprev[1] + enew[1];   // column sums J = cj, cj+1
                const WV sl = w_up_raw(s1);                                     // column J = cj-1
                val[k0][0] = mk[0] * (val[k0][0] + q25 * (sl + s0));
                val[k0][2] = mk[2] * (val[k0][2] + q25 * (s0 + s1));
                eprev[0] = enew[0];
                eprev[1] = enew[1];
            } else {       // odd row 2I+1, red = odd columns: coincide with coarse (I, J)
                val[k0][1] = mk[1] * (val[k0][1] + eprev[0]);
                val[k0][3] = mk[3] * (val[k0][3] + eprev[1]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) val[k0][e] = bb[k0][e] = zero;
        }
        // ---- smoothing stages 1 .. NS (black first) on rows tc-1 .. tc-NS
#pragma unroll
        for (int s = 1; s <= NS; ++s) {
            const int r = tc - s;
            if (FAST || (r >= ts && r <= last)) {
                const int ks = (tt - s + 2 * W) % W, ku = (tt - s - 1 + 2 * W) % W, kd = (tt - s + 1 + 2 * W) % W;
                const bool lastrow = !FAST && r == last;
                const WV upv = (P == 0) ? w_up_raw(val[ks][3]) : zero;
                const WV dnv = (P == 1) ? w_down_raw(val[ks][0]) : zero;
                const WV sa = (val[ku][P] + val[kd][P]) + PSN_LR(val[ks], P, upv, dnv);
                const WV sb = (val[ku][P + 2] + val[kd][P + 2]) + PSN_LR(val[ks], P + 2, upv, dnv);
                val[ks][P] = (bb[ks][P] + sa) * (lastrow ? rdB[P] : rdA[P]);
                val[ks][P + 2] = (bb[ks][P + 2] + sb) * (lastrow ? rdB[P + 2] : rdA[P + 2]);
            }
        }
        // ---- row tc-NS is final
        {
            const int ro = tc - NS;
            if (FAST || (ro >= i0 && ro < i1)) {
                const int ko = (tt - NS + 2 * W) % W;
                w_stq(ug + (long long)ro * n, col0, so[0], so[1], val[ko]);
                if (want_dot)
                    dot = dot + w_sel(so[0], w_cvt<double>(val[ko][0]) * w_cvt<double>(bb[ko][0]) + w_cvt<double>(val[ko][1]) * w_cvt<double>(bb[ko][1]), zerod) +
                          w_sel(so[1], w_cvt<double>(val[ko][2]) * w_cvt<double>(bb[ko][2]) + w_cvt<double>(val[ko][3]) * w_cvt<double>(bb[ko][3]), zerod);
            }
        }
    }
}

template <int NU, class TB, class TV, class Feeder>
PSN_FN void psn_strip_post_body(const PsnWarp& w, const PsnPlan& plan, const PsnStripArgs& a, long long b, int band, int strip, Feeder& fd) {
    if (a.state && a.state[b].phase != PSN_PHASE_ITER) return;
    constexpr int NS = 2 * NU, W = NS + 2;
    constexpr int F32 = sizeof(TV) == 4 ? 1 : 0;
    constexpr bool SCALE = F32 && sizeof(TB) == 8;
    const int USE = psn_strip_use_post(NU, F32);
    const int HL = psn_strip_hl_post(NU, F32);
    const PsnLevel& lv = plan.lv[a.level];
    const PsnLevel& lc = plan.lv[a.level + 1];
    const int n = lv.n, nc = lc.n, last = n - 1;
    const int a0 = strip * USE, a1 = psn_imin(a0 + USE, n);
    const int c0 = a0 - HL;       // even; at least NS + 2 halo columns on the left, NS on the right
    const int i0 = band * a.band, i1 = psn_imin(i0 + a.band, n);

    const wi col0 = w_lane(w) * 4 + c0;
    const wi col2 = col0 + 2;
    const wb mp[2] = {(col0 >= 0) && (col0 < n), (col2 >= 0) && (col2 < n)};
    const wb so[2] = {mp[0] && (col0 >= a0) && (col0 < a1), mp[1] && (col2 >= a0) && (col2 < a1)};
    const wi cj = w_lane(w) * 2 + (c0 >> 1);
    const wb mc[2] = {(cj >= 0) && (cj < nc), (cj + 1 >= 0) && (cj + 1 < nc)};
    typedef typename PsnWV<TV>::type WV;
    const WV zero = w_bcastT<TV>(0.0);
    WV rdA[4], mk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const wi ce = col0 + e;
        const wb in = (e < 2) ? mp[0] : mp[1];
        const wb lc_ = ce == last;
        rdA[e] = w_sel(in, w_sel(lc_, w_bcastT<TV>(lv.rd1), w_bcastT<TV>(lv.rd0)), zero);
        mk[e] = w_sel(in, w_bcastT<TV>(1.0), zero);
    }

    const int ts = psn_imax(i0 - NS, 0);
    const int tload = psn_imin(i1 - 1 + NS, last);
    const int tend = psn_imin(i1 - 1, last) + NS;
    const TB* bg = reinterpret_cast<const TB*>(a.b) + b * (long long)n * n;
    const TV* uig = reinterpret_cast<const TV*>(a.uin) + b * (long long)n * n;
    const TV* eg = reinterpret_cast<const TV*>(a.ec) + b * (long long)nc * nc;
    TV* ug = reinterpret_cast<TV*>(a.uout) + b * (long long)n * n;
    const double bs = (SCALE && a.state) ? psn_vcycle_scale(a.state[b], a.fused) : 1.0;
    const wd bsc = w_bcast(bs);

    WV val[W][4], bb[W][4], eprev[2], enext[4];
#pragma unroll
    for (int k = 0; k < W; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) val[k][e] = bb[k][e] = zero;
    fd.set_base0(bg);
    fd.set_base1(uig);
    fd.n = n;
    fd.tload = tload;
    psn_feeder_geom(fd, col0, mp, c0, n);
    fd.prime(w, ts);
    // coarse correction rows: eprev = row ts/2 - 1 (zero above the grid), enext = rows ts/2 and ts/2 + 1
    eprev[0] = eprev[1] = enext[0] = enext[1] = enext[2] = enext[3] = zero;
    wd dot = w_bcast(0.0);
    {
        const int Ip = (ts >> 1) - 1;
        if (Ip >= 0 && Ip < nc) {
            eprev[0] = w_ld1(eg + (long long)Ip * nc, cj, mc[0]);
            eprev[1] = w_ld1(eg + (long long)Ip * nc, cj + 1, mc[1]);
        }
        if (Ip + 1 < nc) {
            enext[0] = w_ld1(eg + (long long)(Ip + 1) * nc, cj, mc[0]);
            enext[1] = w_ld1(eg + (long long)(Ip + 1) * nc, cj + 1, mc[1]);
        }
        if (Ip + 2 < nc) {
            enext[2] = w_ld1(eg + (long long)(Ip + 2) * nc, cj, mc[0]);
            enext[3] = w_ld1(eg + (long long)(Ip + 2) * nc, cj + 1, mc[1]);
        }
    }
    const bool want_dot = a.part != nullptr;

    for (int t = ts; t <= tend; t += W) {
        const bool fast = (t - NS - 1 >= ts) && (t + W - 1 + (W - 1) <= tload) && (t + W - 1 < last) && (((t + W - 1) >> 1) + 2 < nc) &&
                          (t - NS >= i0) && (t + W - 1 - NS < i1);
        const int par = ((t - ts) / W) & 1;
        if (fast)
            psn_strip_post_chunk<NU, true, SCALE, TV>(w, t, ts, tload, tend, last, n, nc, i0, i1, val, bb, fd, par, eprev, enext, dot, rdA, mk, lv,
                                                  bsc, eg, ug, col0, mp, so, cj, mc, want_dot);
        else
            psn_strip_post_chunk<NU, false, SCALE, TV>(w, t, ts, tload, tend, last, n, nc, i0, i1, val, bb, fd, par, eprev, enext, dot, rdA, mk, lv,
                                                   bsc, eg, ug, col0, mp, so, cj, mc, want_dot);
    }
    if (want_dot) {
        // rhs and iterate both carry the scale: <r, z> = <s r, s z> / s^2 (exact, s is a power of two)
        const double s = w_sum(dot) * ((1.0 / bs) * (1.0 / bs));
        if (w_lane0(w)) a.part[b * (long long)(a.nstrips * a.nbands) + (long long)band * a.nstrips + strip] = s;
    }
}


// ----------------------------------------------------------------------------- direction kernel, strip form
// p_new = z + beta p_old;  q = At p_new;  partial <p_new, q>  -- one warp marches down a 128-column strip
// (2 halo columns per side) with a 3-row window of p_new in registers; z and p_old rows arrive through
// the same feeders as the V-cycle strips.
#define PSN_DIR_USE (PSN_STRIP_COLS - 4)
PSN_HDFN int psn_dir_hl(int f32) { return f32 ? 4 : 2; }
PSN_HDFN int psn_dir_use(int f32) { return f32 ? PSN_STRIP_COLS - 8 : PSN_DIR_USE; }

PSN_FN double psn_warp_sum_partials(const PsnWarp& w, const double* part, int count) {
    wd acc = w_bcast(0.0);
    const wi lane = w_lane(w);
    for (int k0 = 0; k0 < count; k0 += 32) acc = acc + w_ld1(part + k0, lane, lane < (count - k0));
    return w_sum(acc);
}

// TZ: storage type of z (float: the value read is s z with s = psn_vscale(||r||^2), undone here)
// TP: storage type of the search direction.  float (needs TZ = float and the strip update kernel): p is kept as the
//     float s_k p_k with s_k = psn_vscale(||r_k||^2); CG stays exact for the ROUNDED direction because q = At p, <p,q>,
//     x += alpha p and r -= alpha q are all formed from that same rounded p in fp64 (rounding p perturbs the search
//     direction like the float-stored preconditioner does, not the solution: same iteration counts, tests).
template <class TZ, class TP, class Feeder>
PSN_FN void psn_dir_strip_body(const PsnWarp& w, const PsnCgArgs& a, long long b, int band, int strip, Feeder& fd) {
    const PsnSampleState st = a.state[b];
    if (st.phase != PSN_PHASE_ITER) return;
    constexpr int W = Feeder::W;
    static_assert(W % 3 == 0, "the 3-row window rotates inside one ring period");
    const double rz = psn_warp_sum_partials(w, a.part_rz + b * (long long)a.ntiles_rz, a.ntiles_rz);
    const bool fresh = st.restart != 0;
    const wd beta = w_bcast(fresh ? 0.0 : rz / st.rz_old);
    const int n = a.n, last = n - 1;
    constexpr int F32 = sizeof(TZ) == 4 ? 1 : 0;
    constexpr bool PF32 = sizeof(TP) == 4;
    static_assert(!PF32 || F32, "a float search direction needs the float-stored (scaled) z");
    const int a0 = strip * psn_dir_use(F32), a1 = psn_imin(a0 + psn_dir_use(F32), n);
    const int c0 = a0 - psn_dir_hl(F32);
    const double sk = F32 ? psn_vscale(st.rr) : 1.0;                 // scale of the direction stored by THIS launch
    const double sz = F32 ? psn_vcycle_scale(st, a.fused) : 1.0;     // scale z was computed and stored with (== sk unless fused)
    const wd zinv = w_bcast(1.0 / sz);
    const wd zr = w_bcast(sk / sz);                                  // power of two (1 when unfused): exact
    // scaled domain: s_k p_k = s_k z_k + beta (s_k / s_{k-1}) (s_{k-1} p_{k-1})   (powers of two: exact)
    const wd beta_s = w_bcast((fresh || !PF32) ? 0.0 : (rz / st.rz_old) * (sk / psn_vscale(st.rr_prev)));
    const int i0 = band * a.dir_band, i1 = psn_imin(i0 + a.dir_band, n);
    const wi col0 = w_lane(w) * 4 + c0;
    const wi col2 = col0 + 2;
    const wb mp[2] = {(col0 >= 0) && (col0 < n), (col2 >= 0) && (col2 < n)};
    const wb so[2] = {mp[0] && (col0 >= a0) && (col0 < a1), mp[1] && (col2 >= a0) && (col2 < a1)};
    const long long base = b * (long long)n * n;
    const wd zero = w_bcast(0.0);

    const int ts = psn_imax(i0 - 1, 0);
    const int tload = psn_imin(i1, last);
    const int tend = i1;   // q(row r) is formed at time r + 1; row n reads as zero
    fd.set_base0(reinterpret_cast<const TZ*>(a.z) + base);
    fd.set_base1(reinterpret_cast<const TP*>(a.p_old) + base);
    fd.n = n;
    fd.tload = tload;
    psn_feeder_geom(fd, col0, mp, c0, n);
    fd.prime(w, ts);

    wd pn[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) pn[k][e] = zero;
    wd dot = zero;
    TP* pg = reinterpret_cast<TP*>(a.p_new) + base;
    double* qg = a.q + base;

    for (int t = ts; t <= tend; t += W) {
        const int par = ((t - ts) / W) & 1;
#pragma unroll
        for (int tt = 0; tt < W; ++tt) {
            const int tc = t + tt;
            if (tc > tend) break;
            const int k0 = tt % 3, k1 = (tt + 2) % 3, k2 = (tt + 1) % 3;   // rows tc, tc-1, tc-2
            typename Feeder::V0 r0[4];
            typename Feeder::V1 r1[4];
            fd.template fetch<false>(w, tt, tc, par, r0, r1);
            if constexpr (PF32) {
                typename PsnWV<TP>::type pr[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const wd zv = w_cvt<double>(r0[e]) * zr;
                    pr[e] = w_cvt<TP>(fresh ? zv : zv + beta_s * w_cvt<double>(r1[e]));
                    pn[k0][e] = w_cvt<double>(pr[e]);   // everything below sees the rounded direction
                }
                if (tc >= i0 && tc < i1) w_stq(pg + (long long)tc * n, col0, so[0], so[1], pr);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    wd zv = w_cvt<double>(r0[e]);
                    if constexpr (F32 != 0) zv = zv * zinv;
                    pn[k0][e] = fresh ? zv : zv + beta * w_cvt<double>(r1[e]);
                }
                if (tc >= i0 && tc < i1) {
                    TP* rowp = pg + (long long)tc * n;
                    w_stp(rowp, col0, so[0], pn[k0][0], pn[k0][1]);
                    w_stp(rowp, col2, so[1], pn[k0][2], pn[k0][3]);
                }
            }
            const int r = tc - 1;
            if (r >= i0 && r < i1) {
                const wd upv = w_up_raw(pn[k1][3]);
                const wd dnv = w_down_raw(pn[k1][0]);
                wd q[4];
                q[0] = 4.0 * pn[k1][0] - ((pn[k0][0] + pn[k2][0]) + (upv + pn[k1][1]));
                q[1] = 4.0 * pn[k1][1] - ((pn[k0][1] + pn[k2][1]) + (pn[k1][0] + pn[k1][2]));
                q[2] = 4.0 * pn[k1][2] - ((pn[k0][2] + pn[k2][2]) + (pn[k1][1] + pn[k1][3]));
                q[3] = 4.0 * pn[k1][3] - ((pn[k0][3] + pn[k2][3]) + (pn[k1][2] + dnv));
                if (a.upd_mode == 0) {   // the strip update kernel recomputes q = At p from p instead of reading it back
                    double* rowq = qg + (long long)r * n;
                    w_stp(rowq, col0, so[0], q[0], q[1]);
                    w_stp(rowq, col2, so[1], q[2], q[3]);
                }
                dot = dot + w_sel(so[0], pn[k1][0] * q[0] + pn[k1][1] * q[1], zero) + w_sel(so[1], pn[k1][2] * q[2] + pn[k1][3] * q[3], zero);
            }
        }
    }
    // with a float direction p and q carry the scale s_k: <p,q> = <s p, s q> / s^2
    const double s = PF32 ? w_sum(dot) * ((1.0 / sk) * (1.0 / sk)) : w_sum(dot);
    if (w_lane0(w)) {
        a.part_pq[b * (long long)a.npq + (long long)band * a.dir_nstrips + strip] = s;
        if (band == 0 && strip == 0) a.state[b].rz = rz;
    }
}


// ----------------------------------------------------------------------------- CG update, strip form
// x += alpha p;  r -= alpha q with q = At p RECOMPUTED from a 3-row window of p (identical expression to the direction
// kernel, so r is bit-identical to the stored-q variant);  partial ||r||^2.  Not storing q saves its 8-byte write in
// the direction kernel and its 8-byte read here: 76 -> 60 bytes per node and iteration for the two kernels together.
// One warp per 128-column strip (2 halo columns for the stencil), p / x / r rows through a 3-array TMA ring.
// feeder of the strip update kernel: p, x, r -- or p, r when x is deferred
template <class TP, bool XDEFER>
struct PsnUpdateFeederSel {
    typedef PsnTmaFeeder<2, TP, double, double> type;
};
template <class TP>
struct PsnUpdateFeederSel<TP, true> {
    typedef PsnTmaFeeder<2, TP, double> type;
};

// TP: storage type of p (float: s_k p_k, see psn_dir_strip_body; x and r are always fp64)
// XDEFER: x is not touched; the step length is recorded in PsnPend instead (feeder streams p and r only)
template <class TP, bool XDEFER, class Feeder>
PSN_FN void psn_update_strip_body(const PsnWarp& w, const PsnCgArgs& a, long long b, int band, int strip, Feeder& fd) {
    const PsnSampleState st = a.state[b];
    if (st.phase != PSN_PHASE_ITER) return;
    constexpr int W = Feeder::W;
    static_assert(W % 3 == 0, "the 3-row window rotates inside one ring period");
    constexpr int PF32 = sizeof(TP) == 4 ? 1 : 0;
    const double pq = psn_warp_sum_partials(w, a.part_pq + b * (long long)a.npq, a.npq);
    // with a float direction the window holds s_k p_k and s_k q_k: fold 1 / s_k into alpha (power of two, exact)
    const wd alpha = w_bcast(((pq > 0.0) ? st.rz / pq : 0.0) * (PF32 ? 1.0 / psn_vscale(st.rr) : 1.0));
    const int n = a.n, last = n - 1;
    const int a0 = strip * psn_dir_use(PF32), a1 = psn_imin(a0 + psn_dir_use(PF32), n);
    const int c0 = a0 - psn_dir_hl(PF32);
    const int i0 = band * a.upd_band, i1 = psn_imin(i0 + a.upd_band, n);
    const wi col0 = w_lane(w) * 4 + c0;
    const wi col2 = col0 + 2;
    const wb mp[2] = {(col0 >= 0) && (col0 < n), (col2 >= 0) && (col2 < n)};
    const wb so[2] = {mp[0] && (col0 >= a0) && (col0 < a1), mp[1] && (col2 >= a0) && (col2 < a1)};
    const long long base = b * (long long)n * n;
    const wd zero = w_bcast(0.0);

    const int ts = psn_imax(i0 - 1, 0);
    const int tload = psn_imin(i1, last);
    const int tend = i1;   // row r is finished at time r + 1; row n reads as zero
    fd.set_base0(reinterpret_cast<const TP*>(a.p_new) + base);
    if constexpr (XDEFER) {
        fd.set_base1(a.r + base);
    } else {
        fd.set_base1(a.x + base);
        fd.set_base2(a.r + base);
    }
    fd.n = n;
    fd.tload = tload;
    psn_feeder_geom(fd, col0, mp, c0, n);
    fd.prime(w, ts);

    wd pn[3][4], xp[4], rp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        xp[e] = rp[e] = zero;
#pragma unroll
        for (int k = 0; k < 3; ++k) pn[k][e] = zero;
    }
    wd acc = zero;
    double* xg = a.x + base;
    double* rg = a.r + base;

    for (int t = ts; t <= tend; t += W) {
        const int par = ((t - ts) / W) & 1;
#pragma unroll
        for (int tt = 0; tt < W; ++tt) {
            const int tc = t + tt;
            if (tc > tend) break;
            const int k0 = tt % 3, k1 = (tt + 2) % 3, k2 = (tt + 1) % 3;   // rows tc, tc-1, tc-2
            wd xc[4], rc[4];
            typename Feeder::V0 pr[4];
            if constexpr (XDEFER) fd.template fetch<false>(w, tt, tc, par, pr, rc);
            else fd.template fetch3<false>(w, tt, tc, par, pr, xc, rc);
#pragma unroll
            for (int e = 0; e < 4; ++e) pn[k0][e] = w_cvt<double>(pr[e]);
            const int r = tc - 1;
            if (r >= i0 && r < i1) {
                const wd upv = w_up_raw(pn[k1][3]);
                const wd dnv = w_down_raw(pn[k1][0]);
                wd q[4];
                q[0] = 4.0 * pn[k1][0] - ((pn[k0][0] + pn[k2][0]) + (upv + pn[k1][1]));
                q[1] = 4.0 * pn[k1][1] - ((pn[k0][1] + pn[k2][1]) + (pn[k1][0] + pn[k1][2]));
                q[2] = 4.0 * pn[k1][2] - ((pn[k0][2] + pn[k2][2]) + (pn[k1][1] + pn[k1][3]));
                q[3] = 4.0 * pn[k1][3] - ((pn[k0][3] + pn[k2][3]) + (pn[k1][2] + dnv));
                wd rn[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) rn[e] = rp[e] - alpha * q[e];
                if constexpr (!XDEFER) {
                    wd xn[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) xn[e] = xp[e] + alpha * pn[k1][e];
                    double* rowx = xg + (long long)r * n;
                    w_stp(rowx, col0, so[0], xn[0], xn[1]);
                    w_stp(rowx, col2, so[1], xn[2], xn[3]);
                }
                double* rowr = rg + (long long)r * n;
                w_stp(rowr, col0, so[0], rn[0], rn[1]);
                w_stp(rowr, col2, so[1], rn[2], rn[3]);
                acc = acc + w_sel(so[0], rn[0] * rn[0] + rn[1] * rn[1], zero) + w_sel(so[1], rn[2] * rn[2] + rn[3] * rn[3], zero);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if constexpr (!XDEFER) xp[e] = xc[e];
                rp[e] = rc[e];
            }
        }
    }
    const double s = w_sum(acc);
    if (w_lane0(w)) {
        a.part_u[b * (long long)(a.upd_nstrips * a.upd_nbands) + (long long)band * a.upd_nstrips + strip] = s;
        if (band == 0 && strip == 0) {
            a.state[b].pq = pq;
            if constexpr (XDEFER) {
                // record the step: x will receive (alpha / s_k) * (s_k p_k) when this sample is flushed
                PsnPend* pd = a.pend + b;
                if (pd->lo == pd->hi) pd->lo = a.it;
                pd->a[a.it % PSN_PEND_K] = ((pq > 0.0) ? st.rz / pq : 0.0) * (1.0 / psn_vscale(st.rr));
                pd->hi = a.it + 1;
            }
        }
    }
}



// ----------------------------------------------------------------------------- fused CG update + level-0 down-leg
// The update kernel ends with r_{k+1} = r_k - alpha A p_k written to HBM and the down-leg of the next V-cycle starts by
// reading it back: 8 of the 67 bytes per node and iteration.  Here one pass does both: the strip marches down with a
// 3-row window of p_k (float ring slot, converted to fp64) one row AHEAD of the smoother's newest row, forms
// q = At p, r_{k+1} (fp64, stored for the owned part), its partial ||r||^2, and hands the row -- scaled and rounded to
// float exactly like the plain down-leg does on load -- to the temporally blocked sweeps.  r is double-buffered
// (PsnCgArgs::r -> r_out): neighbouring tasks still read r_k in their halos.  One more halo column per side than the
// plain strips (q at the strip edge needs p one column further out): 112 useful columns for nu = 2.
// Level 0 of the float-stored configuration only (rhs fp64, iterate float, float search direction, deferred x).
PSN_HDFN int psn_fused_hl(int nu) { return (2 * nu + 1 + 3) & ~3; }
PSN_HDFN int psn_fused_use(int nu) { return (PSN_STRIP_COLS - psn_fused_hl(nu) - (2 * nu + 2)) & ~3; }

struct PsnFusedRow {     // per-task constants and running state of the update part
    wd alpha;            // alpha_k / s_k (the ring slot holds s_k p_k)
    wd bsc;              // V-cycle scale of the NEXT iteration (psn_vcycle_scale as iteration k+1 will see it)
    wd pm[4], pc[4];     // p rows tc-1, tc (fp64 copies of the float slot)
    wd dot;              // partial ||r_{k+1}||^2 over the owned nodes
    double* rout;        // r_{k+1} of this sample
    wb ind[2];           // column pairs inside the domain
    bool edge;           // the strip window leaves the domain on a side (warp-uniform): only then are the selects on `ind` needed
    wd mo[2];            // 1.0 for the column pairs this task owns, else 0.0 (dot-product mask)
    // mode 1 (prologue): r_0 = bt = -(sin_y[i] sin_x[j] / theta_ij) h^2 instead of r_k - alpha q
    wd sxv[4];           // the lane's four sin_x values
    const double* sy;    // sin_y of this sample
    wd h2;
};

template <int NU, int MODE, bool FAST, class Feeder>
PSN_FN void psn_strip_upre_chunk(const PsnWarp& w, int t, int ts, int tload, int tend, int last, int n, int nc,
                                 int i0, int i1, int I0, int I1, wf (&val)[2 * NU + 2][4], wf (&bb)[2 * NU + 2][4],
                                 Feeder& fd, int par, wf (&acc)[2], const wf (&rdA)[4], const wf (&dA)[4], const wf (&mk)[4],
                                 const PsnLevel& lv, PsnFusedRow& fr, float* ug, float* cg, const wi& col0, const wb (&so)[2],
                                 const wi& cj, const wb (&sc)[2]) {
    constexpr int NS = 2 * NU, W = NS + 2;
    const wf zero = w_bcastT<float>(0.0);
    const wf q25 = w_bcastT<float>(0.25);
    const wd zerod = w_bcast(0.0);
    wf rdB[4], dB[4];
    if (!FAST) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const wb lc_ = (col0 + e) == last;
            rdB[e] = mk[e] * w_sel(lc_, w_bcastT<float>(lv.rd2), w_bcastT<float>(lv.rd1));
            dB[e] = w_sel(lc_, w_bcastT<float>(4.0 + 2.0 * lv.gam), w_bcastT<float>(4.0 + lv.gam));
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) rdB[e] = dB[e] = zero;
    }
#pragma unroll
    for (int tt = 0; tt < W; ++tt) {
        const int tc = t + tt;
        if (!FAST && tc > tend) break;
        const int P = tt & 1;
        const int k0 = tt;
        // ---- new row.  mode 0: r_k(tc) and p_k(tc+1) from the feeder; q(tc), r_{k+1}(tc).  mode 1: theta(tc) from the feeder,
        //      r_0(tc) = bt(tc).  Then the first red half sweep from the zero guess.
        {
            wd rn[4];
            typename Feeder::V0 r0[4];
            typename Feeder::V1 r1[4];
            fd.template fetch<FAST>(w, tt, tc, par, r0, r1);
            if constexpr (MODE == 0) {
                wd pn[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) pn[e] = w_cvt<double>(r1[e]);
                if (!FAST && tc + 1 > last) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) pn[e] = zerod;          // row n of p: outside the grid (its copy was skipped)
                }
                const wd upv = w_up_raw(fr.pc[3]);
                const wd dnv = w_down_raw(fr.pc[0]);
                wd q[4];
                q[0] = 4.0 * fr.pc[0] - ((pn[0] + fr.pm[0]) + (upv + fr.pc[1]));
                q[1] = 4.0 * fr.pc[1] - ((pn[1] + fr.pm[1]) + (fr.pc[0] + fr.pc[2]));
                q[2] = 4.0 * fr.pc[2] - ((pn[2] + fr.pm[2]) + (fr.pc[1] + fr.pc[3]));
                q[3] = 4.0 * fr.pc[3] - ((pn[3] + fr.pm[3]) + (fr.pc[2] + dnv));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    rn[e] = r0[e] - fr.alpha * q[e];
                    fr.pm[e] = fr.pc[e];
                    fr.pc[e] = pn[e];
                }
                if (fr.edge) {   // out-of-domain columns next to the domain would carry -alpha q != 0
#pragma unroll
                    for (int e = 0; e < 4; ++e) rn[e] = w_sel(fr.ind[e >> 1], rn[e], zerod);
                }
            } else {
                if (FAST || tc <= last) {
                    const wd syr = w_bcast(fr.sy[tc]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // out-of-domain lanes divide by 1 (the ring's un-copied columns are zero) and are masked
                        const wd th = w_sel(fr.ind[e >> 1], w_cvt<double>(r0[e]), w_bcast(1.0));
                        const wd f = (syr * fr.sxv[e]) * w_rcp(th);          // same expression as psn_rhs
                        rn[e] = w_sel(fr.ind[e >> 1], zerod - (f * fr.h2), zerod);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) rn[e] = zerod;
                }
            }
            if (tc >= i0 && tc < i1) {
                double* rowr = fr.rout + (long long)tc * n;
                w_stp(rowr, col0, so[0], rn[0], rn[1]);
                w_stp(rowr, col0 + 2, so[1], rn[2], rn[3]);
                fr.dot = fr.dot + fr.mo[0] * (rn[0] * rn[0] + rn[1] * rn[1]) + fr.mo[1] * (rn[2] * rn[2] + rn[3] * rn[3]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) bb[k0][e] = w_cvt<float>(rn[e] * fr.bsc);
        }
        if (FAST || tc <= last) {
            const bool lastrow = !FAST && tc == last;
            val[k0][P] = bb[k0][P] * (lastrow ? rdB[P] : rdA[P]);
            val[k0][P + 2] = bb[k0][P + 2] * (lastrow ? rdB[P + 2] : rdA[P + 2]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) bb[k0][e] = zero;
            val[k0][P] = val[k0][P + 2] = zero;
        }
        val[k0][1 - P] = val[k0][3 - P] = zero;
        // ---- smoothing stages 1 .. NS-1 on rows tc-1 .. tc-NS+1
#pragma unroll
        for (int s = 1; s < NS; ++s) {
            const int r = tc - s;
            if (FAST || (r >= ts && r <= last)) {
                const int ks = (tt - s + 2 * W) % W, ku = (tt - s - 1 + 2 * W) % W, kd = (tt - s + 1 + 2 * W) % W;
                const bool lastrow = !FAST && r == last;
                const wf upf = (P == 0) ? w_up_raw(val[ks][3]) : zero;
                const wf dnf = (P == 1) ? w_down_raw(val[ks][0]) : zero;
                const wf sa = (val[ku][P] + val[kd][P]) + PSN_LR(val[ks], P, upf, dnf);
                const wf sb = (val[ku][P + 2] + val[kd][P + 2]) + PSN_LR(val[ks], P + 2, upf, dnf);
                val[ks][P] = (bb[ks][P] + sa) * (lastrow ? rdB[P] : rdA[P]);
                val[ks][P + 2] = (bb[ks][P + 2] + sb) * (lastrow ? rdB[P + 2] : rdA[P + 2]);
            }
        }
        // ---- residual on the red nodes of row tc-NS, accumulate the restriction
        {
            const int r = tc - NS;
            wf ra = zero, rb = zero;
            if (FAST || (r >= ts && r <= last)) {
                const int ks = (tt - NS + 2 * W) % W, ku = (tt - NS - 1 + 2 * W) % W, kd = (tt - NS + 1 + 2 * W) % W;
                const bool lastrow = !FAST && r == last;
                const wf upf = (P == 0) ? w_up_raw(val[ks][3]) : zero;
                const wf dnf = (P == 1) ? w_down_raw(val[ks][0]) : zero;
                const wf sa = (val[ku][P] + val[kd][P]) + PSN_LR(val[ks], P, upf, dnf);
                const wf sb = (val[ku][P + 2] + val[kd][P + 2]) + PSN_LR(val[ks], P + 2, upf, dnf);
                ra = mk[P] * (bb[ks][P] - ((lastrow ? dB[P] : dA[P]) * val[ks][P] - sa));
                rb = mk[P + 2] * (bb[ks][P + 2] - ((lastrow ? dB[P + 2] : dA[P + 2]) * val[ks][P + 2] - sb));
            }
            if (P == 0) {
                const wf c0v = q25 * (ra + rb);
                const wf c1v = q25 * (rb + w_down_raw(ra));
                const int I = (r >> 1) - 1;
                if (FAST || (r >= 2 && I >= I0 && I < I1)) {
                    float* crow = cg + (long long)I * nc;
                    w_st1(crow, cj, sc[0], acc[0] + c0v);
                    w_st1(crow, cj + 1, sc[1], acc[1] + c1v);
                }
                acc[0] = c0v;
                acc[1] = c1v;
            } else {
                acc[0] = acc[0] + ra;
                acc[1] = acc[1] + rb;
            }
        }
        // ---- row tc-NS+1 is final after its last black half sweep
        {
            const int ro = tc - (NS - 1);
            if (FAST || (ro >= i0 && ro < i1)) {
                const int ko = (tt - (NS - 1) + 2 * W) % W;
                w_stq(ug + (long long)ro * n, col0, so[0], so[1], val[ko]);
            }
        }
    }
}

template <int NU, int MODE>
struct PsnUpreFeederSel {
    typedef PsnTmaFeeder<NU, double, float> type;
};
template <int NU>
struct PsnUpreFeederSel<NU, 1> {
    typedef PsnTmaFeeder<NU, double> type;
};

// sum of squares of a vector, one warp, fixed order (every task of a sample gets the identical value)
PSN_FN double psn_warp_sum_sq(const PsnWarp& w, const double* v, int count) {
    wd acc = w_bcast(0.0);
    const wi lane = w_lane(w);
    for (int k0 = 0; k0 < count; k0 += 32) {
        const wd x = w_ld1(v + k0, lane, lane < (count - k0));
        acc = acc + x * x;
    }
    return w_sum(acc);
}

// MODE 0 -- Feeder: PsnTmaFeeder<NU, double /* r_k */, float /* s_k p_k, rows offset by one */>
// MODE 1 -- the solve's PROLOGUE fused with the first down-leg: r_0 = bt formed from theta (Feeder: PsnTmaFeeder<NU, double>)
//           and the sine factors; the sample state does not exist yet (state_init runs afterwards and finds ||bt||^2 in
//           part_u and the scale used here in vs0).
template <int NU, int MODE, class Feeder>
PSN_FN void psn_strip_upre_body(const PsnWarp& w, const PsnPlan& plan, const PsnStripArgs& a, const PsnCgArgs& g, long long b,
                                int band, int strip, Feeder& fd) {
    PsnSampleState st;
    memset(&st, 0, sizeof(st));
    if constexpr (MODE == 0) {
        st = g.state[b];
        if (st.phase != PSN_PHASE_ITER) return;
    }
    constexpr int NS = 2 * NU, W = NS + 2;
    const int USE = psn_fused_use(NU);
    const int HL = psn_fused_hl(NU);
    const PsnLevel& lv = plan.lv[0];
    const PsnLevel& lc = plan.lv[1];
    const int n = lv.n, nc = lc.n, last = n - 1;
    const int a0 = strip * USE, a1 = psn_imin(a0 + USE, n);
    const int c0 = a0 - HL;
    const int i0 = band * a.band, i1 = psn_imin(i0 + a.band, n);
    const int I0 = i0 >> 1, I1 = psn_imin(i1 >> 1, nc);
    const int J0 = a0 >> 1, J1 = psn_imin(a1 >> 1, nc);

    double pq = 0.0, pscale = 1.0, alpha = 0.0;
    if constexpr (MODE == 0) {
        pq = psn_warp_sum_partials(w, g.part_pq + b * (long long)g.npq, g.npq);
        pscale = psn_vscale(st.rr);                                // scale of the stored direction s_k p_k
        alpha = (pq > 0.0) ? st.rz / pq : 0.0;
    } else {
        // ||bt||^2 does not exist yet: scale with ||sin_x||^2 ||sin_y||^2 h^4, which is ||bt||^2 for theta == 1 and within
        // max(theta)^2 / min(theta)^2 of it otherwise -- it only positions the data in the float exponent range
        const double sxx = psn_warp_sum_sq(w, g.rhs.sx + b * (long long)n, n);
        const double syy = psn_warp_sum_sq(w, g.rhs.sy + b * (long long)n, n);
        pscale = psn_vscale(sxx * syy * g.h2 * g.h2);
    }

    const wi col0 = w_lane(w) * 4 + c0;
    const wi col2 = col0 + 2;
    const wb mp[2] = {(col0 >= 0) && (col0 < n), (col2 >= 0) && (col2 < n)};
    const wb so[2] = {mp[0] && (col0 >= a0) && (col0 < a1), mp[1] && (col2 >= a0) && (col2 < a1)};
    const wi cj = w_lane(w) * 2 + (c0 >> 1);
    const wb sc[2] = {(cj >= J0) && (cj < J1), (cj + 1 >= J0) && (cj + 1 < J1)};
    const wf zero = w_bcastT<float>(0.0);
    const wd zerod = w_bcast(0.0);
    wf rdA[4], dA[4], mk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const wi ce = col0 + e;
        const wb in = (e < 2) ? mp[0] : mp[1];
        const wb lc_ = ce == last;
        rdA[e] = w_sel(in, w_sel(lc_, w_bcastT<float>(lv.rd1), w_bcastT<float>(lv.rd0)), zero);
        dA[e] = w_sel(lc_, w_bcastT<float>(4.0 + lv.gam), w_bcastT<float>(4.0));
        mk[e] = w_sel(in, w_bcastT<float>(1.0), zero);
    }

    const int ts = psn_imax(i0 - NS, 0);
    const int tload = psn_imin(i1 + NS, last);
    const int tend = i1 + NS;
    const long long base = b * (long long)n * n;
    const float* pg = reinterpret_cast<const float*>(g.p_new) + base;
    float* ug = reinterpret_cast<float*>(a.uout) + base;
    float* cgp = reinterpret_cast<float*>(a.bc) + b * (long long)nc * nc;

    PsnFusedRow fr;
    fr.alpha = w_bcast(alpha * (1.0 / pscale));
    fr.bsc = w_bcast(pscale);              // = psn_vscale(||r_k||^2): what iteration k+1 finds in rr_prev
    fr.dot = zerod;
    fr.rout = g.r_out + base;
    fr.ind[0] = mp[0];
    fr.ind[1] = mp[1];
    fr.edge = (c0 < 0) || (c0 + PSN_STRIP_COLS > n);
    fr.mo[0] = w_sel(so[0], w_bcast(1.0), zerod);
    fr.mo[1] = w_sel(so[1], w_bcast(1.0), zerod);
    fr.h2 = w_bcast(g.h2);
    fr.sy = nullptr;
#pragma unroll
    for (int e = 0; e < 4; ++e) fr.pm[e] = fr.pc[e] = fr.sxv[e] = zerod;
    if constexpr (MODE == 1) {
        fr.sy = g.rhs.sy + b * (long long)n;
        w_ldp(g.rhs.sx + b * (long long)n, col0, mp[0], fr.sxv[0], fr.sxv[1]);
        w_ldp(g.rhs.sx + b * (long long)n, col2, mp[1], fr.sxv[2], fr.sxv[3]);
    }
    // p rows ts-1 and ts (the feeder delivers row tc+1 at time tc)
    if constexpr (MODE == 0) {
        wf t0[4], t1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) t0[e] = t1[e] = zero;
        if (ts >= 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) t0[e] = w_ld1(pg + (long long)(ts - 1) * n, col0 + e, (e < 2) ? mp[0] : mp[1]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) t1[e] = w_ld1(pg + (long long)ts * n, col0 + e, (e < 2) ? mp[0] : mp[1]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            fr.pm[e] = w_cvt<double>(t0[e]);
            fr.pc[e] = w_cvt<double>(t1[e]);
        }
    }

    wf val[W][4], bb[W][4], acc[2];
#pragma unroll
    for (int k = 0; k < W; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) val[k][e] = bb[k][e] = zero;
    if constexpr (MODE == 0) {
        fd.set_base0(g.r + base);
        fd.set_base1(pg + n);              // array 1 runs one row ahead
        fd.rows1_max = last - 1;           // ... and has no row n
    } else {
        fd.set_base0(g.rhs.theta + base);
    }
    fd.n = n;
    fd.tload = tload;
    psn_feeder_geom(fd, col0, mp, c0, n);
    fd.prime(w, ts);
    acc[0] = acc[1] = zero;

    for (int t = ts; t <= tend; t += W) {
        const bool fast = (t - NS - 1 >= ts) && (t - NS >= 2) && (t + W - 1 + (W - 1) <= tload) && (t + W - 1 < last) &&
                          (t - (NS - 1) >= i0) && (t + W - 1 - (NS - 1) < i1) &&
                          (((t - NS) >> 1) - 1 >= I0) && (((t + W - 1 - NS) >> 1) - 1 < I1);
        const int par = ((t - ts) / W) & 1;
        if (fast)
            psn_strip_upre_chunk<NU, MODE, true>(w, t, ts, tload, tend, last, n, nc, i0, i1, I0, I1, val, bb, fd, par, acc, rdA, dA, mk, lv, fr,
                                           ug, cgp, col0, so, cj, sc);
        else
            psn_strip_upre_chunk<NU, MODE, false>(w, t, ts, tload, tend, last, n, nc, i0, i1, I0, I1, val, bb, fd, par, acc, rdA, dA, mk, lv, fr,
                                            ug, cgp, col0, so, cj, sc);
    }
    const double sdot = w_sum(fr.dot);
    if (w_lane0(w)) {
        g.part_u[b * (long long)(a.nstrips * a.nbands) + (long long)band * a.nstrips + strip] = sdot;
        if (band == 0 && strip == 0) {
            if constexpr (MODE == 0) {
                g.state[b].pq = pq;
                PsnPend* pd = g.pend + b;   // deferred x: record the step, x receives (alpha / s_k) (s_k p_k) at the next flush
                if (pd->lo == pd->hi) pd->lo = g.it;
                pd->a[g.it % PSN_PEND_K] = alpha * (1.0 / pscale);
                pd->hi = g.it + 1;
            } else {
                g.vs0[b] = pscale;
            }
        }
    }
}

// ----------------------------------------------------------------------------- true residual, strip form
// Samples in PSN_PHASE_CHECK: r = bt - At x with bt = -(sy_i sx_j / theta_ij) h^2 (the same operations as psn_rhs, so
// r is bit-identical to the 2-D tile kernel); partial ||r||^2.  Only for the product's standard call form: sine
// factors + theta on the solve grid itself (no master-field stride); other forcings keep the tile kernel.
template <class Feeder>
PSN_FN void psn_true_residual_strip_body(const PsnWarp& w, const PsnCgArgs& a, long long b, int band, int strip, Feeder& fd) {
    if (a.state[b].phase != PSN_PHASE_CHECK) return;
    constexpr int W = Feeder::W;
    static_assert(W % 3 == 0, "the 3-row window rotates inside one ring period");
    const int n = a.n, last = n - 1;
    // same strip decomposition as the update kernel (its partial-sum slots are reused)
    const int a0 = strip * psn_dir_use(a.pf32), a1 = psn_imin(a0 + psn_dir_use(a.pf32), n);
    const int c0 = a0 - psn_dir_hl(a.pf32);
    const int i0 = band * a.upd_band, i1 = psn_imin(i0 + a.upd_band, n);
    const wi col0 = w_lane(w) * 4 + c0;
    const wi col2 = col0 + 2;
    const wb mp[2] = {(col0 >= 0) && (col0 < n), (col2 >= 0) && (col2 < n)};
    const wb so[2] = {mp[0] && (col0 >= a0) && (col0 < a1), mp[1] && (col2 >= a0) && (col2 < a1)};
    const long long base = b * (long long)n * n;
    const wd zero = w_bcast(0.0);
    const wd h2 = w_bcast(a.h2);
    wd sxv[4];
    w_ldp(a.rhs.sx + b * (long long)n, col0, mp[0], sxv[0], sxv[1]);
    w_ldp(a.rhs.sx + b * (long long)n, col2, mp[1], sxv[2], sxv[3]);
    const double* syp = a.rhs.sy + b * (long long)n;

    const int ts = psn_imax(i0 - 1, 0);
    const int tload = psn_imin(i1, last);
    const int tend = i1;
    fd.set_base0(a.x + base);
    fd.set_base1(a.rhs.theta + base);
    fd.n = n;
    fd.tload = tload;
    psn_feeder_geom(fd, col0, mp, c0, n);
    fd.prime(w, ts);

    wd xw[3][4], thp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        thp[e] = w_bcast(1.0);
#pragma unroll
        for (int k = 0; k < 3; ++k) xw[k][e] = zero;
    }
    wd acc = zero;
    double* rg = a.r + base;

    for (int t = ts; t <= tend; t += W) {
        const int par = ((t - ts) / W) & 1;
#pragma unroll
        for (int tt = 0; tt < W; ++tt) {
            const int tc = t + tt;
            if (tc > tend) break;
            const int k0 = tt % 3, k1 = (tt + 2) % 3, k2 = (tt + 1) % 3;   // rows tc, tc-1, tc-2
            wd thc[4];
            fd.template fetch<false>(w, tt, tc, par, xw[k0], thc);
            const int r = tc - 1;
            if (r >= i0 && r < i1) {
                const wd upv = w_up_raw(xw[k1][3]);
                const wd dnv = w_down_raw(xw[k1][0]);
                wd ax[4];
                ax[0] = 4.0 * xw[k1][0] - ((upv + xw[k1][1]) + (xw[k2][0] + xw[k0][0]));
                ax[1] = 4.0 * xw[k1][1] - ((xw[k1][0] + xw[k1][2]) + (xw[k2][1] + xw[k0][1]));
                ax[2] = 4.0 * xw[k1][2] - ((xw[k1][1] + xw[k1][3]) + (xw[k2][2] + xw[k0][2]));
                ax[3] = 4.0 * xw[k1][3] - ((xw[k1][2] + dnv) + (xw[k2][3] + xw[k0][3]));
                const wd syr = w_bcast(syp[r]);
                wd rv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // out-of-domain lanes multiply 0 by 1 / 1 (thp is 1 there: the ring's un-copied columns are zero, see below)
                    const wd f = (syr * sxv[e]) * w_rcp(thp[e]);   // same expression as psn_rhs
                    rv[e] = (zero - (f * h2)) - ax[e];
                }
                double* rowr = rg + (long long)r * n;
                w_stp(rowr, col0, so[0], rv[0], rv[1]);
                w_stp(rowr, col2, so[1], rv[2], rv[3]);
                acc = acc + w_sel(so[0], rv[0] * rv[0] + rv[1] * rv[1], zero) + w_sel(so[1], rv[2] * rv[2] + rv[3] * rv[3], zero);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) thp[e] = w_sel((e < 2) ? mp[0] : mp[1], thc[e], w_bcast(1.0));
        }
    }
    const double s = w_sum(acc);
    if (w_lane0(w)) a.part_u[b * (long long)(a.upd_nstrips * a.upd_nbands) + (long long)band * a.upd_nstrips + strip] = s;
}
