// psn_warp.cuh -- "warp vector" types for kernels written at warp granularity.
//
// Device build (nvcc): wd / wi / wb are one lane's double / int / bool and the helpers are
// warp shuffles and predicated global accesses.  Host build (g++, tests/hostemu only): the same
// names are 32-lane arrays with element-wise operators, so a warp-synchronous kernel body runs
// unchanged on the CPU test tier.  Control flow in such bodies must be warp-uniform; everything
// lane-dependent goes through w_sel().
#pragma once
#include "psn_core.cuh"

#if defined(__CUDACC__)

typedef double wd;
typedef int wi;
typedef bool wb;
struct PsnWarp {
    int lane;
};

PSN_FN wd w_bcast(double s) { return s; }
PSN_FN wi w_lane(const PsnWarp& w) { return w.lane; }
PSN_FN wd w_sel(wb c, wd a, wd b) { return c ? a : b; }
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
PSN_FN wd w_ld1(const double* row, wi idx, wb m) { return m ? row[idx] : 0.0; }
PSN_FN void w_st1(double* row, wi idx, wb m, wd v) {
    if (m) row[idx] = v;
}

#else  // ------------------------------------------------------------------ host emulation

struct wd {
    double v[32];
};
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
inline wd operator+(const wd& a, const wd& b) { wd r; PSN_W_FOR r.v[l_] = a.v[l_] + b.v[l_]; return r; }
inline wd operator-(const wd& a, const wd& b) { wd r; PSN_W_FOR r.v[l_] = a.v[l_] - b.v[l_]; return r; }
inline wd operator*(const wd& a, const wd& b) { wd r; PSN_W_FOR r.v[l_] = a.v[l_] * b.v[l_]; return r; }
inline wd operator*(double a, const wd& b) { wd r; PSN_W_FOR r.v[l_] = a * b.v[l_]; return r; }
inline wi operator*(const wi& a, int b) { wi r; PSN_W_FOR r.v[l_] = a.v[l_] * b; return r; }
inline wi operator+(const wi& a, int b) { wi r; PSN_W_FOR r.v[l_] = a.v[l_] + b; return r; }
inline wb operator>=(const wi& a, int b) { wb r; PSN_W_FOR r.v[l_] = a.v[l_] >= b; return r; }
inline wb operator<(const wi& a, int b) { wb r; PSN_W_FOR r.v[l_] = a.v[l_] < b; return r; }
inline wb operator==(const wi& a, int b) { wb r; PSN_W_FOR r.v[l_] = a.v[l_] == b; return r; }
inline wb operator&&(const wb& a, const wb& b) { wb r; PSN_W_FOR r.v[l_] = a.v[l_] && b.v[l_]; return r; }
inline wd w_sel(const wb& c, const wd& a, const wd& b) { wd r; PSN_W_FOR r.v[l_] = c.v[l_] ? a.v[l_] : b.v[l_]; return r; }
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
inline wd w_ld1(const double* row, const wi& idx, const wb& m) { wd r; PSN_W_FOR r.v[l_] = m.v[l_] ? row[idx.v[l_]] : 0.0; return r; }
inline void w_st1(double* row, const wi& idx, const wb& m, const wd& v) { PSN_W_FOR if (m.v[l_]) row[idx.v[l_]] = v.v[l_]; }

#endif

// ----------------------------------------------------------------------------- strip kernels
// One warp owns a strip of 64 columns (lane l: columns c0+2l, c0+2l+1, c0 even) and marches down
// the rows of its band.  A rolling window of W = 2*NU+2 rows lives in registers; at time t the
// pipeline applies
//     pre : R(t) [zero guess], B(t-1), R(t-2), ..., B(t-2NU+1), residual(t-2NU) -> restriction
//     post: prolong(t), B(t-1), R(t-2), ..., R(t-2NU)                           -> <b,u> partial
// so every sweep of a V-cycle leg is done in ONE pass over the rows, with vertical neighbours in
// registers and horizontal neighbours through one shuffle per stage.  Row parity works out so that
// at time t every stage touches the SAME register slot P = t & 1 of its row.  The validity window
// shrinks by one column per stage from the strip edges (halo), exactly as in the tiled version, so
// the result is identical to the untiled algorithm.
struct PsnStripArgs {
    int level;
    int band;            // output rows per task (even)
    int nstrips, nbands; // tasks per sample = nstrips * nbands
    const double* b;     // [B, n, n]
    const double* uin;   // post
    double* uout;
    double* bc;          // pre:  [B, nc, nc] out
    const double* ec;    // post: [B, nc, nc] in
    double* part;        // post, level 0: [B, nstrips*nbands] partial <b,u>
    const PsnSampleState* state;
};

PSN_HDFN int psn_strip_use_pre(int nu) { return 64 - (4 * nu + 2); }
// (the up-leg needs two more halo columns on the left: the prolongation onto an even column
// reads the coarse value of the lane to the left)
PSN_HDFN int psn_strip_use_post(int nu) { return 64 - (4 * nu + 2); }

template <int NU>
PSN_FN void psn_strip_pre_body(const PsnWarp& w, const PsnPlan& plan, const PsnStripArgs& a, long long b, int band, int strip) {
    if (a.state && a.state[b].phase != PSN_PHASE_ITER) return;
    constexpr int NS = 2 * NU;   // smoothing stages; the residual is stage NS
    constexpr int W = NS + 2;    // window rows
    constexpr int PF = W / 2;    // rows prefetched ahead
    const int USE = 64 - (2 * NS + 2);
    const PsnLevel& lv = plan.lv[a.level];
    const PsnLevel& lc = plan.lv[a.level + 1];
    const int n = lv.n, nc = lc.n, last = n - 1;
    const int a0 = strip * USE, a1 = psn_imin(a0 + USE, n);
    const int c0 = a0 - NS;
    const int i0 = band * a.band, i1 = psn_imin(i0 + a.band, n);
    const int I0 = i0 >> 1, I1 = psn_imin(i1 >> 1, nc);
    const int J0 = a0 >> 1, J1 = psn_imin(a1 >> 1, nc);
    const bool vec = (n & 1) == 0;

    const wi col0 = w_lane(w) * 2 + c0;
    const wi col1 = col0 + 1;
    const wb m[2] = {(col0 >= 0) && (col0 < n), (col1 >= 0) && (col1 < n)};
    const wb lastc[2] = {col0 == last, col1 == last};
    const wd rdA[2] = {w_sel(lastc[0], w_bcast(lv.rd1), w_bcast(lv.rd0)), w_sel(lastc[1], w_bcast(lv.rd1), w_bcast(lv.rd0))};
    const wd rdB[2] = {w_sel(lastc[0], w_bcast(lv.rd2), w_bcast(lv.rd1)), w_sel(lastc[1], w_bcast(lv.rd2), w_bcast(lv.rd1))};
    const wd dA[2] = {w_sel(lastc[0], w_bcast(4.0 + lv.gam), w_bcast(4.0)), w_sel(lastc[1], w_bcast(4.0 + lv.gam), w_bcast(4.0))};
    const wd dB[2] = {w_sel(lastc[0], w_bcast(4.0 + 2.0 * lv.gam), w_bcast(4.0 + lv.gam)), w_sel(lastc[1], w_bcast(4.0 + 2.0 * lv.gam), w_bcast(4.0 + lv.gam))};
    // store masks
    const wb so[2] = {m[0] && (col0 >= a0) && (col0 < a1), m[1] && (col1 >= a0) && (col1 < a1)};
    const wi cj = w_lane(w) + (c0 >> 1);   // coarse column held by this lane (c0 even, may be negative)
    const wb sc = (cj >= J0) && (cj < J1);

    const int ts = psn_imax(i0 - NS, 0);
    const int tload = psn_imin(i1 + NS, last);
    const int tend = i1 + NS;
    const double* bg = a.b + b * (long long)n * n;
    double* ug = a.uout + b * (long long)n * n;
    double* cg = a.bc + b * (long long)nc * nc;
    const wd zero = w_bcast(0.0);

    wd val[W][2], bb[W][2], pf[PF][2];
#pragma unroll
    for (int k = 0; k < W; ++k) val[k][0] = val[k][1] = bb[k][0] = bb[k][1] = zero;
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        pf[k][0] = pf[k][1] = zero;
        if (ts + k <= tload) w_ld2(bg + (long long)(ts + k) * n, col0, m[0], m[1], vec, pf[k][0], pf[k][1]);
    }
    wd acc = zero;

    for (int t = ts; t <= tend; t += W) {
#pragma unroll
        for (int tt = 0; tt < W; ++tt) {
            const int tc = t + tt;
            if (tc > tend) break;
            constexpr int dummy = 0;
            (void)dummy;
            const int P = tt & 1;  // ts and W are even: parity of row tc
            const int k0 = tt;     // window slot of row tc
            // ---- new row: right-hand side from the prefetch ring, first red half sweep (zero guess)
            bb[k0][0] = pf[tt % PF][0];
            bb[k0][1] = pf[tt % PF][1];
            pf[tt % PF][0] = pf[tt % PF][1] = zero;
            if (tc + PF <= tload) w_ld2(bg + (long long)(tc + PF) * n, col0, m[0], m[1], vec, pf[tt % PF][0], pf[tt % PF][1]);
            if (tc <= last) {
                val[k0][P] = w_sel(m[P], bb[k0][P] * (tc == last ? rdB[P] : rdA[P]), zero);
            } else {
                bb[k0][0] = bb[k0][1] = zero;
                val[k0][P] = zero;
            }
            val[k0][1 - P] = zero;
            // ---- smoothing stages 1 .. NS-1 on rows tc-1 .. tc-NS+1
#pragma unroll
            for (int s = 1; s < NS; ++s) {
                const int r = tc - s;
                if (r >= ts && r <= last) {
                    const int ks = (tt - s + 2 * W) % W, ku = (tt - s - 1 + 2 * W) % W, kd = (tt - s + 1 + 2 * W) % W;
                    const wd other = val[ks][1 - P];
                    const wd lr = (P == 0) ? (w_up(w, other) + other) : (other + w_down(w, other));
                    const wd s4 = (val[ku][P] + val[kd][P]) + lr;
                    val[ks][P] = w_sel(m[P], (bb[ks][P] + s4) * (r == last ? rdB[P] : rdA[P]), zero);
                }
            }
            // ---- residual on the red node of row tc-NS, accumulate the restriction
            {
                const int r = tc - NS;
                wd rho = zero;
                if (r >= ts && r <= last) {
                    const int ks = (tt - NS + 2 * W) % W, ku = (tt - NS - 1 + 2 * W) % W, kd = (tt - NS + 1 + 2 * W) % W;
                    const wd other = val[ks][1 - P];
                    const wd lr = (P == 0) ? (w_up(w, other) + other) : (other + w_down(w, other));
                    const wd s4 = (val[ku][P] + val[kd][P]) + lr;
                    rho = w_sel(m[P], bb[ks][P] - ((r == last ? dB[P] : dA[P]) * val[ks][P] - s4), zero);
                }
                if (P == 0) {  // even row 2I: corner weights of coarse rows I-1 (bottom) and I (top)
                    const wd cs = 0.25 * (rho + w_down(w, rho));
                    const int I = (r >> 1) - 1;
                    if (r >= 2 && I >= I0 && I < I1) w_st1(cg + (long long)I * nc, cj, sc, acc + cs);
                    acc = cs;
                } else {       // odd row 2I+1: centre of coarse row I
                    acc = acc + rho;
                }
            }
            // ---- row tc-NS+1 is final after its last black half sweep
            {
                const int ro = tc - (NS - 1);
                if (ro >= i0 && ro < i1) {
                    const int ko = (tt - (NS - 1) + 2 * W) % W;
                    w_st2(ug + (long long)ro * n, col0, so[0], so[1], vec, val[ko][0], val[ko][1]);
                }
            }
        }
    }
}

template <int NU>
PSN_FN void psn_strip_post_body(const PsnWarp& w, const PsnPlan& plan, const PsnStripArgs& a, long long b, int band, int strip) {
    if (a.state && a.state[b].phase != PSN_PHASE_ITER) return;
    constexpr int NS = 2 * NU;
    constexpr int W = NS + 2;
    constexpr int PF = W / 2;
    const int USE = 64 - (2 * NS + 2);
    const PsnLevel& lv = plan.lv[a.level];
    const PsnLevel& lc = plan.lv[a.level + 1];
    const int n = lv.n, nc = lc.n, last = n - 1;
    const int a0 = strip * USE, a1 = psn_imin(a0 + USE, n);
    const int c0 = a0 - NS - 2;   // even; NS + 2 halo columns on the left, NS on the right
    const int i0 = band * a.band, i1 = psn_imin(i0 + a.band, n);
    const bool vec = (n & 1) == 0;

    const wi col0 = w_lane(w) * 2 + c0;
    const wi col1 = col0 + 1;
    const wb m[2] = {(col0 >= 0) && (col0 < n), (col1 >= 0) && (col1 < n)};
    const wb lastc[2] = {col0 == last, col1 == last};
    const wd rdA[2] = {w_sel(lastc[0], w_bcast(lv.rd1), w_bcast(lv.rd0)), w_sel(lastc[1], w_bcast(lv.rd1), w_bcast(lv.rd0))};
    const wd rdB[2] = {w_sel(lastc[0], w_bcast(lv.rd2), w_bcast(lv.rd1)), w_sel(lastc[1], w_bcast(lv.rd2), w_bcast(lv.rd1))};
    const wb so[2] = {m[0] && (col0 >= a0) && (col0 < a1), m[1] && (col1 >= a0) && (col1 < a1)};
    const wi cj = w_lane(w) + (c0 >> 1);
    const wb mc = (cj >= 0) && (cj < nc);

    const int ts = psn_imax(i0 - NS, 0);
    const int tload = psn_imin(i1 - 1 + NS, last);
    const int tend = psn_imin(i1 - 1, last) + NS;
    const double* bg = a.b + b * (long long)n * n;
    const double* uig = a.uin + b * (long long)n * n;
    const double* eg = a.ec + b * (long long)nc * nc;
    double* ug = a.uout + b * (long long)n * n;
    const wd zero = w_bcast(0.0);

    wd val[W][2], bb[W][2], pfb[PF][2], pfu[PF][2];
#pragma unroll
    for (int k = 0; k < W; ++k) val[k][0] = val[k][1] = bb[k][0] = bb[k][1] = zero;
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        pfb[k][0] = pfb[k][1] = pfu[k][0] = pfu[k][1] = zero;
        if (ts + k <= tload) {
            w_ld2(bg + (long long)(ts + k) * n, col0, m[0], m[1], vec, pfb[k][0], pfb[k][1]);
            w_ld2(uig + (long long)(ts + k) * n, col0, m[0], m[1], vec, pfu[k][0], pfu[k][1]);
        }
    }
    // coarse correction rows: eprev = row ts/2 - 1 (zero above the grid)
    wd eprev = zero, dot = zero;
    {
        const int Ip = (ts >> 1) - 1;
        if (Ip >= 0 && Ip < nc) eprev = w_ld1(eg + (long long)Ip * nc, cj, mc);
    }

    for (int t = ts; t <= tend; t += W) {
#pragma unroll
        for (int tt = 0; tt < W; ++tt) {
            const int tc = t + tt;
            if (tc > tend) break;
            const int P = tt & 1;
            const int k0 = tt;
            // ---- new row: rhs and pre-smoothed iterate, then add the prolongated correction on the red node
            bb[k0][0] = pfb[tt % PF][0];
            bb[k0][1] = pfb[tt % PF][1];
            val[k0][0] = pfu[tt % PF][0];
            val[k0][1] = pfu[tt % PF][1];
            pfb[tt % PF][0] = pfb[tt % PF][1] = pfu[tt % PF][0] = pfu[tt % PF][1] = zero;
            if (tc + PF <= tload) {
                w_ld2(bg + (long long)(tc + PF) * n, col0, m[0], m[1], vec, pfb[tt % PF][0], pfb[tt % PF][1]);
                w_ld2(uig + (long long)(tc + PF) * n, col0, m[0], m[1], vec, pfu[tt % PF][0], pfu[tt % PF][1]);
            }
            if (tc <= last) {
                if (P == 0) {  // even row 2I, red = even column: mean of coarse (I-1,J-1), (I-1,J), (I,J-1), (I,J)
                    const int I = tc >> 1;
                    wd enew = zero;
                    if (I < nc) enew = w_ld1(eg + (long long)I * nc, cj, mc);
                    const wd corr = 0.25 * ((w_up(w, eprev) + eprev) + (w_up(w, enew) + enew));
                    val[k0][0] = w_sel(m[0], val[k0][0] + corr, zero);
                    eprev = enew;
                } else {       // odd row 2I+1, red = odd column: coincides with coarse (I, J)
                    val[k0][1] = w_sel(m[1], val[k0][1] + eprev, zero);
                }
            } else {
                val[k0][0] = val[k0][1] = bb[k0][0] = bb[k0][1] = zero;
            }
            // ---- smoothing stages 1 .. NS (black first) on rows tc-1 .. tc-NS
#pragma unroll
            for (int s = 1; s <= NS; ++s) {
                const int r = tc - s;
                if (r >= ts && r <= last) {
                    const int ks = (tt - s + 2 * W) % W, ku = (tt - s - 1 + 2 * W) % W, kd = (tt - s + 1 + 2 * W) % W;
                    const wd other = val[ks][1 - P];
                    const wd lr = (P == 0) ? (w_up(w, other) + other) : (other + w_down(w, other));
                    const wd s4 = (val[ku][P] + val[kd][P]) + lr;
                    val[ks][P] = w_sel(m[P], (bb[ks][P] + s4) * (r == last ? rdB[P] : rdA[P]), zero);
                }
            }
            // ---- row tc-NS is final
            {
                const int ro = tc - NS;
                if (ro >= i0 && ro < i1) {
                    const int ko = (tt - NS + 2 * W) % W;
                    w_st2(ug + (long long)ro * n, col0, so[0], so[1], vec, val[ko][0], val[ko][1]);
                    if (a.part) dot = dot + (w_sel(so[0], val[ko][0] * bb[ko][0], zero) + w_sel(so[1], val[ko][1] * bb[ko][1], zero));
                }
            }
        }
    }
    if (a.part) {
        const double s = w_sum(dot);
        if (w_lane0(w)) a.part[b * (long long)(a.nstrips * a.nbands) + (long long)band * a.nstrips + strip] = s;
    }
}
