// psn_stream_driver.h -- host-side orchestration of the streamed (large-grid) MG-PCG solve.
//
// Plain C++, templated over a backend `BK` that knows how to run each kernel body over a
// batch: the CUDA backend (psn_api.cu) launches kernels on a stream, the host-emulation
// backend (tests/hostemu, TEST ONLY) loops over the blocks on the CPU.  The driver never
// synchronises with the device inside an iteration: alpha/beta live in per-sample device state,
// converged samples are masked out per CTA, and the number of still-active samples is read
// back with a lag of PSN_COUNTER_LAG iteration(s): the host waits for the counter of iteration it-1 after it has
// enqueued iteration it, so the device always has a full iteration of work queued.
#pragma once
#include "psn_warp.cuh"

#define PSN_COUNTER_LAG 1

struct PsnStreamCfg {
    int ls;       // first level handled by the shared-memory coarse kernel (levels 0..ls-1 are streamed)
    int coarse_block = 0;   // level ls is 80 x 80 and goes to the register-blocked V-cycle kernel (psn_block_vcycle_body)
    int TH, TW;   // interior tile of the shared-memory tile kernels (mode 0)
    int mode;     // V-cycle legs: 0 = shared-memory tiles, 1 = warp strips fed by register prefetch,
                  //               2 = warp strips fed by TMA bulk copies through a shared-memory ring
    int band[PSN_MAX_LEVELS];  // mode 1: output rows per warp task on each streamed level (even)
    int dir_band; // rows per warp task of the direction kernel (0: band[0])
    int upd_band; // rows per warp task of the strip update kernel (0: 2 * dir_band)
    int upd_strip; // mode 2 only: the update kernel recomputes q = At p in strips and q is never stored
    int pf32;      // with vf32 and upd_strip: the search direction p is stored as (scaled) float as well
    int xdefer;    // with pf32: x is accumulated from the ring of stored directions when a sample is checked (PsnPend)
    int fuse_band; // rows per warp task of the fused kernel (0: band[0]); its own wave model: 3 CTAs per SM, not 4
    //             fuse: 0 off, 1 update + down-leg, 2 (default) also prologue + first down-leg for the standard call form
    int fuse;      // with xdefer: the CG update of iteration k and the level-0 down-leg of iteration k+1 are ONE strip kernel
                   // (psn_strip_upre_body): r is not re-read, 33 -> 25 bytes per node and iteration for the two
    int vf32;     // mode 2 only: V-cycle intermediates (iterates, restricted residuals, z) are STORED as float; all
                  // arithmetic, the CG vectors x r p q and every inner product stay fp64.  The preconditioner only has
                  // to be an approximate inverse, so CG converges in the same number of iterations (tests), while the
                  // V-cycle legs move 30-35 % fewer bytes.  Needs n % 4 == 0 on every streamed level.
};

// one leg (down or up) of the V-cycle on one streamed level
struct PsnLegArgs {
    int level;
    const double* b;    // [B, n_l, n_l] right-hand side
    const double* uin;  // up-leg: pre-smoothed iterate
    double* uout;
    double* bc;         // down-leg: restricted residual out
    const double* ec;   // up-leg: coarse correction in
    double* part;       // up-leg on level 0: partial <b,u> per task
    const PsnSampleState* state;
    int fused = 0;        // fused schedule (see PsnStreamCfg::fuse): V-cycle scale from rr_prev
    int only_restart = 0; // down-leg on level 0, fused schedule: only samples whose residual was rebuilt
};

// the warp-strip kernels use 128-bit pair accesses: every streamed level must have an even size
static inline bool psn_strip_ok(const PsnPlan& plan, int ls) {
    for (int l = 0; l < ls; ++l)
        if (plan.lv[l].n & 1) return false;
    return true;
}
// float rows are accessed in 16-byte quads: every streamed level needs n % 4 == 0
static inline bool psn_vf32_ok(const PsnPlan& plan, int ls) {
    for (int l = 0; l < ls; ++l)
        if (plan.lv[l].n & 3) return false;
    return true;
}
static inline int psn_strip_count(int n, int use) { return (n + use - 1) / use; }
static inline int psn_band_count(int n, int band) { return (n + band - 1) / band; }
// upper bound of the <p,q> partial count per sample (any band choice >= 16 rows, or the 2-D tiles)
static inline int psn_pq_max(int n) {
    const int strips = ((n + psn_dir_use(1) - 1) / psn_dir_use(1)) * ((n + 15) / 16);
    const int tiles = ((n + PSN_GT_COLS - 1) / PSN_GT_COLS) * ((n + PSN_GT_ROWS - 1) / PSN_GT_ROWS);
    return strips > tiles ? strips : tiles;
}
// upper bound of the ||r||^2 partial count per sample: the strip update / true-residual kernels (as <p,q>), or the fused
// update + down-leg kernel (112-column strips for nu = 2, bands of >= 16 rows)
static inline int psn_pu_max(int n, int nu) {
    const int fused = psn_strip_count(n, psn_fused_use(nu)) * psn_band_count(n, 16);
    const int pq = psn_pq_max(n);
    return fused > pq ? fused : pq;
}
// number of <r,z> partials per sample produced by the level-0 up-leg
static inline int psn_rz_count(const PsnPlan& plan, const PsnStreamCfg& cfg) {
    const int n = plan.n;
    if (cfg.mode >= 1) return psn_strip_count(n, psn_strip_use_post(plan.nu, cfg.vf32)) * psn_band_count(n, cfg.band[0]);
    return ((n + cfg.TW - 1) / cfg.TW) * ((n + cfg.TH - 1) / cfg.TH);
}

// device arrays for one chunk of Bc samples
struct PsnStreamWs {
    long long cap;               // samples this workspace was sized for
    double* r;                   // level 0: residual (= rhs of the V-cycle)
    double* r2;                  // fused schedule: second residual buffer (the update writes r[cur ^ 1])
    int rcur;                    // fused schedule: buffer that holds the current residual
    double* vs0;                 // fused schedule: [cap] V-cycle scale the fused prologue used (read by state_init)
    double* q;                   // level 0: A p  (doubles as the pre-smoothed iterate)
    double* z;                   // level 0: preconditioned residual
    double* p[2];                // level 0: search direction, ping-pong
    double* bl[PSN_MAX_LEVELS];  // levels 1..ls: rhs
    double* ua[PSN_MAX_LEVELS];  // levels 1..ls-1: pre-smoothed iterate
    double* ub[PSN_MAX_LEVELS];  // levels 1..ls: corrected iterate
    double* part_rz;
    double* part_a;
    double* part_pq;
    double* part_b;
    double* part_u;              // ||r||^2 partials of the strip update kernel
    float* pring;                // xdefer: ring of PSN_PEND_K float search directions, [K][cap, n, n]
    PsnPend* pend;               // xdefer: [cap]
    PsnSampleState* state;
    int* counters;               // [max_iter + 8] active-sample count per outer iteration
};

static inline void psn_tile_counts(int n, int TH, int TW, int* ntx, int* nty) {
    *ntx = (n + TW - 1) / TW;
    *nty = (n + TH - 1) / TH;
}

static inline PsnTileArgs psn_leg_to_tile(const PsnPlan& plan, const PsnStreamCfg& cfg, const PsnLegArgs& g) {
    PsnTileArgs t;
    t.level = g.level;
    t.TH = cfg.TH;
    t.TW = cfg.TW;
    psn_tile_counts(plan.lv[g.level].n, cfg.TH, cfg.TW, &t.ntx, &t.nty);
    t.b = g.b; t.uin = g.uin; t.uout = g.uout; t.bc = g.bc; t.ec = g.ec; t.part = g.part; t.state = g.state;
    return t;
}
static inline PsnStripArgs psn_leg_to_strip(const PsnPlan& plan, const PsnStreamCfg& cfg, const PsnLegArgs& g, bool post) {
    PsnStripArgs a;
    const int n = plan.lv[g.level].n;
    a.level = g.level;
    a.band = cfg.band[g.level];
    const int nu_l = plan.nul[g.level];
    a.nstrips = psn_strip_count(n, post ? psn_strip_use_post(nu_l, cfg.vf32) : psn_strip_use_pre(nu_l, cfg.vf32));
    a.nbands = psn_band_count(n, a.band);
    a.b = g.b; a.uin = g.uin; a.uout = g.uout; a.bc = g.bc; a.ec = g.ec; a.part = g.part; a.state = g.state;
    a.fused = g.fused; a.only_restart = g.only_restart;
    return a;
}
static inline bool psn_fuse_ok(const PsnStreamCfg& cfg) { return cfg.mode == 2 && cfg.vf32 && cfg.upd_strip && cfg.pf32 && cfg.xdefer && cfg.fuse; }

// x: [B, n, n] device array that receives the solution.
// start_it == 0: fresh solve (prologue + state init); > 0: resume an earlier call at that outer iteration.
// fixed_iters == 0: poll the active-sample counter (lag PSN_COUNTER_LAG) and stop when every sample is done;
// fixed_iters  > 0: enqueue exactly that many outer iterations WITHOUT waiting on the device (iterations after a
//                   sample's convergence are no-ops for it); the caller checks counter[returned - 1] later.
// Returns the index of the next outer iteration.
template <class BK>
int psn_stream_solve(BK& bk, const PsnPlan& plan, const PsnStreamCfg& cfg, const PsnRhs& rhs, long long B,
                     double rtol, int max_iter, double* x, PsnStreamWs& ws, int start_it = 0, int fixed_iters = 0) {
    const int n = plan.n;
    PsnCgArgs cg;
    cg.n = n;
    cg.ntx = (n + PSN_GT_COLS - 1) / PSN_GT_COLS;
    cg.nty = (n + PSN_GT_ROWS - 1) / PSN_GT_ROWS;
    cg.ntiles_rz = psn_rz_count(plan, cfg);
    cg.rhs = rhs;
    cg.h2 = plan.h2;
    cg.x = x;
    cg.r = ws.r;
    cg.z = ws.z;
    cg.q = ws.q;
    cg.p_old = ws.p[0];
    cg.p_new = ws.p[1];
    cg.part_rz = ws.part_rz;
    cg.part_a = ws.part_a;
    cg.part_pq = ws.part_pq;
    cg.dir_mode = (cfg.mode >= 1 && (n & 1) == 0) ? cfg.mode : 0;
    cg.dir_band = cfg.dir_band >= 16 ? cfg.dir_band : cfg.band[0];
    cg.vf32 = (cg.dir_mode == 2) ? cfg.vf32 : 0;
    cg.dir_nstrips = (n + psn_dir_use(cg.vf32) - 1) / psn_dir_use(cg.vf32);
    cg.dir_nbands = psn_band_count(n, cg.dir_band);
    cg.npq = cg.dir_mode ? cg.dir_nstrips * cg.dir_nbands : cg.ntx * cg.nty;
    cg.upd_mode = (cg.dir_mode == 2 && cfg.upd_strip) ? 2 : 0;
    cg.upd_band = cfg.upd_band >= 16 ? cfg.upd_band : 2 * cg.dir_band;
    cg.pf32 = (cg.upd_mode == 2 && cg.vf32 && cfg.pf32) ? 1 : 0;
    cg.upd_nstrips = (n + psn_dir_use(cg.pf32) - 1) / psn_dir_use(cg.pf32);
    cg.upd_nbands = psn_band_count(n, cg.upd_band);
    cg.part_u = ws.part_u;
    cg.xdefer = (cg.pf32 && cfg.xdefer && ws.pring && ws.pend) ? 1 : 0;
    cg.pring = ws.pring;
    cg.ring_stride = ws.cap * (long long)n * n;
    cg.pend = ws.pend;
    cg.it = start_it;
    cg.chk_mode = (cg.upd_mode == 2 && rhs.sx && rhs.theta && rhs.step == 1 && rhs.mn == n) ? 2 : 0;
    const int fuse = (psn_fuse_ok(cfg) && cg.xdefer && ws.r2) ? 1 : 0;
    PsnStripArgs fa;                      // fused update + down-leg: level-0 strip task geometry
    memset(&fa, 0, sizeof(fa));
    fa.level = 0;
    fa.band = cfg.fuse_band >= 16 ? cfg.fuse_band : cfg.band[0];
    fa.nstrips = psn_strip_count(n, psn_fused_use(plan.nu));
    fa.nbands = psn_band_count(n, fa.band);
    fa.uout = ws.q;
    fa.bc = ws.bl[1];
    fa.state = ws.state;
    fa.fused = 1;
    cg.fused = fuse;
    cg.fu_nparts = fa.nstrips * fa.nbands;
    cg.r_out = nullptr;
    double* rbuf[2] = {ws.r, fuse ? ws.r2 : ws.r};
    if (start_it == 0) ws.rcur = 0;
    cg.r = rbuf[ws.rcur];
    cg.part_b = ws.part_b;
    cg.state = ws.state;
    cg.rtol = rtol;
    cg.max_iter = max_iter;

    // standard call form (sine factors + theta on the solve grid): the prologue runs inside the fused kernel as well
    const int profuse = (fuse && cg.chk_mode == 2 && ws.vs0 && cfg.fuse >= 2) ? 1 : 0;
    cg.pro_fused = profuse;
    cg.vs0 = ws.vs0;
    if (start_it == 0) {
        bk.zero_counters(ws.counters, max_iter + 8);
        if (profuse) {
            cg.r_out = rbuf[0];
            bk.fused_prologue(plan, fa, cg, B);
        } else {
            bk.prologue(cg, B);
        }
        bk.state_init(cg, B);
    }

    const int max_outer = max_iter + 4;  // every check/restart costs at most one extra outer pass
    // The check block (x flush, true residual, state transition, counter) is four launches that do nothing until the
    // first sample converges.  It is skipped for the first outer iterations: a sample that converges earlier simply
    // waits in PSN_PHASE_CHECK (nothing touches it) until the first check.  Never later than PSN_PEND_K - 2, so that a
    // full direction ring is always flushed in time.
    int check_from = start_it;
    if (start_it == 0) {
        check_from = fixed_iters > 0 ? fixed_iters - 4 : 3;
        if (check_from > PSN_PEND_K - 2) check_from = PSN_PEND_K - 2;
        if (check_from > max_iter - 1) check_from = max_iter - 1;   // samples that stop at max_iter are checked at once
        if (check_from < 0) check_from = 0;
    }
    int it = start_it;
    // fused schedule: some samples may carry a residual that no fused kernel has smoothed yet
    bool rebuilt = !(profuse && start_it == 0);
    for (; it < max_outer; ++it) {
        // ---- z = M^{-1} r : one V(nu,nu) cycle
        for (int l = 0; l < cfg.ls; ++l) {
            if (l == 0 && fuse && !rebuilt) continue;   // done by the previous iteration's fused update kernel
            PsnLegArgs g;
            g.level = l;
            g.fused = fuse;
            g.only_restart = (l == 0 && fuse) ? 1 : 0;
            g.b = (l == 0) ? rbuf[ws.rcur] : ws.bl[l];
            g.uin = nullptr;
            g.uout = (l == 0) ? ws.q : ws.ua[l];
            g.bc = ws.bl[l + 1];
            g.ec = nullptr;
            g.part = nullptr;
            g.state = ws.state;
            bk.leg_pre(plan, cfg, g, B);
        }
        bk.coarse_vcycle(plan, cfg, ws.bl[cfg.ls], ws.ub[cfg.ls], ws.state, B);
        for (int l = cfg.ls - 1; l >= 0; --l) {
            PsnLegArgs g;
            g.level = l;
            g.fused = fuse;
            g.b = (l == 0) ? rbuf[ws.rcur] : ws.bl[l];
            g.uin = (l == 0) ? ws.q : ws.ua[l];
            g.uout = (l == 0) ? ws.z : ws.ub[l];
            g.bc = nullptr;
            g.ec = ws.ub[l + 1];
            g.part = (l == 0) ? ws.part_rz : nullptr;
            g.state = ws.state;
            bk.leg_post(plan, cfg, g, B);
        }
        // ---- CG update
        cg.it = it;
        if (cg.xdefer) {   // the directions of the last PSN_PEND_K iterations stay in their ring slots until x has absorbed them
            cg.p_old = reinterpret_cast<double*>(ws.pring + (long long)((it + PSN_PEND_K - 1) % PSN_PEND_K) * cg.ring_stride);
            cg.p_new = reinterpret_cast<double*>(ws.pring + (long long)(it % PSN_PEND_K) * cg.ring_stride);
        } else {
            cg.p_old = ws.p[it & 1];
            cg.p_new = ws.p[(it & 1) ^ 1];
        }
        bk.dir_apply(cg, B);
        if (fuse) {
            cg.r = rbuf[ws.rcur];
            cg.r_out = rbuf[ws.rcur ^ 1];
            bk.fused_update(plan, fa, cg, B);
            ws.rcur ^= 1;
            cg.r = rbuf[ws.rcur];          // flush / true residual / the next up-leg work on the new residual
        } else {
            bk.update(cg, B);
        }
        bk.state_after_update(cg, B);
        rebuilt = it >= check_from;        // the check block below may restart samples from a recomputed residual
        if (it >= check_from) {
            if (cg.xdefer) bk.flush_x(cg, B);
            bk.true_residual(cg, B);
            bk.state_after_check(cg, B, ws.counters + it);
            bk.fetch_counter_async(ws.counters + it, it);
        }
        if (fixed_iters > 0) {
            if (it + 1 >= start_it + fixed_iters) {
                ++it;
                break;
            }
        } else if (it - PSN_COUNTER_LAG >= check_from && bk.wait_counter(it - PSN_COUNTER_LAG) == 0) {
            ++it;
            break;
        }
    }
    return it;
}
