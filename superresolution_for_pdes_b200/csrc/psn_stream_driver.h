// psn_stream_driver.h -- host-side orchestration of the streamed (large-grid) MG-PCG solve.
//
// Plain C++, templated over a backend `BK` that knows how to run each kernel body over a
// batch: the CUDA backend (psn_api.cu) launches kernels on a stream, the host-emulation
// backend (tests/hostemu, TEST ONLY) loops over the blocks on the CPU.  The driver never
// synchronises with the device inside an iteration: alpha/beta live in per-sample device state,
// converged samples are masked out per CTA, and the number of still-active samples is read
// back with a lag of PSN_COUNTER_LAG iterations.
#pragma once
#include "psn_core.cuh"

#define PSN_COUNTER_LAG 2

struct PsnStreamCfg {
    int ls;       // first level handled by the shared-memory coarse kernel (levels 0..ls-1 are streamed)
    int TH, TW;   // interior tile of the pre / post kernels
};

// device arrays for one chunk of Bc samples
struct PsnStreamWs {
    long long cap;               // samples this workspace was sized for
    double* r;                   // level 0: residual (= rhs of the V-cycle)
    double* q;                   // level 0: A p  (doubles as the pre-smoothed iterate)
    double* z;                   // level 0: preconditioned residual
    double* p[2];                // level 0: search direction, ping-pong
    double* bl[PSN_MAX_LEVELS];  // levels 1..ls: rhs
    double* ua[PSN_MAX_LEVELS];  // levels 1..ls-1: pre-smoothed iterate
    double* ub[PSN_MAX_LEVELS];  // levels 1..ls: corrected iterate
    double* part_rz;
    double* part_a;
    double* part_b;
    PsnSampleState* state;
    int* counters;               // [max_iter + 8] active-sample count per outer iteration
};

static inline void psn_tile_counts(int n, int TH, int TW, int* ntx, int* nty) {
    *ntx = (n + TW - 1) / TW;
    *nty = (n + TH - 1) / TH;
}

// x: [B, n, n] device array that receives the solution.  Returns the number of outer iterations run.
template <class BK>
int psn_stream_solve(BK& bk, const PsnPlan& plan, const PsnStreamCfg& cfg, const PsnRhs& rhs, long long B,
                     double rtol, int max_iter, double* x, PsnStreamWs& ws) {
    const int n = plan.n;
    PsnCgArgs cg;
    cg.n = n;
    cg.ntx = (n + PSN_GT_COLS - 1) / PSN_GT_COLS;
    cg.nty = (n + PSN_GT_ROWS - 1) / PSN_GT_ROWS;
    int t0x, t0y;
    psn_tile_counts(n, cfg.TH, cfg.TW, &t0x, &t0y);
    cg.ntiles_rz = t0x * t0y;
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
    cg.part_b = ws.part_b;
    cg.state = ws.state;
    cg.rtol = rtol;
    cg.max_iter = max_iter;

    bk.zero_counters(ws.counters, max_iter + 8);
    bk.prologue(cg, B);
    bk.state_init(cg, B);

    const int max_outer = max_iter + 4;  // every check/restart costs at most one extra outer pass
    int it = 0;
    for (; it < max_outer; ++it) {
        // ---- z = M^{-1} r : one V(nu,nu) cycle
        for (int l = 0; l < cfg.ls; ++l) {
            PsnTileArgs t;
            t.level = l;
            t.TH = cfg.TH;
            t.TW = cfg.TW;
            psn_tile_counts(plan.lv[l].n, cfg.TH, cfg.TW, &t.ntx, &t.nty);
            t.b = (l == 0) ? ws.r : ws.bl[l];
            t.uin = nullptr;
            t.uout = (l == 0) ? ws.q : ws.ua[l];
            t.bc = ws.bl[l + 1];
            t.ec = nullptr;
            t.part = nullptr;
            t.state = ws.state;
            bk.tile_pre(plan, t, B);
        }
        bk.coarse_vcycle(plan, cfg.ls, ws.bl[cfg.ls], ws.ub[cfg.ls], ws.state, B);
        for (int l = cfg.ls - 1; l >= 0; --l) {
            PsnTileArgs t;
            t.level = l;
            t.TH = cfg.TH;
            t.TW = cfg.TW;
            psn_tile_counts(plan.lv[l].n, cfg.TH, cfg.TW, &t.ntx, &t.nty);
            t.b = (l == 0) ? ws.r : ws.bl[l];
            t.uin = (l == 0) ? ws.q : ws.ua[l];
            t.uout = (l == 0) ? ws.z : ws.ub[l];
            t.bc = nullptr;
            t.ec = ws.ub[l + 1];
            t.part = (l == 0) ? ws.part_rz : nullptr;
            t.state = ws.state;
            bk.tile_post(plan, t, B);
        }
        // ---- CG update
        cg.p_old = ws.p[it & 1];
        cg.p_new = ws.p[(it & 1) ^ 1];
        bk.dir_apply(cg, B);
        bk.update(cg, B);
        bk.state_after_update(cg, B);
        bk.true_residual(cg, B);
        bk.state_after_check(cg, B, ws.counters + it);
        bk.fetch_counter_async(ws.counters + it, it);
        if (it >= PSN_COUNTER_LAG && bk.wait_counter(it - PSN_COUNTER_LAG) == 0) {
            ++it;
            break;
        }
    }
    return it;
}
