// psn_hostemu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the kernel bodies of superresolution_for_pdes_b200/csrc/psn_core.cuh and the
// streamed-path driver (psn_stream_driver.h) as plain host C++ with a single virtual thread per
// CTA, so that the CPU-only test tier can check indexing, halo widths, the temporal-blocking
// schedule and the per-sample state machine without a GPU.  The product package never loads
// this library; the shipped path is the CUDA build (libpsn_b200.so) and has no CPU fallback.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../superresolution_for_pdes_b200/csrc/psn_stream_driver.h"
#include "../../superresolution_for_pdes_b200/csrc/psn_block.cuh"

namespace {

long long g_fused_prologues = 0;
long long g_fused_launches = 0;   // how often the fused update + down-leg backend entry ran (asserted by the tests)

struct HostPlan {
    PsnPlan plan;
    std::vector<double> inv;
    HostPlan(int n, int nu) {
        // nu = 10 * nu_coarse + nu_level0 selects different sweep counts on the coarse levels (e.g. 32: 2 on level 0, 3 below)
        psn_plan_levels(&plan, n, nu % 10, 0, nu / 10);
        inv = psn_coarse_inverse(plan);
        plan.coarse_inv = inv.data();
    }
};

struct HostBackend {
    std::vector<double> smem;
    std::vector<int> counters_host;
    HostBackend() : smem(1 << 20) {}
    void zero_counters(int* c, int n) { memset(c, 0, sizeof(int) * n); counters_host.assign(n, 0); }
    void prologue(const PsnCgArgs& a, long long B) {
        PsnHostCtx ctx;
        for (long long b = 0; b < B; ++b)
            for (int ty = 0; ty < a.nty; ++ty)
                for (int tx = 0; tx < a.ntx; ++tx) psn_prologue_body(ctx, a, b, ty, tx);
    }
    void state_init(const PsnCgArgs& a, long long B) { for (long long b = 0; b < B; ++b) psn_state_init(a, b); }
    template <int NU, class TB, class TV>
    void strip_leg_t(const PsnPlan& plan, const PsnStripArgs& a, long long B, bool post, bool tma) {
        PsnWarp w;
        std::vector<unsigned char> ring(PsnTmaFeeder<NU, double, double>::smem_bytes() + 256);
        for (long long b = 0; b < B; ++b)
            for (int band = 0; band < a.nbands; ++band)
                for (int strip = 0; strip < a.nstrips; ++strip) {
                    std::fill(ring.begin(), ring.end(), (unsigned char)0xff);   // NaN-poison the staging ring
                    if (post) {
                        if (tma) { PsnTmaFeeder<NU, TB, TV> fd; fd.attach(ring.data()); psn_strip_post_body<NU, TB, TV>(w, plan, a, b, band, strip, fd); }
                        else if constexpr (sizeof(TV) == 8) { PsnRegFeeder<NU, 2> fd; psn_strip_post_body<NU, TB, TV>(w, plan, a, b, band, strip, fd); }
                    } else {
                        if (tma) { PsnTmaFeeder<NU, TB> fd; fd.attach(ring.data()); psn_strip_pre_body<NU, TB, TV>(w, plan, a, b, band, strip, fd); }
                        else if constexpr (sizeof(TV) == 8) { PsnRegFeeder<NU, 1> fd; psn_strip_pre_body<NU, TB, TV>(w, plan, a, b, band, strip, fd); }
                    }
                }
    }
    template <int NU>
    void strip_leg(const PsnPlan& plan, const PsnStripArgs& a, long long B, bool post, bool tma, bool f32) {
        if (tma && f32) {
            if (a.level == 0) strip_leg_t<NU, double, float>(plan, a, B, post, true);
            else strip_leg_t<NU, float, float>(plan, a, B, post, true);
        } else {
            strip_leg_t<NU, double, double>(plan, a, B, post, tma);
        }
    }
    void strip_dispatch(const PsnPlan& plan, const PsnStripArgs& a, long long B, bool post, bool tma, bool f32) {
        const int nu_l = plan.nul[a.level];
        if (nu_l == 1) strip_leg<1>(plan, a, B, post, tma, f32);
        else if (nu_l == 2) strip_leg<2>(plan, a, B, post, tma, f32);
        else strip_leg<3>(plan, a, B, post, tma, f32);
    }
    void leg_pre(const PsnPlan& plan, const PsnStreamCfg& cfg, const PsnLegArgs& g, long long B) {
        if (cfg.mode >= 1) { strip_dispatch(plan, psn_leg_to_strip(plan, cfg, g, false), B, false, cfg.mode == 2, cfg.mode == 2 && cfg.vf32); return; }
        const PsnTileArgs t = psn_leg_to_tile(plan, cfg, g);
        PsnHostCtx ctx;
        for (long long b = 0; b < B; ++b)
            for (int ty = 0; ty < t.nty; ++ty)
                for (int tx = 0; tx < t.ntx; ++tx) {
                    // poison the scratch so that reads of never-written halo cells show up as NaN
                    std::fill(smem.begin(), smem.begin() + psn_tile_pre_doubles(t.TH, t.TW, plan.nul[t.level]), std::nan(""));
                    psn_tile_pre_body(ctx, plan, t, b, ty, tx, smem.data());
                }
    }
    void leg_post(const PsnPlan& plan, const PsnStreamCfg& cfg, const PsnLegArgs& g, long long B) {
        if (cfg.mode >= 1) { strip_dispatch(plan, psn_leg_to_strip(plan, cfg, g, true), B, true, cfg.mode == 2, cfg.mode == 2 && cfg.vf32); return; }
        const PsnTileArgs t = psn_leg_to_tile(plan, cfg, g);
        PsnHostCtx ctx;
        for (long long b = 0; b < B; ++b)
            for (int ty = 0; ty < t.nty; ++ty)
                for (int tx = 0; tx < t.ntx; ++tx) {
                    std::fill(smem.begin(), smem.begin() + psn_tile_post_doubles(t.TH, t.TW, plan.nul[t.level]), std::nan(""));
                    psn_tile_post_body(ctx, plan, t, b, ty, tx, smem.data());
                }
    }
    void coarse_vcycle(const PsnPlan& plan, const PsnStreamCfg& cfg, const double* bg, double* ug, const PsnSampleState* st, long long B) {
        PsnHostCtx ctx;
        if (cfg.coarse_block && psn_block_ok_at<20>(plan, cfg.ls)) {   // register-blocked V-cycle on the 80 x 80 level and below
            std::vector<double> sm(psn_block_vcycle_smem_doubles<20>(plan, cfg.ls) + PsnBlockGeom<20>::XS + 64);
            for (long long b = 0; b < B; ++b) {
                std::fill(sm.begin(), sm.end(), std::nan(""));
                double* base = sm.data() + PsnBlockGeom<20>::XS;   // (the body hands the solver "base - XS")
                if (cfg.mode == 2 && cfg.vf32)
                    psn_block_vcycle_body<20>(ctx, plan, cfg.ls, reinterpret_cast<const float*>(bg), reinterpret_cast<float*>(ug), st, b, base,
                                              (const double*)nullptr, false);
                else
                    psn_block_vcycle_body<20>(ctx, plan, cfg.ls, bg, ug, st, b, base, (const double*)nullptr, false);
            }
            return;
        }
        for (long long b = 0; b < B; ++b) {
            if (cfg.mode == 2 && cfg.vf32)
                psn_coarse_vcycle_body(ctx, plan, cfg.ls, reinterpret_cast<const float*>(bg), reinterpret_cast<float*>(ug), st, b, smem.data());
            else
                psn_coarse_vcycle_body(ctx, plan, cfg.ls, bg, ug, st, b, smem.data());
        }
    }
    void dir_apply(const PsnCgArgs& a, long long B) {
        if (a.dir_mode >= 1) {
            PsnWarp w;
            std::vector<unsigned char> ring(PsnTmaFeeder<2, double, double>::smem_bytes() + 256);
            for (long long b = 0; b < B; ++b)
                for (int band = 0; band < a.dir_nbands; ++band)
                    for (int strip = 0; strip < a.dir_nstrips; ++strip) {
                        std::fill(ring.begin(), ring.end(), (unsigned char)0xff);
                        if (a.dir_mode == 2 && a.vf32 && a.pf32) { PsnTmaFeeder<2, float, float> fd; fd.attach(ring.data()); psn_dir_strip_body<float, float>(w, a, b, band, strip, fd); }
                        else if (a.dir_mode == 2 && a.vf32) { PsnTmaFeeder<2, float, double> fd; fd.attach(ring.data()); psn_dir_strip_body<float, double>(w, a, b, band, strip, fd); }
                        else if (a.dir_mode == 2) { PsnTmaFeeder<2, double, double> fd; fd.attach(ring.data()); psn_dir_strip_body<double, double>(w, a, b, band, strip, fd); }
                        else { PsnRegFeeder<2, 2> fd; psn_dir_strip_body<double, double>(w, a, b, band, strip, fd); }
                    }
            return;
        }
        PsnHostCtx ctx;
        for (long long b = 0; b < B; ++b)
            for (int ty = 0; ty < a.nty; ++ty)
                for (int tx = 0; tx < a.ntx; ++tx) psn_dir_apply_body(ctx, a, b, ty, tx);
    }
    void update(const PsnCgArgs& a, long long B) {
        if (a.upd_mode == 2) {
            typedef PsnTmaFeeder<2, double, double, double> Feeder;
            PsnWarp w;
            std::vector<unsigned char> ring(Feeder::smem_bytes() + 256);
            for (long long b = 0; b < B; ++b)
                for (int band = 0; band < a.upd_nbands; ++band)
                    for (int strip = 0; strip < a.upd_nstrips; ++strip) {
                        std::fill(ring.begin(), ring.end(), (unsigned char)0xff);
                        if (a.pf32 && a.xdefer) { PsnTmaFeeder<2, float, double> fd; fd.attach(ring.data()); psn_update_strip_body<float, true>(w, a, b, band, strip, fd); }
                        else if (a.pf32) { PsnTmaFeeder<2, float, double, double> fd; fd.attach(ring.data()); psn_update_strip_body<float, false>(w, a, b, band, strip, fd); }
                        else { Feeder fd; fd.attach(ring.data()); psn_update_strip_body<double, false>(w, a, b, band, strip, fd); }
                    }
            return;
        }
        PsnHostCtx ctx;
        for (long long b = 0; b < B; ++b)
            for (int ty = 0; ty < a.nty; ++ty)
                for (int tx = 0; tx < a.ntx; ++tx) psn_update_body(ctx, a, b, ty, tx);
    }
    template <int NU, int MODE>
    void fused_t(const PsnPlan& plan, const PsnStripArgs& a, const PsnCgArgs& cg, long long B) {
        typedef typename PsnUpreFeederSel<NU, MODE>::type Feeder;
        PsnWarp w;
        std::vector<unsigned char> ring(Feeder::smem_bytes() + 256);
        for (long long b = 0; b < B; ++b)
            for (int band = 0; band < a.nbands; ++band)
                for (int strip = 0; strip < a.nstrips; ++strip) {
                    std::fill(ring.begin(), ring.end(), (unsigned char)0xff);
                    Feeder fd; fd.attach(ring.data());
                    psn_strip_upre_body<NU, MODE>(w, plan, a, cg, b, band, strip, fd);
                }
    }
    void fused_update(const PsnPlan& plan, const PsnStripArgs& a, const PsnCgArgs& cg, long long B) {
        ++g_fused_launches;
        if (plan.nu == 1) fused_t<1, 0>(plan, a, cg, B);
        else if (plan.nu == 2) fused_t<2, 0>(plan, a, cg, B);
        else fused_t<3, 0>(plan, a, cg, B);
    }
    void fused_prologue(const PsnPlan& plan, const PsnStripArgs& a, const PsnCgArgs& cg, long long B) {
        ++g_fused_prologues;
        if (plan.nu == 1) fused_t<1, 1>(plan, a, cg, B);
        else if (plan.nu == 2) fused_t<2, 1>(plan, a, cg, B);
        else fused_t<3, 1>(plan, a, cg, B);
    }
    void flush_x(const PsnCgArgs& a, long long B) {
        PsnHostCtx ctx;
        for (long long b = 0; b < B; ++b)
            for (int blk = 0; blk < 3; ++blk) psn_flush_body(ctx, a, b, blk, 3);
    }
    void true_residual(const PsnCgArgs& a, long long B) {
        if (a.chk_mode == 2) {
            typedef PsnTmaFeeder<2, double, double> Feeder;
            PsnWarp w;
            std::vector<unsigned char> ring(Feeder::smem_bytes() + 256);
            for (long long b = 0; b < B; ++b)
                for (int band = 0; band < a.upd_nbands; ++band)
                    for (int strip = 0; strip < a.upd_nstrips; ++strip) {
                        std::fill(ring.begin(), ring.end(), (unsigned char)0xff);
                        Feeder fd; fd.attach(ring.data());
                        psn_true_residual_strip_body(w, a, b, band, strip, fd);
                    }
            return;
        }
        PsnHostCtx ctx;
        for (long long b = 0; b < B; ++b)
            for (int ty = 0; ty < a.nty; ++ty)
                for (int tx = 0; tx < a.ntx; ++tx) psn_true_residual_body(ctx, a, b, ty, tx);
    }
    void state_after_update(const PsnCgArgs& a, long long B) { for (long long b = 0; b < B; ++b) psn_state_after_update(a, b); }
    void state_after_check(const PsnCgArgs& a, long long B, int* counter) {
        for (long long b = 0; b < B; ++b) *counter += psn_state_after_check(a, b);
    }
    void fetch_counter_async(const int* c, int it) { counters_host[it] = *c; }
    int wait_counter(int it) { return counters_host[it]; }
};

PsnRhs make_rhs(const double* theta, const double* f, const double* sx, const double* sy, const double* k1,
                const double* k2, int mn, int step) {
    PsnRhs r;
    r.theta = theta; r.f = f; r.sx = sx; r.sy = sy; r.k1 = k1; r.k2 = k2; r.mn = mn; r.step = step;
    return r;
}

}  // namespace

extern "C" {

long long hostemu_fused_launches(void) { return g_fused_launches; }
long long hostemu_fused_prologues(void) { return g_fused_prologues; }

// levels / gammas of the plan, for direct comparison with the numpy prototype
int hostemu_plan(int n, int nu, int* ns, double* gams, double* inv_out) {
    HostPlan hp(n, nu);
    for (int l = 0; l < hp.plan.nlev; ++l) { ns[l] = hp.plan.lv[l].n; gams[l] = hp.plan.lv[l].gam; }
    if (inv_out) memcpy(inv_out, hp.inv.data(), hp.inv.size() * sizeof(double));
    return hp.plan.nlev;
}

// resident (one-CTA-per-sample) solver
int hostemu_solve_smem(int n, long long B, const double* theta, const double* f, const double* sx, const double* sy,
                       const double* k1, const double* k2, int mn, int step, double rtol, int max_iter, int nu,
                       double* u, int32_t* iters, double* relres, int32_t* status,
                       int sub, const int32_t* start_x, const int32_t* start_y,
                       double* u_fine, double* f_fine, double* u_coarse, double* f_coarse) {
    if (n * n > 8192) return -1;
    HostPlan hp(n, nu);
    PsnSolveArgs a;
    a.rhs = make_rhs(theta, f, sx, sy, k1, k2, mn, step);
    a.rtol = rtol; a.max_iter = max_iter; a.u = u; a.iters = iters; a.relres = relres; a.status = status;
    a.sub = sub; a.start_x = start_x; a.start_y = start_y;
    a.u_fine = u_fine; a.f_fine = f_fine; a.u_coarse = u_coarse; a.f_coarse = f_coarse;
    a.th_fine = nullptr; a.th_coarse = nullptr;
    std::vector<double> smem(psn_smem_solver_doubles(hp.plan) + 64);
    PsnHostCtx ctx;
    for (long long b = 0; b < B; ++b) {
        std::fill(smem.begin(), smem.end(), std::nan(""));
        psn_smem_solve_body<8192>(ctx, hp.plan, a, b, smem.data());
    }
    return 0;
}

// streamed solver (tile kernels + global CG kernels + coarse shared-memory V-cycle)
int hostemu_solve_stream(int n, long long B, const double* theta, const double* f, const double* sx, const double* sy,
                         const double* k1, const double* k2, int mn, int step, double rtol, int max_iter, int nu,
                         int TH, int TW, int ls, int mode, int band, double* u, int32_t* iters, double* relres, int32_t* status) {
    HostPlan hp(n, nu);
    const PsnPlan& plan = hp.plan;
    if (ls < 1 || ls >= plan.nlev) return -1;
    PsnStreamCfg cfg;
    // mode 3 = mode 2 (TMA-fed strips) with float storage of the V-cycle intermediates; mode 4 = mode 3 + fused update / down-leg;
    // mode 5 = mode 4 + fused prologue / first down-leg
    // mode 6 = mode 5 with the 80 x 80 level (ls) and everything below in the register-blocked V-cycle kernel
    cfg.ls = ls; cfg.TH = TH; cfg.TW = TW; cfg.mode = (mode >= 1 && psn_strip_ok(plan, ls)) ? (mode >= 3 ? 2 : mode) : 0;
    cfg.coarse_block = (mode >= 6 && psn_block_ok_at<20>(plan, ls)) ? 1 : 0;
    cfg.vf32 = (mode >= 3 && cfg.mode == 2 && psn_vf32_ok(plan, ls)) ? 1 : 0;
    cfg.dir_band = 0;
    cfg.upd_band = 0;
    cfg.upd_strip = (cfg.mode == 2) ? 1 : 0;
    cfg.pf32 = cfg.vf32;
    cfg.xdefer = cfg.pf32;
    cfg.fuse = (mode >= 4 && cfg.xdefer) ? (mode >= 5 ? 2 : 1) : 0;
    cfg.fuse_band = (band >= 32) ? ((band / 2 + 1) & ~1) : 0;   // a different split than the other level-0 kernels
    if (mode >= 3 && !cfg.vf32) return -2;
    for (int l = 0; l < PSN_MAX_LEVELS; ++l) cfg.band[l] = band;
    PsnStreamWs ws;
    memset(&ws, 0, sizeof(ws));
    std::vector<std::vector<double>> keep;
    auto alloc = [&](size_t cnt) { keep.emplace_back(cnt, std::nan("")); return keep.back().data(); };
    const size_t nn = (size_t)n * n;
    ws.cap = B;
    if (cfg.fuse) { ws.r2 = alloc(B * nn); ws.vs0 = alloc(B); }
    ws.r = alloc(B * nn); ws.q = alloc(B * nn); ws.z = alloc(B * nn); ws.p[0] = alloc(B * nn); ws.p[1] = alloc(B * nn);
    for (int l = 1; l <= ls; ++l) {
        size_t m = (size_t)plan.lv[l].n * plan.lv[l].n;
        ws.bl[l] = alloc(B * m); ws.ub[l] = alloc(B * m);
        if (l < ls) ws.ua[l] = alloc(B * m);
    }
    const int gx = (n + PSN_GT_COLS - 1) / PSN_GT_COLS, gy = (n + PSN_GT_ROWS - 1) / PSN_GT_ROWS;
    ws.part_rz = alloc((size_t)B * psn_rz_count(plan, cfg));
    ws.part_a = alloc((size_t)B * gx * gy);
    ws.part_pq = alloc((size_t)B * psn_pq_max(n));
    ws.part_b = alloc((size_t)B * gx * gy);
    ws.part_u = alloc((size_t)B * psn_pu_max(n, plan.nu));
    std::vector<float> pring;
    std::vector<PsnPend> pend(B);
    if (cfg.xdefer) {
        pring.assign((size_t)PSN_PEND_K * B * nn, std::nanf(""));
        ws.pring = pring.data();
        ws.pend = pend.data();
    }
    std::vector<PsnSampleState> st(B);
    ws.state = st.data();
    std::vector<int> counters(max_iter + 8);
    ws.counters = counters.data();
    HostBackend bk;
    PsnRhs rhs = make_rhs(theta, f, sx, sy, k1, k2, mn, step);
    int outer = psn_stream_solve(bk, plan, cfg, rhs, B, rtol, max_iter, u, ws);
    for (long long b = 0; b < B; ++b) {
        iters[b] = st[b].iters;
        status[b] = st[b].phase == PSN_PHASE_DONE ? st[b].status : -1;
        relres[b] = st[b].bb > 0 ? sqrt(st[b].rr / st[b].bb) : 0.0;
    }
    return outer;
}

// one V-cycle z = M^{-1} r on the full grid (scaled form), for comparison with the numpy prototype
int hostemu_vcycle(int n, int nu, const double* rin, double* zout) {
    HostPlan hp(n, nu);
    const PsnLevel& l0 = hp.plan.lv[0];
    std::vector<double> smem(2 * l0.size + psn_pool_doubles(hp.plan, 0), 0.0);
    double* sb = smem.data();
    double* su = sb + l0.size;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) sb[(i + 1) * l0.pitch + j + 1] = rin[i * n + j];
    PsnHostCtx ctx;
    psn_vcycle_smem(ctx, hp.plan, 0, sb, su, su + l0.size);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) zout[i * n + j] = su[(i + 1) * l0.pitch + j + 1];
    return 0;
}

// one streamed V-cycle z = M^{-1} r (levels 0..ls-1 by tile or strip kernels, the rest by the
// shared-memory coarse kernel), single sample; for unit comparison with hostemu_vcycle
int hostemu_stream_vcycle(int n, int nu, int TH, int TW, int ls, int mode, int band, const double* rin, double* zout, double* rz) {
    HostPlan hp(n, nu);
    const PsnPlan& plan = hp.plan;
    if (ls < 1 || ls >= plan.nlev) return -1;
    PsnStreamCfg cfg;
    // mode 3 = mode 2 (TMA-fed strips) with float storage of the V-cycle intermediates
    cfg.ls = ls; cfg.TH = TH; cfg.TW = TW; cfg.mode = (mode >= 1 && psn_strip_ok(plan, ls)) ? (mode == 3 ? 2 : mode) : 0;
    cfg.vf32 = (mode == 3 && cfg.mode == 2 && psn_vf32_ok(plan, ls)) ? 1 : 0;
    cfg.dir_band = 0;
    cfg.upd_band = 0;
    cfg.upd_strip = (cfg.mode == 2) ? 1 : 0;
    cfg.pf32 = cfg.vf32;
    cfg.xdefer = cfg.pf32;
    cfg.fuse = 0;
    cfg.fuse_band = 0;
    if (mode == 3 && !cfg.vf32) return -2;
    for (int l = 0; l < PSN_MAX_LEVELS; ++l) cfg.band[l] = band;
    std::vector<std::vector<double>> keep;
    auto alloc = [&](size_t cnt) { keep.emplace_back(cnt, std::nan("")); return keep.back().data(); };
    double* bl[PSN_MAX_LEVELS]; double* ua[PSN_MAX_LEVELS]; double* ub[PSN_MAX_LEVELS];
    for (int l = 0; l <= ls; ++l) {
        size_t m = (size_t)plan.lv[l].n * plan.lv[l].n;
        bl[l] = alloc(m); ua[l] = alloc(m); ub[l] = alloc(m);
    }
    memcpy(bl[0], rin, sizeof(double) * n * n);
    std::vector<double> part(psn_rz_count(plan, cfg), std::nan(""));
    HostBackend bk;
    for (int l = 0; l < ls; ++l) {
        PsnLegArgs g{l, bl[l], nullptr, ua[l], bl[l + 1], nullptr, nullptr, nullptr};
        bk.leg_pre(plan, cfg, g, 1);
    }
    bk.coarse_vcycle(plan, cfg, bl[ls], ub[ls], nullptr, 1);
    for (int l = ls - 1; l >= 0; --l) {
        PsnLegArgs g{l, bl[l], ua[l], ub[l], nullptr, ub[l + 1], l == 0 ? part.data() : nullptr, nullptr};
        bk.leg_post(plan, cfg, g, 1);
    }
    if (cfg.vf32) {
        const float* zf = reinterpret_cast<const float*>(ub[0]);
        for (int i = 0; i < n * n; ++i) zout[i] = (double)zf[i];
    } else {
        memcpy(zout, ub[0], sizeof(double) * n * n);
    }
    double s = 0; for (double v : part) s += v;
    *rz = s;
    return 0;
}

// order in which the emulated threads of the register-blocked solver run inside a phase (0 forward, 1 backward)
void hostemu_block_thread_order(int reverse) { psn_bk_reverse_order() = reverse; }

// register-blocked resident solver (n = 80), all threads of the CTA emulated phase by phase
int hostemu_solve_block(int n, long long B, const double* theta, const double* f, const double* sx, const double* sy,
                        int mn, int step, double rtol, int max_iter, int nu,
                        double* u, int32_t* iters, double* relres, int32_t* status,
                        int sub, const int32_t* start_x, const int32_t* start_y,
                        double* u_fine, double* f_fine, double* u_coarse, double* f_coarse) {
    if (n != 80 && n != 40 && n != 20) return -1;
    HostPlan hp(n, nu);
    if (n == 20 ? !psn_block_ok<5>(hp.plan) : n == 40 ? !psn_block_ok<10>(hp.plan) : !psn_block_ok<20>(hp.plan)) return -2;
    PsnSolveArgs a;
    a.rhs = make_rhs(theta, f, sx, sy, nullptr, nullptr, mn, step);
    a.rtol = rtol; a.max_iter = max_iter; a.u = u; a.iters = iters; a.relres = relres; a.status = status;
    a.sub = sub; a.start_x = start_x; a.start_y = start_y;
    a.u_fine = u_fine; a.f_fine = f_fine; a.u_coarse = u_coarse; a.f_coarse = f_coarse;
    a.th_fine = nullptr; a.th_coarse = nullptr;
    std::vector<double> smem((n == 20 ? psn_block_smem_doubles<5>(hp.plan)
                                      : n == 40 ? psn_block_smem_doubles<10>(hp.plan) : psn_block_smem_doubles<20>(hp.plan)) + 64);
    PsnHostCtx ctx;
    for (long long b = 0; b < B; ++b) {
        std::fill(smem.begin(), smem.end(), std::nan(""));
        if (n == 20)
            psn_block_solve_body<5>(ctx, hp.plan, a, b, smem.data());
        else if (n == 40)
            psn_block_solve_body<10>(ctx, hp.plan, a, b, smem.data());
        else
            psn_block_solve_body<20>(ctx, hp.plan, a, b, smem.data());
    }
    return 0;
}
}
