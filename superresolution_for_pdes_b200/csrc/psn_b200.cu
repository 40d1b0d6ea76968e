// psn_b200.cu -- CUDA kernels (sm_100a) and the C ABI of libpsn_b200.so.
//
// Two regimes:
//   resident  n*n <= 8192 and the hierarchy fits 227 KB of shared memory (n <= ~86):
//             ONE kernel, one CTA per sample, the whole MG-PCG solve in shared memory/registers.
//   streamed  larger grids: batched kernels over [B, n, n] arrays in HBM, temporally blocked
//             shared-memory tiles for the V-cycle legs, per-sample scalars and convergence state
//             on the device, no host synchronisation inside an iteration.
// The numerical bodies live in psn_core.cuh (shared with the CPU-only test harness).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <chrono>
#include <vector>

#include "../../include/psn_b200.h"
#include "psn_stream_driver.h"
#include "psn_block.cuh"
#include "psn_mt19937.cuh"

#define PSN_VERSION 100

// ----------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

static int psn_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                            \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            cudaGetLastError();                                                                 \
            return psn_fail(e_ == cudaErrorMemoryAllocation ? PSN_ERR_NOMEM : PSN_ERR_CUDA,     \
                            "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                       \
    } while (0)

// ----------------------------------------------------------------------------- kernels
// Programmatic dependent launch: every kernel of the streamed solve waits here for the grid launched before it on the
// stream (full completion + memory flush) and immediately allows the NEXT grid to be launched, so that the next
// kernel's CTAs are already resident (waiting at this instruction) while this kernel's last CTAs drain.  Launched
// without the attribute (option "pdl" = 0, or the profiler's event pairs in between) both instructions are no-ops.
__device__ __forceinline__ void psn_pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <int T, int XPT>
__global__ void __launch_bounds__(T) psn_smem_solve_kernel(const __grid_constant__ PsnPlan plan,
                                                           const __grid_constant__ PsnSolveArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, T, red, 0};
    psn_smem_solve_body<XPT>(ctx, plan, a, (long long)blockIdx.x, smem);
}

// register-blocked resident solver (n = 4*BX): one CTA per SM, persistent over the samples b = blockIdx.x, + gridDim.x, ...;
// 4x4 nodes per thread in registers.  inv_doubles > 0: the dense inverse of the coarsest level is copied into shared
// memory once per CTA (behind the solver's arrays) -- read through L2 it was the top stall of the kernel
// (long scoreboard 3.2 of 11 stall cycles per issue).
// (128 registers: the register file is split per scheduler, 16 K each, and one scheduler holds 4 of the 13 warps --
//  __maxnreg__(152) compiles but cannot be launched)
// BX = 10 (40 x 40): 128 threads and ~46 KB per sample, four samples per SM (three with 168 registers and no spills: 6.2M
// instead of 6.6M solves/s).
// BX = 5 (20 x 20): one warp and ~18 KB per sample.  Eight samples per SM: 206 registers and no spills.  Shared memory would
// allow eleven, but then the register cap is 168, the kernel spills 128 B per thread, and with 220 KB of the SM's 256 KB
// carved out as shared memory the spill traffic no longer fits the remaining L1: 3.7M - 17M solves/s from run to run
// instead of a steady 25M (10 per SM: 25.3M with 120 B of spills; 9: 23.5M).
#ifndef PSN_BLOCK_CTAS5
#define PSN_BLOCK_CTAS5 8
#endif
#define PSN_BLOCK_CTAS(BX) ((BX) >= 20 ? 1 : (BX) >= 10 ? 4 : PSN_BLOCK_CTAS5)
template <int BX>
__global__ void __launch_bounds__(PsnBlockGeom<BX>::NTPAD, PSN_BLOCK_CTAS(BX)) psn_block_solve_kernel(const __grid_constant__ PsnPlan plan,
                                                                                     const __grid_constant__ PsnSolveArgs a,
                                                                                     long long B, int inv_doubles) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, PsnBlockGeom<BX>::NTPAD, red, 0};
    double* inv_s = nullptr;
    if (inv_doubles > 0) {
        inv_s = smem + psn_block_smem_doubles<BX>(plan);
        for (int i = threadIdx.x; i < inv_doubles; i += PsnBlockGeom<BX>::NTPAD) inv_s[i] = __ldg(plan.coarse_inv + i);
        __syncthreads();
    }
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        psn_block_solve_body<BX>(ctx, plan, a, b, smem, inv_s);
        __syncthreads();   // the next sample re-initialises the shared arrays
    }
}

// Coarse tail of the streamed solver when it has an 80 x 80 level: the register-blocked V-cycle of psn_block.cuh on that level
// and everything below it, one persistent CTA per SM; the dense inverse and the zeroed edge planes are set up before the
// programmatic-dependent-launch wait, i.e. under the tail of the kernel in front.
template <int BX, class TV>
__global__ void __launch_bounds__(PsnBlockGeom<BX>::NTPAD, PSN_BLOCK_CTAS(BX)) psn_block_vcycle_kernel(const __grid_constant__ PsnPlan plan, int L0,
                                                                                                      const TV* bg, TV* ug,
                                                                                                      const PsnSampleState* state, long long B,
                                                                                                      int inv_doubles) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, PsnBlockGeom<BX>::NTPAD, red, 0};
    const int own = psn_block_vcycle_smem_doubles<BX>(plan, L0);
    double* inv_s = nullptr;
    if (inv_doubles > 0) {
        inv_s = smem + own;
        for (int i = threadIdx.x; i < inv_doubles; i += PsnBlockGeom<BX>::NTPAD) inv_s[i] = __ldg(plan.coarse_inv + i);
    }
    for (int i = threadIdx.x; i < own; i += PsnBlockGeom<BX>::NTPAD) smem[i] = 0.0;
    __syncthreads();
    psn_pdl_enter();
    for (long long b = blockIdx.x; b < B; b += gridDim.x) psn_block_vcycle_body<BX>(ctx, plan, L0, bg, ug, state, b, smem, inv_s, true);
}

#define PSN_TILE_THREADS 256
#define PSN_COARSE_THREADS 512
#define PSN_CG_THREADS 256

__global__ void __launch_bounds__(PSN_TILE_THREADS) psn_tile_pre_kernel(const __grid_constant__ PsnPlan plan,
                                                                        const __grid_constant__ PsnTileArgs a) {
    psn_pdl_enter();
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, PSN_TILE_THREADS, red, 0};
    psn_tile_pre_body(ctx, plan, a, (long long)blockIdx.z, (int)blockIdx.y, (int)blockIdx.x, smem);
}

__global__ void __launch_bounds__(PSN_TILE_THREADS) psn_tile_post_kernel(const __grid_constant__ PsnPlan plan,
                                                                         const __grid_constant__ PsnTileArgs a) {
    psn_pdl_enter();
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, PSN_TILE_THREADS, red, 0};
    psn_tile_post_body(ctx, plan, a, (long long)blockIdx.z, (int)blockIdx.y, (int)blockIdx.x, smem);
}

// register-resident warp-strip V-cycle legs: one warp per (sample, band, strip) task, no shared memory
#define PSN_STRIP_WARPS 4
#ifndef PSN_STRIP_MINB
#define PSN_STRIP_MINB 2
#endif
#ifndef PSN_STRIP_MINB_TMA
#define PSN_STRIP_MINB_TMA 3
#endif
#ifndef PSN_STRIP_MINB_F32
#define PSN_STRIP_MINB_F32 4   // float strips hold half the register window: 4 CTAs (16 warps) per SM
#endif
// TB / TV: element types of the right-hand side and of the V-cycle intermediates (see PsnStripArgs)
template <int NU, bool POST, bool TMA, class TB, class TV>
__global__ void __launch_bounds__(PSN_STRIP_WARPS * 32, TMA ? (sizeof(TV) == 4 ? PSN_STRIP_MINB_F32 : PSN_STRIP_MINB_TMA) : PSN_STRIP_MINB)
psn_strip_kernel(const __grid_constant__ PsnPlan plan, const __grid_constant__ PsnStripArgs a, long long ntasks) {
    psn_pdl_enter();
    extern __shared__ __align__(128) unsigned char strip_smem[];
    // broadcast from lane 0 so that the compiler treats the warp index (and the task, strip, band and every TMA
    // address derived from it) as warp-uniform: uniform registers for UBLKCP operands instead of an election loop
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const long long task = (long long)blockIdx.x * PSN_STRIP_WARPS + warp;
    if (task >= ntasks) return;
    const int per = a.nstrips * a.nbands;
    const long long b = task / per;
    const int rem = (int)(task - b * per);
    const int band = rem / a.nstrips, strip = rem - band * a.nstrips;
    const PsnWarp w{(int)(threadIdx.x & 31)};
    constexpr int NARR = POST ? 2 : 1;
    if constexpr (TMA) {
        typedef typename PsnStripFeederSel<NU, POST, TB, TV>::type Feeder;
        Feeder fd;
        fd.attach(strip_smem + (size_t)warp * Feeder::smem_bytes());
        if constexpr (POST) psn_strip_post_body<NU, TB, TV>(w, plan, a, b, band, strip, fd);
        else psn_strip_pre_body<NU, TB, TV>(w, plan, a, b, band, strip, fd);
    } else {
        PsnRegFeeder<NU, NARR> fd;
        if constexpr (POST) psn_strip_post_body<NU, TB, TV>(w, plan, a, b, band, strip, fd);
        else psn_strip_pre_body<NU, TB, TV>(w, plan, a, b, band, strip, fd);
    }
}

template <bool TMA, class TZ, class TP>
__global__ void __launch_bounds__(PSN_STRIP_WARPS * 32, 3)
psn_dir_strip_kernel(const __grid_constant__ PsnCgArgs a, long long ntasks) {
    psn_pdl_enter();
    extern __shared__ __align__(128) unsigned char strip_smem[];
    // broadcast from lane 0 so that the compiler treats the warp index (and the task, strip, band and every TMA
    // address derived from it) as warp-uniform: uniform registers for UBLKCP operands instead of an election loop
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const long long task = (long long)blockIdx.x * PSN_STRIP_WARPS + warp;
    if (task >= ntasks) return;
    const int per = a.dir_nstrips * a.dir_nbands;
    const long long b = task / per;
    const int rem = (int)(task - b * per);
    const int band = rem / a.dir_nstrips, strip = rem - band * a.dir_nstrips;
    const PsnWarp w{(int)(threadIdx.x & 31)};
    if constexpr (TMA) {
        PsnTmaFeeder<2, TZ, TP> fd;
        fd.attach(strip_smem + (size_t)warp * PsnTmaFeeder<2, TZ, TP>::smem_bytes());
        psn_dir_strip_body<TZ, TP>(w, a, b, band, strip, fd);
    } else {
        PsnRegFeeder<2, 2> fd;
        psn_dir_strip_body<TZ, TP>(w, a, b, band, strip, fd);
    }
}

// CG update of iteration k fused with the level-0 down-leg of iteration k+1 (psn_strip_upre_body)
#ifndef PSN_UPRE_MINB
#define PSN_UPRE_MINB 3
#endif
template <int NU, int MODE>
__global__ void __launch_bounds__(PSN_STRIP_WARPS * 32, PSN_UPRE_MINB)
psn_upre_strip_kernel(const __grid_constant__ PsnPlan plan, const __grid_constant__ PsnStripArgs a, const __grid_constant__ PsnCgArgs cg,
                      long long ntasks) {
    psn_pdl_enter();
    extern __shared__ __align__(128) unsigned char strip_smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const long long task = (long long)blockIdx.x * PSN_STRIP_WARPS + warp;
    if (task >= ntasks) return;
    const int per = a.nstrips * a.nbands;
    const long long b = task / per;
    const int rem = (int)(task - b * per);
    const int band = rem / a.nstrips, strip = rem - band * a.nstrips;
    const PsnWarp w{(int)(threadIdx.x & 31)};
    typedef typename PsnUpreFeederSel<NU, MODE>::type Feeder;
    Feeder fd;
    fd.attach(strip_smem + (size_t)warp * Feeder::smem_bytes());
    psn_strip_upre_body<NU, MODE>(w, plan, a, cg, b, band, strip, fd);
}

__global__ void __launch_bounds__(256) psn_flush_x_kernel(const __grid_constant__ PsnCgArgs a) {
    psn_pdl_enter();
    PsnDevCtx ctx{(int)threadIdx.x, 256, nullptr, 0};
    psn_flush_body(ctx, a, (long long)blockIdx.y, (int)blockIdx.x, (int)gridDim.x);
}

template <class TP, bool XDEFER>
__global__ void __launch_bounds__(PSN_STRIP_WARPS * 32, 3)
psn_update_strip_kernel(const __grid_constant__ PsnCgArgs a, long long ntasks) {
    psn_pdl_enter();
    extern __shared__ __align__(128) unsigned char strip_smem[];
    // broadcast from lane 0 so that the compiler treats the warp index (and the task, strip, band and every TMA
    // address derived from it) as warp-uniform: uniform registers for UBLKCP operands instead of an election loop
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const long long task = (long long)blockIdx.x * PSN_STRIP_WARPS + warp;
    if (task >= ntasks) return;
    const int per = a.upd_nstrips * a.upd_nbands;
    const long long b = task / per;
    const int rem = (int)(task - b * per);
    const int band = rem / a.upd_nstrips, strip = rem - band * a.upd_nstrips;
    const PsnWarp w{(int)(threadIdx.x & 31)};
    typedef typename PsnUpdateFeederSel<TP, XDEFER>::type Feeder;
    Feeder fd;
    fd.attach(strip_smem + (size_t)warp * Feeder::smem_bytes());
    psn_update_strip_body<TP, XDEFER>(w, a, b, band, strip, fd);
}

__global__ void __launch_bounds__(PSN_STRIP_WARPS * 32, 3)
psn_true_residual_strip_kernel(const __grid_constant__ PsnCgArgs a, long long ntasks) {
    psn_pdl_enter();
    extern __shared__ __align__(128) unsigned char strip_smem[];
    // broadcast from lane 0 so that the compiler treats the warp index (and the task, strip, band and every TMA
    // address derived from it) as warp-uniform: uniform registers for UBLKCP operands instead of an election loop
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const long long task = (long long)blockIdx.x * PSN_STRIP_WARPS + warp;
    if (task >= ntasks) return;
    const int per = a.upd_nstrips * a.upd_nbands;
    const long long b = task / per;
    const int rem = (int)(task - b * per);
    const int band = rem / a.upd_nstrips, strip = rem - band * a.upd_nstrips;
    const PsnWarp w{(int)(threadIdx.x & 31)};
    typedef PsnTmaFeeder<2, double, double> Feeder;
    Feeder fd;
    fd.attach(strip_smem + (size_t)warp * Feeder::smem_bytes());
    psn_true_residual_strip_body(w, a, b, band, strip, fd);
}

template <class TV>
__global__ void __launch_bounds__(PSN_COARSE_THREADS) psn_coarse_vcycle_kernel(const __grid_constant__ PsnPlan plan, int l0,
                                                                               const TV* bg, TV* ug,
                                                                               const PsnSampleState* state) {
    psn_pdl_enter();
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, PSN_COARSE_THREADS, red, 0};
    psn_coarse_vcycle_body(ctx, plan, l0, bg, ug, state, (long long)blockIdx.x, smem);
}

__global__ void __launch_bounds__(PSN_CG_THREADS) psn_prologue_kernel(const __grid_constant__ PsnCgArgs a) {
    psn_pdl_enter();
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, PSN_CG_THREADS, red, 0};
    psn_prologue_body(ctx, a, (long long)blockIdx.z, (int)blockIdx.y, (int)blockIdx.x);
}
__global__ void __launch_bounds__(PSN_CG_THREADS) psn_dir_apply_kernel(const __grid_constant__ PsnCgArgs a) {
    psn_pdl_enter();
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, PSN_CG_THREADS, red, 0};
    psn_dir_apply_body(ctx, a, (long long)blockIdx.z, (int)blockIdx.y, (int)blockIdx.x);
}
__global__ void __launch_bounds__(PSN_CG_THREADS) psn_update_kernel(const __grid_constant__ PsnCgArgs a) {
    psn_pdl_enter();
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, PSN_CG_THREADS, red, 0};
    psn_update_body(ctx, a, (long long)blockIdx.z, (int)blockIdx.y, (int)blockIdx.x);
}
__global__ void __launch_bounds__(PSN_CG_THREADS) psn_true_residual_kernel(const __grid_constant__ PsnCgArgs a) {
    psn_pdl_enter();
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, PSN_CG_THREADS, red, 0};
    psn_true_residual_body(ctx, a, (long long)blockIdx.z, (int)blockIdx.y, (int)blockIdx.x);
}
__global__ void psn_state_init_kernel(const __grid_constant__ PsnCgArgs a, long long B) {
    psn_pdl_enter();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) psn_state_init(a, b);
}
__global__ void psn_state_after_update_kernel(const __grid_constant__ PsnCgArgs a, long long B) {
    psn_pdl_enter();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) psn_state_after_update(a, b);
}
__global__ void psn_state_after_check_kernel(const __grid_constant__ PsnCgArgs a, long long B, int* counter) {
    psn_pdl_enter();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int active = 0;
    if (b < B) active = psn_state_after_check(a, b);
    const unsigned m = __ballot_sync(0xffffffffu, active);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(counter, __popc(m));
}
// The convergence counter goes to the host through mapped pinned memory, not through a D2H copy: a copy would
// queue behind the multi-hundred-MB result copies of the previous chunk on the same copy engine and stall the solve.
__global__ void psn_publish_counter_kernel(const int* counter, int* host_slot) {
    psn_pdl_enter();
    *reinterpret_cast<volatile int*>(host_slot) = *counter;
    __threadfence_system();
}
__global__ void psn_zero_ints_kernel(int* p, int n) {
    psn_pdl_enter();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0;
}
__global__ void psn_finalize_kernel(const PsnSampleState* state, long long B, int32_t* iters, double* relres, int32_t* status) {
    psn_pdl_enter();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const PsnSampleState st = state[b];
    if (iters) iters[b] = st.iters;
    if (relres) relres[b] = st.bb > 0.0 ? sqrt(st.rr / st.bb) : 0.0;
    if (status) status[b] = st.phase == PSN_PHASE_DONE ? st.status : PSN_ST_MAXITER;
}
// Sub-domain epilogue of the STREAMED path (the resident kernel fuses it): window extract + stride-2
// downsample of u, f, theta for one sample per blockIdx.y  (src/enhanced_data_generation.py:141-148)
__global__ void psn_window_kernel(const __grid_constant__ PsnSolveArgs a, int n, const double* u) {
    psn_pdl_enter();
    const long long b = blockIdx.y;
    const int s = a.sub, sc = (s + 1) >> 1;
    const int x0 = a.start_x[b], y0 = a.start_y[b];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < s * s; idx += gridDim.x * blockDim.x) {
        const int wi = idx / s, wj = idx - wi * s;
        const int i = y0 + wi, j = x0 + wj;
        const double uv = u[(b * n + i) * (long long)n + j];
        const double fv = psn_forcing(a.rhs, b, i, j);
        const double tv = a.rhs.theta ? a.rhs.theta[(b * a.rhs.mn + (long long)i * a.rhs.step) * a.rhs.mn + (long long)j * a.rhs.step] : 1.0;
        const long long o = (b * s + wi) * s + wj;
        if (a.u_fine) a.u_fine[o] = uv;
        if (a.f_fine) a.f_fine[o] = fv;
        if (a.th_fine) a.th_fine[o] = tv;
        if (((wi | wj) & 1) == 0) {
            const long long oc = (b * sc + (wi >> 1)) * sc + (wj >> 1);
            if (a.u_coarse) a.u_coarse[oc] = uv;
            if (a.f_coarse) a.f_coarse[oc] = fv;
            if (a.th_coarse) a.th_coarse[oc] = tv;
        }
    }
}
__global__ void psn_fill_kernel(double* p, long long cnt, double v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}


// ----------------------------------------------------------------------------- consumer prologue (row f1)
// moments of the float32-cast data: per-block partial (sum, sum of squares) in double, fixed-order final sum
#define PSN_MOM_BLOCKS 592
__global__ void __launch_bounds__(256) psn_moments_kernel(const double* a, long long cnt, double* partial) {
    __shared__ double red[64];
    PsnDevCtx ctx{(int)threadIdx.x, 256, red, 0};
    double s = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < cnt; i += (long long)gridDim.x * 256) {
        const double v = (double)(float)a[i];
        s += v;
        s2 += v * v;
    }
    const double ts = ctx.block_sum(s);
    const double ts2 = ctx.block_sum(s2);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = ts;
        partial[2 * blockIdx.x + 1] = ts2;
    }
}
// stats[2k] = mean, stats[2k+1] = unbiased std of array k (k = 0: u_fine, 1: f_fine, 2: theta_fine)
__global__ void psn_moments_final_kernel(const double* partial, int nblocks, long long cnt, double* stats) {
    const int k = threadIdx.x;
    if (k >= 3) return;
    const double* p = partial + (size_t)k * 2 * nblocks;
    double s = 0.0, s2 = 0.0;
    for (int b = 0; b < nblocks; ++b) {
        s += p[2 * b];
        s2 += p[2 * b + 1];
    }
    const double mean = s / (double)cnt;
    const double var = cnt > 1 ? fmax(0.0, (s2 - (double)cnt * mean * mean) / (double)(cnt - 1)) : 0.0;
    stats[2 * k] = mean;
    stats[2 * k + 1] = sqrt(var);
    if (k == 2) {
        const bool constant = sqrt(var) < 1e-6;
        stats[6] = constant ? 1.0 : 0.0;
        stats[7] = 0.0;
        if (constant) {   // src/models.py:163-165: theta passes through unchanged
            stats[4] = 0.0;
            stats[5] = 1.0;
        }
    }
}
__global__ void __launch_bounds__(256) psn_tensor_kernel(long long n, int nc, int nf, const double* uc, const double* uf,
                                                         const double* ff, const double* th, const double* stats,
                                                         float* x, float* y) {
    const long long total = n * nf * nf;
    const float um = (float)stats[0], us = (float)stats[1], fm = (float)stats[2], fs = (float)stats[3];
    const float tm = (float)stats[4], ts = (float)stats[5];
    const bool tconst = stats[6] != 0.0;
    const float scale = nf > 1 ? (float)(nc - 1) / (float)(nf - 1) : 0.f;   // align_corners = True
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const long long b = idx / (nf * nf);
        const int rem = (int)(idx - b * nf * nf);
        const int i = rem / nf, j = rem - i * nf;
        // bilinear sample of the normalised coarse solution
        const float sy = scale * (float)i, sx = scale * (float)j;
        const int i0 = (int)sy, j0 = (int)sx;
        const int i1 = min(i0 + 1, nc - 1), j1 = min(j0 + 1, nc - 1);
        const float ly = sy - (float)i0, lx = sx - (float)j0;
        const double* c = uc + b * (long long)nc * nc;
        const float v00 = ((float)c[i0 * nc + j0] - um) / us, v01 = ((float)c[i0 * nc + j1] - um) / us;
        const float v10 = ((float)c[i1 * nc + j0] - um) / us, v11 = ((float)c[i1 * nc + j1] - um) / us;
        const float up = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
        const float tv = (float)th[idx];
        float* xo = x + (b * 3) * (long long)nf * nf + rem;
        xo[0] = up;
        xo[(long long)nf * nf] = tconst ? tv : (tv - tm) / ts;
        xo[2ll * nf * nf] = ((float)ff[idx] - fm) / fs;
        y[idx] = ((float)uf[idx] - um) / us;
    }
}


// ----------------------------------------------------------------------------- interpolation baselines (row f4)
__device__ __forceinline__ void psn_cubic_coeffs(float t, float (&w)[4]) {
    const float A = -0.75f;
    float x = t + 1.f;
    w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
    x = t;
    w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 1.f - t;
    w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 2.f - t;
    w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}
// one x2 level, align_corners = True; SRC = double for the first level (cast to float32 on load), float after
template <class SRC>
__global__ void __launch_bounds__(256) psn_upscale2_kernel(int mode, int n, long long batch, const SRC* in, float* out) {
    const int m = 2 * n;
    const long long total = batch * m * m;
    const float scale = (float)(n - 1) / (float)(m - 1);
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const long long b = idx / ((long long)m * m);
        const int rem = (int)(idx - b * (long long)m * m);
        const int oy = rem / m, ox = rem - oy * m;
        const SRC* p = in + b * (long long)n * n;
        const float ry = scale * (float)oy, rx = scale * (float)ox;
        float v;
        if (mode == 0) {
            const int y0 = (int)ry, x0 = (int)rx;
            const int y1 = min(y0 + 1, n - 1), x1 = min(x0 + 1, n - 1);
            const float ly = ry - (float)y0, lx = rx - (float)x0;
            v = (1.f - ly) * ((1.f - lx) * (float)p[y0 * n + x0] + lx * (float)p[y0 * n + x1]) +
                ly * ((1.f - lx) * (float)p[y1 * n + x0] + lx * (float)p[y1 * n + x1]);
        } else {
            const int iy = (int)floorf(ry), ix = (int)floorf(rx);
            float wy[4], wx[4];
            psn_cubic_coeffs(ry - (float)iy, wy);
            psn_cubic_coeffs(rx - (float)ix, wx);
            v = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int yy = min(max(iy - 1 + a, 0), n - 1);
                float row = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int xx = min(max(ix - 1 + c, 0), n - 1);
                    row += wx[c] * (float)p[yy * n + xx];
                }
                v += wy[a] * row;
            }
        }
        out[idx] = v;
    }
}

// ----------------------------------------------------------------------------- library context
namespace {

struct PlanEntry {
    PsnPlan plan;
    double* d_inv;
};

#define PSN_NSLOTS 6   // staging slots that exist; Context::host_slots of them are used

struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
};

struct Slot {
    cudaStream_t stream = nullptr;
    Arena arena;      // staging buffers (per chunk)
    Arena ws_arena;   // streamed-path workspace
    int* h_counters = nullptr;  // pinned, mapped into the device address space
    int* d_counters_map = nullptr;   // device alias of h_counters
    int h_counters_cap = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t staging_free = nullptr;   // recorded after the last copy of the chunk that used the staging arena
    cudaEvent_t h2d_done = nullptr;       // recorded on the upload stream after the chunk's last H2D copy
    cudaEvent_t solve_done = nullptr;     // recorded on the solve stream after the chunk's last kernel (copy_streams)
};

// few streams on purpose: a process has CUDA_DEVICE_MAX_CONNECTIONS (default 8) hardware queues, streams beyond that
// share a queue and serialise falsely (seen as solves waiting for another chunk's H2D copy)
#define PSN_NCHUNK_STREAMS 4

struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
};

struct Context {
    // Host-buffer pipeline: chunk c runs on chunk_stream[c % 4], created with DEcreasing priority, so that of two
    // overlapping chunks the earlier one wins the SMs (it finishes, and its result copy starts, as early as it
    // would alone) while the later one fills the gaps.
    cudaStream_t chunk_stream[PSN_NCHUNK_STREAMS] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t h2d_stream = nullptr;   // uploads of all chunks, in chunk order (host_h2d_stream)
    cudaStream_t d2h_stream = nullptr;   // result copies of all chunks, in chunk order (copy_streams)
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    size_t smem_optin = 0;
    Slot slots[PSN_NSLOTS];
    cudaEvent_t ws_free[2] = {nullptr, nullptr};   // recorded after the last kernel that touches streamed workspace k
    char* h_info = nullptr;          // pinned bounce buffer for iters / relres / status of host callers
    size_t h_info_cap = 0;   // recorded after the last kernel that touches the shared streamed workspace
    std::map<long long, PlanEntry> plans;  // key = (n * 16 + nu) * 64 + coarsest_max
    int nu = 2;
    int nu_coarse = 3;      // RB-GS sweeps per leg on levels >= 1 (0: as on level 0)
    int coarsest_max = 0;   // largest level solved with the dense inverse (option "coarsest_max"; 0: psn_default_coarsest(n))
    int tile = 64;
    long long ws_bytes = 24ll << 30;
    int smem_max_n = 0;
    int coarse_max_n = 40; // largest level handed to the shared-memory coarse V-cycle kernel (0: as large as fits)
    int coarse_block = 1;  // hierarchies with an 80 x 80 level: that level and below go to the register-blocked V-cycle kernel
    int serialize_solves = 0;
    int res_block = 1;     // use the register-blocked resident kernel where one is compiled (n = 80)
    int res_threads = 0;   // 0: smallest variant that fits; else force this CTA size for the resident kernel
    int res_persist = 1;   // register-blocked resident kernel: one persistent CTA per SM loops over the samples
    int res_waves = 1;     // persistent CTAs per SM (the kernel holds 1 CTA per SM: more only queue up)
    int inv_smem = 1;      // persistent kernel: keep the coarsest-level inverse in shared memory when it fits
    int fuse_upd = 2;      // with x_defer: 1 = CG update fused with the level-0 down-leg of the next V-cycle (PsnStreamCfg::fuse),
                           // 2 = also the prologue with the first down-leg (sine-factor forcing + theta on the solve grid)
    int x_defer = 1;       // with p_f32: x is accumulated from the ring of stored directions when a sample is checked
    int p_f32 = 1;         // with vcycle_f32 and upd_strip: the search direction is stored as (scaled) float too
    int upd_strip = 1;     // mode 2: strip update kernel that recomputes q = At p (q is never written to HBM)
    int upd_band = 64;     // rows per warp task of the strip update kernel (0: twice the direction kernel's)
    int pdl = 1;           // programmatic dependent launch between the kernels of the streamed solve
    int band0 = 0;         // rows per warp task on level 0 (0 = automatic)
    int fuse_band = 0;     // rows per warp task of the fused update + down-leg kernel (0 = automatic)
    int band_target = 8;   // largest number of bands per column the band split considers (choose_bands)
    int dir_band = 64;     // rows per warp task of the direction kernel (1 halo row per band: short bands balance better); 0: as level 0
    int host_edge_div = 2; // first / last chunk of the host pipeline = regular chunk / host_edge_div
    int vcycle_f32 = 1;    // streamed path: store V-cycle intermediates as float where the grid sizes allow (PsnStreamCfg::vf32)
    int host_chunks = 0;   // target number of pipeline chunks when staging host buffers (0: 6 streamed / 8 resident)
    int host_slots = 5;    // staging slots in use (3..PSN_NSLOTS): chunk c's upload waits for chunk c - host_slots to leave its slot
    int host_prefetch = 3; // uploads enqueued this many chunks ahead of the solve being enqueued (<= host_slots - 2, <= 3)
    int host_h2d_stream = 1;  // 1: uploads go through their own stream instead of the chunk's solve stream
    int copy_streams = 1;     // streamed grids: uploads and result copies of the host pipeline each go through ONE dedicated stream
                              // (prefetch depth unchanged).  With the copies on the four chunk streams, an upload and a result copy
                              // of different chunks were at times serialised on one copy engine (5 ms bubbles in the 640^2 timeline).
    int vcycle_mode = 2;   // 0: shared-memory tile kernels, 1: warp strips + register prefetch, 2: warp strips + TMA ring
    // Every call shares the staging arenas, the streamed workspaces and the plan cache.  Host threads are serialised by
    // g_mu; on the device, a call on a different stream than the previous one first waits for that call's last
    // operation (device-pointer calls return without synchronising, so their kernels may still be in flight).
    Arena mt_arena;                        // psn_mt19937_uniform_blocks: jumped states, polynomials, staging
    Arena master[2];                       // psn_multires: device copies of host master fields (theta, sine factors)
    cudaEvent_t master_done[2] = {nullptr, nullptr};
    Arena multi_out;                       // psn_multires with host outputs: device results of every resolution of a chunk
    cudaEvent_t multi_solved = nullptr;    // recorded behind each resolution's solve; the result-copy stream waits for it
    int multires_pipeline = 1;             // option: 0 = one blocking host-output solve per resolution (the round-1 behaviour)
    cudaEvent_t last_call = nullptr;
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    bool attrs_set = false;
    bool profiling = false;
    std::vector<ProfRec> prof;
};

Context g;
std::recursive_mutex g_mu;   // recursive: entry points call each other (ensure_init -> psn_init)

// order a call on stream `s` after the previous call's device work
int call_begin(cudaStream_t s) {
    if (g.have_last && g.last_stream != s) CU_TRY(cudaStreamWaitEvent(s, g.last_call, 0));
    return PSN_OK;
}
int call_end(cudaStream_t s) {
    CU_TRY(cudaEventRecord(g.last_call, s));
    g.last_stream = s;
    g.have_last = true;
    return PSN_OK;
}

int arena_reserve(Arena& a, size_t bytes) {
    if (bytes > a.cap) {
        if (a.base) CU_TRY(cudaFree(a.base));
        a.base = nullptr;
        a.cap = 0;
        CU_TRY(cudaMalloc(&a.base, bytes));
        a.cap = bytes;
    }
    a.off = 0;
    return PSN_OK;
}
template <class T>
T* arena_take(Arena& a, size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    if (a.off + bytes > a.cap) return nullptr;
    T* p = reinterpret_cast<T*>(a.base + a.off);
    a.off += bytes;
    return p;
}
inline size_t pad256(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

enum MemKind { MEM_NULL = 0, MEM_HOST = 1, MEM_DEVICE = 2 };
MemKind classify(const void* p) {
    if (!p) return MEM_NULL;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return MEM_HOST;
    }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? MEM_DEVICE : MEM_HOST;
}

// pageable host memory: cudaMemcpyAsync to / from it holds the calling thread until the copy has run
bool is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

int ensure_init() {
    if (g.ready) return PSN_OK;
    return psn_init(-1);
}

int get_plan(int n, PlanEntry** out) {
    const long long key = (((long long)n * 16 + g.nu) * 64 + g.coarsest_max) * 8 + g.nu_coarse;
    auto it = g.plans.find(key);
    if (it == g.plans.end()) {
        PlanEntry e;
        psn_plan_levels(&e.plan, n, g.nu, g.coarsest_max, g.nu_coarse);
        std::vector<double> inv = psn_coarse_inverse(e.plan);
        e.d_inv = nullptr;
        CU_TRY(cudaMalloc(&e.d_inv, inv.size() * sizeof(double)));
        CU_TRY(cudaMemcpy(e.d_inv, inv.data(), inv.size() * sizeof(double), cudaMemcpyHostToDevice));
        e.plan.coarse_inv = e.d_inv;
        it = g.plans.emplace(key, e).first;
    }
    *out = &it->second;
    return PSN_OK;
}

// resident-kernel configuration for a plan; T == 0 -> does not fit
struct ResidentCfg {
    int T, XPT;
    size_t smem;
};
ResidentCfg resident_cfg(const PsnPlan& plan) {
    ResidentCfg c{0, 0, 0};
    const int nn = plan.n * plan.n;
    if (g.smem_max_n > 0 && plan.n > g.smem_max_n) return c;
    if (g.smem_max_n < 0) return c;
    c.smem = (size_t)psn_smem_solver_doubles(plan) * sizeof(double);
    if (c.smem + 1024 > g.smem_optin) return c;
    // compiled (threads, values-per-thread) variants; "res_threads" forces a thread count for experiments
    // order = preference (measured on B200: 20^2 best with 128 threads, 40^2 with 512x4, 80^2 with 512x16)
    static const int variants[][2] = {{128, 4}, {256, 4}, {512, 4}, {256, 8}, {512, 8}, {1024, 4}, {512, 16}, {1024, 8}};
    for (auto& v : variants) {
        if (g.res_threads > 0 && v[0] != g.res_threads) continue;
        if (nn <= v[0] * v[1]) { c.T = v[0]; c.XPT = v[1]; break; }
    }
    if (c.T == 0 && g.res_threads > 0)
        for (auto& v : variants)
            if (nn <= v[0] * v[1]) { c.T = v[0]; c.XPT = v[1]; break; }
    return c;
}

// first level that the shared-memory coarse kernel can take
int stream_ls(const PsnPlan& plan) {
    if (g.coarse_block)   // an 80 x 80 level: the register-blocked V-cycle kernel takes it and everything below
        for (int l = 1; l < plan.nlev; ++l)
            if (psn_block_ok_at<20>(plan, l)) return l;
    for (int l = 1; l < plan.nlev; ++l) {
        size_t bytes = ((size_t)2 * plan.lv[l].size + psn_pool_doubles(plan, l)) * sizeof(double);
        if (g.coarse_max_n > 0 && plan.lv[l].n > g.coarse_max_n && l + 1 < plan.nlev) continue;
        if (bytes + 1024 <= g.smem_optin) return l;
    }
    return -1;
}

int set_kernel_attrs() {
    if (g.attrs_set) return PSN_OK;
    const int mx = (int)g.smem_optin - 1024;  // leave room for the static reduction scratch
#define PSN_RES_ATTR(T_, X_) CU_TRY(cudaFuncSetAttribute(psn_smem_solve_kernel<T_, X_>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))
    CU_TRY(cudaFuncSetAttribute(psn_block_solve_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CU_TRY(cudaFuncSetAttribute(psn_block_solve_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CU_TRY(cudaFuncSetAttribute(psn_block_solve_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    PSN_RES_ATTR(128, 4); PSN_RES_ATTR(256, 4); PSN_RES_ATTR(256, 8); PSN_RES_ATTR(512, 4);
    PSN_RES_ATTR(512, 8); PSN_RES_ATTR(1024, 4); PSN_RES_ATTR(512, 16); PSN_RES_ATTR(1024, 8);
#undef PSN_RES_ATTR
    CU_TRY(cudaFuncSetAttribute(psn_tile_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CU_TRY(cudaFuncSetAttribute(psn_tile_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CU_TRY(cudaFuncSetAttribute(psn_coarse_vcycle_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CU_TRY(cudaFuncSetAttribute(psn_coarse_vcycle_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CU_TRY(cudaFuncSetAttribute(psn_block_vcycle_kernel<20, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CU_TRY(cudaFuncSetAttribute(psn_block_vcycle_kernel<20, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    g.attrs_set = true;
    return PSN_OK;
}

// ---- profiling helpers
struct ProfScope {
    cudaStream_t s;
    ProfRec rec;
    bool on;
    ProfScope(const char* name, cudaStream_t stream) : s(stream), on(g.profiling) {
        if (on) {
            rec.name = name;
            cudaEventCreate(&rec.e0);
            cudaEventCreate(&rec.e1);
            cudaEventRecord(rec.e0, s);
        }
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(rec.e1, s);
            g.prof.push_back(rec);
        }
    }
};

// ---- chunk timeline of the host-buffer pipeline (PSN_TRACE=1: one line per chunk on stderr, ms since call start)
struct TraceMark {
    const char* what;
    long long chunk;
    cudaEvent_t ev;
    double host_ms;
};
struct Trace {
    bool on = false;
    std::vector<TraceMark> marks;
    std::chrono::steady_clock::time_point t0;
    void begin(cudaStream_t s) {
        const char* e = getenv("PSN_TRACE");
        on = e && atoi(e) > 0;
        if (!on) return;
        marks.clear();
        t0 = std::chrono::steady_clock::now();
        mark("call", -1, s);
    }
    void mark(const char* what, long long chunk, cudaStream_t s) {
        if (!on) return;
        TraceMark m{what, chunk, nullptr, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()};
        cudaEventCreate(&m.ev);
        cudaEventRecord(m.ev, s);
        marks.push_back(m);
    }
    void dump() {
        if (!on) return;
        cudaDeviceSynchronize();
        const double host_end = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[0].ev, marks[i].ev);
            fprintf(stderr, "[psn trace] chunk %2lld %-10s device %8.3f ms   enqueued at host %8.3f ms\n", marks[i].chunk, marks[i].what, ms,
                    marks[i].host_ms);
        }
        fprintf(stderr, "[psn trace] call returns at host %8.3f ms\n", host_end);
        for (auto& m : marks) cudaEventDestroy(m.ev);
        marks.clear();
    }
};

// ---- CUDA backend of the streamed driver
struct CudaBackend {
    cudaStream_t s;
    Slot* slot;
    cudaError_t err = cudaSuccess;
    void chk() {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && err == cudaSuccess) err = e;
    }
    dim3 cg_grid(const PsnCgArgs& a, long long B) { return dim3(a.ntx, a.nty, (unsigned)B); }
    // kernel launch with the programmatic-stream-serialization attribute (see psn_pdl_enter)
    template <class... KA, class... A>
    void launch(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, A&&... args) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = (g.pdl && !g.profiling) ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
    }
    void zero_counters(int* c, int cnt) {
        launch(psn_zero_ints_kernel, dim3(1), dim3(128), 0, c, cnt);
        chk();
    }
    void prologue(const PsnCgArgs& a, long long B) {
        ProfScope p("prologue", s);
        launch(psn_prologue_kernel, dim3(cg_grid(a, B)), dim3(PSN_CG_THREADS), 0, a);
        chk();
    }
    void state_init(const PsnCgArgs& a, long long B) {
        ProfScope p("state", s);
        launch(psn_state_init_kernel, dim3((unsigned)((B + 127) / 128)), dim3(128), 0, a, B);
        chk();
    }
    template <int NU, bool POST, bool TMA, class TB, class TV>
    void strip_launch_t(const PsnPlan& plan, const PsnStripArgs& a, long long B) {
        const long long ntasks = B * a.nstrips * a.nbands;
        const unsigned grid = (unsigned)((ntasks + PSN_STRIP_WARPS - 1) / PSN_STRIP_WARPS);
        size_t smem = 0;
        if constexpr (TMA) {
            smem = (size_t)PSN_STRIP_WARPS * PsnStripFeederSel<NU, POST, TB, TV>::type::smem_bytes();
            static bool attr_done = false;   // one flag per template instantiation
            if (!attr_done) {
                cudaFuncSetAttribute(psn_strip_kernel<NU, POST, TMA, TB, TV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                attr_done = true;
            }
        }
        launch(psn_strip_kernel<NU, POST, TMA, TB, TV>, dim3(grid), dim3(PSN_STRIP_WARPS * 32), smem, plan, a, ntasks);
    }
    template <int NU, bool POST>
    void strip_launch_nu(const PsnPlan& plan, const PsnStripArgs& a, long long B, bool tma, bool f32) {
        if (tma && f32) {
            if (a.level == 0) strip_launch_t<NU, POST, true, double, float>(plan, a, B);
            else strip_launch_t<NU, POST, true, float, float>(plan, a, B);
        } else if (tma) {
            strip_launch_t<NU, POST, true, double, double>(plan, a, B);
        } else {
            strip_launch_t<NU, POST, false, double, double>(plan, a, B);
        }
    }
    template <bool POST>
    void strip_launch(const PsnPlan& plan, const PsnStripArgs& a, long long B, bool tma, bool f32) {
        const int nu_l = plan.nul[a.level];
        if (nu_l == 1) strip_launch_nu<1, POST>(plan, a, B, tma, f32);
        else if (nu_l == 2) strip_launch_nu<2, POST>(plan, a, B, tma, f32);
        else strip_launch_nu<3, POST>(plan, a, B, tma, f32);
    }
    void leg_pre(const PsnPlan& plan, const PsnStreamCfg& cfg, const PsnLegArgs& g, long long B) {
        if (cfg.mode >= 1) {
            ProfScope p(g.level == 0 ? "strip_pre_l0" : "strip_pre_coarse", s);
            strip_launch<false>(plan, psn_leg_to_strip(plan, cfg, g, false), B, cfg.mode == 2, cfg.mode == 2 && cfg.vf32);
        } else {
            const PsnTileArgs t = psn_leg_to_tile(plan, cfg, g);
            ProfScope p(t.level == 0 ? "tile_pre_l0" : "tile_pre_coarse", s);
            const size_t smem = (size_t)psn_tile_pre_doubles(t.TH, t.TW, plan.nul[t.level]) * sizeof(double);
            launch(psn_tile_pre_kernel, dim3(dim3(t.ntx, t.nty, (unsigned)B)), dim3(PSN_TILE_THREADS), smem, plan, t);
        }
        chk();
    }
    void leg_post(const PsnPlan& plan, const PsnStreamCfg& cfg, const PsnLegArgs& g, long long B) {
        if (cfg.mode >= 1) {
            ProfScope p(g.level == 0 ? "strip_post_l0" : "strip_post_coarse", s);
            strip_launch<true>(plan, psn_leg_to_strip(plan, cfg, g, true), B, cfg.mode == 2, cfg.mode == 2 && cfg.vf32);
        } else {
            const PsnTileArgs t = psn_leg_to_tile(plan, cfg, g);
            ProfScope p(t.level == 0 ? "tile_post_l0" : "tile_post_coarse", s);
            const size_t smem = (size_t)psn_tile_post_doubles(t.TH, t.TW, plan.nul[t.level]) * sizeof(double);
            launch(psn_tile_post_kernel, dim3(dim3(t.ntx, t.nty, (unsigned)B)), dim3(PSN_TILE_THREADS), smem, plan, t);
        }
        chk();
    }
    void coarse_vcycle(const PsnPlan& plan, const PsnStreamCfg& cfg, const double* bg, double* ug, const PsnSampleState* st, long long B) {
        ProfScope p("coarse_vcycle", s);
        const int l0 = cfg.ls;
        if (cfg.coarse_block && psn_block_ok_at<20>(plan, l0)) {
            size_t bsm = (size_t)psn_block_vcycle_smem_doubles<20>(plan, l0) * sizeof(double);
            int invd = psn_coarse_inv_doubles(plan);
            if (!g.inv_smem || bsm + (size_t)invd * sizeof(double) + 1024 > g.smem_optin) invd = 0;
            bsm += (size_t)invd * sizeof(double);
            const unsigned grid = (unsigned)std::min<long long>(B, (long long)g.sm_count);
            if (cfg.mode == 2 && cfg.vf32)
                launch(psn_block_vcycle_kernel<20, float>, dim3(grid), dim3(PsnBlockGeom<20>::NTPAD), bsm, plan, l0,
                       reinterpret_cast<const float*>(bg), reinterpret_cast<float*>(ug), st, B, invd);
            else
                launch(psn_block_vcycle_kernel<20, double>, dim3(grid), dim3(PsnBlockGeom<20>::NTPAD), bsm, plan, l0, bg, ug, st, B, invd);
            chk();
            return;
        }
        const size_t smem = ((size_t)2 * plan.lv[l0].size + psn_pool_doubles(plan, l0)) * sizeof(double);
        if (cfg.mode == 2 && cfg.vf32)
            launch(psn_coarse_vcycle_kernel<float>, dim3((unsigned)B), dim3(PSN_COARSE_THREADS), smem, plan, l0, reinterpret_cast<const float*>(bg),
                                                                                         reinterpret_cast<float*>(ug), st);
        else
            launch(psn_coarse_vcycle_kernel<double>, dim3((unsigned)B), dim3(PSN_COARSE_THREADS), smem, plan, l0, bg, ug, st);
        chk();
    }
    template <bool TMA, class TZ, class TP>
    void dir_launch(const PsnCgArgs& a, unsigned grid, long long ntasks) {
        const size_t smem = (size_t)PSN_STRIP_WARPS * PsnTmaFeeder<2, TZ, TP>::smem_bytes();
        static bool attr_done = false;   // one flag per template instantiation
        if (!attr_done) {
            cudaFuncSetAttribute(psn_dir_strip_kernel<TMA, TZ, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_done = true;
        }
        launch(psn_dir_strip_kernel<TMA, TZ, TP>, dim3(grid), dim3(PSN_STRIP_WARPS * 32), smem, a, ntasks);
    }
    template <class TP, bool XDEFER>
    void update_launch(const PsnCgArgs& a, unsigned grid, long long ntasks) {
        typedef typename PsnUpdateFeederSel<TP, XDEFER>::type Feeder;
        const size_t smem = (size_t)PSN_STRIP_WARPS * Feeder::smem_bytes();
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(psn_update_strip_kernel<TP, XDEFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_done = true;
        }
        launch(psn_update_strip_kernel<TP, XDEFER>, dim3(grid), dim3(PSN_STRIP_WARPS * 32), smem, a, ntasks);
    }
    template <int NU, int MODE>
    void fused_launch(const PsnPlan& plan, const PsnStripArgs& a, const PsnCgArgs& cg, long long B) {
        typedef typename PsnUpreFeederSel<NU, MODE>::type Feeder;
        const long long ntasks = B * a.nstrips * a.nbands;
        const unsigned grid = (unsigned)((ntasks + PSN_STRIP_WARPS - 1) / PSN_STRIP_WARPS);
        const size_t smem = (size_t)PSN_STRIP_WARPS * Feeder::smem_bytes();
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(psn_upre_strip_kernel<NU, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_done = true;
        }
        launch(psn_upre_strip_kernel<NU, MODE>, dim3(grid), dim3(PSN_STRIP_WARPS * 32), smem, plan, a, cg, ntasks);
    }
    void fused_update(const PsnPlan& plan, const PsnStripArgs& a, const PsnCgArgs& cg, long long B) {
        ProfScope p("update_pre_l0", s);
        if (plan.nu == 1) fused_launch<1, 0>(plan, a, cg, B);
        else if (plan.nu == 2) fused_launch<2, 0>(plan, a, cg, B);
        else fused_launch<3, 0>(plan, a, cg, B);
        chk();
    }
    void fused_prologue(const PsnPlan& plan, const PsnStripArgs& a, const PsnCgArgs& cg, long long B) {
        ProfScope p("prologue_pre_l0", s);
        if (plan.nu == 1) fused_launch<1, 1>(plan, a, cg, B);
        else if (plan.nu == 2) fused_launch<2, 1>(plan, a, cg, B);
        else fused_launch<3, 1>(plan, a, cg, B);
        chk();
    }
    void flush_x(const PsnCgArgs& a, long long B) {
        ProfScope p("flush_x", s);
        launch(psn_flush_x_kernel, dim3(PSN_FLUSH_BLOCKS, (unsigned)B), dim3(256), 0, a);
        chk();
    }
    void dir_apply(const PsnCgArgs& a, long long B) {
        ProfScope p("dir_apply_dot", s);
        if (a.dir_mode >= 1) {
            const long long ntasks = B * a.dir_nstrips * a.dir_nbands;
            const unsigned grid = (unsigned)((ntasks + PSN_STRIP_WARPS - 1) / PSN_STRIP_WARPS);
            if (a.dir_mode == 2 && a.vf32 && a.pf32) {
                dir_launch<true, float, float>(a, grid, ntasks);
            } else if (a.dir_mode == 2 && a.vf32) {
                dir_launch<true, float, double>(a, grid, ntasks);
            } else if (a.dir_mode == 2) {
                dir_launch<true, double, double>(a, grid, ntasks);
            } else {
                launch(psn_dir_strip_kernel<false, double, double>, dim3(grid), dim3(PSN_STRIP_WARPS * 32), 0, a, ntasks);
            }
        } else {
            launch(psn_dir_apply_kernel, dim3(cg_grid(a, B)), dim3(PSN_CG_THREADS), 0, a);
        }
        chk();
    }
    void update(const PsnCgArgs& a, long long B) {
        ProfScope p("update_norm", s);
        if (a.upd_mode == 2) {
            const long long ntasks = B * a.upd_nstrips * a.upd_nbands;
            const unsigned grid = (unsigned)((ntasks + PSN_STRIP_WARPS - 1) / PSN_STRIP_WARPS);
            if (a.pf32 && a.xdefer) update_launch<float, true>(a, grid, ntasks);
            else if (a.pf32) update_launch<float, false>(a, grid, ntasks);
            else update_launch<double, false>(a, grid, ntasks);
            chk();
            return;
        }
        launch(psn_update_kernel, dim3(cg_grid(a, B)), dim3(PSN_CG_THREADS), 0, a);
        chk();
    }
    void true_residual(const PsnCgArgs& a, long long B) {
        ProfScope p("true_residual", s);
        if (a.chk_mode == 2) {
            typedef PsnTmaFeeder<2, double, double> Feeder;
            const long long ntasks = B * a.upd_nstrips * a.upd_nbands;
            const unsigned grid = (unsigned)((ntasks + PSN_STRIP_WARPS - 1) / PSN_STRIP_WARPS);
            const size_t smem = (size_t)PSN_STRIP_WARPS * Feeder::smem_bytes();
            static bool attr_done = false;
            if (!attr_done) {
                cudaFuncSetAttribute(psn_true_residual_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                attr_done = true;
            }
            launch(psn_true_residual_strip_kernel, dim3(grid), dim3(PSN_STRIP_WARPS * 32), smem, a, ntasks);
            chk();
            return;
        }
        launch(psn_true_residual_kernel, dim3(cg_grid(a, B)), dim3(PSN_CG_THREADS), 0, a);
        chk();
    }
    void state_after_update(const PsnCgArgs& a, long long B) {
        ProfScope p("state", s);
        launch(psn_state_after_update_kernel, dim3((unsigned)((B + 127) / 128)), dim3(128), 0, a, B);
        chk();
    }
    void state_after_check(const PsnCgArgs& a, long long B, int* counter) {
        ProfScope p("state", s);
        launch(psn_state_after_check_kernel, dim3((unsigned)((B + 127) / 128)), dim3(128), 0, a, B, counter);
        chk();
    }
    void fetch_counter_async(const int* c, int it) {
        launch(psn_publish_counter_kernel, dim3(1), dim3(1), 0, c, slot->d_counters_map + it);
        chk();
        cudaError_t e = cudaEventRecord(slot->ev[it & 3], s);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
    }
    int wait_counter(int it) {
        cudaError_t e = cudaEventSynchronize(slot->ev[it & 3]);
        if (e != cudaSuccess) {
            if (err == cudaSuccess) err = e;
            return 0;  // stop iterating on error
        }
        return slot->h_counters[it];
    }
};

// ---- one batched problem as seen by the ABI
struct Problem {
    int n;
    long long B;
    const double *theta, *f, *sx, *sy, *k1, *k2;
    int mn, step;
    double rtol;
    int max_iter;
    double* u;
    int32_t* iters;
    double* relres;
    int32_t* status;
    int sub;
    const int32_t *start_x, *start_y;
    double *u_fine, *f_fine, *th_fine, *u_coarse, *f_coarse, *th_coarse;
    cudaStream_t stream;
};

// stage one input slice: returns device pointer valid on `s`
template <class T>
int stage_in(Arena& ar, cudaStream_t s, const T* src, MemKind kind, long long offset, size_t count, const T** out) {
    if (kind == MEM_NULL) { *out = nullptr; return PSN_OK; }
    if (kind == MEM_DEVICE) { *out = src + offset; return PSN_OK; }
    T* d = arena_take<T>(ar, count);
    if (!d) return psn_fail(PSN_ERR_NOMEM, "staging arena exhausted");
    CU_TRY(cudaMemcpyAsync(d, src + offset, count * sizeof(T), cudaMemcpyHostToDevice, s));
    *out = d;
    return PSN_OK;
}
template <class T>
int stage_out(Arena& ar, T* dst, MemKind kind, long long offset, size_t count, T** out, bool always) {
    if (kind == MEM_DEVICE) { *out = dst + offset; return PSN_OK; }
    if (kind == MEM_NULL && !always) { *out = nullptr; return PSN_OK; }
    T* d = arena_take<T>(ar, count);
    if (!d) return psn_fail(PSN_ERR_NOMEM, "staging arena exhausted");
    *out = d;
    return PSN_OK;
}
template <class T>
int copy_back(cudaStream_t s, T* host_dst, MemKind kind, long long offset, size_t count, const T* dev) {
    if (kind != MEM_HOST) return PSN_OK;
    CU_TRY(cudaMemcpyAsync(host_dst + offset, dev, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    return PSN_OK;
}

// upper bound of the <r,z> partial count over every band choice (smallest band = 16 rows)
int psn_rz_max(const PsnPlan& plan, const PsnStreamCfg& cfg) {
    if (cfg.mode < 1) return psn_rz_count(plan, cfg);
    return psn_strip_count(plan.n, psn_strip_use_post(plan.nu, cfg.vf32)) * psn_band_count(plan.n, 16);
}

// Rows per warp task ("band") on every streamed level.  A strip kernel runs ceil(CTAs / resident CTAs) waves and a
// task costs (band + halo rows) row steps, so the split minimises waves x (band + 2*(2 nu)): long bands when the batch
// alone fills the GPU, shorter ones (a little halo re-computation) when that evens out the last wave or fills the SMs.
// Measured on 512 / 256 / 103 x 640^2: the model picks 214-row bands, 6 % / 2 % / 5 % faster than whole columns.
void choose_bands(const PsnPlan& plan, PsnStreamCfg* cfg, long long Bc) {
    for (int l = 0; l < PSN_MAX_LEVELS; ++l) cfg->band[l] = 16;
    const long long slots = (long long)g.sm_count * (cfg->vf32 ? PSN_STRIP_MINB_F32 : PSN_STRIP_MINB_TMA);
    for (int l = 0; l < cfg->ls; ++l) {
        const int n = plan.lv[l].n;
        const int halo = 4 * plan.nul[l];
        const long long strips = psn_strip_count(n, psn_strip_use_pre(plan.nul[l], cfg->vf32));
        int best_band = (n + 1) & ~1;
        long long best_cost = -1;
        const int nb_max = g.band_target > 0 ? g.band_target : 1;
        for (int nb = 1; nb <= nb_max; ++nb) {
            int band = (n + nb - 1) / nb;
            band = (band + 1) & ~1;
            if (band < 16) break;
            const long long ctas = (Bc * strips * psn_band_count(n, band) + PSN_STRIP_WARPS - 1) / PSN_STRIP_WARPS;
            const long long waves = (ctas + slots - 1) / slots;
            const long long cost = waves * (band + halo);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best_band = band;
            }
        }
        cfg->band[l] = best_band;
    }
    // The fused update + down-leg kernel follows the level-0 split unless "fuse_band" says otherwise: a split from its own wave
    // model (3 CTAs per SM) measured the same at 640^2 (797-819 us for bands of 128-320 rows) and 28 % slower at 160^2 x 8192.
    cfg->fuse_band = g.fuse_band >= 16 ? ((g.fuse_band + 1) & ~1) : 0;
    if (g.band0 >= 16 && cfg->ls >= 1) cfg->band[0] = (g.band0 + 1) & ~1;   // experiment knob
    // direction kernel: bands of about "dir_band" rows, equalised over the column (160 rows as 64 + 64 + 32 left the last band of
    // every strip half empty: 499 us per launch at 160^2 x 8192 against 444 us for 80 + 80; 320 and 640 divide evenly anyway)
    if (g.dir_band >= 16) {
        const int nb = std::max(1, plan.n / g.dir_band);
        cfg->dir_band = ((plan.n + nb - 1) / nb + 1) & ~1;
    } else {
        cfg->dir_band = cfg->band[0];
    }
    cfg->upd_band = g.upd_band >= 16 ? ((g.upd_band + 1) & ~1) : 0;
    cfg->upd_strip = g.upd_strip;
    cfg->pf32 = (cfg->vf32 && g.upd_strip && g.p_f32) ? 1 : 0;
    cfg->xdefer = (cfg->pf32 && g.x_defer) ? 1 : 0;
    cfg->fuse = cfg->xdefer ? g.fuse_upd : 0;
}

size_t stream_ws_bytes_per_sample(const PsnPlan& plan, const PsnStreamCfg& cfg) {
    const size_t nn = (size_t)plan.n * plan.n;
    size_t b = 5 * nn * sizeof(double);
    if (cfg.xdefer) b += (size_t)PSN_PEND_K * nn * sizeof(float) - 2 * nn * sizeof(double) + sizeof(PsnPend);
    if (cfg.fuse) b += nn * sizeof(double) + sizeof(double);   // second residual buffer, scale of the fused prologue
    for (int l = 1; l <= cfg.ls; ++l) {
        const size_t m = (size_t)plan.lv[l].n * plan.lv[l].n * sizeof(double);
        b += (l < cfg.ls ? 3 : 2) * m;
    }
    const int gx = (plan.n + PSN_GT_COLS - 1) / PSN_GT_COLS, gy = (plan.n + PSN_GT_ROWS - 1) / PSN_GT_ROWS;
    b += ((size_t)psn_rz_max(plan, cfg) + (size_t)psn_pq_max(plan.n) + (size_t)psn_pu_max(plan.n, plan.nu) + 2 * (size_t)gx * gy) * sizeof(double) +
         sizeof(PsnSampleState);
    return b + 4096 / 16;
}

int carve_stream_ws(Arena& ar, const PsnPlan& plan, const PsnStreamCfg& cfg, long long Bc, int max_iter, PsnStreamWs* ws) {
    const size_t nn = (size_t)plan.n * plan.n;
    memset(ws, 0, sizeof(*ws));
    ws->cap = Bc;
    ws->r = arena_take<double>(ar, Bc * nn);
    ws->q = arena_take<double>(ar, Bc * nn);
    ws->z = arena_take<double>(ar, Bc * nn);
    bool ok = true;
    if (cfg.xdefer) {   // the ring of float directions replaces the fp64 ping-pong pair
        ws->pring = arena_take<float>(ar, (size_t)PSN_PEND_K * Bc * nn);
        ws->pend = arena_take<PsnPend>(ar, Bc);
        ok = ws->pring && ws->pend;
    } else {
        ws->p[0] = arena_take<double>(ar, Bc * nn);
        ws->p[1] = arena_take<double>(ar, Bc * nn);
        ok = ws->p[0] && ws->p[1];
    }
    ok = ok && ws->r && ws->q && ws->z;
    if (cfg.fuse) {
        ws->r2 = arena_take<double>(ar, Bc * nn);
        ws->vs0 = arena_take<double>(ar, Bc);
        ok = ok && ws->r2 && ws->vs0;
    }
    for (int l = 1; l <= cfg.ls && ok; ++l) {
        const size_t m = (size_t)plan.lv[l].n * plan.lv[l].n;
        ws->bl[l] = arena_take<double>(ar, Bc * m);
        ws->ub[l] = arena_take<double>(ar, Bc * m);
        ok = ws->bl[l] && ws->ub[l];
        if (l < cfg.ls) {
            ws->ua[l] = arena_take<double>(ar, Bc * m);
            ok = ok && ws->ua[l];
        }
    }
    const int gx = (plan.n + PSN_GT_COLS - 1) / PSN_GT_COLS, gy = (plan.n + PSN_GT_ROWS - 1) / PSN_GT_ROWS;
    ws->part_rz = arena_take<double>(ar, (size_t)Bc * psn_rz_max(plan, cfg));
    ws->part_a = arena_take<double>(ar, (size_t)Bc * gx * gy);
    ws->part_pq = arena_take<double>(ar, (size_t)Bc * psn_pq_max(plan.n));
    ws->part_b = arena_take<double>(ar, (size_t)Bc * gx * gy);
    ws->part_u = arena_take<double>(ar, (size_t)Bc * psn_pu_max(plan.n, plan.nu));
    ws->state = arena_take<PsnSampleState>(ar, Bc);
    ws->counters = arena_take<int>(ar, max_iter + 8);
    ok = ok && ws->part_rz && ws->part_a && ws->part_pq && ws->part_b && ws->part_u && ws->state && ws->counters;
    return ok ? PSN_OK : psn_fail(PSN_ERR_NOMEM, "workspace arena exhausted");
}

int solve_impl(const Problem& P);
int solve_internal(const Problem& P) {
    int rc = ensure_init();
    if (rc) return rc;
    if (P.B <= 0 || P.n < 2) return solve_impl(P);   // argument errors / empty batch: no device work
    if ((rc = call_begin(P.stream))) return rc;
    if ((rc = solve_impl(P))) return rc;
    return call_end(P.stream);
}

int solve_impl(const Problem& P) {
    int rc = ensure_init();
    if (rc) return rc;
    if (P.n < 2) return psn_fail(PSN_ERR_INVALID, "n must be >= 2 (got %d)", P.n);
    if (P.B < 0) return psn_fail(PSN_ERR_INVALID, "batch must be >= 0");
    if (P.B == 0) return PSN_OK;
    if (P.step < 1 || P.mn < (P.n - 1) * P.step + 1)
        return psn_fail(PSN_ERR_INVALID, "master_n=%d too small for n=%d step=%d", P.mn, P.n, P.step);
    const int nforce = (P.f ? 1 : 0) + ((P.sx || P.sy) ? 1 : 0) + ((P.k1 || P.k2) ? 1 : 0);
    if (nforce != 1 || (!P.sx != !P.sy) || (!P.k1 != !P.k2))
        return psn_fail(PSN_ERR_INVALID, "give exactly one forcing: f, (sin_x and sin_y), or (k1 and k2)");
    if (!(P.rtol > 0.0) || P.max_iter < 1) return psn_fail(PSN_ERR_INVALID, "rtol must be > 0 and max_iter >= 1");
    if (!P.u && P.sub == 0) return psn_fail(PSN_ERR_INVALID, "output u is NULL");
    if (P.sub < 0 || P.sub > P.n || (P.sub > 0 && (!P.start_x || !P.start_y)))
        return psn_fail(PSN_ERR_INVALID, "bad sub-domain request");
    rc = set_kernel_attrs();
    if (rc) return rc;

    PlanEntry* pe = nullptr;
    rc = get_plan(P.n, &pe);
    if (rc) return rc;
    const PsnPlan& plan = pe->plan;
    const ResidentCfg rcfg = resident_cfg(plan);
    const bool resident = rcfg.T != 0;
    PsnStreamCfg scfg;
    memset(&scfg, 0, sizeof(scfg));
    scfg.TH = scfg.TW = g.tile;
    scfg.mode = g.vcycle_mode;
    if (!resident) {
        scfg.ls = stream_ls(plan);
        scfg.coarse_block = (g.coarse_block && scfg.ls >= 1 && psn_block_ok_at<20>(plan, scfg.ls)) ? 1 : 0;
        if (scfg.ls >= 1 && !psn_strip_ok(plan, scfg.ls)) scfg.mode = 0;   // odd level sizes: shared-memory tile kernels
        scfg.vf32 = (scfg.mode == 2 && g.vcycle_f32 && scfg.ls >= 1 && psn_vf32_ok(plan, scfg.ls)) ? 1 : 0;
        scfg.upd_strip = g.upd_strip;
        scfg.pf32 = (scfg.vf32 && g.upd_strip && g.p_f32) ? 1 : 0;
        scfg.xdefer = (scfg.pf32 && g.x_defer) ? 1 : 0;
        scfg.fuse = scfg.xdefer ? g.fuse_upd : 0;
        if (scfg.ls < 1) return psn_fail(PSN_ERR_UNSUPPORTED, "no shared-memory coarse level for n=%d", P.n);
    }

    const MemKind k_theta = classify(P.theta), k_f = classify(P.f), k_sx = classify(P.sx), k_sy = classify(P.sy),
                  k_k1 = classify(P.k1), k_k2 = classify(P.k2), k_u = classify(P.u), k_it = classify(P.iters),
                  k_rr = classify(P.relres), k_st = classify(P.status), k_x0 = classify(P.start_x), k_y0 = classify(P.start_y),
                  k_uf = classify(P.u_fine), k_ff = classify(P.f_fine), k_tf = classify(P.th_fine),
                  k_uc = classify(P.u_coarse), k_fc = classify(P.f_coarse), k_tc = classify(P.th_coarse);
    const MemKind kinds[] = {k_theta, k_f, k_sx, k_sy, k_k1, k_k2, k_u, k_it, k_rr, k_st, k_x0, k_y0, k_uf, k_ff, k_tf, k_uc, k_fc, k_tc};
    bool any_host = false;
    for (MemKind k : kinds) any_host = any_host || (k == MEM_HOST);
    {
        // Caller-owned DEVICE arrays are read and written in place with 128-bit accesses and bulk (TMA) copies:
        // they must be 16-byte aligned (host arrays are staged through the library's own aligned buffers).
        const struct { const void* p; MemKind k; const char* name; } dev_arrays[] = {
            {P.theta, k_theta, "theta"}, {P.f, k_f, "f"}, {P.sx, k_sx, "sin_x"}, {P.sy, k_sy, "sin_y"}, {P.u, k_u, "u"}};
        for (const auto& d : dev_arrays)
            if (d.k == MEM_DEVICE && (reinterpret_cast<uintptr_t>(d.p) & 15) != 0)
                return psn_fail(PSN_ERR_INVALID, "device array %s (%p) is not 16-byte aligned", d.name, d.p);
    }
    // The deep upload pipeline (host_slots / host_prefetch / host_h2d_stream) enqueues the copies of several chunks before
    // the solve of the current one; a copy from or to pageable memory would hold the host inside that enqueue until the
    // copies queued in front of it have run, so callers with pageable buffers keep the shallow 3-slot pipeline.
    // (iters / relres / status of host callers go through the pinned bounce buffer.)
    // Resident grids only: the streamed sizes are bound by the PCIe link at any depth on one GPU (640^2: 44.4 ms per 3.36 GB
    // either way), and with two ranks sharing the host fabric the uploads running ahead took bandwidth from the result
    // copies on the critical path (640^2 at 2 GPUs: 18.1k -> 14.3k solves/s end to end), while 80^2 gained at 1 and 2 GPUs.
    const void* bulk[] = {P.theta, P.f, P.sx, P.sy, P.k1, P.k2, P.u, P.start_x, P.start_y,
                          P.u_fine, P.f_fine, P.th_fine, P.u_coarse, P.f_coarse, P.th_coarse};
    bool deep = any_host && resident;
    for (const void* q : bulk) deep = deep && !is_pageable(q);

    const size_t nn = (size_t)P.n * P.n, mm = (size_t)P.mn * P.mn;
    const int sc = (P.sub + 1) / 2;
    const size_t ss = (size_t)P.sub * P.sub, scc = (size_t)sc * sc;

    // bytes of staging per sample
    size_t stage_ps = 64;
    if (k_theta == MEM_HOST) stage_ps += mm * 8;
    if (k_f == MEM_HOST) stage_ps += mm * 8;
    if (k_sx == MEM_HOST) stage_ps += (size_t)P.mn * 8;
    if (k_sy == MEM_HOST) stage_ps += (size_t)P.mn * 8;
    if (k_u == MEM_HOST || (!resident && k_u == MEM_NULL)) stage_ps += nn * 8;
    if (k_uf == MEM_HOST) stage_ps += ss * 8;
    if (k_ff == MEM_HOST) stage_ps += ss * 8;
    if (k_tf == MEM_HOST) stage_ps += ss * 8;
    if (k_uc == MEM_HOST) stage_ps += scc * 8;
    if (k_fc == MEM_HOST) stage_ps += scc * 8;
    if (k_tc == MEM_HOST) stage_ps += scc * 8;

    // chunk size
    long long Bc = P.B;
    const size_t ws_ps = resident ? 0 : stream_ws_bytes_per_sample(plan, scfg);
    {
        const long long budget = g.ws_bytes;
        const long long per = (long long)(ws_ps + stage_ps);
        long long cap = per > 0 ? budget / per : P.B;
        if (cap < 1) cap = 1;
        if (cap > 65535 && !resident) cap = 65535;   // gridDim.z
        if (any_host) {
            // at least ~8 chunks so that copies overlap the solve, but keep chunks large enough to fill the GPU
            const int hc = g.host_chunks > 0 ? g.host_chunks : (resident ? 8 : 6);
            long long want = (P.B + hc - 1) / hc;
            const long long floor_ = resident ? 4ll * g.sm_count : 64;
            if (want < floor_) want = floor_;
            if (cap > want) cap = want;
        }
        if (Bc > cap) Bc = cap;
    }
    const int nslots = any_host ? (deep ? g.host_slots : 3) : 1;
    const bool copy_dedicated = !resident && g.copy_streams && P.sub == 0;
    if (any_host) CU_TRY(cudaStreamSynchronize(P.stream));
    Trace trace;
    if (any_host) trace.begin(g.chunk_stream[0]);  // device inputs may still be in flight on the caller's stream

    // Software pipeline over chunks (host buffers): the H2D staging of chunk c+1 is enqueued on its own
    // slot stream BEFORE the (host-blocking) solve of chunk c, and the D2H of chunk c-1 drains on a third
    // stream, so copies in both directions overlap the solve.  The streamed-path workspace is shared by
    // all slots (solves are serialised by an event chain); staging buffers are per slot.
    struct ChunkCtx {
        long long b0, bc;
        Slot* slot;
        cudaStream_t s;
        PsnRhs rhs;
        double* d_u;
        int32_t* d_it;
        double* d_rr;
        int32_t* d_st;
        PsnSolveArgs a;
    };
    // chunk schedule: equal chunks of Bc for device callers; for host callers the first and last chunks
    // are shorter (1/4 of Bc) so that pipeline fill (first H2D) and drain (last D2H) are short
    std::vector<long long> bounds;
    bounds.push_back(0);
    if (any_host && P.B > 2 * Bc) {
        // short first / last chunk (pipeline fill = first H2D, drain = last D2H), equal chunks <= Bc in between
        const long long e = Bc / g.host_edge_div > 0 ? Bc / g.host_edge_div : 1;
        const long long mid = P.B - 2 * e;
        const long long parts = (mid + Bc - 1) / Bc;
        bounds.push_back(e);
        for (long long k = 1; k < parts; ++k) bounds.push_back(e + (mid * k) / parts);
        bounds.push_back(P.B - e);
        bounds.push_back(P.B);
    } else {
        for (long long pos = Bc; pos < P.B; pos += Bc) bounds.push_back(pos);
        bounds.push_back(P.B);
    }
    const long long nchunks = (long long)bounds.size() - 1;
    // Streamed path with host buffers: two workspaces (slots 0 / 1) alternate between chunks, so that a chunk can
    // be enqueued completely (expected iteration count, no polling) while the previous one is still running;
    // its convergence is verified one chunk later, before its workspace is reused.
    const int nws = (!resident && any_host) ? 2 : 1;
    if (!resident) {
        for (int k = 0; k < nws; ++k) {
            Slot& wsl = g.slots[k];
            rc = arena_reserve(wsl.ws_arena, ws_ps * Bc + 64 * 256 + pad256((size_t)(P.max_iter + 8) * 4));
            if (rc) return rc;
            if (wsl.h_counters_cap < P.max_iter + 8) {
                if (wsl.h_counters) cudaFreeHost(wsl.h_counters);
                wsl.h_counters = nullptr;
                CU_TRY(cudaHostAlloc(&wsl.h_counters, sizeof(int) * (P.max_iter + 8), cudaHostAllocMapped));
                CU_TRY(cudaHostGetDevicePointer(&wsl.d_counters_map, wsl.h_counters, 0));
                wsl.h_counters_cap = P.max_iter + 8;
            }
        }
    }
    struct Pending {
        bool active = false;
        long long b0 = 0, bc = 0;
        cudaStream_t s = nullptr;
        PsnRhs rhs;
        double* d_u = nullptr;
        int32_t* d_it = nullptr;
        double* d_rr = nullptr;
        int32_t* d_st = nullptr;
        PsnStreamWs ws;
        PsnStreamCfg cfg;
        int it_end = 0;
        Slot* staging = nullptr;
    } pending[2];
    int est_iters = 0;   // outer iterations the previous chunks needed (0: unknown -> poll)

    // Small per-sample outputs of host callers go through a pinned bounce buffer: an async D2H into
    // pageable memory would block the host until the whole chunk (including its multi-MB copies) has
    // drained, which serialises the pipeline.
    int32_t* hb_it = nullptr;
    double* hb_rr = nullptr;
    int32_t* hb_st = nullptr;
    if (any_host && (k_it == MEM_HOST || k_rr == MEM_HOST || k_st == MEM_HOST)) {
        const size_t need = (size_t)P.B * 16;
        if (g.h_info_cap < need) {
            if (g.h_info) cudaFreeHost(g.h_info);
            g.h_info = nullptr;
            g.h_info_cap = 0;
            CU_TRY(cudaMallocHost(&g.h_info, need));
            g.h_info_cap = need;
        }
        hb_rr = reinterpret_cast<double*>(g.h_info);
        hb_it = reinterpret_cast<int32_t*>(g.h_info + (size_t)P.B * 8);
        hb_st = reinterpret_cast<int32_t*>(g.h_info + (size_t)P.B * 12);
    }

    auto prepare = [&](long long c, ChunkCtx& cx) -> int {
        int rc2;
        cx.b0 = bounds[c];
        cx.bc = bounds[c + 1] - bounds[c];
        cx.slot = &g.slots[c % nslots];
        cx.s = any_host ? g.chunk_stream[c % PSN_NCHUNK_STREAMS] : P.stream;
        const long long b0 = cx.b0, bc = cx.bc;
        const cudaStream_t solve_s = cx.s;
        // uploads either ride the chunk's own stream or a dedicated upload stream (FIFO over chunks) that the solve waits for
        cudaStream_t s = ((deep && g.host_h2d_stream) || (copy_dedicated && any_host)) ? g.h2d_stream : solve_s;
        if (c < nslots) {
            rc2 = arena_reserve(cx.slot->arena, pad256(stage_ps * Bc) + 64 * 256 + (size_t)Bc * 64);
            if (rc2) return rc2;
        } else {
            cx.slot->arena.off = 0;
            CU_TRY(cudaStreamWaitEvent(s, cx.slot->staging_free, 0));   // chunk c - nslots has left the staging buffers
        }
        Arena& ar = cx.slot->arena;
        trace.mark("h2d_begin", c, s);
        cx.rhs.mn = P.mn;
        cx.rhs.step = P.step;
        if ((rc2 = stage_in(ar, s, P.theta, k_theta, b0 * (long long)mm, bc * mm, &cx.rhs.theta))) return rc2;
        if ((rc2 = stage_in(ar, s, P.f, k_f, b0 * (long long)mm, bc * mm, &cx.rhs.f))) return rc2;
        if ((rc2 = stage_in(ar, s, P.sx, k_sx, b0 * (long long)P.mn, bc * (size_t)P.mn, &cx.rhs.sx))) return rc2;
        if ((rc2 = stage_in(ar, s, P.sy, k_sy, b0 * (long long)P.mn, bc * (size_t)P.mn, &cx.rhs.sy))) return rc2;
        if ((rc2 = stage_in(ar, s, P.k1, k_k1, b0, (size_t)bc, &cx.rhs.k1))) return rc2;
        if ((rc2 = stage_in(ar, s, P.k2, k_k2, b0, (size_t)bc, &cx.rhs.k2))) return rc2;
        cx.d_u = nullptr; cx.d_it = nullptr; cx.d_rr = nullptr; cx.d_st = nullptr;
        if ((rc2 = stage_out(ar, P.u, k_u, b0 * (long long)nn, bc * nn, &cx.d_u, !resident))) return rc2;
        if ((rc2 = stage_out(ar, P.iters, k_it, b0, (size_t)bc, &cx.d_it, false))) return rc2;
        if ((rc2 = stage_out(ar, P.relres, k_rr, b0, (size_t)bc, &cx.d_rr, false))) return rc2;
        if ((rc2 = stage_out(ar, P.status, k_st, b0, (size_t)bc, &cx.d_st, false))) return rc2;
        if (resident || P.sub > 0) {
            PsnSolveArgs& a = cx.a;
            memset(&a, 0, sizeof(a));
            a.rhs = cx.rhs;
            a.rtol = P.rtol;
            a.max_iter = P.max_iter;
            a.u = cx.d_u;
            a.iters = cx.d_it;
            a.relres = cx.d_rr;
            a.status = cx.d_st;
            a.sub = P.sub;
            if (P.sub > 0) {
                if ((rc2 = stage_in(ar, s, P.start_x, k_x0, b0, (size_t)bc, &a.start_x))) return rc2;
                if ((rc2 = stage_in(ar, s, P.start_y, k_y0, b0, (size_t)bc, &a.start_y))) return rc2;
                if ((rc2 = stage_out(ar, P.u_fine, k_uf, b0 * (long long)ss, bc * ss, &a.u_fine, false))) return rc2;
                if ((rc2 = stage_out(ar, P.f_fine, k_ff, b0 * (long long)ss, bc * ss, &a.f_fine, false))) return rc2;
                if ((rc2 = stage_out(ar, P.th_fine, k_tf, b0 * (long long)ss, bc * ss, &a.th_fine, false))) return rc2;
                if ((rc2 = stage_out(ar, P.u_coarse, k_uc, b0 * (long long)scc, bc * scc, &a.u_coarse, false))) return rc2;
                if ((rc2 = stage_out(ar, P.f_coarse, k_fc, b0 * (long long)scc, bc * scc, &a.f_coarse, false))) return rc2;
                if ((rc2 = stage_out(ar, P.th_coarse, k_tc, b0 * (long long)scc, bc * scc, &a.th_coarse, false))) return rc2;
            }
        }
        trace.mark("h2d_end", c, s);
        if (s != solve_s) {
            CU_TRY(cudaEventRecord(cx.slot->h2d_done, s));
            CU_TRY(cudaStreamWaitEvent(solve_s, cx.slot->h2d_done, 0));
        }
        return PSN_OK;
    };

    auto run = [&](long long c, ChunkCtx& cx) -> int {
        int rc2;
        cudaStream_t s = cx.s;
        const long long b0 = cx.b0, bc = cx.bc;
        if (resident) trace.mark("solve_begin", c, s);
        if (resident) {
            const PsnSolveArgs& a = cx.a;
            {
                ProfScope pr("smem_solve", s);
                const unsigned grid = (unsigned)bc;
                const int bxk = !g.res_block ? 0 : psn_block_ok<20>(plan) ? 20 : psn_block_ok<10>(plan) ? 10 : psn_block_ok<5>(plan) ? 5 : 0;
                if (bxk) {
                    size_t bsm = (size_t)(bxk == 20 ? psn_block_smem_doubles<20>(plan)
                                                    : bxk == 10 ? psn_block_smem_doubles<10>(plan) : psn_block_smem_doubles<5>(plan)) * sizeof(double);
                    int invd = psn_coarse_inv_doubles(plan);
                    if (!g.inv_smem || bsm + (size_t)invd * sizeof(double) + 1024 > g.smem_optin) invd = 0;
                    bsm += (size_t)invd * sizeof(double);
                    const int per_sm = bxk == 20 ? PSN_BLOCK_CTAS(20) : bxk == 10 ? PSN_BLOCK_CTAS(10) : PSN_BLOCK_CTAS(5);
                    const unsigned pgrid =
                        (unsigned)std::min<long long>(bc, (long long)g.sm_count * per_sm * (g.res_waves > 0 ? g.res_waves : 1));
                    const unsigned gr = g.res_persist ? pgrid : grid;
                    if (bxk == 20)
                        psn_block_solve_kernel<20><<<gr, PsnBlockGeom<20>::NTPAD, bsm, s>>>(plan, a, bc, invd);
                    else if (bxk == 10)
                        psn_block_solve_kernel<10><<<gr, PsnBlockGeom<10>::NTPAD, bsm, s>>>(plan, a, bc, invd);
                    else
                        psn_block_solve_kernel<5><<<gr, PsnBlockGeom<5>::NTPAD, bsm, s>>>(plan, a, bc, invd);
                } else {
#define PSN_RES_LAUNCH(T_, X_) \
    if (rcfg.T == T_ && rcfg.XPT == X_) psn_smem_solve_kernel<T_, X_><<<grid, T_, rcfg.smem, s>>>(plan, a)
                PSN_RES_LAUNCH(128, 4); PSN_RES_LAUNCH(256, 4); PSN_RES_LAUNCH(256, 8); PSN_RES_LAUNCH(512, 4);
                PSN_RES_LAUNCH(512, 8); PSN_RES_LAUNCH(1024, 4); PSN_RES_LAUNCH(512, 16); PSN_RES_LAUNCH(1024, 8);
#undef PSN_RES_LAUNCH
                }
            }
            CU_TRY(cudaGetLastError());
            if (P.sub > 0) {
                if ((rc2 = copy_back(s, P.u_fine, k_uf, b0 * (long long)ss, bc * ss, a.u_fine))) return rc2;
                if ((rc2 = copy_back(s, P.f_fine, k_ff, b0 * (long long)ss, bc * ss, a.f_fine))) return rc2;
                if ((rc2 = copy_back(s, P.th_fine, k_tf, b0 * (long long)ss, bc * ss, a.th_fine))) return rc2;
                if ((rc2 = copy_back(s, P.u_coarse, k_uc, b0 * (long long)scc, bc * scc, a.u_coarse))) return rc2;
                if ((rc2 = copy_back(s, P.f_coarse, k_fc, b0 * (long long)scc, bc * scc, a.f_coarse))) return rc2;
                if ((rc2 = copy_back(s, P.th_coarse, k_tc, b0 * (long long)scc, bc * scc, a.th_coarse))) return rc2;
            }
        } else {
            const int wk = (nws == 2) ? (int)(c & 1) : 0;
            Slot& wsl = g.slots[wk];
            // workspace wk is free once the last kernel of the chunk that used it before has run
            if (c >= nws && any_host) CU_TRY(cudaStreamWaitEvent(s, g.ws_free[wk], 0));
            // solves run back to back (the other workspace's chunk first): interleaving two chunks' kernels was
            // measured slower than keeping each chunk's working set to itself
            if (c >= 1 && any_host && nws == 2 && g.serialize_solves) CU_TRY(cudaStreamWaitEvent(s, g.ws_free[wk ^ 1], 0));
            wsl.ws_arena.off = 0;
            trace.mark("solve_begin", c, s);
            PsnStreamWs ws;
            if ((rc2 = carve_stream_ws(wsl.ws_arena, plan, scfg, bc, P.max_iter, &ws))) return rc2;
            choose_bands(plan, &scfg, bc);
            CudaBackend bk{s, &wsl};
            // After the first chunk has told us how many outer iterations this workload needs, later chunks are
            // enqueued with that count (+1 spare) without the host waiting on the device; stragglers are caught by
            // the deferred check (resolve) and resumed.
            const bool speculative = any_host && nws == 2 && P.sub == 0 && est_iters > 0;
            const int it_end = psn_stream_solve(bk, plan, scfg, cx.rhs, bc, P.rtol, P.max_iter, cx.d_u, ws, 0,
                                                speculative ? est_iters + 1 : 0);
            if (bk.err != cudaSuccess) return psn_fail(PSN_ERR_CUDA, "streamed solve failed: %s", cudaGetErrorString(bk.err));
            if (!speculative && it_end - PSN_COUNTER_LAG > est_iters) est_iters = it_end - PSN_COUNTER_LAG;
            {
                ProfScope pr("state", s);
                psn_finalize_kernel<<<(unsigned)((bc + 127) / 128), 128, 0, s>>>(ws.state, bc, cx.d_it, cx.d_rr, cx.d_st);
            }
            CU_TRY(cudaGetLastError());
            if (speculative) {
                Pending& pd = pending[wk];
                pd.active = true;
                pd.b0 = b0; pd.bc = bc; pd.s = s; pd.rhs = cx.rhs;
                pd.d_u = cx.d_u; pd.d_it = cx.d_it; pd.d_rr = cx.d_rr; pd.d_st = cx.d_st;
                pd.ws = ws; pd.cfg = scfg; pd.it_end = it_end; pd.staging = cx.slot;
            }
            if (P.sub > 0) {
                const PsnSolveArgs& a = cx.a;
                psn_window_kernel<<<dim3((unsigned)((P.sub * P.sub + 255) / 256), (unsigned)bc), 256, 0, s>>>(a, P.n, cx.d_u);
                CU_TRY(cudaGetLastError());
                if ((rc2 = copy_back(s, P.u_fine, k_uf, b0 * (long long)ss, bc * ss, a.u_fine))) return rc2;
                if ((rc2 = copy_back(s, P.f_fine, k_ff, b0 * (long long)ss, bc * ss, a.f_fine))) return rc2;
                if ((rc2 = copy_back(s, P.th_fine, k_tf, b0 * (long long)ss, bc * ss, a.th_fine))) return rc2;
                if ((rc2 = copy_back(s, P.u_coarse, k_uc, b0 * (long long)scc, bc * scc, a.u_coarse))) return rc2;
                if ((rc2 = copy_back(s, P.f_coarse, k_fc, b0 * (long long)scc, bc * scc, a.f_coarse))) return rc2;
                if ((rc2 = copy_back(s, P.th_coarse, k_tc, b0 * (long long)scc, bc * scc, a.th_coarse))) return rc2;
            }
            if (any_host) CU_TRY(cudaEventRecord(g.ws_free[wk], s));
        }
        trace.mark("solve_end", c, s);
        cudaStream_t cs = s;
        if (copy_dedicated && any_host) {       // result copies on their own stream, behind this chunk's last kernel
            CU_TRY(cudaEventRecord(cx.slot->solve_done, s));
            CU_TRY(cudaStreamWaitEvent(g.d2h_stream, cx.slot->solve_done, 0));
            cs = g.d2h_stream;
        }
        if ((rc2 = copy_back(cs, P.u, k_u, b0 * (long long)nn, bc * nn, cx.d_u))) return rc2;
        if ((rc2 = copy_back(cs, hb_it ? hb_it : P.iters, k_it, b0, (size_t)bc, cx.d_it))) return rc2;
        if ((rc2 = copy_back(cs, hb_rr ? hb_rr : P.relres, k_rr, b0, (size_t)bc, cx.d_rr))) return rc2;
        if ((rc2 = copy_back(cs, hb_st ? hb_st : P.status, k_st, b0, (size_t)bc, cx.d_st))) return rc2;
        trace.mark("d2h_end", c, cs);
        if (any_host) CU_TRY(cudaEventRecord(cx.slot->staging_free, cs));
        return PSN_OK;
    };

    // deferred convergence check of a speculatively enqueued chunk; resumes it (polling) if samples are still active
    auto resolve = [&](int k) -> int {
        Pending& pd = pending[k];
        if (!pd.active) return PSN_OK;
        pd.active = false;
        Slot& wsl = g.slots[k];
        CudaBackend bk{pd.s, &wsl};
        const int left = bk.wait_counter(pd.it_end - 1);
        if (bk.err != cudaSuccess) return psn_fail(PSN_ERR_CUDA, "streamed solve failed: %s", cudaGetErrorString(bk.err));
        if (left == 0) return PSN_OK;
        const int it2 = psn_stream_solve(bk, plan, pd.cfg, pd.rhs, pd.bc, P.rtol, P.max_iter, pd.d_u, pd.ws, pd.it_end, 0);
        if (bk.err != cudaSuccess) return psn_fail(PSN_ERR_CUDA, "streamed solve failed: %s", cudaGetErrorString(bk.err));
        if (it2 - PSN_COUNTER_LAG > est_iters) est_iters = it2 - PSN_COUNTER_LAG;
        psn_finalize_kernel<<<(unsigned)((pd.bc + 127) / 128), 128, 0, pd.s>>>(pd.ws.state, pd.bc, pd.d_it, pd.d_rr, pd.d_st);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(g.ws_free[k], pd.s));
        int rc2;
        if ((rc2 = copy_back(pd.s, P.u, k_u, pd.b0 * (long long)nn, pd.bc * nn, pd.d_u))) return rc2;
        if ((rc2 = copy_back(pd.s, hb_it ? hb_it : P.iters, k_it, pd.b0, (size_t)pd.bc, pd.d_it))) return rc2;
        if ((rc2 = copy_back(pd.s, hb_rr ? hb_rr : P.relres, k_rr, pd.b0, (size_t)pd.bc, pd.d_rr))) return rc2;
        if ((rc2 = copy_back(pd.s, hb_st ? hb_st : P.status, k_st, pd.b0, (size_t)pd.bc, pd.d_st))) return rc2;
        CU_TRY(cudaEventRecord(pd.staging->staging_free, pd.s));
        return PSN_OK;
    };

    // Uploads run `ahead` chunks in front of the solve being enqueued.  Chunk c + ahead reuses the staging slot of chunk
    // c + ahead - nslots, whose last copy is enqueued by run() or, for a resumed straggler, by resolve() two iterations
    // later: ahead <= nslots - 2 keeps every staging_free record ahead (in host order) of the wait that needs it.  Without
    // the upload stream the copy rides the solve stream of chunk c + ahead, which chunk c + ahead - PSN_NCHUNK_STREAMS
    // used: ahead < PSN_NCHUNK_STREAMS keeps it out of the way of a solve that is yet to be enqueued.
    int ahead = 1;
    if (deep) {
        ahead = g.host_prefetch;
        if (ahead > nslots - 2) ahead = nslots - 2;
        if (ahead > PSN_NCHUNK_STREAMS - 1) ahead = PSN_NCHUNK_STREAMS - 1;
        if (ahead < 1) ahead = 1;
    }
    ChunkCtx ring[PSN_NCHUNK_STREAMS];   // chunks c .. c + ahead
    long long prepared = 0;
    for (long long c = 0; c < nchunks; ++c) {
        // chunk c-2 used the workspace (and chunk c+1 reuses the staging slot) that are about to be recycled
        if (nws == 2 && (rc = resolve((int)(c & 1)))) return rc;
        for (; prepared <= c + ahead && prepared < nchunks; ++prepared)
            if ((rc = prepare(prepared, ring[prepared % PSN_NCHUNK_STREAMS]))) return rc;
        if ((rc = run(c, ring[c % PSN_NCHUNK_STREAMS]))) return rc;
    }
    if (nws == 2) {
        if ((rc = resolve(0))) return rc;
        if ((rc = resolve(1))) return rc;
    }
    if (any_host) {
        for (int i = 0; i < PSN_NCHUNK_STREAMS; ++i) CU_TRY(cudaStreamSynchronize(g.chunk_stream[i]));
        CU_TRY(cudaStreamSynchronize(g.h2d_stream));
        CU_TRY(cudaStreamSynchronize(g.d2h_stream));
        if (hb_it && k_it == MEM_HOST) memcpy(P.iters, hb_it, (size_t)P.B * 4);
        if (hb_rr && k_rr == MEM_HOST) memcpy(P.relres, hb_rr, (size_t)P.B * 8);
        if (hb_st && k_st == MEM_HOST) memcpy(P.status, hb_st, (size_t)P.B * 4);
        trace.dump();
    }
    return PSN_OK;
}

}  // namespace

// ----------------------------------------------------------------------------- C ABI
extern "C" {

int psn_version(void) { return PSN_VERSION; }

int psn_get_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return c;
}

const char* psn_last_error(void) { return g_err; }

int psn_init(int device) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    int count = psn_get_device_count();
    if (count <= 0) return psn_fail(PSN_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device >= count) return psn_fail(PSN_ERR_INVALID, "device %d out of range (%d devices)", device, count);
    if (device >= 0) CU_TRY(cudaSetDevice(device));
    int cur = 0;
    CU_TRY(cudaGetDevice(&cur));
    if (g.ready && g.device == cur) return PSN_OK;
    if (g.ready) psn_shutdown();
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, cur));
    g.device = cur;
    g.sm_count = prop.multiProcessorCount;
    g.smem_optin = prop.sharedMemPerBlockOptin;
    for (int k = 0; k < 2; ++k) CU_TRY(cudaEventCreateWithFlags(&g.ws_free[k], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&g.last_call, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) CU_TRY(cudaEventCreateWithFlags(&g.master_done[k], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&g.multi_solved, cudaEventDisableTiming));
    g.have_last = false;
    for (int i = 0; i < PSN_NSLOTS; ++i) {
        CU_TRY(cudaEventCreateWithFlags(&g.slots[i].staging_free, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&g.slots[i].h2d_done, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&g.slots[i].solve_done, cudaEventDisableTiming));
        for (int k = 0; k < 4; ++k) CU_TRY(cudaEventCreateWithFlags(&g.slots[i].ev[k], cudaEventDisableTiming));
    }
    {
        int prio_lo = 0, prio_hi = 0;   // numerically lower = higher priority
        CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        for (int i = 0; i < PSN_NCHUNK_STREAMS; ++i) {
            int pr = prio_hi + i;
            if (pr > prio_lo) pr = prio_lo;
            CU_TRY(cudaStreamCreateWithPriority(&g.chunk_stream[i], cudaStreamNonBlocking, pr));
        }
        CU_TRY(cudaStreamCreateWithPriority(&g.h2d_stream, cudaStreamNonBlocking, prio_hi));
        CU_TRY(cudaStreamCreateWithPriority(&g.d2h_stream, cudaStreamNonBlocking, prio_hi));
    }
    const char* env = getenv("PSN_WS_BYTES");
    if (env) g.ws_bytes = atoll(env);
    g.attrs_set = false;
    g.ready = true;
    return PSN_OK;
}

void psn_shutdown(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!g.ready) return;
    cudaDeviceSynchronize();
    for (auto& kv : g.plans) cudaFree(kv.second.d_inv);
    g.plans.clear();
    for (int k = 0; k < 2; ++k) {
        if (g.ws_free[k]) cudaEventDestroy(g.ws_free[k]);
        g.ws_free[k] = nullptr;
    }
    if (g.last_call) cudaEventDestroy(g.last_call);
    g.last_call = nullptr;
    if (g.mt_arena.base) cudaFree(g.mt_arena.base);
    g.mt_arena = Arena();
    for (int k = 0; k < 2; ++k) {
        if (g.master_done[k]) cudaEventDestroy(g.master_done[k]);
        g.master_done[k] = nullptr;
        if (g.master[k].base) cudaFree(g.master[k].base);
        g.master[k] = Arena();
    }
    if (g.multi_solved) cudaEventDestroy(g.multi_solved);
    g.multi_solved = nullptr;
    if (g.multi_out.base) cudaFree(g.multi_out.base);
    g.multi_out = Arena();
    g.have_last = false;
    if (g.h_info) cudaFreeHost(g.h_info);
    g.h_info = nullptr;
    g.h_info_cap = 0;
    for (int i = 0; i < PSN_NCHUNK_STREAMS; ++i) {
        if (g.chunk_stream[i]) cudaStreamDestroy(g.chunk_stream[i]);
        g.chunk_stream[i] = nullptr;
    }
    if (g.h2d_stream) cudaStreamDestroy(g.h2d_stream);
    g.h2d_stream = nullptr;
    if (g.d2h_stream) cudaStreamDestroy(g.d2h_stream);
    g.d2h_stream = nullptr;
    for (int i = 0; i < PSN_NSLOTS; ++i) {
        Slot& s = g.slots[i];
        if (s.arena.base) cudaFree(s.arena.base);
        if (s.ws_arena.base) cudaFree(s.ws_arena.base);
        if (s.h_counters) cudaFreeHost(s.h_counters);
        for (int k = 0; k < 4; ++k)
            if (s.ev[k]) cudaEventDestroy(s.ev[k]);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.staging_free) cudaEventDestroy(s.staging_free);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.solve_done) cudaEventDestroy(s.solve_done);
        s = Slot();
    }
    for (auto& r : g.prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g.prof.clear();
    g.ready = false;
    cudaGetLastError();
}

int psn_solve_batch(int n, int64_t batch, const double* theta, const double* f, const double* sin_x, const double* sin_y,
                    const double* k1, const double* k2, int master_n, int step, int operator_mode, double rtol, int max_iter,
                    double* u, int32_t* iters, double* relres, int32_t* status, void* stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_err[0] = 0;
    if (operator_mode != PSN_OPERATOR_REFERENCE)
        return psn_fail(PSN_ERR_UNSUPPORTED, "operator_mode %d not implemented (only the reference operator diag(theta) L u = f)", operator_mode);
    Problem P;
    memset(&P, 0, sizeof(P));
    P.n = n; P.B = batch; P.theta = theta; P.f = f; P.sx = sin_x; P.sy = sin_y; P.k1 = k1; P.k2 = k2;
    P.mn = master_n; P.step = step; P.rtol = rtol; P.max_iter = max_iter;
    P.u = u; P.iters = iters; P.relres = relres; P.status = status; P.stream = (cudaStream_t)stream;
    return solve_internal(P);
}

int psn_subdomain_batch(int n_super, int n_fine, int64_t batch, const double* sin_x, const double* sin_y,
                        const int32_t* start_x, const int32_t* start_y, double rtol, int max_iter,
                        double* u_fine, double* f_fine, double* theta_fine, double* u_coarse, double* f_coarse, double* theta_coarse,
                        int32_t* iters, double* relres, int32_t* status, void* stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_err[0] = 0;
    if (n_fine < 1 || n_fine > n_super) return psn_fail(PSN_ERR_INVALID, "need 1 <= n_fine <= n_super");
    if (!sin_x || !sin_y) return psn_fail(PSN_ERR_INVALID, "sin_x / sin_y are required");
    Problem P;
    memset(&P, 0, sizeof(P));
    P.n = n_super; P.B = batch; P.sx = sin_x; P.sy = sin_y; P.mn = n_super; P.step = 1;
    P.rtol = rtol; P.max_iter = max_iter; P.iters = iters; P.relres = relres; P.status = status;
    P.sub = n_fine; P.start_x = start_x; P.start_y = start_y;
    P.u_fine = u_fine; P.f_fine = f_fine; P.th_fine = theta_fine;
    P.u_coarse = u_coarse; P.f_coarse = f_coarse; P.th_coarse = theta_coarse;
    P.stream = (cudaStream_t)stream;
    return solve_internal(P);
}

int psn_multires(int n_finest, const int32_t* resolutions, int n_res, int64_t batch, const double* theta_finest,
                 const double* sin_x, const double* sin_y, double rtol, int max_iter, double* const* u_out,
                 int32_t* iters, double* relres, int32_t* status, void* stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_err[0] = 0;
    if (!resolutions || n_res < 1 || !u_out || !sin_x || !sin_y) return psn_fail(PSN_ERR_INVALID, "resolutions, u_out, sin_x, sin_y are required");
    for (int k = 0; k < n_res; ++k) {
        const int r = resolutions[k];
        if (r < 2 || r > n_finest || n_finest % r != 0) return psn_fail(PSN_ERR_INVALID, "resolution %d does not divide n_finest=%d", r, n_finest);
    }
    int rc = ensure_init();
    if (rc) return rc;
    if (batch < 0) return psn_fail(PSN_ERR_INVALID, "batch must be >= 0");
    if (batch == 0) return PSN_OK;
    // Host outputs: every resolution is solved into device buffers on the caller's stream, largest grid first, and copied
    // back on the result-copy stream while the next resolution is being solved (before: solve, copy, synchronise per
    // resolution -- for 64 examples x 40..640 the copies were half of the call).
    bool pipelined = g.multires_pipeline != 0;
    for (int k = 0; k < n_res; ++k) pipelined = pipelined && classify(u_out[k]) == MEM_HOST;
    if (iters && classify(iters) != MEM_HOST) pipelined = false;
    if (relres && classify(relres) != MEM_HOST) pipelined = false;
    if (status && classify(status) != MEM_HOST) pipelined = false;
    std::vector<int> order(n_res);
    for (int k = 0; k < n_res; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return resolutions[a] > resolutions[b]; });
    size_t out_per = 0;   // bytes of device results per example
    for (int k = 0; k < n_res; ++k) out_per += pad256((size_t)resolutions[k] * resolutions[k] * 8) + 16;
    auto solve_range_pipelined = [&](const double* th, const double* sx, const double* sy, long long b0, long long bc) -> int {
        const cudaStream_t s = (cudaStream_t)stream;
        int rc2 = arena_reserve(g.multi_out, out_per * (size_t)bc + (size_t)n_res * 4 * 256);
        if (rc2) return rc2;
        std::vector<double*> d_u(n_res);
        std::vector<int32_t*> d_it(n_res), d_st(n_res);
        std::vector<double*> d_rr(n_res);
        for (int k = 0; k < n_res; ++k) {
            d_u[k] = arena_take<double>(g.multi_out, (size_t)bc * resolutions[k] * resolutions[k]);
            d_it[k] = arena_take<int32_t>(g.multi_out, (size_t)bc);
            d_rr[k] = arena_take<double>(g.multi_out, (size_t)bc);
            d_st[k] = arena_take<int32_t>(g.multi_out, (size_t)bc);
            if (!d_u[k] || !d_it[k] || !d_rr[k] || !d_st[k]) return psn_fail(PSN_ERR_CUDA, "multires result arena too small");
        }
        for (int o = 0; o < n_res; ++o) {
            const int k = order[o];
            Problem P;
            memset(&P, 0, sizeof(P));
            const size_t rr = (size_t)resolutions[k] * resolutions[k];
            P.n = resolutions[k]; P.B = bc; P.theta = th; P.sx = sx; P.sy = sy;
            P.mn = n_finest; P.step = n_finest / resolutions[k]; P.rtol = rtol; P.max_iter = max_iter;
            P.u = d_u[k]; P.iters = d_it[k]; P.relres = d_rr[k]; P.status = d_st[k];
            P.stream = s;
            if ((rc2 = solve_internal(P))) return rc2;
            CU_TRY(cudaEventRecord(g.multi_solved, s));
            CU_TRY(cudaStreamWaitEvent(g.d2h_stream, g.multi_solved, 0));
            CU_TRY(cudaMemcpyAsync(u_out[k] + (size_t)b0 * rr, d_u[k], (size_t)bc * rr * 8, cudaMemcpyDeviceToHost, g.d2h_stream));
        }
        CU_TRY(cudaStreamSynchronize(g.d2h_stream));   // (behind the last solve: the caller's stream is complete as well)
        for (int k = 0; k < n_res; ++k) {              // small, and the callers' arrays are usually pageable
            if (iters) CU_TRY(cudaMemcpy(iters + (size_t)k * batch + b0, d_it[k], (size_t)bc * 4, cudaMemcpyDeviceToHost));
            if (relres) CU_TRY(cudaMemcpy(relres + (size_t)k * batch + b0, d_rr[k], (size_t)bc * 8, cudaMemcpyDeviceToHost));
            if (status) CU_TRY(cudaMemcpy(status + (size_t)k * batch + b0, d_st[k], (size_t)bc * 4, cudaMemcpyDeviceToHost));
        }
        return PSN_OK;
    };
    auto solve_range = [&](const double* th, const double* sx, const double* sy, long long b0, long long bc) -> int {
        if (pipelined) return solve_range_pipelined(th, sx, sy, b0, bc);
        for (int k = 0; k < n_res; ++k) {
            Problem P;
            memset(&P, 0, sizeof(P));
            const size_t rr = (size_t)resolutions[k] * resolutions[k];
            P.n = resolutions[k]; P.B = bc; P.theta = th; P.sx = sx; P.sy = sy;
            P.mn = n_finest; P.step = n_finest / resolutions[k]; P.rtol = rtol; P.max_iter = max_iter;
            P.u = u_out[k] + (size_t)b0 * rr;
            P.iters = iters ? iters + (size_t)k * batch + b0 : nullptr;
            P.relres = relres ? relres + (size_t)k * batch + b0 : nullptr;
            P.status = status ? status + (size_t)k * batch + b0 : nullptr;
            P.stream = (cudaStream_t)stream;
            int rc2 = solve_internal(P);
            if (rc2) return rc2;
        }
        return PSN_OK;
    };
    const MemKind k_th = classify(theta_finest), k_sx = classify(sin_x), k_sy = classify(sin_y);
    if (k_th != MEM_HOST && k_sx != MEM_HOST && k_sy != MEM_HOST) {
        if (!pipelined) return solve_range(theta_finest, sin_x, sin_y, 0, batch);
        long long Bo = (long long)((4ull << 30) / out_per);   // examples per result arena
        if (Bo < 1) Bo = 1;
        for (long long b0 = 0; b0 < batch; b0 += Bo) {
            const long long bc = std::min<long long>(Bo, batch - b0);
            if ((rc = solve_range(theta_finest + (size_t)b0 * n_finest * n_finest, sin_x + (size_t)b0 * n_finest,
                                  sin_y + (size_t)b0 * n_finest, b0, bc)))
                return rc;
        }
        return PSN_OK;
    }

    // Host master fields: every resolution reads the SAME theta / sine factors through a stride gather, so they are
    // uploaded once per chunk of examples (two alternating device buffers: the upload of chunk c+1 runs on the upload
    // stream while the resolutions of chunk c are solved) instead of once per resolution.
    const size_t mm = (size_t)n_finest * n_finest;
    const size_t per = (k_th == MEM_HOST ? mm * 8 : 0) + 2 * pad256((size_t)n_finest * 8) + 64;
    long long Bm = (long long)((4ull << 30) / (per + (pipelined ? out_per : 0)));
    if (Bm < 1) Bm = 1;
    if (Bm > batch) Bm = batch;
    const long long nchunks = (batch + Bm - 1) / Bm;
    const double *d_th[2] = {nullptr, nullptr}, *d_sx[2] = {nullptr, nullptr}, *d_sy[2] = {nullptr, nullptr};
    auto upload = [&](long long c) -> int {
        const int k = (int)(c & 1);
        const long long b0 = c * Bm, bc = std::min<long long>(Bm, batch - b0);
        Arena& ar = g.master[k];
        int rc2 = arena_reserve(ar, per * (size_t)Bm + 4 * 256);
        if (rc2) return rc2;
        if ((rc2 = stage_in(ar, g.h2d_stream, theta_finest, k_th, b0 * (long long)mm, (size_t)bc * mm, &d_th[k]))) return rc2;
        if ((rc2 = stage_in(ar, g.h2d_stream, sin_x, k_sx, b0 * (long long)n_finest, (size_t)bc * n_finest, &d_sx[k]))) return rc2;
        if ((rc2 = stage_in(ar, g.h2d_stream, sin_y, k_sy, b0 * (long long)n_finest, (size_t)bc * n_finest, &d_sy[k]))) return rc2;
        CU_TRY(cudaEventRecord(g.master_done[k], g.h2d_stream));
        return PSN_OK;
    };
    CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));   // device inputs of the caller may still be in flight
    if ((rc = call_begin(g.h2d_stream))) return rc;
    if ((rc = upload(0))) return rc;
    for (long long c = 0; c < nchunks; ++c) {
        // buffer (c+1)&1 was last read by the solves of chunk c-1: complete for host outputs (those calls return
        // synchronised), possibly still in flight on the caller's stream when every output is a device array
        if (c >= 1 && c + 1 < nchunks) CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
        if (c + 1 < nchunks && (rc = upload(c + 1))) return rc;
        CU_TRY(cudaEventSynchronize(g.master_done[c & 1]));
        const long long b0 = c * Bm, bc = std::min<long long>(Bm, batch - b0);
        if ((rc = solve_range(d_th[c & 1], d_sx[c & 1], d_sy[c & 1], b0, bc))) return rc;
    }
    CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));   // all-device outputs: the master buffers must outlive the solves
    return PSN_OK;
}

int psn_dataset_tensors(int64_t n_samples, int n_coarse, int n_fine, const double* u_coarse, const double* u_fine,
                        const double* f_fine, const double* theta_fine, float* x, float* y, double* stats, void* stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_err[0] = 0;
    int rc = ensure_init();
    if (rc) return rc;
    if (n_samples < 0 || n_coarse < 1 || n_fine < 1) return psn_fail(PSN_ERR_INVALID, "bad sizes");
    if (!u_coarse || !u_fine || !f_fine || !theta_fine || !x || !y) return psn_fail(PSN_ERR_INVALID, "NULL array");
    if (n_samples == 0) return PSN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = call_begin(s))) return rc;
    const size_t cc = (size_t)n_samples * n_coarse * n_coarse, cf = (size_t)n_samples * n_fine * n_fine;
    const MemKind k_uc = classify(u_coarse), k_uf = classify(u_fine), k_ff = classify(f_fine), k_th = classify(theta_fine),
                  k_x = classify(x), k_y = classify(y), k_s = classify(stats);
    Arena& ar = g.slots[0].arena;
    size_t need = 64 * 256 + pad256((size_t)3 * 2 * PSN_MOM_BLOCKS * 8) + 256;
    if (k_uc == MEM_HOST) need += pad256(cc * 8);
    if (k_uf == MEM_HOST) need += pad256(cf * 8);
    if (k_ff == MEM_HOST) need += pad256(cf * 8);
    if (k_th == MEM_HOST) need += pad256(cf * 8);
    if (k_x == MEM_HOST) need += pad256(cf * 3 * 4);
    if (k_y == MEM_HOST) need += pad256(cf * 4);
    if ((rc = arena_reserve(ar, need))) return rc;
    const double *d_uc, *d_uf, *d_ff, *d_th;
    if ((rc = stage_in(ar, s, u_coarse, k_uc, 0, cc, &d_uc))) return rc;
    if ((rc = stage_in(ar, s, u_fine, k_uf, 0, cf, &d_uf))) return rc;
    if ((rc = stage_in(ar, s, f_fine, k_ff, 0, cf, &d_ff))) return rc;
    if ((rc = stage_in(ar, s, theta_fine, k_th, 0, cf, &d_th))) return rc;
    float *d_x, *d_y;
    if ((rc = stage_out(ar, x, k_x, 0, cf * 3, &d_x, true))) return rc;
    if ((rc = stage_out(ar, y, k_y, 0, cf, &d_y, true))) return rc;
    double* d_part = arena_take<double>(ar, (size_t)3 * 2 * PSN_MOM_BLOCKS);
    double* d_stats = (k_s == MEM_DEVICE) ? stats : arena_take<double>(ar, 8);
    if (!d_part || !d_stats) return psn_fail(PSN_ERR_NOMEM, "staging arena exhausted");
    const double* arrs[3] = {d_uf, d_ff, d_th};
    for (int k = 0; k < 3; ++k) {
        ProfScope pr("dataset_moments", s);
        psn_moments_kernel<<<PSN_MOM_BLOCKS, 256, 0, s>>>(arrs[k], (long long)cf, d_part + (size_t)k * 2 * PSN_MOM_BLOCKS);
    }
    psn_moments_final_kernel<<<1, 32, 0, s>>>(d_part, PSN_MOM_BLOCKS, (long long)cf, d_stats);
    {
        ProfScope pr("dataset_tensors", s);
        const unsigned grid = (unsigned)std::min<size_t>((cf + 255) / 256, (size_t)g.sm_count * 16);
        psn_tensor_kernel<<<grid, 256, 0, s>>>((long long)n_samples, n_coarse, n_fine, d_uc, d_uf, d_ff, d_th, d_stats, d_x, d_y);
    }
    CU_TRY(cudaGetLastError());
    if ((rc = copy_back(s, x, k_x, 0, cf * 3, d_x))) return rc;
    if ((rc = copy_back(s, y, k_y, 0, cf, d_y))) return rc;
    if (k_s == MEM_HOST) CU_TRY(cudaMemcpyAsync(stats, d_stats, 8 * sizeof(double), cudaMemcpyDeviceToHost, s));
    const MemKind kinds[] = {k_uc, k_uf, k_ff, k_th, k_x, k_y, k_s};
    bool any_host = false;
    for (MemKind k : kinds) any_host = any_host || (k == MEM_HOST);
    if (any_host) CU_TRY(cudaStreamSynchronize(s));
    return call_end(s);
}

int psn_upscale_cascade(int mode, int n_in, int levels, int64_t batch, const double* u_in, float* out, void* stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_err[0] = 0;
    int rc = ensure_init();
    if (rc) return rc;
    if ((mode != 0 && mode != 1) || n_in < 2 || levels < 1 || levels > 8 || batch < 0) return psn_fail(PSN_ERR_INVALID, "bad arguments");
    if (!u_in || !out) return psn_fail(PSN_ERR_INVALID, "NULL array");
    if (batch == 0) return PSN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = call_begin(s))) return rc;
    const long long n_out = (long long)n_in << levels;
    const size_t c_in = (size_t)batch * n_in * n_in, c_out = (size_t)batch * n_out * n_out;
    const MemKind k_in = classify(u_in), k_out = classify(out);
    Arena& ar = g.slots[0].arena;
    size_t need = 64 * 256 + 2 * pad256(c_out * 4);
    if (k_in == MEM_HOST) need += pad256(c_in * 8);
    if ((rc = arena_reserve(ar, need))) return rc;
    const double* d_in;
    if ((rc = stage_in(ar, s, u_in, k_in, 0, c_in, &d_in))) return rc;
    float* buf[2] = {arena_take<float>(ar, c_out), arena_take<float>(ar, c_out)};
    if (!buf[0] || !buf[1]) return psn_fail(PSN_ERR_NOMEM, "staging arena exhausted");
    int n = n_in;
    const float* cur = nullptr;
    for (int l = 0; l < levels; ++l) {
        const bool last = l == levels - 1;
        float* dst = (last && k_out == MEM_DEVICE) ? out : buf[l & 1];
        const size_t cnt = (size_t)batch * (2 * n) * (2 * n);
        const unsigned grid = (unsigned)std::min<size_t>((cnt + 255) / 256, (size_t)g.sm_count * 32);
        ProfScope pr("upscale2", s);
        if (l == 0) psn_upscale2_kernel<double><<<grid, 256, 0, s>>>(mode, n, (long long)batch, d_in, dst);
        else psn_upscale2_kernel<float><<<grid, 256, 0, s>>>(mode, n, (long long)batch, cur, dst);
        cur = dst;
        n *= 2;
    }
    CU_TRY(cudaGetLastError());
    if (k_out == MEM_HOST) {
        CU_TRY(cudaMemcpyAsync(out, cur, c_out * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    if (k_in == MEM_HOST || k_out == MEM_HOST) CU_TRY(cudaStreamSynchronize(s));
    return call_end(s);
}

int psn_mt19937_jump_host(const uint32_t* key, int pos, int64_t n_words, uint32_t* key_out, int* pos_out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_err[0] = 0;
    if (!key || !key_out || !pos_out || pos < 0 || pos > psn_mt::N || n_words < 0) return psn_fail(PSN_ERR_INVALID, "bad arguments");
    if (!psn_mt::chi_ok()) return psn_fail(PSN_ERR_UNSUPPORTED, "characteristic polynomial of MT19937 could not be derived");
    if (n_words == 0) {
        memcpy(key_out, key, psn_mt::N * sizeof(uint32_t));
        *pos_out = pos;
        return PSN_OK;
    }
    // keep pos (the window moves by n_words); pos == 0: move the window by n_words - 624 and report pos = 624, so that the
    // undetermined low bits of the jumped window's first word are never output
    int64_t shift = n_words;
    int pos2 = pos;
    if (pos == 0) {
        if (n_words < psn_mt::N) return psn_fail(PSN_ERR_UNSUPPORTED, "a state with pos == 0 can only jump by >= 624 words");
        shift = n_words - psn_mt::N;
        pos2 = psn_mt::N;
    }
    psn_mt::host_apply(psn_mt::jump_poly((uint64_t)shift, 0), key, key_out);
    *pos_out = pos2;
    return PSN_OK;
}

int psn_mt19937_uniform_blocks(const uint32_t* key, int pos, int64_t n_blocks, int64_t skip, int64_t count, double low, double high,
                               double* out, double* skipped, uint32_t* key_out, int* pos_out, void* stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_err[0] = 0;
    int rc = ensure_init();
    if (rc) return rc;
    using namespace psn_mt;
    if (!key || pos < 0 || pos > N || n_blocks < 0 || skip < 0 || count < 0) return psn_fail(PSN_ERR_INVALID, "bad arguments");
    if (n_blocks > 0 && count > 0 && !out) return psn_fail(PSN_ERR_INVALID, "out is NULL");
    if (classify(key) == MEM_DEVICE || classify(key_out) == MEM_DEVICE || classify(pos_out) == MEM_DEVICE)
        return psn_fail(PSN_ERR_INVALID, "key / key_out / pos_out must be host arrays");
    const int64_t J = 2 * (skip + count);   // words per block
    if (n_blocks == 0 || J == 0) {
        if (key_out) memcpy(key_out, key, N * sizeof(uint32_t));
        if (pos_out) *pos_out = pos;
        return PSN_OK;
    }
    if (n_blocks > (1ll << 31) / 2 || J > (1ll << 40)) return psn_fail(PSN_ERR_INVALID, "stream too long");
    if (!chi_ok()) return psn_fail(PSN_ERR_UNSUPPORTED, "characteristic polynomial of MT19937 could not be derived");
    // state index b = the state block b starts from; index n_blocks = the state after the call.  pos > 0: every state
    // keeps pos and state b = t^(bJ) state 0.  pos == 0: states b >= 1 are windows 624 words further back with
    // pos = 624 (see psn_mt19937_jump_host), state 1 = t^(J - 624) state 0.
    const int first = (pos == 0) ? 1 : 0;
    if (pos == 0 && J < N) return psn_fail(PSN_ERR_UNSUPPORTED, "a state with pos == 0 needs blocks of >= 312 doubles");
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = call_begin(s))) return rc;
    const MemKind k_out = classify(out), k_sk = classify(skipped);
    const long long nstates = n_blocks + 1;
    int rounds = 0;
    for (long long have = 1; have < nstates - first; have *= 2) ++rounds;
    size_t need = pad256((size_t)nstates * N * 4) + pad256((size_t)(rounds + 1) * CHUNKS * 4) + 8 * 256;
    if (k_out == MEM_HOST) need += pad256((size_t)n_blocks * count * 8);
    if (k_sk == MEM_HOST) need += pad256((size_t)n_blocks * skip * 8);
    Arena& ar = g.mt_arena;
    if ((rc = arena_reserve(ar, need))) return rc;
    uint32_t* d_states = arena_take<uint32_t>(ar, (size_t)nstates * N);
    uint32_t* d_polys = arena_take<uint32_t>(ar, (size_t)(rounds + 1) * CHUNKS);
    double* d_out = (k_out == MEM_HOST) ? arena_take<double>(ar, (size_t)n_blocks * count) : out;
    double* d_sk = (k_sk == MEM_HOST) ? arena_take<double>(ar, (size_t)n_blocks * skip) : skipped;
    if (!d_states || !d_polys || (k_out == MEM_HOST && !d_out) || (k_sk == MEM_HOST && !d_sk)) return psn_fail(PSN_ERR_NOMEM, "arena exhausted");
    CU_TRY(cudaMemsetAsync(d_states, 0, (size_t)nstates * N * 4, s));   // the jump kernel accumulates with atomicXor
    CU_TRY(cudaMemcpyAsync(d_states, key, N * 4, cudaMemcpyHostToDevice, s));
    {
        static bool attr_done = false;
        if (!attr_done) {
            CU_TRY(cudaFuncSetAttribute(psn_mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psn_mt_jump_smem()));
            attr_done = true;
        }
    }
    // polynomials: [0 .. rounds) = t^(J 2^k); [rounds] = t^(J - 624) (pos == 0 only)
    std::vector<uint32_t> h_polys((size_t)(rounds + 1) * CHUNKS, 0u);
    for (int k = 0; k < rounds; ++k) to_chunks(jump_poly((uint64_t)J, k), h_polys.data() + (size_t)k * CHUNKS);
    if (first) to_chunks(jump_poly((uint64_t)(J - N), 0), h_polys.data() + (size_t)rounds * CHUNKS);
    CU_TRY(cudaMemcpyAsync(d_polys, h_polys.data(), h_polys.size() * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaStreamSynchronize(s));   // h_polys and key are host temporaries
    if (first) {
        ProfScope pr("mt_jump", s);
        psn_mt_jump_kernel<<<dim3(1, PSN_MT_SPLIT), PSN_MT_THREADS, psn_mt_jump_smem(), s>>>(d_states, 0, 1, d_polys + (size_t)rounds * CHUNKS);
    }
    {
        int k = 0;
        for (long long have = 1; have < nstates - first; have *= 2, ++k) {
            const long long cnt = std::min<long long>(have, nstates - first - have);
            ProfScope pr("mt_jump", s);
            psn_mt_jump_kernel<<<dim3((unsigned)cnt, PSN_MT_SPLIT), PSN_MT_THREADS, psn_mt_jump_smem(), s>>>(d_states, first, first + have,
                                                                                                           d_polys + (size_t)k * CHUNKS);
        }
    }
    CU_TRY(cudaGetLastError());
    {
        ProfScope pr("mt_uniform", s);
        if (first) {
            psn_mt_uniform_kernel<<<1, PSN_MT_GEN_THREADS, 0, s>>>(d_states, 0, skip, count, low, high - low, d_out, d_sk);
            if (n_blocks > 1)
                psn_mt_uniform_kernel<<<(unsigned)(n_blocks - 1), PSN_MT_GEN_THREADS, 0, s>>>(d_states + N, N, skip, count, low, high - low,
                                                                                           d_out + count, d_sk ? d_sk + skip : nullptr);
        } else {
            psn_mt_uniform_kernel<<<(unsigned)n_blocks, PSN_MT_GEN_THREADS, 0, s>>>(d_states, pos, skip, count, low, high - low, d_out, d_sk);
        }
    }
    CU_TRY(cudaGetLastError());
    if (k_out == MEM_HOST) CU_TRY(cudaMemcpyAsync(out, d_out, (size_t)n_blocks * count * 8, cudaMemcpyDeviceToHost, s));
    if (k_sk == MEM_HOST) CU_TRY(cudaMemcpyAsync(skipped, d_sk, (size_t)n_blocks * skip * 8, cudaMemcpyDeviceToHost, s));
    if (key_out) CU_TRY(cudaMemcpyAsync(key_out, d_states + (size_t)n_blocks * N, N * 4, cudaMemcpyDeviceToHost, s));
    if (pos_out) *pos_out = first ? N : pos;
    if (k_out == MEM_HOST || k_sk == MEM_HOST || key_out) CU_TRY(cudaStreamSynchronize(s));
    return call_end(s);
}

int psn_set_option(const char* name, int64_t value) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_err[0] = 0;
    if (!name) return psn_fail(PSN_ERR_INVALID, "option name is NULL");
    if (!strcmp(name, "nu")) {
        if (value < 1 || value > 3) return psn_fail(PSN_ERR_INVALID, "nu must be in 1..3");
        g.nu = (int)value;
    } else if (!strcmp(name, "nu_coarse")) {
        if (value < 0 || value > 3) return psn_fail(PSN_ERR_INVALID, "nu_coarse must be 0 (as level 0) or in 1..3");
        g.nu_coarse = (int)value;
    } else if (!strcmp(name, "coarsest_max")) {
        if (value != 0 && (value < 2 || value > PSN_COARSEST_MAX)) return psn_fail(PSN_ERR_INVALID, "coarsest_max must be 0 (automatic) or in 2..%d", PSN_COARSEST_MAX);
        g.coarsest_max = (int)value;
    } else if (!strcmp(name, "tile")) {
        if (value < 8 || value > 96 || (value & 1)) return psn_fail(PSN_ERR_INVALID, "tile must be even, in 8..96");
        g.tile = (int)value;
    } else if (!strcmp(name, "ws_bytes")) {
        if (value < (1 << 20)) return psn_fail(PSN_ERR_INVALID, "ws_bytes must be >= 1 MiB");
        g.ws_bytes = value;
    } else if (!strcmp(name, "smem_max_n")) {
        g.smem_max_n = (int)value;
    } else if (!strcmp(name, "coarse_block")) {
        g.coarse_block = value ? 1 : 0;
    } else if (!strcmp(name, "coarse_max_n")) {
        if (value < 0) return psn_fail(PSN_ERR_INVALID, "coarse_max_n must be >= 0");
        g.coarse_max_n = (int)value;
    } else if (!strcmp(name, "serialize_solves")) {
        g.serialize_solves = value ? 1 : 0;
    } else if (!strcmp(name, "res_block")) {
        g.res_block = value ? 1 : 0;
    } else if (!strcmp(name, "res_persist")) {
        g.res_persist = value ? 1 : 0;
    } else if (!strcmp(name, "inv_smem")) {
        g.inv_smem = value ? 1 : 0;
    } else if (!strcmp(name, "res_threads")) {
        if (value != 0 && value != 128 && value != 256 && value != 512 && value != 1024) return psn_fail(PSN_ERR_INVALID, "res_threads must be 0, 128, 256, 512 or 1024");
        g.res_threads = (int)value;
    } else if (!strcmp(name, "fuse_upd")) {
        if (value < 0 || value > 2) return psn_fail(PSN_ERR_INVALID, "fuse_upd must be 0, 1 or 2");
        g.fuse_upd = (int)value;
    } else if (!strcmp(name, "x_defer")) {
        if (value < 0 || value > 1) return psn_fail(PSN_ERR_INVALID, "x_defer must be 0 or 1");
        g.x_defer = (int)value;
    } else if (!strcmp(name, "p_f32")) {
        if (value < 0 || value > 1) return psn_fail(PSN_ERR_INVALID, "p_f32 must be 0 or 1");
        g.p_f32 = (int)value;
    } else if (!strcmp(name, "upd_strip")) {
        if (value < 0 || value > 1) return psn_fail(PSN_ERR_INVALID, "upd_strip must be 0 or 1");
        g.upd_strip = (int)value;
    } else if (!strcmp(name, "upd_band")) {
        if (value < 0 || value > 65536) return psn_fail(PSN_ERR_INVALID, "upd_band must be 0 (automatic) or >= 16");
        g.upd_band = (int)value;
    } else if (!strcmp(name, "pdl")) {
        if (value < 0 || value > 1) return psn_fail(PSN_ERR_INVALID, "pdl must be 0 or 1");
        g.pdl = (int)value;
    } else if (!strcmp(name, "host_edge_div")) {
        if (value < 1 || value > 64) return psn_fail(PSN_ERR_INVALID, "host_edge_div must be in 1..64");
        g.host_edge_div = (int)value;
    } else if (!strcmp(name, "band_target")) {
        if (value < 1 || value > 1024) return psn_fail(PSN_ERR_INVALID, "band_target (max bands per column) must be in 1..1024");
        g.band_target = (int)value;
    } else if (!strcmp(name, "dir_band")) {
        if (value < 0 || value > 65536) return psn_fail(PSN_ERR_INVALID, "dir_band must be 0 (automatic) or >= 16");
        g.dir_band = (int)value;
    } else if (!strcmp(name, "fuse_band")) {
        if (value < 0 || value > 65536) return psn_fail(PSN_ERR_INVALID, "fuse_band must be 0 (automatic) or >= 16");
        g.fuse_band = (int)value;
    } else if (!strcmp(name, "band0")) {
        if (value < 0 || value > 65536) return psn_fail(PSN_ERR_INVALID, "band0 must be 0 (automatic) or >= 16");
        g.band0 = (int)value;
    } else if (!strcmp(name, "vcycle_f32")) {
        if (value < 0 || value > 1) return psn_fail(PSN_ERR_INVALID, "vcycle_f32 must be 0 or 1");
        g.vcycle_f32 = (int)value;
    } else if (!strcmp(name, "host_slots")) {
        if (value < 3 || value > PSN_NSLOTS) return psn_fail(PSN_ERR_INVALID, "host_slots must be in 3..%d", PSN_NSLOTS);
        g.host_slots = (int)value;
    } else if (!strcmp(name, "host_prefetch")) {
        if (value < 1 || value > PSN_NCHUNK_STREAMS - 1) return psn_fail(PSN_ERR_INVALID, "host_prefetch must be in 1..%d", PSN_NCHUNK_STREAMS - 1);
        g.host_prefetch = (int)value;
    } else if (!strcmp(name, "multires_pipeline")) {
        if (value < 0 || value > 1) return psn_fail(PSN_ERR_INVALID, "multires_pipeline must be 0 or 1");
        g.multires_pipeline = (int)value;
    } else if (!strcmp(name, "copy_streams")) {
        g.copy_streams = value ? 1 : 0;
    } else if (!strcmp(name, "host_h2d_stream")) {
        g.host_h2d_stream = value ? 1 : 0;
    } else if (!strcmp(name, "host_chunks")) {
        if (value < 0 || value > 1024) return psn_fail(PSN_ERR_INVALID, "host_chunks must be in 0..1024");
        g.host_chunks = (int)value;
    } else if (!strcmp(name, "vcycle_mode")) {
        if (value < 0 || value > 2) return psn_fail(PSN_ERR_INVALID, "vcycle_mode must be 0 (tiles), 1 (warp strips, register prefetch) or 2 (warp strips, TMA ring)");
        g.vcycle_mode = (int)value;
    } else {
        return psn_fail(PSN_ERR_INVALID, "unknown option '%s'", name);
    }
    return PSN_OK;
}

int psn_profile_begin(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    int rc = ensure_init();
    if (rc) return rc;
    for (auto& r : g.prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g.prof.clear();
    g.profiling = true;
    return PSN_OK;
}

int psn_profile_read(char* buf, int64_t buf_bytes) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (cudaDeviceSynchronize() != cudaSuccess) return psn_fail(PSN_ERR_CUDA, "device synchronize failed");
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : g.prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        auto& a = agg[r.name];
        a.first += 1;
        a.second += ms;
    }
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
        first = false;
    }
    out += "}";
    if ((int64_t)out.size() + 1 > buf_bytes || !buf) return -(int)(out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return (int)out.size();
}

int psn_profile_end(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g.profiling = false;
    return PSN_OK;
}

int psn_describe(int n, int64_t* info) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_err[0] = 0;
    if (!info || n < 2) return psn_fail(PSN_ERR_INVALID, "bad arguments");
    PsnPlan plan;
    psn_plan_levels(&plan, n, g.nu, g.coarsest_max, g.nu_coarse);
    if (!g.ready) g.smem_optin = 232448;  // sm_100 opt-in limit; lets the CPU-only tier query the layout
    const ResidentCfg rc = resident_cfg(plan);
    memset(info, 0, 8 * sizeof(int64_t));
    info[1] = plan.nlev;
    if (rc.T) {
        info[0] = 1;
        info[2] = 0;
        info[3] = rc.T;
        info[4] = (int64_t)rc.smem;
        info[5] = 0;
        info[6] = 0;
    } else {
        PsnStreamCfg sc;
        memset(&sc, 0, sizeof(sc));
        sc.ls = stream_ls(plan); sc.TH = sc.TW = g.tile; sc.mode = g.vcycle_mode;
        sc.coarse_block = (g.coarse_block && sc.ls >= 1 && psn_block_ok_at<20>(plan, sc.ls)) ? 1 : 0;
        if (sc.ls >= 1 && !psn_strip_ok(plan, sc.ls)) sc.mode = 0;
        sc.vf32 = (sc.mode == 2 && g.vcycle_f32 && sc.ls >= 1 && psn_vf32_ok(plan, sc.ls)) ? 1 : 0;
        info[7] = sc.vf32 | ((sc.mode == 2 && g.upd_strip && (plan.n & 1) == 0) ? 2 : 0);
        if ((info[7] & 3) == 3 && g.p_f32) info[7] |= 4;
        if ((info[7] & 4) && g.x_defer) info[7] |= 8;
        if ((info[7] & 8) && g.fuse_upd) info[7] |= 16;
        sc.upd_strip = g.upd_strip; sc.pf32 = (info[7] & 4) ? 1 : 0; sc.xdefer = (info[7] & 8) ? 1 : 0; sc.fuse = (info[7] & 16) ? 1 : 0;
        info[0] = 0;
        info[2] = sc.ls;
        info[3] = PSN_TILE_THREADS;
        info[4] = (int64_t)psn_tile_post_doubles(sc.TH, sc.TW, plan.nu) * 8;
        info[5] = g.tile;
        info[6] = sc.ls >= 1 ? (int64_t)stream_ws_bytes_per_sample(plan, sc) : 0;
    }
    return PSN_OK;
}

}  // extern "C"
