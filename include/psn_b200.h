/* psn_b200.h -- C ABI of the B200-native batched Poisson solver (libpsn_b200.so).
 *
 * Drop-in boundary for ONE hot path of tahmidawal/Superresolution_for_PDEs: the per-sample
 * sparse assemble-and-solve that manufactures every training / evaluation sample.  The reference
 * has no FFI of its own (it is pure Python); each entry point below names the reference
 * interface it replaces (paths relative to the reference repository root).  The Python shims in
 * superresolution_for_pdes_b200/ bind these symbols with ctypes and keep the reference's class
 * and method signatures.
 *
 * Conventions
 *  - all arrays are float64 (int32 where stated), C-contiguous, row index = y, column = x;
 *  - every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer on the
 *    library's device; the library detects which (cudaPointerGetAttributes) and stages host
 *    data through its own device buffers in overlapped chunks; DEVICE arrays must be 16-byte aligned (they are
 *    accessed in place with 128-bit loads / stores and bulk copies; PSN_ERR_INVALID otherwise), host arrays
 *    need only their natural alignment;
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  When every
 *    data pointer is a device pointer the call only ENQUEUES work on `stream` and returns
 *    without synchronising, except for small D2H reads of convergence counters on the
 *    large-grid path; with host pointers the call returns after the results are in place;
 *  - return value: 0 on success, negative PSN_ERR_* otherwise (psn_last_error() has the text).
 *    No exceptions cross the ABI.  There is no CPU fallback: without a CUDA device every
 *    compute entry point fails with PSN_ERR_CUDA.
 *
 * Operator (reference: src/data_generation.py:35-58 `_create_laplacian`, :79-104 `solve_poisson`;
 * src/enhanced_data_generation.py:47-68 `solve_poisson_superfine`):
 *     diag(theta) * L * u = f,   L = 5-point Laplacian / h^2,  h = 1/(n-1),
 *     all n*n nodes unknown, out-of-array neighbours are zero.
 * Solved matrix-free in fp64 by multigrid-preconditioned CG to a relative residual `rtol`
 * (true residual recomputed at exit).
 *
 * status[] values per sample:
 *   0 PSN_ST_CONVERGED   true relative residual <= rtol
 *   1 PSN_ST_FP64_FLOOR  residual stagnated at the fp64 round-off floor above rtol but below 1e-10 (as accurate
 *                        as a backward-stable direct solve; happens for very smooth f on large grids)
 *   2 PSN_ST_MAXITER     max_iter reached without convergence
 *   3 PSN_ST_BREAKDOWN   non-finite / non-SPD data (e.g. theta == 0, NaN input)
 *   4 PSN_ST_STAGNATED   a restart from the true residual did not reduce it 4x and it is above 1e-10:
 *                        not converged (callers must treat 2, 3 and 4 as failures)
 */
#ifndef PSN_B200_H
#define PSN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSN_OK 0
#define PSN_ERR_INVALID (-1)     /* bad argument (shape, NULL output, unsupported size) */
#define PSN_ERR_CUDA (-2)        /* CUDA runtime error, or no CUDA device */
#define PSN_ERR_NOMEM (-3)       /* device allocation failed */
#define PSN_ERR_UNSUPPORTED (-4) /* valid request this build does not implement */

#define PSN_OPERATOR_REFERENCE 0 /* diag(theta) L u = f -- what the reference computes (default) */

/* Library version (major*10000 + minor*100 + patch). */
int psn_version(void);

/* Number of CUDA devices visible to the process (0 if none / no driver). */
int psn_get_device_count(void);

/* Bind the library to `device` (-1: keep the current device), create its streams and caches.
 * Replaces nothing in the reference (which has no device); called once per process (one process
 * per GPU).  Idempotent for the same device. */
int psn_init(int device);

/* Release every device buffer, stream and cached plan. */
void psn_shutdown(void);

/* Text of the last error on this thread's most recent failing call ("" if none). */
const char* psn_last_error(void);

/* Solve `batch` independent n x n problems.
 *
 * Replaces: PoissonSolver.solve_poisson (src/data_generation.py:79-104) and
 * EnhancedPoissonSolver.solve_poisson_superfine (src/enhanced_data_generation.py:47-68), batched;
 * with sin_x/sin_y it also replaces generate_forcing_term (src/data_generation.py:60-77).
 *
 *   theta  [batch, master_n, master_n] or NULL (theta == 1, as in every dataset path)
 *   forcing: exactly one of
 *     f      [batch, master_n, master_n]
 *     sin_x, sin_y  [batch, master_n] each: f[i][j] = sin_y[i] * sin_x[j]  (host-evaluated sines
 *                   => f is bit-identical to the reference's numpy expression)
 *     k1, k2 [batch]: f = sin(2 pi k1 x_j) sin(2 pi k2 y_i) evaluated on the device
 *                   (synthetic-throughput mode; not bit-identical to numpy's sin)
 *   master_n, step: grid node (i, j) reads master node (i*step, j*step); master_n >= (n-1)*step+1.
 *                   (master_n = n, step = 1 for plain inputs; step > 1 reproduces the
 *                   `field[::step, ::step]` gathers of src/resolution_comparison.py:56-58)
 *   operator_mode  PSN_OPERATOR_REFERENCE
 *   rtol, max_iter relative residual target (e.g. 1e-12) and iteration cap (e.g. 100)
 *   u      [batch, n, n] out
 *   iters  [batch] int32 out (may be NULL), relres [batch] out (may be NULL),
 *   status [batch] int32 out (may be NULL)
 */
int psn_solve_batch(int n, int64_t batch,
                    const double* theta, const double* f,
                    const double* sin_x, const double* sin_y,
                    const double* k1, const double* k2,
                    int master_n, int step,
                    int operator_mode, double rtol, int max_iter,
                    double* u, int32_t* iters, double* relres, int32_t* status,
                    void* stream);

/* Fused super-fine solve + window extract + stride-2 downsample (theta == 1).
 *
 * Replaces the body of EnhancedPoissonSolver.generate_subdomain_dataset's loop
 * (src/enhanced_data_generation.py:121-163): solve_poisson_superfine (:133), extract_subdomain
 * (:141-143 -> :83, rows = y, cols = x) and downsample (:146-148 -> :96).
 *
 *   sin_x, sin_y [batch, n_super];  start_x, start_y [batch] int32 (window origin, column / row)
 *   outputs (each may be NULL): u_fine, f_fine, theta_fine [batch, n_fine, n_fine];
 *                               u_coarse, f_coarse, theta_coarse [batch, ceil(n_fine/2), ceil(n_fine/2)]
 */
int psn_subdomain_batch(int n_super, int n_fine, int64_t batch,
                        const double* sin_x, const double* sin_y,
                        const int32_t* start_x, const int32_t* start_y,
                        double rtol, int max_iter,
                        double* u_fine, double* f_fine, double* theta_fine,
                        double* u_coarse, double* f_coarse, double* theta_coarse,
                        int32_t* iters, double* relres, int32_t* status,
                        void* stream);

/* Multi-resolution reference solutions from one master field per example.
 *
 * Replaces the solve loop of solve_multi_resolution (src/resolution_comparison.py:50-73; timed
 * variant src/resolution_comparison_statistical.py:59-90): for each resolution r the inputs are
 * the stride gathers master[::n_finest/r, ::n_finest/r].
 *
 *   resolutions [n_res] int32 (each must divide n_finest);  theta_finest [batch, n_finest, n_finest]
 *   or NULL;  sin_x, sin_y [batch, n_finest];  u_out[k] -> [batch, r_k, r_k];
 *   iters / relres / status: [n_res, batch] (may be NULL)
 * Host master fields are uploaded once per chunk of examples.  Host outputs: every resolution is solved into device
 * buffers, largest grid first, and copied back on the result-copy stream while the next one is solved (option
 * "multires_pipeline", default 1; 0 = one blocking host-output solve per resolution; same bits either way).
 */
int psn_multires(int n_finest, const int32_t* resolutions, int n_res, int64_t batch,
                 const double* theta_finest, const double* sin_x, const double* sin_y,
                 double rtol, int max_iter,
                 double* const* u_out, int32_t* iters, double* relres, int32_t* status,
                 void* stream);

/* Training-tensor prologue of the path's immediate consumer ("next" row f1 of SURVEY.md section 8).
 *
 * Replaces PDEDataset.__init__ / __getitem__ (src/models.py:132-203): float32 cast, global mean and unbiased
 * std of u_fine / f_fine / theta_fine, z-score (theta is passed through unchanged when its std < 1e-6,
 * src/models.py:162-179), bilinear upsample of the normalised coarse solution n_coarse -> n_fine with
 * align_corners = True (src/models.py:182-187; the reference hard-codes 40x40) and the channel concat
 * [u_coarse_up, theta, f] (src/models.py:192-203).
 *
 *   u_coarse [n, n_coarse, n_coarse], u_fine / f_fine / theta_fine [n, n_fine, n_fine]  float64, host or device
 *   x [n, 3, n_fine, n_fine] float32 out, y [n, 1, n_fine, n_fine] float32 out          host or device
 *   stats [8] float64 out (host or device, may be NULL):
 *         u_mean, u_std, f_mean, f_std, theta_mean, theta_std, theta_is_constant (0/1), 0
 */
int psn_dataset_tensors(int64_t n_samples, int n_coarse, int n_fine,
                        const double* u_coarse, const double* u_fine, const double* f_fine, const double* theta_fine,
                        float* x, float* y, double* stats, void* stream);

/* Interpolation baselines of the evaluation scripts ("next" row f4 of SURVEY.md section 8).
 *
 * Replaces bilinear_multi_level_upscale / cubic_multi_level_upscale
 * (src/resolution_comparison_enhanced.py:19-65): `levels` successive x2 upsamplings with
 * F.interpolate(mode = 'bilinear' | 'bicubic', align_corners = True) in float32 (the reference casts the
 * float64 solution to float32 first and every level rounds to float32).
 *
 *   mode 0 = bilinear, 1 = bicubic (cubic convolution, A = -0.75, border-clamped taps: torch's definition)
 *   u_in [batch, n_in, n_in] float64;  out [batch, n_in * 2^levels, n_in * 2^levels] float32;  host or device
 */
int psn_upscale_cascade(int mode, int n_in, int levels, int64_t batch, const double* u_in, float* out, void* stream);

/* The theta fields of the evaluation-script callers, produced on the device bit-identically to numpy's legacy global
 * RNG (MT19937).
 *
 * Replaces the host draws  k1 = np.random.uniform(*k_range); k2 = np.random.uniform(*k_range);
 * theta = np.random.uniform(0.5, 2.0, (n, n))  of solve_multi_resolution (src/resolution_comparison.py:27-28,42), of
 * solve_large_scale / solve_multi_level (src/subdomain_upscaling.py:19-20,31 and :284-285,299) and the theta block of
 * generate_test_data (src/compare_test_cases.py:39), for `n_blocks` consecutive examples at once.
 *
 *   key [624] uint32, pos in 0..624: the generator state, np.random.get_state()[1:3]  (host)
 *   the stream is consumed as n_blocks x (skip + count) doubles: per block first `skip` raw doubles d in [0, 1)
 *   (returned in `skipped` [n_blocks, skip], may be NULL: the caller forms low + (high - low) * d for the wave
 *   numbers exactly like numpy), then `count` values low + (high - low) * d written to out [n_blocks, count];
 *   out / skipped: host or device.  key_out [624] / pos_out (host, may be NULL): the state after the call, to be
 *   handed back with np.random.set_state.  A state with pos == 0 (only reachable through set_state) needs
 *   skip + count >= 312.  The values are those of numpy's x86-64 builds (multiply, then add: no FMA contraction).
 */
int psn_mt19937_uniform_blocks(const uint32_t* key, int pos, int64_t n_blocks, int64_t skip, int64_t count,
                               double low, double high, double* out, double* skipped,
                               uint32_t* key_out, int* pos_out, void* stream);

/* Host-only companion (no device work): the generator state after `n_words` more 32-bit outputs, by the same
 * jump-ahead polynomials the device path uses (one double = 2 words).  Used by the tests and to skip draws. */
int psn_mt19937_jump_host(const uint32_t* key, int pos, int64_t n_words, uint32_t* key_out, int* pos_out);

/* Tuning knobs (all optional).  name: "nu" (RB-GS sweeps per leg on level 0, default 2), "nu_coarse" (sweeps per leg on the
 * coarser levels, default 3; 0 = as on level 0), "coarsest_max" (0 = automatic: 10 for grids
 * of 32 x 32 and more, 6 below; else 2..10: largest level n x n that is solved with the precomputed dense inverse instead
 * of being smoothed), "tile" (interior
 * tile edge of the streamed kernels, default 64), "ws_bytes" (workspace budget per pipeline slot),
 * "smem_max_n" (largest n solved fully in shared memory; 0 = as large as fits), "vcycle_mode" (streamed V-cycle
 * legs: 0 shared-memory tiles, 1 register-fed warp strips, 2 TMA-fed warp strips = default), "vcycle_f32"
 * (default 1: with vcycle_mode 2 and grid sizes divisible by 4 on every streamed level, the preconditioner's
 * intermediate arrays are stored as float; x, r, p, q, every inner product and all arithmetic stay fp64, so the
 * accuracy of the solution is unchanged), "upd_strip" (default 1: with vcycle_mode 2 the CG update kernel recomputes
 * q = A p from p instead of reading it back, q is never written to HBM), "p_f32" (default 1: with vcycle_f32 and
 * upd_strip the search direction is stored as a power-of-two-scaled float; q, <p,q>, x and r are formed in fp64 from
 * that same rounded direction, so the recurrences stay exact), "x_defer" (default 1: with p_f32 the last 12 directions
 * stay in a ring and x = sum alpha_k p_k is formed in one fp64 pass when a sample is checked or the ring is full,
 * instead of a read-modify-write of x every iteration), "fuse_upd" (default 2; with x_defer, 1: the CG update of iteration k and the
 * level-0 down-leg of iteration k+1 are one kernel, the residual is double-buffered and not re-read; 2: also the prologue and the
 * first down-leg when the forcing comes as sine factors with theta on the solve grid; 0: separate kernels), "pdl" (default 1: programmatic dependent
 * launch between the kernels of a solve), the host-buffer pipeline of resident grids with page-locked buffers:
 * "host_slots" (default 5, 3..6: staging slots),
 * "host_prefetch" (default 3, 1..3: uploads are enqueued this many chunks ahead of the solve being enqueued),
 * "host_h2d_stream" (default 1: uploads of all chunks go through one dedicated stream, in chunk order),
 * "coarse_block" (default 1: an 80 x 80 level of a streamed hierarchy and everything below it run in the register-blocked V-cycle kernel),
 * "coarse_max_n", "host_chunks", "host_edge_div", "dir_band", "upd_band",
 * "band_target", "res_block", "serialize_solves" (pipeline / decomposition tuning, see psn_b200.cu). */
int psn_set_option(const char* name, int64_t value);

/* Per-kernel device timing (CUDA events around every launch).  Off by default because the event
 * pairs serialise the stream.  psn_profile_read writes a JSON object
 * {"kernel": {"launches": L, "ms": total}, ...} into buf and returns its length (or the
 * required size, negated, if buf is too small). */
int psn_profile_begin(void);
int psn_profile_read(char* buf, int64_t buf_bytes);
int psn_profile_end(void);

/* Which path a grid size takes and what it costs: fills info[0..7] =
 * {resident(1)/streamed(0), levels, first shared-memory level, threads per CTA,
 *  dynamic shared memory bytes of the dominant kernel, tile edge, workspace bytes per sample,
 *  bit 0: the V-cycle intermediates of this size are stored as float (option "vcycle_f32", streamed path),
 *  bit 1: the CG update recomputes q = A p in strips instead of storing it (option "upd_strip"),
 *  bit 2: the search direction p is stored as scaled float (option "p_f32"; x, r, q and every dot stay fp64),
 *  bit 3: x is accumulated from the ring of stored directions when a sample is checked (option "x_defer"),
 *  bit 4: the CG update is fused with the level-0 down-leg of the next V-cycle (option "fuse_upd")}. */
int psn_describe(int n, int64_t* info);

#ifdef __cplusplus
}
#endif
#endif /* PSN_B200_H */
