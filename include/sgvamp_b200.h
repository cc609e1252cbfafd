/*
 * sgvamp_b200.h -- C ABI of libsgvamp_b200.so, the B200 (sm_100a) implementation of the
 * sgVAMP hot path (VAMP.infer and everything it calls, reference src/sgvamp.py:196-389).
 *
 * The reference has no FFI layer: the path sits behind the Python class `VAMP`
 * (src/sgvamp.py:14).  These entry points are what a ctypes binding for that class binds;
 * `sgvamp-py_b200/sgvamp.py` is that binding and INTEGRATION.md shows the stub a maintainer of
 * the reference would add.  Each function cites the reference lines it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; sgv_last_error() gives the message
 *     (reference: numerical failures are log-only, bad input raises, src/main.py:90-97).
 *   - host pointers are owned by the caller for the duration of the call; all device memory is
 *     owned by the library behind the opaque handle.  One host thread per handle.
 *   - vectors are fp64; LD values are stored fp32 in HBM (arithmetic: fp32 value x fp64 vector,
 *     fp64 accumulation).  There is no CPU fallback: without a CUDA device sgv_create fails.
 *   - a handle owns the marker rows [row_lo,row_hi) of every cohort's LD matrix (the whole
 *     matrix on one GPU).
 */
#ifndef SGVAMP_B200_H
#define SGVAMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgv_ctx* sgv_handle;

#define SGV_MAX_K 8   /* cohorts per handle   */
#define SGV_MAX_L 8   /* mixture components   */

/* LD storage layouts in HBM */
enum { SGV_LAYOUT_AUTO = 0, SGV_LAYOUT_DENSE = 1, SGV_LAYOUT_DIA = 2, SGV_LAYOUT_BLOCKDIAG = 3, SGV_LAYOUT_CSR = 4,
       SGV_LAYOUT_DSYM = 5 /* symmetric half band: diagonals 0..w of the upper triangle only */ };
/* element types of host LD values */
enum { SGV_F32 = 0, SGV_F64 = 1 };
/* vectors readable / writable through sgv_get_vec / sgv_set_vec (per cohort unless noted) */
enum { SGV_VEC_XHAT1 = 0 /* shared */, SGV_VEC_R1 = 1, SGV_VEC_XHAT2 = 2, SGV_VEC_SIGMA2U = 3,
       SGV_VEC_R2 = 4, SGV_VEC_XTY = 5 };

int         sgv_version(void);
const char* sgv_last_error(void);

/* Library/device lifetime.  `stream` is a cudaStream_t to launch on (NULL: the library creates
 * its own non-blocking stream). */
int sgv_create(int device, void* stream, sgv_handle* out);
int sgv_destroy(sgv_handle h);
int sgv_sync(sgv_handle h);

/* Problem shape: M markers, K cohorts (reference VAMP.__init__, src/sgvamp.py:15-21).
 * Allocates all per-cohort state vectors. */
int sgv_configure(sgv_handle h, int64_t M, int K);

/* ---- multi-GPU (no reference counterpart: the reference has one rank per cohort and no data
 * parallelism).  One handle per GPU owns the marker rows [row_lo,row_hi) of every
 * cohort; all vectors are local.  Every grid reduction is completed across ranks inside the
 * kernels: partial sums are written into every rank's inbox through peer memory (NVLink) and
 * added in rank order, so all ranks obtain identical scalars without host collectives.
 * halo == 1: banded LD - the SpMM reads the w halo entries of its input vector directly from the
 * neighbouring ranks' memory.  halo == 0: block-diagonal LD sharded by block (scalar-only exchange).
 * halo == 2: dense LD partitioned by rows (sgv_ld_upload_dense_rows): every product gathers the input vector pair of
 * all ranks from their memory; outputs and all other vectors stay with the owner of the rows.
 * All ranks must issue the same sequence of calls. ---- */
int sgv_configure_part(sgv_handle h, int64_t M, int K, int rank, int world, int64_t row_lo, int64_t row_hi, int halo);
int sgv_ipc_export(sgv_handle h, void* handle64);                                   /* 64-byte CUDA IPC handle of the arena */
int sgv_ipc_import(sgv_handle h, int peer_rank, const void* handle64, int64_t peer_rows);
int sgv_peer_attach_local(sgv_handle h, int peer_rank, sgv_handle other);           /* same-process variant */
/* Ranks that share one GPU (tests on a box with fewer GPUs than ranks, same process only): complete
 * every cross-rank reduction with a host barrier instead of an in-kernel wait, so that no kernel ever
 * waits for another rank's kernel.  Must be set identically on all ranks. */
int sgv_set_host_barrier(sgv_handle h, int enable);
int sgv_device_id(sgv_handle h, char* pci_bus_id, int len);   /* PCI bus id of the handle's GPU */
int sgv_ld_set_bandwidth_hint(sgv_handle h, int64_t w);   /* common half-bandwidth of the DIA layout across ranks */
int sgv_partition_info(sgv_handle h, int64_t* M, int64_t* rows, int64_t* row_lo, int* rank, int* world);
/* Device address of the K x rows block of r1 vectors (cohort k at k * stride doubles).  Rank-per-cohort deployments
 * (src/sgvamp.py:228-233: every rank broadcasts its cohort's r1 and gam1) all-gather straight into it with one
 * collective; the caller orders that collective against the handle's stream. */
int sgv_r1_block(sgv_handle h, void** dev_ptr, int64_t* stride);

/* ---- LD matrices (reference: R argument of VAMP.infer, src/sgvamp.py:196; Rused formed at
 * src/main.py:265).  `s` applies Rused = (1-s) R + s I at upload (pass 0 for an R that is
 * already regularised).  Values are converted to fp32 on the device. ---- */
int sgv_ld_upload_dense(sgv_handle h, int cohort, const void* R, int dtype, int64_t ld, double s);
int sgv_ld_upload_csr(sgv_handle h, int cohort, const int64_t* indptr, const int32_t* indices,
                      const void* data, int dtype, int64_t nnz, double s, int layout_hint);
/* LD in scipy's DIA format (the natural container of banded LD; `scipy.sparse.load_npz` returns it when the
 * matrix was saved so, src/main.py:199-200): data[k*ldd + (j - col0)] = R[j - offsets[k]][j], the entry of diagonal
 * offsets[k] in COLUMN j, for the column window j in [col0, col0+ldd) (col0 = 0, ldd >= M: the whole matrix; a
 * rank of a row partition may pass just the window its rows touch).  No index arrays travel, and for the
 * symmetric half-band layout only the diagonals >= 0 are copied; unless assume_symmetric, the diagonals < 0
 * are streamed through the device once and compared with their mirrors (mismatch: full band / error). */
int sgv_ld_upload_dia(sgv_handle h, int cohort, const void* data, int dtype, int64_t ldd, int64_t col0,
                      const int64_t* offsets, int ndiag, double s, int layout_hint, int assume_symmetric);
/* Adopt LD already resident in HBM (benchmarks: inputs generated on the device).  The library
 * does not take ownership; the buffers must outlive the handle's use of them.
 * dia: band[d*ldb + i] = Rused[i][i+d-w], d in [0,2w]; ldb multiple of 4 elements, base 16B aligned.
 * dense: row-major M x M, ld multiple of 4, base 16B aligned. */
int sgv_ld_adopt_dia(sgv_handle h, int cohort, const float* band_dev, int64_t w, int64_t ldb);
/* dsym: symmetric half band.  Holds Rused[i][i+d] for d in [1,w] and HALF of Rused[i][i] for d = 0 (the
 * kernel applies every stored value twice: to row i and to row i+d) at storage row j = i + ext, with
 * Dp = roundup(w+1,4) diagonals (padding diagonals zero), zero where i+d is outside the matrix, and
 * ldb = a multiple of 128 >= rows + ext storage rows (padding rows zero).  Tiled so that each block of 128
 * rows is one contiguous stream of 2 KB groups of 4 diagonals:
 *     U[ ((j/128)*(Dp/4) + d/4)*512 + (d%4)*128 + j%128 ]
 * `ext` (from sgv_dsym_extension: 0 on a single GPU / rank 0) is the number of rows BEFORE this rank's
 * first row that the buffer also holds (their couplings to this rank's rows; anything else in them zero). */
int sgv_ld_adopt_dsym(sgv_handle h, int cohort, const float* U_dev, int64_t w, int64_t ldb, int64_t ext);
int sgv_dsym_extension(sgv_handle h, int64_t w, int64_t* ext);
int sgv_ld_adopt_dense(sgv_handle h, int cohort, const float* R_dev, int64_t ld);
/* Dense LD partitioned by ROWS over the ranks (sgv_configure_part with halo = 2; no reference counterpart: the
 * reference holds a dense cohort on one rank, src/main.py:255-268).  `rows`: this rank's rows [row_lo, row_hi) of R,
 * row-major with M columns.  The device store is the column panel P[j][i] = Rused[row_lo + i][j] (M rows of
 * roundup(Ml,4) floats), so a product is exactly Rused v also for a non-symmetric R; the vector pair of all ranks is
 * gathered from the peers' memory before each product (the CG direction update fused into the gather).  `adopt`:
 * the column panel already in HBM (fp32, regularised, pad columns readable).  General sparse LD on such a handle:
 * sgv_ld_upload_csr with this rank's rows and GLOBAL column indices (layout hint auto / csr). */
int sgv_ld_upload_dense_rows(sgv_handle h, int cohort, const void* rows, int dtype, int64_t ld_src, double s);
int sgv_ld_adopt_dense_colpanel(sgv_handle h, int cohort, const float* P_dev, int64_t ld);
/* block-diagonal LD (per-chromosome LD blocks, BASELINE.json configs[2]) already in HBM: block b holds the rows
 * [starts[b], starts[b+1]) of this handle (starts[nblocks] = local rows) as a dense row-major panel at
 * panels_dev + offs[b] with leading dimension lds[b] (offs, lds multiples of 4 elements).  Symmetry is verified. */
int sgv_ld_adopt_blockdiag(sgv_handle h, int cohort, const float* panels_dev, int nblocks, const int64_t* starts,
                           const int64_t* offs, const int* lds);
/* On-GPU LD construction (simulation/sim_gen_phen_mult.py:39-55: X column-standardised genotypes / sqrt(N), R = X^T X,
 * r = X^T y; scripts/plink2np.py builds the same kind of matrix from PLINK output): banded R with half-bandwidth w from
 * int8 genotypes in {0,1,2}, marker-major (one row of N samples per marker, as in a .bed file; rows `ldg` bytes apart,
 * ldg a multiple of 16, padding zero).  G holds the markers [g0, g0+nmark), which must cover this handle's rows, the
 * w markers after them and (ranks > 0 of a row partition) the sgv_dsym_extension rows before them; host or device
 * pointer (on_device).  The integer Gram sums are exact; standardisation, the optional Bartlett taper
 * 1-|i-j|/(w+1), Rused = (1-s) R + s I and the fp32 rounding are applied once, and the result is written directly in
 * the symmetric half-band layout (see sgv_ld_adopt_dsym).  y (host, N entries) != NULL: xty_out (host, local rows)
 * receives r = X^T y of the same standardised X. */
int sgv_ld_build_banded(sgv_handle h, int cohort, const int8_t* G, int on_device, int64_t g0, int64_t nmark, int64_t N,
                        int64_t ldg, int64_t w, double s, int taper, const double* y, double* xty_out);
/* copy a cohort's half band (tiled layout of sgv_ld_adopt_dsym, roundup(w+1,4) x ldb floats) into a device buffer of the
 * caller, e.g. to adopt one constructed matrix in several handles; dst_dev == NULL only returns w / ldb / ext */
int sgv_ld_copy_band(sgv_handle h, int cohort, float* dst_dev, int64_t nfloats, int64_t* w, int64_t* ldb, int64_t* ext);
/* layout actually chosen + stored bytes + algorithmic bytes of one SpMM pass at nrhs */
int sgv_ld_info(sgv_handle h, int cohort, int* layout, int64_t* nnz_stored, int64_t* bandwidth,
                int64_t* nblocks, double* bytes_per_pass_nrhs2);

/* XTy vector r of one cohort (src/sgvamp.py:203) and state reset (src/sgvamp.py:199-217:
 * r1 <- r, xhat1 = xhat2 = Sigma2_u_prev = 0). */
int sgv_set_xty(sgv_handle h, int cohort, const double* r);
int sgv_reset_state(sgv_handle h);
int sgv_get_vec(sgv_handle h, int cohort, int which, double* dst);
int sgv_set_vec(sgv_handle h, int cohort, int which, const double* src);
/* asynchronous variant into pinned memory obtained from sgv_pinned_alloc; completes at the next
 * sgv_sync / sgv_wait_copies */
int sgv_get_vec_async(sgv_handle h, int cohort, int which, double scale, double* pinned_dst);
int sgv_wait_copies(sgv_handle h);
int sgv_pinned_alloc(sgv_handle h, int64_t bytes, void** out);
int sgv_pinned_free(sgv_handle h, void* p);

/* ---- prior (src/sgvamp.py:21-28): lam, omegas[L-1], sigmas[L-1] (already times Nt), cohort
 * weights a[K] (src/main.py:287) ---- */
int sgv_set_prior(sgv_handle h, int L, double lam, const double* omegas, const double* sigmas);
int sgv_set_weights(sgv_handle h, const double* a);

/* Denoiser + derivative + damping, one fused kernel over all markers (denoiser_meta /
 * der_denoiser_meta, src/sgvamp.py:93-114, called at :273,:285; damping :275-276).
 * xhat1 <- rho*eta(r1s) + (1-rho)*xhat1 if damp else eta(r1s).
 * *dfac_mean = mean_j d(j) with  d eta/d r_k (j) = a[k]*gam1s[k]*d(j). */
int sgv_denoise(sgv_handle h, const double* gam1s, double rho, int damp, double* dfac_mean);

/* EM prior learning loop (prior_update_em, src/sgvamp.py:116-136, driver loop :250-257).
 * Updates the handle's prior; returns lam, omegas, #steps, final relative error. */
int sgv_prior_em(sgv_handle h, const double* gam1s, int maxit, double tol,
                 double* lam, double* omegas, int* steps, double* relerr);

/* MLE Lagrangian residual (Lagrangian_der, src/sgvamp.py:139-160); the root finder
 * (scipy.optimize.fsolve, src/sgvamp.py:180) stays on the host.  x has L+1 entries. */
int sgv_lagrangian(sgv_handle h, const double* gam1s, const double* x, const double* omega0,
                   const double* sigma2, double* y);

/* LMMSE step of one cohort (src/sgvamp.py:301-374): r2, mu2, the two CG solves of
 * (gamw*R + gam2*I) x = b batched as one 2-RHS system with scipy.sparse.linalg.cg semantics
 * (rtol 1e-5, test at loop top, warm starts), Hutchinson dots and the gamw statistics. */
typedef struct {
    double gamw, gam2, alpha1, rho;   /* in */
    int    cg_maxit, lmmse_damp, learn_gamw, x0_zero;
} sgv_lmmse_in;
typedef struct {
    double u_sigma2u;        /* u^T Sigma2 u            (src/sgvamp.py:338) */
    double xhat2_r;          /* xhat2^T r               (:352) */
    double xhat2_R_xhat2;    /* xhat2^T R xhat2         (:352) */
    double u_R_sigma2u;      /* u^T R Sigma2 u          (:359) */
    int    cg_iters[2];      /* updates performed by each solve */
    int    cg_info[2];       /* 0 converged, cg_maxit exhausted (scipy `info`) */
    int    spmm_passes;      /* SpMM launches that did work */
} sgv_lmmse_out;
int sgv_lmmse(sgv_handle h, int cohort, const sgv_lmmse_in* in, const int8_t* probe, sgv_lmmse_out* out);
/* Additional Hutchinson probes (no reference counterpart: src/sgvamp.py:326-340 draws ONE probe per cohort and iteration;
 * averaging P probes divides the variance of alpha2 by P).  Solves (gamw R + gam2 I) s = u from s = 0 for u = probe_a and
 * probe_b (probe_b may be NULL) with the same CG as sgv_lmmse and returns u.s (the trace term of :338) and u^T R s (the
 * trace term of the gamw update, :359) per probe.  The solver state of the cohort (xhat2, Sigma2_u) is not touched; call
 * it after sgv_lmmse of the same iteration and average with that call's u_sigma2u / u_R_sigma2u. */
typedef struct {
    double u_s[2], u_R_s[2];
    int    cg_iters[2], cg_info[2];
    int    spmm_passes;
} sgv_probe_out;
int sgv_probe_pair(sgv_handle h, int cohort, double gamw, double gam2, int cg_maxit, const int8_t* probe_a,
                   const int8_t* probe_b, sgv_probe_out* out);
/* r1 <- (xhat2 - alpha2*r2)/(1-alpha2)  (src/sgvamp.py:348) */
int sgv_update_r1(sgv_handle h, int cohort, double alpha2);

/* ---- fused VAMP iteration (src/sgvamp.py:222-387 for every cohort of this process).  The scalar chain gam1 ->
 * alpha1 -> gam2 -> alpha2 -> gam1', gamw (:285-374, Python floats in the reference) is kept on the device and advanced
 * by the kernels' finalisers with separately rounded IEEE operations in the reference's order, so one VAMP iteration is
 * enqueued without a host round trip (prior update by EM only; MLE needs scipy's fsolve on the host and uses the
 * stepwise entry points above).  Usage: sgv_set_prior + sgv_set_weights + sgv_set_xty + sgv_reset_state, then
 * sgv_vamp_begin once, then per iteration: fill the probe buffer of a slot (K x local rows, int8, cohort-major),
 * sgv_iteration_enqueue, and - possibly after enqueueing the next iteration on another slot - sgv_iteration_wait. ---- */
#define SGV_ITER_SLOTS 4
typedef struct {
    int    it;              /* iteration index: damping for it > 0 (:275,:290), x0 = 0 warm start for it == 0 */
    int    update_prior;    /* run the EM loop first (:250-257) */
    int    em_maxit;
    double em_tol;          /* the reference uses 1e-6 */
    double rho;
    int    cg_maxit, lmmse_damp, learn_gamw, want_metrics;
} sgv_iter_in;
typedef struct {
    double row[7];          /* it, gamw, gam1, gam2, alpha1, alpha2, lam: the CSV row of :377 */
    int    cg_iters[2], cg_info[2], spmm_passes;
} sgv_iter_cohort;
typedef struct {
    double lam, omegas[SGV_MAX_L], em_relerr, metrics[4];   /* metrics: see sgv_metrics */
    int    em_steps;
    sgv_iter_cohort coh[SGV_MAX_K];
} sgv_iter_out;
int sgv_iteration_supported(sgv_handle h);   /* 1 unless ranks share a GPU (host-barrier mode) or cooperative launch is missing */
int sgv_vamp_begin(sgv_handle h, const double* gam1, const double* gamw, const double* N);   /* K entries each (:210-216, :352) */
int sgv_vamp_set_alphas(sgv_handle h, const double* alpha1, const double* alpha2);   /* resume: previous alpha1 / alpha2 (K each) */
int sgv_set_truth(sgv_handle h, const double* x0);
int sgv_iteration_probe_buffer(sgv_handle h, int slot, int8_t** buf);
/* xhat_pinned / r1_pinned[k]: pinned destinations (sgv_pinned_alloc) of the iteration's xhat1 and incoming r1 dumps, or NULL */
int sgv_iteration_enqueue(sgv_handle h, const sgv_iter_in* in, double* xhat_pinned, double* const* r1_pinned, int slot);
int sgv_iteration_wait(sgv_handle h, int slot, sgv_iter_out* out);

/* metrics vs truth (src/sgvamp.py:379-382): dots[0]=xhat1.x0, [1]=xhat1.xhat1, [2]=x0.x0, [3]=|xhat1-x0|^2 */
int sgv_metrics(sgv_handle h, const double* x0, double* dots);

/* Unit-test / benchmark hook: Y = alpha*(R X) + beta*X for nrhs in {1,2}; X, Y host, column-major M x nrhs */
int sgv_spmm(sgv_handle h, int cohort, const double* X, double* Y, int nrhs, double alpha, double beta);
/* two-phase variant for multi-rank runs (upload; the caller synchronises the ranks; multiply) */
int sgv_spmm_stage(sgv_handle h, const double* X, int nrhs);
int sgv_spmm_run(sgv_handle h, int cohort, double* Y, int nrhs, double alpha, double beta);
/* Same on device-resident interleaved vectors already inside the handle; launches `reps` times
 * and returns the average device time per launch in ms (CUDA events on the handle's stream). */
int sgv_spmm_bench(sgv_handle h, int cohort, int reps, float* ms_per_launch);

/* Per-launch device timing of the SpMM kernels: CUDA event pairs recorded on the handle's stream
 * around every SpMM launch while enabled.  sgv_profile(h,1) resets and enables, (h,0) disables;
 * sgv_profile_read synchronises and returns the summed device time and the number of launches. */
int sgv_profile(sgv_handle h, int enable);
int sgv_profile_read(sgv_handle h, double* total_ms, int64_t* launches);

/* number of kernels launched by this handle since creation */
int64_t sgv_launch_count(sgv_handle h);

#ifdef __cplusplus
}
#endif
#endif
