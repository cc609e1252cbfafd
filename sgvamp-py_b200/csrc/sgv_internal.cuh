// Internal declarations shared by the translation units of libsgvamp_b200.so.
// sm_100a only; fp32 LD values, fp64 vectors / accumulation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include <vector>
#include "../../include/sgvamp_b200.h"

// ---------------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------------
void sgv_set_error(const char* fmt, ...);
#define SGV_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            sgv_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)
#define SGV_CHECK(cond, ...)                                                                   \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            sgv_set_error(__VA_ARGS__);                                                        \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)
#define SGV_TRY(expr)                                                                          \
    do {                                                                                       \
        int rc_ = (expr);                                                                      \
        if (rc_ != 0) return rc_;                                                              \
    } while (0)

#define SGV_MAX_RANKS 8
#define SGV_INBOX_SLOTS 4
#define SGV_MAX_PARTIAL_VALUES 16

// ---------------------------------------------------------------------------------------------
// device-resident CG / reduction state (one per handle, reused by every cohort in turn)
// ---------------------------------------------------------------------------------------------
// device-resident state of the EM prior-learning loop (src/sgvamp.py:250-257): the finaliser of every
// EM pass updates it and raises `done`, so the host can enqueue passes in batches without reading back
struct EmState {
    double lam, omegas[SGV_MAX_L];
    double a[SGV_MAX_K], asum, Mtot, tol, relerr;
    int    K, Lm1, steps, maxit, done;
};

// Device-resident scalar chain of the VAMP loop (src/sgvamp.py:285-374 keeps these in Python floats): the finalisers of
// the denoiser and of the LMMSE post-processing update them, so a whole VAMP iteration is enqueued without a host
// round trip; `IterLog` is what the host reads back once per iteration (CSV row, CG counts, prior, metrics).
struct VampScal {
    double gam1[SGV_MAX_K], gamw[SGV_MAX_K], alpha1[SGV_MAX_K], alpha2[SGV_MAX_K], gam2[SGV_MAX_K];
    double a[SGV_MAX_K], N[SGV_MAX_K];     // cohort weights (src/main.py:287) and sample sizes (gamw update :352-363)
    double sigmas[SGV_MAX_L];              // slab variances (already times Nt, :27)
    int    K, Lm1;
};
struct IterLog {
    double row[SGV_MAX_K][7];              // it, gamw, gam1, gam2, alpha1, alpha2, lam  (:377)
    int    cg_iters[SGV_MAX_K][2], cg_info[SGV_MAX_K][2], passes[SGV_MAX_K];
    double lam, omegas[SGV_MAX_L], em_relerr, dmean, metrics[4];
    int    em_steps, error;
};

struct CgState {
    double rho[2], rho_prev[2], pq[2], bnorm2[2];
    double stats[16];      // results of non-CG reductions (read back by the host)
    int    done[2], iters[2], info[2], zero_b[2];
    double alpha[2];       // fused CG (DSYM): step length of the update x += alpha p, r -= alpha q still to be applied (0: none)
    int    step;           // CG updates performed (0 -> p = r)
    int    ambig[2];       // fused CG: the extrapolated |r|^2 was too close to the threshold to decide; the next pass's exact sum decides
    int    maxit;
    double band_eps;       // relative half-width of the "ambiguous" band around the stopping threshold (~1e3 eps)
    int    error;          // != 0: a cross-rank wait timed out; bit q set = rank q's partial never arrived
    unsigned long long err_seq;   // sequence number of the reduction that timed out
    EmState em;
    VampScal vs;
    IterLog  log;
};

// Cross-rank reduction mailbox.  Lives at the start of every rank's symmetric arena; rank r writes its
// partial sums of exchange `seq` (counted on the device, see red_next_seq) into slot seq % SLOTS, row r, of EVERY rank's inbox with peer stores
// over NVLink.  Every value travels with the sequence number in ONE 16-byte store, laid out as in NCCL's
// low-latency (LL) protocol: each 8-byte half carries 32 bits of the double and its own 32-bit copy of the
// sequence number, so the scheme only relies on 8-byte store atomicity.  A reader that sees the right
// sequence number in both halves has the value: neither side needs a system-scope fence and there is no
// separate flag round trip.  Consumers add the rows in rank order, so all ranks obtain bit-identical totals.
struct __align__(16) InboxEntry {
    unsigned lo, flag_lo, hi, flag_hi;
};
struct Inbox {
    InboxEntry e[SGV_INBOX_SLOTS][SGV_MAX_RANKS][SGV_MAX_PARTIAL_VALUES];
};

// What to do with the totals of a grid-wide (and, for world > 1, cross-rank) reduction.
enum { AP_STATS = 0, AP_PQ = 1, AP_RESID = 2, AP_SETUP = 3, AP_CGUPDATE = 4, AP_EM = 5, AP_CGFUSED = 6,
       AP_DENOISE = 7 /* alpha1, gam2 of every cohort */, AP_POST = 8 /* alpha2, gam1, gamw, CSV row of one cohort */,
       AP_METRICS = 9 };
enum { SKIP_NEVER = 0, SKIP_CG_DONE = 1, SKIP_EM_DONE = 2 };
struct ApplyArgs {
    int kind, nv, off;      // AP_STATS: stats[off + k] = total[k]
    int maxit, x0_zero;     // AP_SETUP
    int is_min;             // combine with min instead of +
    // device-resident scalar chain (AP_DENOISE / AP_POST / AP_METRICS; AP_SETUP reads maxit from here as well)
    int    cohort, it, lmmse_damp, learn_gamw;
    double rho, Mtot;
};

struct RedCtx {
    double*            partials;   // per-block partial sums (this rank)
    unsigned*          counter;    // "last block" ticket
    CgState*           st;
    Inbox*             inbox[SGV_MAX_RANKS];   // every rank's inbox, mapped into this rank's address space
    unsigned long long seq;        // host-side launch ordinal (diagnostics only)
    unsigned long long* pubseq;    // device counter of completed cross-rank exchanges: numbers the inbox slots (sgv_device.cuh)
    int                world, rank;
    int                inline_resolve; // world > 1, one GPU per rank: the finalising block itself waits for the other ranks
    int                skip_if_done;   // SKIP_*: the reducing kernel exits early in this state, and so does the resolve
    ApplyArgs          ap;
};

// Work item of the dense-panel kernel: out_part[slot][i] = sum_{j in [j0,j0+nj)} P[j][i] v[j]
struct PanelItem {
    int64_t off;   // element offset of P[j0][i0] inside the panel store (multiple of 4)
    int     ld;    // leading dimension of the panel (multiple of 4)
    int     i0, ni;   // output rows [i0, i0+ni)  (local marker index)
    int     j0, nj;   // input  rows [j0, j0+nj)
    int     navail;   // readable floats from column i0 to the end of the padded row (multiple of 4)
    int     slot;
};

struct LdMatrix {
    int      layout = 0;
    bool     owned = false;
    int64_t  nnz_stored = 0;   // fp32 values read by one pass
    // DIA
    const float* band = nullptr;   // DIA: 2w+1 diagonals; DSYM: roundup(w+1,4) upper diagonals
    int64_t  w = 0, ldb = 0;
    int64_t  ext = 0;              // DSYM: extension rows stored before the first own row
    // dense panels (DENSE: one block; BLOCKDIAG: one per LD block)
    const float* panels = nullptr;
    PanelItem*   items = nullptr;
    int      n_items = 0, s_cross = 1, panel_rw = 4;
    bool     panel_sym = true;             // false: the store holds R^T of a non-symmetric R (full-panel kernel only)
    bool     rowpart = false;              // DENSE, rows partition: the store is the column panel R[:, row_lo:row_hi) (M x Ml),
                                           // the product reads the vector pair of ALL ranks (gathered into sgv_ctx::vfull)
    struct SymItem* sym_items = nullptr;   // upper-triangle work items of the symmetric panel kernel (spmm_psym.cu)
    int*     rowmeta = nullptr;            // per row: strip, strips of its block, forward slots of its strip
    int      n_sym_items = 0;
    int64_t  nblocks = 0;
    // CSR
    int64_t* indptr = nullptr;
    int32_t* indices = nullptr;
    float*   vals = nullptr;
    int64_t  nnz = 0;
};

struct Cohort {
    LdMatrix ld;
    double *xty = nullptr, *r1 = nullptr, *r2 = nullptr, *xhat2 = nullptr, *sig = nullptr;
    double2* rxs = nullptr;   // (R xhat2, R Sigma2_u), recovered from the CG recursion: A x = b - r  (vamp.cu)
    int      last_cg_iters = 0; // iterations the previous solve needed: size of the first batch enqueued by the next one
    bool     rxs_valid = true;  // false after xhat2 / Sigma2_u were overwritten from outside (sgv_set_vec)
    int8_t* probe = nullptr;
};

struct PriorParams {
    int    K, L;
    double lam;
    double omegas[SGV_MAX_L];   // L-1 used
    double sigmas[SGV_MAX_L];   // L-1 used
    double a[SGV_MAX_K];
    double gam1s[SGV_MAX_K];
};

// One rank's view of a peer's symmetric arena: [Inbox | xx | rr | pp0 | pp1], vectors of the
// peer's local length.
struct PeerView {
    char*    base = nullptr;
    int64_t  Ml = 0;
    bool     ipc = false;      // opened with cudaIpcOpenMemHandle (must be closed)
};

struct sgv_ctx {
    int          device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool         own_stream = false;
    int          sm_count = 148;
    int64_t      M = 0;              // global number of markers
    int64_t      Ml = 0;             // markers owned by this rank (rows [row_lo, row_lo+Ml))
    int64_t      row_lo = 0;
    int          rank = 0, world = 1;
    int64_t      bandwidth_hint = 0; // common half-bandwidth agreed by all ranks (0: detect)
    int          halo = 0;           // banded row partition: SpMM reads w-element halos from the neighbours
    bool         rowpart = false;    // dense row partition (sgv_configure_part halo = 2): every rank holds Ml rows of a dense R
    int8_t*      probe_b = nullptr;  // second probe of sgv_probe_pair (Ml bytes)
    double2*     vfull = nullptr;    // rows partition: the input vector pair of all ranks, gathered before each product (M entries)
    int          K = 0;
    Cohort       coh[SGV_MAX_K];
    double*      r1_all = nullptr;   // K x Ml, cohort k's r1 at r1_all + k*Ml (coh[k].r1 aliases it)
    double*      xhat1 = nullptr;
    double*      truth = nullptr;
    bool         truth_set = false;
    // symmetric arena (CG work vectors shared by all cohorts + the inbox), peer-mapped for world > 1
    char*        arena = nullptr;
    size_t       arena_bytes = 0;
    double2 *bb = nullptr;                            // private
    // inside the arena; r, p, q are double-buffered for the fused CG step (a kernel reads buffer `prev`
    // - also the neighbours' - and writes buffer `cur`).  rr / qq: the buffers the classic path uses.
    double2 *xx = nullptr, *rr2[2] = {nullptr, nullptr}, *pp[2] = {nullptr, nullptr}, *qq2[2] = {nullptr, nullptr};
    double2 *rr = nullptr, *qq = nullptr;
    PeerView     peer[SGV_MAX_RANKS];
    // Ranks that share one GPU (tests on a box with fewer GPUs than ranks): kernels of different ranks
    // are not guaranteed to run concurrently, so no kernel may wait for another rank's kernel.  In this
    // mode every cross-rank reduction is completed by a HOST barrier between the reducing kernel and the
    // resolve kernel (the partial sums are already in the inbox when it starts; it never spins).
    bool         host_barrier = false;
    sgv_ctx*     peer_ctx[SGV_MAX_RANKS] = {};
    std::atomic<unsigned long long> host_seq{0};
    unsigned long long seq = 0;      // reductions issued so far (identical on all ranks)
    unsigned long long* pubseq = nullptr;   // device: cross-rank exchanges completed (zeroed with the arena)
    // fused VAMP iteration (sgv_iteration_*): ring of pinned log slots, probe staging
    static const int NLOG = 4;
    IterLog*     log_host[NLOG] = {};
    cudaEvent_t  log_ev[NLOG] = {};
    int8_t*      probe_pin[NLOG] = {};
    int64_t      probe_pin_bytes = 0;
    bool         vamp_begun = false;
    int          vs_active = -1;    // cohort whose device-resident gamw / gam2 the SpMM kernels read (-1: by-value arguments)
    double       cg_band_eps = 2.0e-13;   // CgState::band_eps (SGV_CG_BAND overrides: tests force the postponed path)
    bool         coop_ok = false;      // device supports cooperative launches (persistent EM loop kernel)
    int          em_loop_blocks_per_sm = 0;
    double*      em_cache = nullptr;   // EM loop kernel: the pass-invariant exponentials (Ml x K x L doubles)
    int64_t      em_cache_cap = 0;
    int          last_em_steps = 0;  // EM passes the previous prior update needed (first batch of the next one)
    double2*     ds_ypart = nullptr; // DSYM kernel: per-row partial sums and per-tile tails
    double2*     ds_tails = nullptr;
    int64_t      ds_ypart_cap = 0, ds_tails_cap = 0;
    // persistent DSYM kernel (spmm_dsymp.cu): head-row partial sums / carry-out per row range, hand-off flags
    double2*     dsp_yhead = nullptr;
    double2*     dsp_tails = nullptr;
    int64_t      dsp_cap = 0;
    unsigned long long* dsp_flags = nullptr;
    unsigned long long  dsp_epoch = 0;
    unsigned long long* dsp_dbg = nullptr;   // SGV_DS_DEBUG=1: phase clock of the whole-solve kernel (printed by sgv_destroy)
    double2*     ypart = nullptr;    // cross-CTA partial outputs of the panel kernel
    int64_t      ypart_cap = 0;      // in double2 elements
    double2*     ypartT = nullptr;   // transposed partial outputs of the symmetric panel kernel, per strip
    int64_t      ypartT_cap = 0;
    PriorParams  prior{};
    // reductions
    double*      partials = nullptr;   // per-block partial sums
    int64_t      partials_cap = 0;     // in doubles
    unsigned*    counter = nullptr;    // ticket for "last block finalises"
    CgState*     cg = nullptr;         // device
    CgState*     cg_host = nullptr;    // pinned mirror
    double*      host_scal = nullptr;  // pinned scratch (64 doubles)
    void*        stage = nullptr;      // device staging buffer for uploads
    int64_t      stage_bytes = 0;
    cudaEvent_t  ev_a = nullptr, ev_b = nullptr, ev_copy = nullptr;
    // ring of device snapshot buffers for asynchronous vector read-back
    static const int NSNAP = 8;
    double*      snap[NSNAP] = {};
    cudaEvent_t  snap_ev[NSNAP] = {};
    int          snap_next = 0;
    int64_t      launches = 0;
    bool         prof = false;
    std::vector<cudaEvent_t> prof_ev;   // start/stop pairs
    size_t       prof_n = 0;             // events used
};

int sgv_reset_cg_state(sgv_ctx* c);   // api.cu: zero the device state, keep the constants

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// arena layout for a rank with Ml local markers
static inline size_t arena_vec_bytes(int64_t Ml) { return (size_t)round_up(Ml, 64) * sizeof(double2); }
static inline size_t arena_off_xx(int64_t) { return (size_t)round_up((int64_t)sizeof(Inbox), 4096); }
static inline size_t arena_off_rr(int64_t Ml, int i) { return arena_off_xx(Ml) + (size_t)(1 + i) * arena_vec_bytes(Ml); }
static inline size_t arena_off_pp(int64_t Ml, int i) { return arena_off_xx(Ml) + (size_t)(3 + i) * arena_vec_bytes(Ml); }
static inline size_t arena_off_qq(int64_t Ml, int i) { return arena_off_xx(Ml) + (size_t)(5 + i) * arena_vec_bytes(Ml); }
static inline size_t arena_size(int64_t Ml) { return arena_off_qq(Ml, 2); }

// epilogues of the SpMM kernels
enum { EPI_Q = 0, EPI_RESID = 1, EPI_STATS = 2, EPI_PLAIN = 3, EPI_CG = 4 /* fused single-reduction CG step (DSYM) */ };
// which symmetric vector an SpMM reads (so that the halos can be fetched from the neighbours)
enum { VEC_XX = 0, VEC_PP0 = 1, VEC_PP1 = 2 };

struct SpmmArgs {
    // input vector pair in local coordinates (index 0 = first owned marker); halos come from
    // v_left (the left neighbour's last entries) / v_right (the right neighbour's first
    // entries); null = outside the matrix (zeros)
    const double2* v;
    const double2* v_left;
    const double2* v_right;
    int64_t        n_left;     // local length of the left neighbour
    // fused CG direction update (DIA layout): v := r + beta * p_old on the fly, written to p_new
    int            fused_p;
    const double2 *r, *r_left, *r_right;
    double2*       p_new;
    double2*       out;    // EPI_Q: qq ; EPI_RESID: rr ; EPI_PLAIN: y
    const double2* bb;     // EPI_RESID: right-hand sides ; EPI_STATS: col1 = probe u
    double         gamw, gam2;
    int64_t        M;      // local number of markers
    int            check_done;   // 1: exit immediately when both CG columns are done
    int            vs_cohort;    // >= 0: gamw / gam2 are read from the device-resident scalar chain (CgState::vs) of this cohort
    // fused CG step (EPI_CG): r_new = r - alpha q (pending update), p_new = r_new + beta p, x += alpha p are formed
    // while the window is staged; v = p_old, r = r_old, q = q_old (each with the neighbours' halos)
    const double2 *q, *q_left, *q_right;
    double2 *r_new, *x;
    RedCtx         rc;
};

// spmm.cu
int sgv_launch_spmm(sgv_ctx* c, Cohort& co, int epi, int vec, double2* out, double gamw, double gam2, int check_done,
                    int fused_p);
size_t sgv_dia_smem_bytes(int64_t w, int rw, int s);
int    sgv_preload_spmm();   // load all kernels of the TU on the current device (see spmm.cu)
int    sgv_preload_vamp();
// spmm_psym.cu
int    sgv_preload_psym();
int    sgv_build_psym_items(sgv_ctx* c, LdMatrix& ld, const std::vector<int64_t>& starts, const std::vector<int64_t>& offs,
                            const std::vector<int>& lds);
int    sgv_launch_psym(sgv_ctx* c, const LdMatrix& ld, int epi, SpmmArgs& a);
// spmm_dsym.cu
int    sgv_preload_dsym();
size_t sgv_dsym_smem_bytes(int64_t w, int rw, int s, int nst);
bool   sgv_dsym_feasible(int64_t w);
int    sgv_dsym_ensure_scratch(sgv_ctx* c, const LdMatrix& ld);
int    sgv_launch_dsym(sgv_ctx* c, const LdMatrix& ld, int epi, SpmmArgs& a);
int    sgv_launch_dsym_cg(sgv_ctx* c, Cohort& co, int n, double gamw, double gam2);
// spmm_dsymp.cu (persistent form of the same product: one kernel per pass)
int    sgv_preload_dsymp();
bool   sgv_dsymp_feasible(int64_t w);
size_t sgv_dsymp_smem_bytes(int64_t w, int rw, int s, int nst);
int    sgv_dsymp_ensure_scratch(sgv_ctx* c, const LdMatrix& ld);
int    sgv_launch_dsymp(sgv_ctx* c, const LdMatrix& ld, int epi, SpmmArgs& a);
bool   sgv_dsymp_solve_usable(const sgv_ctx* c, const LdMatrix& ld);
int    sgv_launch_dsymp_solve(sgv_ctx* c, const LdMatrix& ld, SpmmArgs& a, int max_steps);
int    sgv_launch_dsym_solve(sgv_ctx* c, Cohort& co, double gamw, double gam2, int max_steps);   // spmm_dsym.cu: whole CG solve   // CG step n (reads buffers (n+1)&1, writes n&1)
// element offset of diagonal d (0..Dp-1) at storage row j in the tiled DSYM layout (ngr = Dp/4)
__host__ __device__ static inline int64_t sgv_dsym_index(int64_t j, int64_t d, int64_t ngr) {
    return ((j >> 7) * ngr + (d >> 2)) * 512 + (d & 3) * 128 + (j & 127);
}
static inline int64_t sgv_dsym_ext(const sgv_ctx* c, int64_t w) { return (c->halo && c->rank > 0) ? round_up(w, 256) : 0; }
bool   sgv_dia_feasible(int64_t w);
// ld_formats.cu
void sgv_ld_free(LdMatrix& ld);
int  sgv_build_panel_items(sgv_ctx* c, LdMatrix& ld, const std::vector<int64_t>& starts,
                           const std::vector<int64_t>& offs, const std::vector<int>& lds);
int  sgv_ensure_stage(sgv_ctx* c, int64_t bytes);
int  sgv_ensure_partials(sgv_ctx* c, int64_t nblocks);
// api.cu: reduction plumbing
RedCtx sgv_red_begin(sgv_ctx* c, int kind, int nv, int off = 0, int maxit = 0, int x0_zero = 0, int is_min = 0);
int    sgv_red_end(sgv_ctx* c, const RedCtx& rc);   // launches the cross-rank resolve kernel when world > 1
