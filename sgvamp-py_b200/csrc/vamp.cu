// Per-iteration VAMP kernels: denoiser (+derivative, damping), EM / MLE prior reductions,
// LMMSE set-up, the 2-RHS conjugate-gradient driver with scipy semantics, and the post-CG
// Hutchinson / gamw statistics.  Reference: src/sgvamp.py:93-160 and :196-389.
#include <cmath>
#include <cstring>
#include "sgv_device.cuh"

static int check_ready(sgv_ctx* c) {
    SGV_CHECK(c != nullptr, "null handle");
    SGV_CHECK(c->M > 0, "sgv_configure has not been called");
    SGV_CUDA(cudaSetDevice(c->device));   // several handles (one per GPU) may be driven from one process
    return 0;
}

// Read the device-resident reduction / CG state back (one small pinned copy) and surface a
// cross-rank time-out as an error instead of a hang.
static int fetch_state(sgv_ctx* c) {
    SGV_CUDA(cudaMemcpyAsync(c->cg_host, c->cg, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    const CgState* hs = c->cg_host;
    SGV_CHECK(hs->error == 0, "cross-rank reduction %llu timed out on rank %d of %d (missing-rank mask 0x%x)",
              (unsigned long long)hs->err_seq, c->rank, c->world, hs->error);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// denoiser: posterior mean + derivative factor + damping (src/sgvamp.py:93-114, :273-276, :285)
// ---------------------------------------------------------------------------------------------
struct DenoiseConsts {
    int    K, Lm1;
    double w[SGV_MAX_K];
    double s2[SGV_MAX_L], sq[SGV_MAX_L], lw[SGV_MAX_L];   // sigma2_meta, sqrt(s2/sigma), lam*omega
    double one_minus_lam;
};

// the same constants from the device-resident scalar chain (fused VAMP iteration): gam1s from CgState::vs, the prior
// from CgState::em; separately rounded operations in fill_denoise_consts' order
__device__ void dev_denoise_consts(const CgState* st, DenoiseConsts& k) {
    const VampScal& v = st->vs;
    k.K = v.K;
    k.Lm1 = v.Lm1;
    double W = 0.0;
    for (int q = 0; q < v.K; ++q) {
        k.w[q] = __dmul_rn(v.a[q], v.gam1[q]);
        W = __dadd_rn(W, k.w[q]);
    }
    for (int l = 0; l < v.Lm1; ++l) {
        k.s2[l] = __ddiv_rn(1.0, __dadd_rn(W, __ddiv_rn(1.0, v.sigmas[l])));
        k.sq[l] = sqrt(__ddiv_rn(k.s2[l], v.sigmas[l]));
        k.lw[l] = __dmul_rn(st->em.lam, st->em.omegas[l]);
    }
    k.one_minus_lam = __dsub_rn(1.0, st->em.lam);
}

__global__ void __launch_bounds__(256)
k_denoise(int64_t M, const double* __restrict__ r1_all, double* __restrict__ xhat1, DenoiseConsts kval, double rho,
          int damp, RedCtx rc, int dev) {
    __shared__ double red[32];
    __shared__ DenoiseConsts k;
    if (threadIdx.x == 0) {
        if (dev) dev_denoise_consts(rc.st, k);
        else k = kval;
    }
    __syncthreads();
    double dsum[1] = {0.0};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        double sw = 0.0;
        for (int q = 0; q < k.K; ++q) sw += r1_all[(int64_t)q * M + j] * k.w[q];   // np.inner(rs, a*gam1s)  :96
        double mu[SGV_MAX_L];
        int mi = 0;
        double best = -1.0;
        for (int l = 0; l < k.Lm1; ++l) {
            mu[l] = sw * k.s2[l];
            const double score = mu[l] * mu[l] / k.s2[l];                              // :97
            if (l == 0 || score > best) { best = score; mi = l; }
        }
        const double mum = mu[mi], s2m = k.s2[mi];
        double num = 0.0, den = 0.0, dnum = 0.0, dden = 0.0;
        for (int l = 0; l < k.Lm1; ++l) {
            const double e = exp(0.5 * (mu[l] * mu[l] * s2m - mum * mum * k.s2[l]) / (k.s2[l] * s2m));   // :98
            const double t = k.lw[l] * e * k.sq[l];
            num += t * mu[l];                          // :99
            den += t;                                  // :101
            dnum += t * (mu[l] * mu[l] + k.s2[l]);     // :112 (without a_k*gam1_k)
            dden += t * mu[l];                         // :113 (without a_k*gam1_k)
        }
        den += k.one_minus_lam * exp(-0.5 * (mum * mum / s2m));                        // :100-101
        double xh = num / den;
        dsum[0] += (dnum * den - dden * num) / (den * den);                            // :114
        if (damp) xh = rho * xh + (1.0 - rho) * xhat1[j];                              // :276
        xhat1[j] = xh;
    }
    grid_reduce<1>(dsum, rc, red);
}

static void fill_denoise_consts(const sgv_ctx* c, const double* gam1s, DenoiseConsts& k) {
    const PriorParams& p = c->prior;
    k.K = p.K;
    k.Lm1 = p.L - 1;
    double W = 0.0;
    for (int q = 0; q < p.K; ++q) {
        k.w[q] = p.a[q] * gam1s[q];
        W += k.w[q];                                   // sum(self.a * gam1s), left to right
    }
    for (int l = 0; l < k.Lm1; ++l) {
        k.s2[l] = 1.0 / (W + 1.0 / p.sigmas[l]);       // :95
        k.sq[l] = std::sqrt(k.s2[l] / p.sigmas[l]);
        k.lw[l] = p.lam * p.omegas[l];
    }
    k.one_minus_lam = 1.0 - p.lam;
}

extern "C" int sgv_denoise(sgv_handle c, const double* gam1s, double rho, int damp, double* dfac_mean) {
    SGV_TRY(check_ready(c));
    SGV_CHECK(c->prior.L >= 2, "prior not set");
    DenoiseConsts k;
    fill_denoise_consts(c, gam1s, k);
    const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * 8);
    SGV_TRY(sgv_ensure_partials(c, grid));
    RedCtx rc = sgv_red_begin(c, AP_STATS, 1, 0);
    k_denoise<<<grid, 256, 0, c->stream>>>(c->Ml, c->r1_all, c->xhat1, k, rho, damp, rc, 0);
    c->launches++;
    SGV_CUDA(cudaGetLastError());
    SGV_TRY(sgv_red_end(c, rc));
    SGV_TRY(fetch_state(c));
    *dfac_mean = c->cg_host->stats[0] / (double)c->M;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// EM prior update (src/sgvamp.py:116-136): one pass = one grid reduction of K + (L-1) + 1 sums
// ---------------------------------------------------------------------------------------------
struct EmConsts {
    int    K, Lm1;
    double a[SGV_MAX_K];
    double mhg[SGV_MAX_K];                     // -gam_k / 2
    double sqg[SGV_MAX_K];                     // sqrt(gam_k) = 1 / sqrt(1/gam_k)
    double ce[SGV_MAX_K][SGV_MAX_L];           // -1 / (2 (sigma_l + 1/gam_k))
    double isq[SGV_MAX_K][SGV_MAX_L];          // 1 / sqrt(1/gam_k + sigma_l)
};

// One EM pass (src/sgvamp.py:116-136).  lam / omegas come from the device-resident EmState (updated by the
// previous pass's finaliser).  The reference's quotients with loop-invariant denominators are
// multiplications by reciprocals formed once on the host, and pi and xi~ share one division:
//   pi = 1/(1 + t/S) = S/(S+t),   a pi xi_l/S = a xi_l/(S+t)     (S = sum_l xi_l, t = the spike term)
__global__ void __launch_bounds__(256)
k_em(int64_t M, const double* __restrict__ r1_all, EmConsts k, RedCtx rc) {
    const EmState& em = rc.st->em;
    if (em.done) return;
    __shared__ double red[16 * 32];
    double acc[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) acc[t] = 0.0;
    double lo[SGV_MAX_L];
    for (int l = 0; l < k.Lm1; ++l) lo[l] = em.lam * em.omegas[l];      // lam * omega_l
    const double one_minus_lam = 1.0 - em.lam;
    // acc[0..K-1] = sum_j pi_kj ; acc[8..8+Lm1-1] = sum a_k pi xi~_l ; acc[15] = sum a_k pi
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        for (int q = 0; q < k.K; ++q) {
            const double r = r1_all[(int64_t)q * M + j];
            const double r2 = r * r;
            double e[SGV_MAX_L], emax = 0.0;
            for (int l = 0; l < k.Lm1; ++l) {
                e[l] = r2 * k.ce[q][l];                                             // :127
                if (l == 0 || e[l] > emax) emax = e[l];
            }
            double xi[SGV_MAX_L], sum_xi = 0.0;
            for (int l = 0; l < k.Lm1; ++l) {
                xi[l] = lo[l] * exp(e[l] - emax) * k.isq[q][l];                     // :128
                sum_xi += xi[l];
            }
            const double t = one_minus_lam * exp(r2 * k.mhg[q] - emax) * k.sqg[q];  // spike term of :131
            const double inv = 1.0 / (sum_xi + t);
            const double pi = sum_xi * inv;                                          // :131
            const double ainv = k.a[q] * inv;
#pragma unroll
            for (int tt = 0; tt < SGV_MAX_K; ++tt)
                if (tt == q) acc[tt] += pi;
#pragma unroll
            for (int l = 0; l < SGV_MAX_L - 1; ++l)
                if (l < k.Lm1) acc[8 + l] += ainv * xi[l];                          // a pi xi~_l  :130,:136
            acc[15] += k.a[q] * pi;
        }
    }
    grid_reduce<16>(acc, rc, red);
}

__device__ void dev_em_consts(const CgState* st, EmConsts& k) {
    const VampScal& v = st->vs;
    k.K = v.K;
    k.Lm1 = v.Lm1;
    for (int q = 0; q < v.K; ++q) {
        const double ginv = __ddiv_rn(1.0, v.gam1[q]);
        k.a[q] = v.a[q];
        k.mhg[q] = __dmul_rn(-0.5, v.gam1[q]);
        k.sqg[q] = __ddiv_rn(1.0, sqrt(ginv));
        for (int l = 0; l < v.Lm1; ++l) {
            k.ce[q][l] = __ddiv_rn(-0.5, __dadd_rn(v.sigmas[l], ginv));
            k.isq[q][l] = __ddiv_rn(1.0, sqrt(__dadd_rn(ginv, v.sigmas[l])));
        }
    }
}

// start of a prior update inside the fused iteration: loop state reset, lam / omegas stay where the last update left them
__global__ void k_em_begin(CgState* st, int maxit, double tol) {
    EmState& e = st->em;
    e.steps = 0;
    e.maxit = maxit;
    e.tol = tol;
    e.relerr = 0.0;
    e.done = maxit <= 0;
}

// The whole EM loop in ONE persistent (cooperative) kernel: per pass every block sums its markers, a grid
// barrier, block 0 adds the block partials in index order, completes the reduction across ranks (LL inbox)
// and applies the update + convergence test (AP_EM), a second grid barrier, next pass.  No kernel launch, no
// host round trip between passes: on 8 GPUs a pass costs ~10 us instead of ~40.
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire_gpu_u32(bar) < target) {
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
k_em_loop(int64_t M, const double* __restrict__ r1_all, EmConsts kval, RedCtx rc, unsigned* __restrict__ bar, int maxit,
          int dev, double* __restrict__ cache) {
    __shared__ double red[16 * 32];
    __shared__ EmConsts k;
    if (threadIdx.x == 0) {
        if (dev) dev_em_consts(rc.st, k);
        else k = kval;
    }
    __syncthreads();
    // The exponentials of a pass do not depend on (lam, omegas), the only things the passes change: they are evaluated once
    // per prior update and kept in `cache` (L values per marker and cohort; every thread reads back only what it wrote, so
    // no barrier is needed).  The passes multiply the same doubles in the same order as the direct evaluation: bit-identical.
    const int Lc = k.Lm1 + 1;
    if (cache != nullptr && maxit > 0 && !rc.st->em.done) {
        for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
            for (int q = 0; q < k.K; ++q) {
                const double r = r1_all[(int64_t)q * M + j];
                const double r2 = r * r;
                double e[SGV_MAX_L], emax = 0.0;
                for (int l = 0; l < k.Lm1; ++l) {
                    e[l] = r2 * k.ce[q][l];
                    if (l == 0 || e[l] > emax) emax = e[l];
                }
                for (int l = 0; l < k.Lm1; ++l) cache[((int64_t)q * Lc + l) * M + j] = exp(e[l] - emax);
                cache[((int64_t)q * Lc + k.Lm1) * M + j] = exp(r2 * k.mhg[q] - emax);
            }
        }
    }
    const volatile EmState* em = &rc.st->em;
    const unsigned nblk = gridDim.x;
    unsigned epoch = 0;
    for (int pass = 0; pass < maxit; ++pass) {
        if (em->done) break;                       // identical on all blocks: written before the last barrier
        double acc[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) acc[t] = 0.0;
        const double lam = em->lam;
        double lo[SGV_MAX_L];
        for (int l = 0; l < k.Lm1; ++l) lo[l] = lam * em->omegas[l];
        const double one_minus_lam = 1.0 - lam;
        for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
            for (int q = 0; q < k.K; ++q) {
                double xi[SGV_MAX_L], sum_xi = 0.0, es;
                if (cache != nullptr) {
                    for (int l = 0; l < k.Lm1; ++l) {
                        xi[l] = lo[l] * cache[((int64_t)q * Lc + l) * M + j] * k.isq[q][l];
                        sum_xi += xi[l];
                    }
                    es = cache[((int64_t)q * Lc + k.Lm1) * M + j];
                } else {
                    const double r = r1_all[(int64_t)q * M + j];
                    const double r2 = r * r;
                    double e[SGV_MAX_L], emax = 0.0;
                    for (int l = 0; l < k.Lm1; ++l) {
                        e[l] = r2 * k.ce[q][l];
                        if (l == 0 || e[l] > emax) emax = e[l];
                    }
                    for (int l = 0; l < k.Lm1; ++l) {
                        xi[l] = lo[l] * exp(e[l] - emax) * k.isq[q][l];
                        sum_xi += xi[l];
                    }
                    es = exp(r2 * k.mhg[q] - emax);
                }
                const double t = one_minus_lam * es * k.sqg[q];
                const double inv = 1.0 / (sum_xi + t);
                const double pi = sum_xi * inv;
                const double ainv = k.a[q] * inv;
#pragma unroll
                for (int tt = 0; tt < SGV_MAX_K; ++tt)
                    if (tt == q) acc[tt] += pi;
#pragma unroll
                for (int l = 0; l < SGV_MAX_L - 1; ++l)
                    if (l < k.Lm1) acc[8 + l] += ainv * xi[l];
                acc[15] += k.a[q] * pi;
            }
        }
        block_reduce<16>(acc, red);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int t = 0; t < 16; ++t) rc.partials[(size_t)blockIdx.x * 16 + t] = acc[t];
        }
        ++epoch;
        grid_barrier(bar, epoch * nblk);
        if (blockIdx.x == 0) {
            double tot[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) tot[t] = 0.0;
            for (unsigned b = threadIdx.x; b < nblk; b += blockDim.x) {
#pragma unroll
                for (int t = 0; t < 16; ++t) tot[t] += __ldcg(&rc.partials[(size_t)b * 16 + t]);
            }
            block_reduce<16>(tot, red);
            if (threadIdx.x == 0 && rc.world == 1) apply_totals(rc.ap, rc.st, tot);
            if (rc.world > 1 && threadIdx.x < 32) {     // totals are valid in every lane of warp 0
                const unsigned long long seq = red_next_seq(rc);
                publish_warp<16>(tot, rc, threadIdx.x, seq);
                __syncwarp();
                resolve_warp(rc, threadIdx.x, seq);
            }
            __threadfence();
        }
        ++epoch;
        grid_barrier(bar, epoch * nblk);
    }
}

// scratch of the EM loop kernel: L exponentials per marker and cohort; optional (without it the kernel re-evaluates them every pass)
static int ensure_em_cache(sgv_ctx* c) {
    const int64_t need = c->Ml * (int64_t)c->prior.K * (int64_t)c->prior.L;
    if (need <= 0 || c->em_cache_cap >= need) return 0;
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    if (c->em_cache) cudaFree(c->em_cache);
    c->em_cache = nullptr;
    c->em_cache_cap = 0;
    if (cudaMalloc(&c->em_cache, (size_t)need * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        c->em_cache = nullptr;
        return 0;
    }
    c->em_cache_cap = need;
    return 0;
}

extern "C" int sgv_prior_em(sgv_handle c, const double* gam1s, int maxit, double tol, double* lam_out,
                            double* omegas_out, int* steps_out, double* relerr_out) {
    SGV_TRY(check_ready(c));
    PriorParams& p = c->prior;
    SGV_CHECK(p.L >= 2, "prior not set");
    const int Lm1 = p.L - 1;
    const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * 4);
    SGV_TRY(sgv_ensure_partials(c, grid + 1));
    EmConsts k;
    k.K = p.K;
    k.Lm1 = Lm1;
    for (int q = 0; q < p.K; ++q) {
        const double ginv = 1.0 / gam1s[q];
        k.a[q] = p.a[q];
        k.mhg[q] = -0.5 * gam1s[q];
        k.sqg[q] = 1.0 / std::sqrt(ginv);
        for (int l = 0; l < Lm1; ++l) {
            k.ce[q][l] = -0.5 / (p.sigmas[l] + ginv);
            k.isq[q][l] = 1.0 / std::sqrt(ginv + p.sigmas[l]);
        }
    }
    // loop state on the device: passes are enqueued in batches, every pass after convergence exits at once
    EmState* he = &c->cg_host->em;
    memset(he, 0, sizeof(EmState));
    he->lam = p.lam;
    he->asum = 0.0;
    for (int l = 0; l < Lm1; ++l) he->omegas[l] = p.omegas[l];
    for (int q = 0; q < p.K; ++q) {
        he->a[q] = p.a[q];
        he->asum += p.a[q];
    }
    he->Mtot = (double)c->M;
    he->tol = tol;
    he->K = p.K;
    he->Lm1 = Lm1;
    he->maxit = maxit;
    he->done = maxit <= 0;
    SGV_CUDA(cudaMemcpyAsync(&c->cg->em, he, sizeof(EmState), cudaMemcpyHostToDevice, c->stream));
    if (maxit > 0 && c->coop_ok && !(c->world > 1 && c->host_barrier)) {
        // persistent loop kernel (all blocks co-resident: cooperative launch)
        if (c->em_loop_blocks_per_sm == 0) {
            int nb = 0;
            SGV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_em_loop, 256, 0));
            c->em_loop_blocks_per_sm = std::max(1, std::min(nb, 4));
        }
        const unsigned lgrid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * c->em_loop_blocks_per_sm);
        SGV_TRY(sgv_ensure_partials(c, lgrid + 1));
        unsigned* bar = c->counter + 4;            // spare word of the ticket block
        SGV_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), c->stream));
        RedCtx rc = sgv_red_begin(c, AP_EM, 16, 0);   // sequence number of pass 0; pass i uses seq + i
        int64_t Ml = c->Ml;
        const double* r1 = c->r1_all;
        int mi = maxit, dev0 = 0;
        SGV_TRY(ensure_em_cache(c));
        double* cache = c->em_cache;
        void* args[] = {&Ml, &r1, &k, &rc, &bar, &mi, &dev0, &cache};
        SGV_CUDA(cudaLaunchCooperativeKernel((const void*)k_em_loop, dim3(lgrid), dim3(256), args, 0, c->stream));
        c->launches++;
        SGV_TRY(fetch_state(c));
        SGV_CHECK(he->done, "EM loop kernel ended without its done flag");
        maxit = 0;                                  // skip the batched path below
        c->last_em_steps = he->steps;
        p.lam = he->lam;
        for (int l = 0; l < Lm1; ++l) p.omegas[l] = he->omegas[l];
        *lam_out = p.lam;
        for (int l = 0; l < Lm1; ++l) omegas_out[l] = p.omegas[l];
        if (steps_out) *steps_out = he->steps;
        if (relerr_out) *relerr_out = he->relerr;
        return 0;
    }
    // first batch: what the previous update needed plus a margin (the counts drift slowly from one VAMP
    // iteration to the next); a pass enqueued after convergence exits at once
    int launched = 0, batch = c->last_em_steps > 0 ? c->last_em_steps + 3 : 12;
    while (launched < maxit) {
        const int nb = std::min(batch, maxit - launched);
        for (int b = 0; b < nb; ++b) {
            RedCtx rc = sgv_red_begin(c, AP_EM, 16, 0);
            rc.skip_if_done = SKIP_EM_DONE;
            k_em<<<grid, 256, 0, c->stream>>>(c->Ml, c->r1_all, k, rc);
            c->launches++;
            SGV_TRY(sgv_red_end(c, rc));
        }
        SGV_CUDA(cudaGetLastError());
        launched += nb;
        SGV_TRY(fetch_state(c));
        if (he->done) break;
        batch = std::min(48, batch * 2);
    }
    c->last_em_steps = he->steps;
    if (maxit > 0) {
        p.lam = he->lam;
        for (int l = 0; l < Lm1; ++l) p.omegas[l] = he->omegas[l];
    }
    *lam_out = p.lam;
    for (int l = 0; l < Lm1; ++l) omegas_out[l] = p.omegas[l];
    if (steps_out) *steps_out = he->steps;
    if (relerr_out) *relerr_out = he->relerr;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// MLE Lagrangian residual (src/sgvamp.py:139-160)
// ---------------------------------------------------------------------------------------------
struct LagConsts {
    int    K, L;
    double a[SGV_MAX_K];
    double den[SGV_MAX_K][SGV_MAX_L];     // sigma2_l + 1/gam_k
    double sqden[SGV_MAX_K][SGV_MAX_L];
    double omega[SGV_MAX_L];
};

__global__ void __launch_bounds__(256)
k_min_r2(int64_t M, int K, const double* __restrict__ r1_all, RedCtx rc) {
    __shared__ double red[8 * 32];
    double mn[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) mn[q] = __longlong_as_double(0x7ff0000000000000LL);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < K) {
                const double r = r1_all[(int64_t)q * M + j];
                mn[q] = fmin(mn[q], r * r);
            }
    }
    grid_reduce<8, true>(mn, rc, red);
}

__global__ void __launch_bounds__(256)
k_lagrangian(int64_t M, const double* __restrict__ r1_all, LagConsts k, RedCtx rc) {
    const CgState* st = rc.st;
    __shared__ double red[8 * 32];
    // global shift exp_max = max_{k,j,l} -r^2/2/(sigma2_l + 1/gam_k)  (:153); e is monotone in r^2, so
    // the maximum over j is attained at min_j r^2 (st->stats[k], from k_min_r2).
    double exp_max = -__longlong_as_double(0x7ff0000000000000LL);
    for (int q = 0; q < k.K; ++q)
        for (int l = 0; l < k.L; ++l) exp_max = fmax(exp_max, -st->stats[q] / 2 / k.den[q][l]);
    double acc[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) acc[l] = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        for (int q = 0; q < k.K; ++q) {
            const double r = r1_all[(int64_t)q * M + j];
            const double r2 = r * r;
            double pr[SGV_MAX_L], den = 0.0;
            for (int l = 0; l < k.L; ++l) {
                pr[l] = exp(-r2 / 2 / k.den[q][l] - exp_max) / k.sqden[q][l];       // :154
                den += pr[l] * k.omega[l];                                           // :156
            }
#pragma unroll
            for (int l = 0; l < 8; ++l)
                if (l < k.L) acc[l] += k.a[q] * pr[l] / den;                         // :155,:158
        }
    }
    grid_reduce<8>(acc, rc, red);
}

extern "C" int sgv_lagrangian(sgv_handle c, const double* gam1s, const double* x, const double* omega0,
                              const double* sigma2, double* y) {
    SGV_TRY(check_ready(c));
    const PriorParams& p = c->prior;
    const int L = p.L;
    SGV_CHECK(L >= 2, "prior not set");
    LagConsts k;
    k.K = p.K;
    k.L = L;
    for (int q = 0; q < p.K; ++q) {
        k.a[q] = p.a[q];
        const double ginv = 1.0 / gam1s[q];
        for (int l = 0; l < L; ++l) {
            k.den[q][l] = sigma2[l] + ginv;
            k.sqden[q][l] = std::sqrt(sigma2[l] + ginv);
        }
    }
    for (int l = 0; l < L; ++l) k.omega[l] = x[l];
    const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * 4);
    SGV_TRY(sgv_ensure_partials(c, grid + 1));
    RedCtx rc1 = sgv_red_begin(c, AP_STATS, 8, 0, 0, 0, 1);
    k_min_r2<<<grid, 256, 0, c->stream>>>(c->Ml, p.K, c->r1_all, rc1);
    SGV_TRY(sgv_red_end(c, rc1));
    RedCtx rc2 = sgv_red_begin(c, AP_STATS, 8, 8);
    k_lagrangian<<<grid, 256, 0, c->stream>>>(c->Ml, c->r1_all, k, rc2);
    c->launches += 2;
    SGV_CUDA(cudaGetLastError());
    SGV_TRY(sgv_red_end(c, rc2));
    SGV_TRY(fetch_state(c));
    double osum = 0.0;
    for (int l = 0; l < L; ++l) {
        y[l] = c->cg_host->stats[8 + l] + (omega0[l] - 1) / x[l] + x[L];                     // :158
        osum += x[l];
    }
    y[L] = osum - 1.0;                                                               // :159
    return 0;
}

// ---------------------------------------------------------------------------------------------
// metrics vs truth (src/sgvamp.py:379-382)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_metrics(int64_t M, const double* __restrict__ xhat1, const double* __restrict__ x0, RedCtx rc) {
    __shared__ double red[4 * 32];
    double acc[4] = {0, 0, 0, 0};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        const double a = xhat1[j], b = x0[j];
        acc[0] += a * b;
        acc[1] += a * a;
        acc[2] += b * b;
        acc[3] += (a - b) * (a - b);
    }
    grid_reduce<4>(acc, rc, red);
}

extern "C" int sgv_metrics(sgv_handle c, const double* x0, double* dots) {
    SGV_TRY(check_ready(c));
    if (x0 != nullptr) {
        SGV_CUDA(cudaMemcpyAsync(c->truth, x0, c->Ml * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        c->truth_set = true;
    }
    SGV_CHECK(c->truth_set, "truth vector not uploaded yet");
    const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * 4);
    SGV_TRY(sgv_ensure_partials(c, grid + 1));
    RedCtx rc = sgv_red_begin(c, AP_STATS, 4, 0);
    k_metrics<<<grid, 256, 0, c->stream>>>(c->Ml, c->xhat1, c->truth, rc);
    c->launches++;
    SGV_CUDA(cudaGetLastError());
    SGV_TRY(sgv_red_end(c, rc));
    SGV_TRY(fetch_state(c));
    for (int i = 0; i < 4; ++i) dots[i] = c->cg_host->stats[i];
    return 0;
}

// ---------------------------------------------------------------------------------------------
// LMMSE: set-up, CG vector kernels, post-processing
// ---------------------------------------------------------------------------------------------
// r2 = (xhat1 - alpha1 r1)/(1-alpha1) (:310); b0 = mu2 = gamw r + gam2 r2 (:313); b1 = u (:326);
// x0 = (xhat2_prev, Sigma2_u_prev) (:316,:332); |b|^2 per column; initialise the CG state.
// Warm start: r = b - A x0 = b - gamw (R x0) - gam2 x0 needs no matrix pass, because R x0 was recovered
// from the previous solve's own recursion (rxs, see k_lmmse_post).
__global__ void __launch_bounds__(256)
k_lmmse_setup(int64_t M, const double* __restrict__ xhat1, const double* __restrict__ r1,
              const double* __restrict__ xty, const int8_t* __restrict__ probe, const double* __restrict__ xhat2,
              const double* __restrict__ sig, const double2* __restrict__ rxs, double* __restrict__ r2,
              double2* __restrict__ bb, double2* __restrict__ xx, double2* __restrict__ rr, double alpha1, double gamw,
              double gam2, int x0_zero, RedCtx rc, int vs_cohort) {
    __shared__ double red[4 * 32];
    if (vs_cohort >= 0) {   // fused iteration: the denoiser's finaliser left alpha1 / gam2 on the device
        alpha1 = rc.st->vs.alpha1[vs_cohort];
        gamw = rc.st->vs.gamw[vs_cohort];
        gam2 = rc.st->vs.gam2[vs_cohort];
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        const double r2j = (xhat1[j] - alpha1 * r1[j]) / (1.0 - alpha1);
        r2[j] = r2j;
        const double2 b = make_double2(gamw * xty[j] + gam2 * r2j, (double)probe[j]);
        bb[j] = b;
        const double2 x0 = make_double2(xhat2[j], sig[j]);
        xx[j] = x0;
        double2 r = b;
        if (!x0_zero) {
            const double2 rx = rxs[j];
            r.x = b.x - (gamw * rx.x + gam2 * x0.x);
            r.y = b.y - (gamw * rx.y + gam2 * x0.y);
        }
        rr[j] = r;
        acc[0] += b.x * b.x;
        acc[1] += b.y * b.y;
        acc[2] += r.x * r.x;
        acc[3] += r.y * r.y;
    }
    grid_reduce<4>(acc, rc, red);
}

// p = r (first step) or p = r + (rho/rho_prev) p      (layouts without the fused update)
__global__ void __launch_bounds__(256)
k_p_update(int64_t M, const double2* __restrict__ rr, double2* __restrict__ pp, const CgState* __restrict__ st) {
    const int d0 = st->done[0], d1 = st->done[1];
    if (d0 && d1) return;
    const bool first = st->step == 0;
    const double b0 = first ? 0.0 : st->rho[0] / st->rho_prev[0];
    const double b1 = first ? 0.0 : st->rho[1] / st->rho_prev[1];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        const double2 r = rr[j];
        double2 p = first ? make_double2(0.0, 0.0) : pp[j];
        // scipy: p *= beta; p += z
        if (!d0) p.x = first ? r.x : (p.x * b0 + r.x);
        if (!d1) p.y = first ? r.y : (p.y * b1 + r.y);
        pp[j] = p;
    }
}

// alpha = rho/(p.q); x += alpha p; r -= alpha q; r.r  (state transition: AP_CGUPDATE)
__global__ void __launch_bounds__(256)
k_cg_update(int64_t M, double2* __restrict__ xx, double2* __restrict__ rr, const double2* __restrict__ pp,
            const double2* __restrict__ qq, RedCtx rc) {
    __shared__ double red[2 * 32];
    const CgState* st = rc.st;
    const int d0 = st->done[0], d1 = st->done[1];
    if (d0 && d1) return;
    const double a0 = d0 ? 0.0 : st->rho[0] / st->pq[0];
    const double a1 = d1 ? 0.0 : st->rho[1] / st->pq[1];
    double acc[2] = {0.0, 0.0};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        double2 x = xx[j], r = rr[j];
        const double2 p = pp[j], q = qq[j];
        if (!d0) { x.x += a0 * p.x; r.x -= a0 * q.x; }
        if (!d1) { x.y += a1 * p.y; r.y -= a1 * q.y; }
        xx[j] = x;
        rr[j] = r;
        acc[0] += r.x * r.x;
        acc[1] += r.y * r.y;
    }
    grid_reduce<2>(acc, rc, red);
}

// xhat2 <- CG col 0 (damped with the previous xhat2 if lmmse_damp, :322-323); Sigma2_u_prev <- col 1
// (:333); dots u.Sigma2_u (:338), xhat2.r (:352).  The products R xhat2 and R Sigma2_u that the gamw
// update (:352,:359) and the next warm start need are recovered from the CG recursion itself,
//     A x = b - r   =>   R x = (b - r - gam2 x) / gamw        (r: the solve's final recursive residual),
// so neither costs a pass over the matrix; xhat2^T R xhat2 and u^T R Sigma2_u are summed here.
// fusedcg: the solve ran as fused steps (spmm_dsym.cu); the last step's update x += alpha p, r -= alpha q is
// still pending and is applied here (buffers (step-1)&1 hold the last r, p, q).
struct CgBufs {
    const double2 *rr[2], *pp[2], *qq[2];
    int fusedcg;
};

__global__ void __launch_bounds__(256)
k_lmmse_post(int64_t M, double2* __restrict__ xx, CgBufs cb, const double2* __restrict__ bb,
             const double* __restrict__ xty, double* __restrict__ xhat2, double* __restrict__ sig,
             double2* __restrict__ rxs, double gamw, double gam2, double rho, int damp, RedCtx rc, int vs_cohort) {
    __shared__ double red[4 * 32];
    if (vs_cohort >= 0) {
        gamw = rc.st->vs.gamw[vs_cohort];
        gam2 = rc.st->vs.gam2[vs_cohort];
    }
    const int z0 = rc.st->zero_b[0], z1 = rc.st->zero_b[1];
    const double igw = 1.0 / gamw;
    const int step = rc.st->step;
    const bool pend = cb.fusedcg && step > 0;
    const int cur = pend ? ((step - 1) & 1) : 1;        // r_0 from the set-up kernel lives in buffer 1
    const double al0 = pend ? rc.st->alpha[0] : 0.0, al1 = pend ? rc.st->alpha[1] : 0.0;
    const double2* __restrict__ rr = cb.rr[cur];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        double2 x = xx[j];
        const double2 b = bb[j];
        double2 r = rr[j];
        if (al0 != 0.0 || al1 != 0.0) {
            const double2 p = cb.pp[cur][j], q = cb.qq[cur][j];
            if (al0 != 0.0) { x.x += al0 * p.x; r.x -= al0 * q.x; }
            if (al1 != 0.0) { x.y += al1 * p.y; r.y -= al1 * q.y; }
        }
        double2 rx;
        rx.x = z0 ? 0.0 : (b.x - r.x - gam2 * x.x) * igw;      // scipy returns x = b (= 0) when |b| = 0
        rx.y = z1 ? 0.0 : (b.y - r.y - gam2 * x.y) * igw;
        if (z0) x.x = b.x;
        if (z1) x.y = b.y;
        if (damp) {
            x.x = rho * x.x + (1.0 - rho) * xhat2[j];
            rx.x = rho * rx.x + (1.0 - rho) * rxs[j].x;
        }
        xhat2[j] = x.x;
        sig[j] = x.y;
        xx[j] = x;
        rxs[j] = rx;
        acc[0] += b.y * x.y;       // u . Sigma2_u
        acc[1] += x.x * xty[j];    // xhat2 . r
        acc[2] += x.x * rx.x;      // xhat2^T R xhat2
        acc[3] += b.y * rx.y;      // u^T R Sigma2_u
    }
    grid_reduce<4>(acc, rc, red);
}

// xx = (xhat2, Sigma2_u) as a vector pair (only when a warm start was injected from outside and R x0 has
// to be formed by a real pass); the reduction doubles as the cross-rank ordering point for the halo reads
__global__ void __launch_bounds__(256)
k_pack_x0(int64_t M, const double* __restrict__ xhat2, const double* __restrict__ sig, double2* __restrict__ xx, RedCtx rc) {
    __shared__ double red[32];
    double acc[1] = {0.0};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        const double2 x = make_double2(xhat2[j], sig[j]);
        xx[j] = x;
        acc[0] += x.x * x.x + x.y * x.y;
    }
    grid_reduce<1>(acc, rc, red);
}

__global__ void __launch_bounds__(256)
k_update_r1(int64_t M, const double* __restrict__ xhat2, const double* __restrict__ r2, double* __restrict__ r1,
            double alpha2, const CgState* __restrict__ st, int vs_cohort) {
    if (vs_cohort >= 0) alpha2 = st->vs.alpha2[vs_cohort];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x)
        r1[j] = (xhat2[j] - alpha2 * r2[j]) / (1.0 - alpha2);                       // :348
}

extern "C" int sgv_update_r1(sgv_handle c, int cohort, double alpha2) {
    SGV_TRY(check_ready(c));
    SGV_CHECK(cohort >= 0 && cohort < c->K, "cohort out of range");
    Cohort& co = c->coh[cohort];
    const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * 8);
    k_update_r1<<<grid, 256, 0, c->stream>>>(c->Ml, co.xhat2, co.r2, co.r1, alpha2, c->cg, -1);
    c->launches++;
    SGV_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_probe_setup(int64_t M, const int8_t* pa, const int8_t* pb, double2* bb, double2* xx, double2* rr, RedCtx rc);
__global__ void k_probe_post(int64_t M, const double2* xx, CgBufs cb, const double2* bb, double gamw, double gam2, RedCtx rc);

int sgv_preload_vamp() {
    cudaFuncAttributes fa;
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_denoise));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_em));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_min_r2));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_lagrangian));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_metrics));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_lmmse_setup));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_p_update));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_cg_update));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_lmmse_post));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_update_r1));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_pack_x0));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_em_loop));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_em_begin));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_probe_setup));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_probe_post));
    return 0;
}

// The CG loop of one 2-RHS solve on A = gamw R + gam2 I, enqueued on the handle's stream after a set-up kernel has left
// b, x0, r_0 and the CG state in place: the whole-solve kernel where it applies, otherwise batches of steps with a state
// read-back between batches.
static int cg_enqueue(sgv_ctx* c, Cohort& co, double gamw, double gam2, int cg_maxit, bool dev, int first_batch,
                      bool* state_fetched) {
    const int64_t M = c->Ml;
    const bool fused = co.ld.layout == SGV_LAYOUT_DIA || co.ld.rowpart;
    const bool fusedcg = co.ld.layout == SGV_LAYOUT_DSYM;
    const unsigned vgrid = (unsigned)std::min<int64_t>((M + 255) / 256, (int64_t)c->sm_count * 8);
    int launched = 0;
    int batch = first_batch;   // see sgv_prior_em: counts drift slowly
    CgState* hs = c->cg_host;
    if (fusedcg && cg_maxit > 0 && sgv_dsymp_solve_usable(c, co.ld)) {
        // the whole solve in one cooperative launch: steps separated by a grid barrier, not by launches
        SGV_TRY(sgv_launch_dsym_solve(c, co, gamw, gam2, cg_maxit));
        if (!dev) {
            SGV_TRY(fetch_state(c));
            *state_fetched = true;
        }
        launched = cg_maxit;
    }
    while (launched < cg_maxit) {
        const int nb = std::min(batch, cg_maxit - launched);
        for (int b = 0; b < nb; ++b) {
            const int n = launched + b;             // == device step while the solve is active
            if (fusedcg) {
                SGV_TRY(sgv_launch_dsym_cg(c, co, n, gamw, gam2));
                continue;
            }
            double2* pcur;
            if (fused) {
                pcur = c->pp[n & 1];
                SGV_TRY(sgv_launch_spmm(c, co, EPI_Q, VEC_PP0 + ((n + 1) & 1), c->qq, gamw, gam2, 1, 1));
            } else {
                pcur = c->pp[0];
                k_p_update<<<vgrid, 256, 0, c->stream>>>(M, c->rr, pcur, c->cg);
                c->launches++;
                SGV_TRY(sgv_launch_spmm(c, co, EPI_Q, VEC_PP0, c->qq, gamw, gam2, 1, 0));
            }
            RedCtx rc = sgv_red_begin(c, AP_CGUPDATE, 2, 0);
            rc.skip_if_done = SKIP_CG_DONE;
            k_cg_update<<<vgrid, 256, 0, c->stream>>>(M, c->xx, c->rr, pcur, c->qq, rc);
            c->launches++;
            SGV_TRY(sgv_red_end(c, rc));
        }
        launched += nb;
        SGV_TRY(fetch_state(c));
        *state_fetched = true;
        if (hs->done[0] && hs->done[1]) break;
        batch = 8;
    }
    return 0;
}

// The LMMSE step of one cohort, enqueued on the handle's stream.  dev == false: scalars by value (reference-style
// stepwise API), the caller reads the state back afterwards.  dev == true (fused iteration): alpha1 / gamw / gam2 are read
// from the device-resident chain, the post kernel's finaliser advances it (AP_POST) and nothing is read back here -
// except by layouts whose CG loop is enqueued in batches (every layout but the whole-solve half-band kernel).
static int lmmse_enqueue(sgv_ctx* c, int cohort, const sgv_lmmse_in* in, bool dev, int it, int* passes_out) {
    Cohort& co = c->coh[cohort];
    SGV_CHECK(co.ld.layout != 0, "cohort %d has no LD matrix", cohort);
    const int64_t M = c->Ml;
    const int vsc = dev ? cohort : -1;
    c->vs_active = vsc;
    // direction update fused into the SpMM staging (DIA) / into the all-gather of the dense rows partition
    const bool fused = co.ld.layout == SGV_LAYOUT_DIA || co.ld.rowpart;
    const bool fusedcg = co.ld.layout == SGV_LAYOUT_DSYM;  // whole CG step in one kernel
    const unsigned vgrid = (unsigned)std::min<int64_t>((M + 255) / 256, (int64_t)c->sm_count * 8);
    SGV_TRY(sgv_ensure_partials(c, vgrid + 1));
    int passes = 0;
    if (in->x0_zero && !co.rxs_valid) {    // the caller states x0 = 0: so is R x0
        SGV_CUDA(cudaMemsetAsync(co.rxs, 0, (size_t)M * sizeof(double2), c->stream));
        co.rxs_valid = true;
    }
    if (!in->x0_zero && !co.rxs_valid) {   // injected warm start: R x0 by a real pass
        c->vs_active = -1;                  // plain product: alpha = 1, beta = 0 by value
        RedCtx rc = sgv_red_begin(c, AP_STATS, 1, 15);
        k_pack_x0<<<vgrid, 256, 0, c->stream>>>(M, co.xhat2, co.sig, c->xx, rc);
        c->launches++;
        SGV_TRY(sgv_red_end(c, rc));
        SGV_TRY(sgv_launch_spmm(c, co, EPI_PLAIN, VEC_XX, co.rxs, 1.0, 0.0, 0, 0));
        if (c->world > 1) {                 // the pass must be complete on every rank before xx is rewritten
            RedCtx rc2 = sgv_red_begin(c, AP_STATS, 1, 15);
            k_pack_x0<<<vgrid, 256, 0, c->stream>>>(M, co.xhat2, co.sig, c->xx, rc2);
            c->launches++;
            SGV_TRY(sgv_red_end(c, rc2));
        }
        co.rxs_valid = true;
        passes++;
        c->vs_active = vsc;
    }
    {   // b, x0, r = b - A x0 (no matrix pass: R x0 is kept from the previous solve), |b|^2, r.r, loop-top test of iteration 0
        RedCtx rc = sgv_red_begin(c, AP_SETUP, 4, 0, in->cg_maxit, in->x0_zero);
        k_lmmse_setup<<<vgrid, 256, 0, c->stream>>>(M, c->xhat1, co.r1, co.xty, co.probe, co.xhat2, co.sig, co.rxs, co.r2,
                                                   c->bb, c->xx, c->rr, in->alpha1, in->gamw, in->gam2, in->x0_zero, rc, vsc);
        c->launches++;
        SGV_TRY(sgv_red_end(c, rc));
    }
    CgState* hs = c->cg_host;
    bool state_fetched = false;
    SGV_TRY(cg_enqueue(c, co, in->gamw, in->gam2, in->cg_maxit, dev, co.last_cg_iters > 0 ? co.last_cg_iters + 2 : 4, &state_fetched));
    if (state_fetched) co.last_cg_iters = std::max(hs->iters[0], hs->iters[1]);
    {
        RedCtx rc = sgv_red_begin(c, dev ? AP_POST : AP_STATS, 4, 0);
        rc.ap.cohort = cohort;
        rc.ap.it = it;
        rc.ap.lmmse_damp = in->lmmse_damp;
        rc.ap.learn_gamw = in->learn_gamw;
        rc.ap.rho = in->rho;
        rc.ap.Mtot = (double)c->M;
        CgBufs cb;
        for (int i = 0; i < 2; ++i) {
            cb.rr[i] = c->rr2[i];
            cb.pp[i] = c->pp[i];
            cb.qq[i] = c->qq2[i];
        }
        cb.fusedcg = fusedcg;
        k_lmmse_post<<<vgrid, 256, 0, c->stream>>>(M, c->xx, cb, c->bb, co.xty, co.xhat2, co.sig, co.rxs, in->gamw, in->gam2,
                                                  in->rho, in->lmmse_damp, rc, vsc);
        c->launches++;
        SGV_TRY(sgv_red_end(c, rc));
    }
    co.rxs_valid = true;
    c->vs_active = -1;
    SGV_CUDA(cudaGetLastError());
    if (passes_out) *passes_out = passes;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Additional Hutchinson probes (no reference counterpart: src/sgvamp.py:326-340 draws exactly one probe per cohort and
// iteration).  Two more probes per call ride the two columns of the 2-RHS solver:  A s = u  from x0 = 0
// for u = probe_a, probe_b; returned are u.s (the trace estimate of :338) and u^T R s (the gamw update's trace, :359),
// recovered from the recursion as in k_lmmse_post.  Nothing of the solver's state (xhat2, Sigma2_u, R x0) is touched.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_probe_setup(int64_t M, const int8_t* __restrict__ pa, const int8_t* __restrict__ pb, double2* __restrict__ bb,
              double2* __restrict__ xx, double2* __restrict__ rr, RedCtx rc) {
    __shared__ double red[4 * 32];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        const double2 b = make_double2((double)pa[j], pb != nullptr ? (double)pb[j] : 0.0);
        bb[j] = b;
        xx[j] = make_double2(0.0, 0.0);
        rr[j] = b;
        acc[0] += b.x * b.x;
        acc[1] += b.y * b.y;
    }
    acc[2] = acc[0];
    acc[3] = acc[1];
    grid_reduce<4>(acc, rc, red);
}

__global__ void __launch_bounds__(256)
k_probe_post(int64_t M, const double2* __restrict__ xx, CgBufs cb, const double2* __restrict__ bb, double gamw, double gam2,
             RedCtx rc) {
    __shared__ double red[4 * 32];
    const int z0 = rc.st->zero_b[0], z1 = rc.st->zero_b[1];
    const double igw = 1.0 / gamw;
    const int step = rc.st->step;
    const bool pend = cb.fusedcg && step > 0;
    const int cur = pend ? ((step - 1) & 1) : 1;
    const double al0 = pend ? rc.st->alpha[0] : 0.0, al1 = pend ? rc.st->alpha[1] : 0.0;
    const double2* __restrict__ rr = cb.rr[cur];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        double2 x = xx[j];
        const double2 b = bb[j];
        double2 r = rr[j];
        if (al0 != 0.0 || al1 != 0.0) {
            const double2 p = cb.pp[cur][j], q = cb.qq[cur][j];
            if (al0 != 0.0) { x.x += al0 * p.x; r.x -= al0 * q.x; }
            if (al1 != 0.0) { x.y += al1 * p.y; r.y -= al1 * q.y; }
        }
        const double rx0 = z0 ? 0.0 : (b.x - r.x - gam2 * x.x) * igw;
        const double rx1 = z1 ? 0.0 : (b.y - r.y - gam2 * x.y) * igw;
        if (z0) x.x = b.x;
        if (z1) x.y = b.y;
        acc[0] += b.x * x.x;       // u_a . s_a
        acc[1] += b.y * x.y;       // u_b . s_b
        acc[2] += b.x * rx0;       // u_a^T R s_a
        acc[3] += b.y * rx1;       // u_b^T R s_b
    }
    grid_reduce<4>(acc, rc, red);
}

extern "C" int sgv_probe_pair(sgv_handle c, int cohort, double gamw, double gam2, int cg_maxit, const int8_t* probe_a,
                              const int8_t* probe_b, sgv_probe_out* out) {
    SGV_TRY(check_ready(c));
    SGV_CHECK(cohort >= 0 && cohort < c->K, "cohort out of range");
    SGV_CHECK(probe_a && out, "null argument");
    Cohort& co = c->coh[cohort];
    SGV_CHECK(co.ld.layout != 0, "cohort %d has no LD matrix", cohort);
    const int64_t M = c->Ml;
    SGV_CUDA(cudaMemcpyAsync(co.probe, probe_a, M, cudaMemcpyHostToDevice, c->stream));
    if (probe_b) SGV_CUDA(cudaMemcpyAsync(c->probe_b, probe_b, M, cudaMemcpyHostToDevice, c->stream));
    c->vs_active = -1;
    const unsigned vgrid = (unsigned)std::min<int64_t>((M + 255) / 256, (int64_t)c->sm_count * 8);
    SGV_TRY(sgv_ensure_partials(c, vgrid + 1));
    {
        RedCtx rc = sgv_red_begin(c, AP_SETUP, 4, 0, cg_maxit, 1);
        k_probe_setup<<<vgrid, 256, 0, c->stream>>>(M, co.probe, probe_b ? c->probe_b : nullptr, c->bb, c->xx, c->rr, rc);
        c->launches++;
        SGV_TRY(sgv_red_end(c, rc));
    }
    bool state_fetched = false;
    SGV_TRY(cg_enqueue(c, co, gamw, gam2, cg_maxit, false, co.last_cg_iters > 0 ? co.last_cg_iters + 2 : 4, &state_fetched));
    {
        RedCtx rc = sgv_red_begin(c, AP_STATS, 4, 0);
        CgBufs cb;
        for (int i = 0; i < 2; ++i) {
            cb.rr[i] = c->rr2[i];
            cb.pp[i] = c->pp[i];
            cb.qq[i] = c->qq2[i];
        }
        cb.fusedcg = co.ld.layout == SGV_LAYOUT_DSYM;
        k_probe_post<<<vgrid, 256, 0, c->stream>>>(M, c->xx, cb, c->bb, gamw, gam2, rc);
        c->launches++;
        SGV_TRY(sgv_red_end(c, rc));
    }
    SGV_CUDA(cudaGetLastError());
    SGV_TRY(fetch_state(c));
    CgState* hs = c->cg_host;
    for (int i = 0; i < 2; ++i) {
        out->u_s[i] = hs->stats[i];
        out->u_R_s[i] = hs->stats[2 + i];
        out->cg_iters[i] = hs->iters[i];
        out->cg_info[i] = hs->done[i] ? hs->info[i] : cg_maxit;
    }
    out->spmm_passes = std::max(hs->iters[0], hs->iters[1]);
    return 0;
}

extern "C" int sgv_lmmse(sgv_handle c, int cohort, const sgv_lmmse_in* in, const int8_t* probe, sgv_lmmse_out* out) {
    SGV_TRY(check_ready(c));
    SGV_CHECK(cohort >= 0 && cohort < c->K, "cohort out of range");
    SGV_CHECK(in && probe && out, "null argument");
    Cohort& co = c->coh[cohort];
    SGV_CUDA(cudaMemcpyAsync(co.probe, probe, c->Ml, cudaMemcpyHostToDevice, c->stream));
    int passes = 0;
    SGV_TRY(lmmse_enqueue(c, cohort, in, false, 0, &passes));
    SGV_TRY(fetch_state(c));
    CgState* hs = c->cg_host;
    out->cg_iters[0] = hs->iters[0];
    out->cg_iters[1] = hs->iters[1];
    co.last_cg_iters = std::max(hs->iters[0], hs->iters[1]);
    // a column that ran out of iterations without ever passing the test reports maxiter (scipy)
    out->cg_info[0] = hs->done[0] ? hs->info[0] : in->cg_maxit;
    out->cg_info[1] = hs->done[1] ? hs->info[1] : in->cg_maxit;
    passes += std::max(hs->iters[0], hs->iters[1]);
    out->u_sigma2u = hs->stats[0];
    out->xhat2_r = hs->stats[1];
    out->xhat2_R_xhat2 = hs->stats[2];      // (:352,:359) - no extra matrix pass, see k_lmmse_post
    out->u_R_sigma2u = hs->stats[3];
    out->spmm_passes = passes;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused VAMP iteration (src/sgvamp.py:222-387 for all cohorts of this process): prior update (EM), denoiser, the
// LMMSE step of every cohort, r1 update, metrics - enqueued without a host round trip; the scalar chain
// (alpha1, gam2, alpha2, gam1, gamw) advances on the device in the kernels' finalisers.  The host reads ONE small
// record per iteration (sgv_iteration_wait) and may enqueue the next iteration before it does.
// ---------------------------------------------------------------------------------------------
extern "C" int sgv_iteration_supported(sgv_handle c) {
    if (c == nullptr) return 0;
    return c->coop_ok && !(c->world > 1 && c->host_barrier) ? 1 : 0;
}

extern "C" int sgv_vamp_begin(sgv_handle c, const double* gam1, const double* gamw, const double* N) {
    SGV_TRY(check_ready(c));
    SGV_CHECK(gam1 && gamw && N, "null argument");
    const PriorParams& p = c->prior;
    SGV_CHECK(p.L >= 2, "prior not set");
    SGV_CHECK(sgv_iteration_supported(c), "fused iteration needs cooperative launches and one GPU per rank");
    SGV_TRY(ensure_em_cache(c));
    for (int i = 0; i < sgv_ctx::NLOG; ++i) {
        if (c->log_host[i] == nullptr) {
            SGV_CUDA(cudaMallocHost(&c->log_host[i], sizeof(IterLog)));
            SGV_CUDA(cudaEventCreateWithFlags(&c->log_ev[i], cudaEventDisableTiming));
        }
    }
    const int64_t pbytes = (int64_t)c->K * c->Ml;
    if (c->probe_pin_bytes < pbytes) {
        for (int i = 0; i < sgv_ctx::NLOG; ++i) {
            if (c->probe_pin[i]) cudaFreeHost(c->probe_pin[i]);
            c->probe_pin[i] = nullptr;
            SGV_CUDA(cudaMallocHost(&c->probe_pin[i], pbytes));
        }
        c->probe_pin_bytes = pbytes;
    }
    CgState* hs = c->cg_host;
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    memset(&hs->vs, 0, sizeof(VampScal));
    memset(&hs->em, 0, sizeof(EmState));
    hs->vs.K = p.K;
    hs->vs.Lm1 = p.L - 1;
    hs->em.K = p.K;
    hs->em.Lm1 = p.L - 1;
    hs->em.lam = p.lam;
    hs->em.Mtot = (double)c->M;
    for (int l = 0; l < p.L - 1; ++l) {
        hs->vs.sigmas[l] = p.sigmas[l];
        hs->em.omegas[l] = p.omegas[l];
    }
    for (int k = 0; k < p.K; ++k) {
        hs->vs.gam1[k] = gam1[k];
        hs->vs.gamw[k] = gamw[k];
        hs->vs.a[k] = p.a[k];
        hs->vs.N[k] = N[k];
        hs->em.a[k] = p.a[k];
        hs->em.asum += p.a[k];
    }
    SGV_CUDA(cudaMemcpyAsync(&c->cg->vs, &hs->vs, sizeof(VampScal), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaMemcpyAsync(&c->cg->em, &hs->em, sizeof(EmState), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    c->vamp_begun = true;
    return 0;
}

// resume (no reference counterpart: the reference cannot restart): the damping terms of :290-291 / :345-346
// need the previous iteration's alpha1 / alpha2
extern "C" int sgv_vamp_set_alphas(sgv_handle c, const double* alpha1, const double* alpha2) {
    SGV_TRY(check_ready(c));
    SGV_CHECK(c->vamp_begun && alpha1 && alpha2, "sgv_vamp_begin has not been called");
    SGV_CUDA(cudaMemcpyAsync(c->cg->vs.alpha1, alpha1, c->K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaMemcpyAsync(c->cg->vs.alpha2, alpha2, c->K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int sgv_set_truth(sgv_handle c, const double* x0) {
    SGV_TRY(check_ready(c));
    SGV_CHECK(x0 != nullptr, "null argument");
    SGV_CUDA(cudaMemcpyAsync(c->truth, x0, c->Ml * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    c->truth_set = true;
    return 0;
}

extern "C" int sgv_iteration_probe_buffer(sgv_handle c, int slot, int8_t** buf) {
    SGV_CHECK(c != nullptr && buf != nullptr && slot >= 0 && slot < sgv_ctx::NLOG, "bad arguments");
    SGV_CHECK(c->vamp_begun && c->probe_pin[slot] != nullptr, "sgv_vamp_begin has not been called");
    *buf = c->probe_pin[slot];
    return 0;
}

extern "C" int sgv_iteration_enqueue(sgv_handle c, const sgv_iter_in* in, double* xhat_pinned, double* const* r1_pinned,
                                     int slot) {
    SGV_TRY(check_ready(c));
    SGV_CHECK(in != nullptr && slot >= 0 && slot < sgv_ctx::NLOG, "bad arguments");
    SGV_CHECK(c->vamp_begun, "sgv_vamp_begin has not been called");
    const int K = c->K;
    const int64_t M = c->Ml;
    const unsigned vgrid = (unsigned)std::min<int64_t>((M + 255) / 256, (int64_t)c->sm_count * 8);
    // ---- prior update: EM loop on the device (src/sgvamp.py:250-257); lam / omegas stay in CgState::em
    if (in->update_prior && in->em_maxit > 0) {
        if (c->em_loop_blocks_per_sm == 0) {
            int nb = 0;
            SGV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_em_loop, 256, 0));
            c->em_loop_blocks_per_sm = std::max(1, std::min(nb, 4));
        }
        // a pass is two grid barriers and a partial-sum sweep: at least 1024 markers per block keeps both short
        const unsigned lgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((M + 1023) / 1024, (int64_t)c->sm_count * c->em_loop_blocks_per_sm));
        SGV_TRY(sgv_ensure_partials(c, lgrid + 1));
        k_em_begin<<<1, 1, 0, c->stream>>>(c->cg, in->em_maxit, in->em_tol);
        c->launches++;
        unsigned* bar = c->counter + 4;
        SGV_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), c->stream));
        RedCtx rc = sgv_red_begin(c, AP_EM, 16, 0);
        int64_t Ml = M;
        const double* r1 = c->r1_all;
        EmConsts kdummy;
        memset(&kdummy, 0, sizeof(kdummy));
        int mi = in->em_maxit, dev1 = 1;
        double* cache = c->em_cache;      // allocated by sgv_vamp_begin
        void* args[] = {&Ml, &r1, &kdummy, &rc, &bar, &mi, &dev1, &cache};
        SGV_CUDA(cudaLaunchCooperativeKernel((const void*)k_em_loop, dim3(lgrid), dim3(256), args, 0, c->stream));
        c->launches++;
    } else {
        k_em_begin<<<1, 1, 0, c->stream>>>(c->cg, 0, 0.0);          // steps = 0 for the log
        c->launches++;
    }
    // ---- denoiser + derivative + damping (:270-293); its finaliser sets alpha1 / gam2 of every cohort
    {
        const unsigned grid = (unsigned)std::min<int64_t>((M + 255) / 256, (int64_t)c->sm_count * 8);
        SGV_TRY(sgv_ensure_partials(c, grid));
        RedCtx rc = sgv_red_begin(c, AP_DENOISE, 1, 0);
        rc.ap.it = in->it;
        rc.ap.rho = in->rho;
        rc.ap.Mtot = (double)c->M;
        DenoiseConsts kdummy;
        memset(&kdummy, 0, sizeof(kdummy));
        k_denoise<<<grid, 256, 0, c->stream>>>(M, c->r1_all, c->xhat1, kdummy, in->rho, in->it > 0, rc, 1);
        c->launches++;
        SGV_CUDA(cudaGetLastError());
        SGV_TRY(sgv_red_end(c, rc));
    }
    // ---- output snapshots (the xhat1 of this iteration, the r1 that entered it: :280-283)
    if (xhat_pinned) SGV_TRY(sgv_get_vec_async(c, 0, SGV_VEC_XHAT1, 1.0, xhat_pinned));
    if (r1_pinned)
        for (int k = 0; k < K; ++k)
            if (r1_pinned[k]) SGV_TRY(sgv_get_vec_async(c, k, SGV_VEC_R1, 1.0, r1_pinned[k]));
    // ---- LMMSE + Hutchinson + gamw + r1 update, cohort by cohort (:301-374)
    for (int k = 0; k < K; ++k) {
        Cohort& co = c->coh[k];
        SGV_CUDA(cudaMemcpyAsync(co.probe, c->probe_pin[slot] + (size_t)k * M, M, cudaMemcpyHostToDevice, c->stream));
        sgv_lmmse_in li;
        memset(&li, 0, sizeof(li));
        li.rho = in->rho;
        li.cg_maxit = in->cg_maxit;
        li.lmmse_damp = in->lmmse_damp;
        li.learn_gamw = in->learn_gamw;
        li.x0_zero = in->it == 0;
        SGV_TRY(lmmse_enqueue(c, k, &li, true, in->it, nullptr));
        k_update_r1<<<vgrid, 256, 0, c->stream>>>(M, co.xhat2, co.r2, co.r1, 0.0, c->cg, k);
        c->launches++;
    }
    // ---- metrics vs truth (:379-387)
    if (in->want_metrics) {
        SGV_CHECK(c->truth_set, "truth vector not uploaded (sgv_set_truth)");
        const unsigned grid = (unsigned)std::min<int64_t>((M + 255) / 256, (int64_t)c->sm_count * 4);
        SGV_TRY(sgv_ensure_partials(c, grid + 1));
        RedCtx rc = sgv_red_begin(c, AP_METRICS, 4, 0);
        k_metrics<<<grid, 256, 0, c->stream>>>(M, c->xhat1, c->truth, rc);
        c->launches++;
        SGV_TRY(sgv_red_end(c, rc));
    }
    SGV_CUDA(cudaGetLastError());
    SGV_CUDA(cudaMemcpyAsync(c->log_host[slot], &c->cg->log, sizeof(IterLog), cudaMemcpyDeviceToHost, c->stream));
    SGV_CUDA(cudaEventRecord(c->log_ev[slot], c->stream));
    return 0;
}

extern "C" int sgv_iteration_wait(sgv_handle c, int slot, sgv_iter_out* out) {
    SGV_CHECK(c != nullptr && out != nullptr && slot >= 0 && slot < sgv_ctx::NLOG, "bad arguments");
    SGV_CUDA(cudaSetDevice(c->device));
    SGV_CUDA(cudaEventSynchronize(c->log_ev[slot]));
    const IterLog* L = c->log_host[slot];
    SGV_CHECK(L->error == 0, "cross-rank reduction timed out on rank %d of %d (mask 0x%x)", c->rank, c->world, L->error);
    memset(out, 0, sizeof(*out));
    out->lam = L->lam;
    out->em_steps = L->em_steps;
    out->em_relerr = L->em_relerr;
    for (int l = 0; l < SGV_MAX_L; ++l) out->omegas[l] = L->omegas[l];
    for (int i = 0; i < 4; ++i) out->metrics[i] = L->metrics[i];
    for (int k = 0; k < c->K; ++k) {
        for (int j = 0; j < 7; ++j) out->coh[k].row[j] = L->row[k][j];
        for (int j = 0; j < 2; ++j) {
            out->coh[k].cg_iters[j] = L->cg_iters[k][j];
            out->coh[k].cg_info[j] = L->cg_info[k][j];
        }
        out->coh[k].spmm_passes = L->passes[k];
    }
    // keep the host copy of the prior in step (the stepwise entry points read it)
    c->prior.lam = L->lam;
    for (int l = 0; l < c->prior.L - 1; ++l) c->prior.omegas[l] = L->omegas[l];
    return 0;
}
