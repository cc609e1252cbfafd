// On-GPU LD construction: banded R = X^T X and r = X^T y from int8 genotypes, written straight into
// the tiled symmetric half-band layout the SpMM kernels stream (sgv_ld_adopt_dsym) - no dense M x M product, no host pass.
//
// Reference recipe (simulation/sim_gen_phen_mult.py:39-55): X = genotypes {0,1,2}, column-standardised with the
// population standard deviation, divided by sqrt(N); R = X^T X, r = X^T y.  With integer genotypes g,
//     S_ij = sum_n g_ni g_nj   (exact int32),   mu_i = sum_n g_ni / N,   sd_i = sqrt(S_ii / N - mu_i^2),
//     R_ij = (S_ij - N mu_i mu_j) / (N sd_i sd_j),        r_j = (sum_n g_nj y_n - mu_j sum_n y_n) / (sd_j sqrt(N)),
// so the whole contraction is an integer Gram product restricted to the band |i - j| <= w; the standardisation, the
// optional Bartlett taper 1 - |i-j|/(w+1) (keeps a truncated band positive semi-definite), Rused = (1-s) R + s I
// (src/main.py:265) and the fp32 rounding happen once, in the epilogue, from exact integers in fp64.
//
// Genotypes are marker-major (one row of N samples per marker, as in a PLINK .bed), int8, rows padded with zeros to a
// multiple of 16 bytes.  Kernel: a CTA owns 64 markers (rows i) x 64 markers (columns j) of the band and sweeps the
// samples in chunks of 256 through shared memory; a thread accumulates a 4 x 4 block of S with IDP4A (4 samples per
// instruction), operands fetched as 128-bit shared-memory loads (8 LDS.128 per 64 IDP4A).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "sgv_device.cuh"

#define LB_T 64        // markers per tile side
#define LB_KC 256      // samples per shared-memory chunk (bytes per marker row in the chunk)
#define LB_PAD 16      // row padding in shared memory (bytes): rows 272 B apart -> conflict-free 128-bit loads

__global__ void __launch_bounds__(256)
k_ld_col_stats(const int8_t* __restrict__ G, int64_t ldg, int64_t nmark, int64_t N, double* __restrict__ mu, double* __restrict__ sd) {
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= nmark) return;
    const int* row = reinterpret_cast<const int*>(G + j * ldg);
    int s1 = 0, s2 = 0;
    for (int64_t k = lane; k < ldg / 4; k += 32) {
        const int v = row[k];
        s1 = __dp4a(v, 0x01010101, s1);
        s2 = __dp4a(v, v, s2);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
        const double m = (double)s1 / (double)N;
        mu[j] = m;
        sd[j] = sqrt(fmax((double)s2 / (double)N - m * m, 0.0));
    }
}

// G: markers [g0, g0 + nmark) of the matrix (global indices), row jb of the buffer = marker g0 + jb.
// Output rows: storage row t <-> global marker row_lo - E + t, t in [0, rows_st); diagonals d in [0, w].
__global__ void __launch_bounds__(256)
k_ld_band_gram(const int8_t* __restrict__ G, int64_t ldg, int64_t g0, int64_t nmark, int64_t N, const double* __restrict__ mu,
               const double* __restrict__ sd, float* __restrict__ U, int64_t ngr, int64_t w, int64_t row_lo, int64_t E,
               int64_t rows_st, int64_t M, double s, int taper) {
    __shared__ __align__(16) int8_t sa[LB_T][LB_KC + LB_PAD];
    __shared__ __align__(16) int8_t sb[LB_T][LB_KC + LB_PAD];
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tj = tid & 15;                       // 16 x 16 threads, 4 x 4 outputs each
    const int64_t t0 = (int64_t)blockIdx.x * LB_T;                // first storage row of the tile
    const int64_t i0 = row_lo - E + t0;                           // its global marker
    const int64_t j0 = i0 + (int64_t)blockIdx.y * LB_T;           // first column marker of the tile
    if (j0 - (i0 + LB_T - 1) > w) return;                         // tile entirely outside the band
    int acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0;
    for (int64_t k0 = 0; k0 < ldg; k0 += LB_KC) {
        const int kc = (int)min((int64_t)LB_KC, ldg - k0);        // multiple of 16
        // stage 64 rows of A and of B: 16-byte pieces, zero outside the buffer / beyond kc
        for (int p = tid; p < LB_T * (LB_KC / 16); p += 256) {
            const int r = p / (LB_KC / 16), c16 = p % (LB_KC / 16);
            int4 va = make_int4(0, 0, 0, 0), vb = va;
            if (c16 * 16 < kc) {
                const int64_t ia = i0 + r - g0, jb = j0 + r - g0;
                if (ia >= 0 && ia < nmark) va = *reinterpret_cast<const int4*>(G + ia * ldg + k0 + c16 * 16);
                if (jb >= 0 && jb < nmark) vb = *reinterpret_cast<const int4*>(G + jb * ldg + k0 + c16 * 16);
            }
            *reinterpret_cast<int4*>(&sa[r][c16 * 16]) = va;
            *reinterpret_cast<int4*>(&sb[r][c16 * 16]) = vb;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < LB_KC; kk += 16) {
            int4 a4[4], b4[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) a4[a] = *reinterpret_cast<const int4*>(&sa[ti + 16 * a][kk]);
#pragma unroll
            for (int b = 0; b < 4; ++b) b4[b] = *reinterpret_cast<const int4*>(&sb[tj + 16 * b][kk]);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    acc[a][b] = __dp4a(a4[a].x, b4[b].x, acc[a][b]);
                    acc[a][b] = __dp4a(a4[a].y, b4[b].y, acc[a][b]);
                    acc[a][b] = __dp4a(a4[a].z, b4[b].z, acc[a][b]);
                    acc[a][b] = __dp4a(a4[a].w, b4[b].w, acc[a][b]);
                }
        }
        __syncthreads();
    }
    // epilogue: standardise, taper, regularise, round once, scatter into the tiled half band
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t t = t0 + ti + 16 * a, i = row_lo - E + t;   // storage row / global marker
        if (t >= rows_st || i < 0 || i >= M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t j = j0 + tj + 16 * b, d = j - i;
            if (d < 0 || d > w || j >= M) continue;
            if (t < E && j < row_lo) continue;                    // extension rows keep only their couplings to own rows
            const double mi = mu[i - g0], mj = mu[j - g0], si = sd[i - g0], sj = sd[j - g0];
            double r = 0.0;
            if (si > 0.0 && sj > 0.0) r = ((double)acc[a][b] - (double)N * mi * mj) / ((double)N * si * sj);
            if (d == 0) r = 1.0;                                   // unit diagonal (a monomorphic marker is left uncoupled)
            if (taper) r *= 1.0 - (double)d / (double)(w + 1);
            r = (1.0 - s) * r + (d == 0 ? s : 0.0);
            if (d == 0) r *= 0.5;                                  // the layout stores half of the diagonal
            U[sgv_dsym_index(t, d, ngr)] = (float)r;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same Gram tile on the 5th-generation tensor cores: tcgen05.mma kind::i8 (int8 x int8 -> int32, exact), accumulator
// in tensor memory.  A CTA of 4 warps owns a 128 x 128 tile of S; per chunk of 128 samples its threads copy the 128 A rows
// and 128 B rows (128 bytes each, K-major) into shared memory in the canonical 128-byte-swizzled UMMA layout (16-byte chunk
// c of row r at chunk c ^ (r & 7) of the 1 KB atom that holds rows 8*(r/8) .. +7), one elected thread issues four
// M128 x N128 x K32 MMAs and commits them to the stage's mbarrier.  Two operand stages: while the tensor core works on one,
// the threads store the next chunk (already in registers, fetched one chunk ahead) into the other and fetch the chunk after
// it; a stage is refilled once the commit of the chunk that used it has arrived.  Epilogue: tcgen05.ld 32 lanes x 32 columns per warp, then the same standardisation / taper / scatter as above.
// ---------------------------------------------------------------------------------------------
#define TC_T 128
__device__ __forceinline__ unsigned tc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long tc_desc_k_sw128(unsigned smem_addr) {
    // K-major, SWIZZLE_128B: start address >> 4, LBO = 16 B (1), SBO = 1024 B between 8-row groups (64), version 1 (sm_100),
    // layout type 2 at bits 61..63
    return (unsigned long long)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128)
k_ld_band_gram_tc(const int8_t* __restrict__ G, int64_t ldg, int64_t g0, int64_t nmark, int64_t N, const double* __restrict__ mu,
                  const double* __restrict__ sd, float* __restrict__ U, int64_t ngr, int64_t w, int64_t row_lo, int64_t E,
                  int64_t rows_st, int64_t M, double s, int taper) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
    // two operand stages of 32 KB (A: 128 rows x 128 B, then B): the tensor core works on one while the threads fill the other
    __shared__ __align__(8) unsigned long long mbar[2];
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    // column-tile offset fastest: the CTAs that share an A tile (and neighbouring B tiles) are resident together and meet in L2
    const int64_t t0 = (int64_t)blockIdx.y * TC_T;
    const int64_t i0 = row_lo - E + t0;
    const int64_t j0 = i0 + (int64_t)blockIdx.x * TC_T;
    if (j0 - (i0 + TC_T - 1) > w) return;
    const unsigned bar0 = tc_smem_u32(&mbar[0]);
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tc_smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = tmem_base_s;
    // instruction descriptor: D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(TC_T >> 3) << 17) | ((unsigned)(TC_T >> 4) << 24);
    // Operand fill: 8 consecutive lanes read the 8 16-byte chunks of one 128-byte row segment (a warp instruction covers 4
    // whole lines), thread tid handles chunk tid % 8 of the rows tid / 8 + 16 * q, q = 0..7, of A and of B
    const int crow = tid >> 3, cchunk = tid & 7;
    unsigned vmask_a = 0, vmask_b = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int64_t ia = i0 + crow + 16 * q - g0, jb = j0 + crow + 16 * q - g0;
        if (ia >= 0 && ia < nmark) vmask_a |= 1u << q;
        if (jb >= 0 && jb < nmark) vmask_b |= 1u << q;
    }
    const int8_t* pa = G + (i0 + crow - g0) * ldg + cchunk * 16;
    const int8_t* pb = G + (j0 + crow - g0) * ldg + cchunk * 16;
    const int64_t rstep = 16 * ldg;
    int4 ra[8], rb[8];                                               // the next chunk's operands, in flight from global memory
    auto fetch = [&](int64_t k0) {
        const bool kin = k0 + cchunk * 16 < ldg;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            ra[q] = make_int4(0, 0, 0, 0);
            rb[q] = ra[q];
            if (kin && ((vmask_a >> q) & 1u)) ra[q] = __ldg(reinterpret_cast<const int4*>(pa + q * rstep + k0));
            if (kin && ((vmask_b >> q) & 1u)) rb[q] = __ldg(reinterpret_cast<const int4*>(pb + q * rstep + k0));
        }
    };
    auto wait_stage = [&](unsigned bar, unsigned parity) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "TC_WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
            "@P1 bra TC_DONE_%=;\n\t"
            "bra TC_WAIT_%=;\n\t"
            "TC_DONE_%=:\n\t"
            "}" ::"r"(bar), "r"(parity)
            : "memory");
    };
    unsigned phase[2] = {0u, 0u};
    int nchunk = 0;
    fetch(0);
    for (int64_t k0 = 0; k0 < ldg; k0 += 128, ++nchunk) {
        const int st = nchunk & 1;
        unsigned char* sA = base + st * 32768;
        unsigned char* sB = sA + 16384;
        if (nchunk >= 2) {                                            // the MMAs of chunk n-2 have consumed this stage
            wait_stage(bar0 + 8 * st, phase[st]);
            phase[st] ^= 1u;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = crow + 16 * q;                                 // r & 7 == crow & 7
            const int off = r * 128 + ((cchunk ^ (r & 7)) << 4);
            *reinterpret_cast<int4*>(sA + off) = ra[q];
            *reinterpret_cast<int4*>(sB + off) = rb[q];
        }
        if (k0 + 128 < ldg) fetch(k0 + 128);                          // overlaps the MMAs issued below
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core (async proxy) reads
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned long long da = tc_desc_k_sw128(tc_smem_u32(sA)), db = tc_desc_k_sw128(tc_smem_u32(sB));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {                            // K = 32 bytes per MMA: + 2 in the 16-byte start address
                const unsigned acc = (nchunk > 0 || kk > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
                    "}\n" ::"r"(tmem_d), "l"(da + 2ull * kk), "l"(db + 2ull * kk), "r"(idesc), "r"(acc)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8 * st) : "memory");
        }
    }
    // the commit of the last chunk covers every MMA issued before it: D is complete
    {
        const int st = (nchunk - 1) & 1;
        wait_stage(bar0 + 8 * st, phase[st]);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes 32w .. 32w+31 (= rows of the tile), 32 columns at a time
    const int64_t t = t0 + tid, i = row_lo - E + t;
    const bool row_ok = t < rows_st && i >= 0 && i < M;
    const double mi = row_ok ? mu[i - g0] : 0.0, si = row_ok ? sd[i - g0] : 0.0;
    for (int c0 = 0; c0 < TC_T; c0 += 32) {
        unsigned v[32];
        const unsigned taddr = tmem_d + ((unsigned)(32 * wid) << 16) + (unsigned)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
            "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row_ok) {
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const int64_t j = j0 + c0 + b, d = j - i;
                if (d < 0 || d > w || j >= M) continue;
                if (t < E && j < row_lo) continue;
                const double mj = mu[j - g0], sj = sd[j - g0];
                double r = 0.0;
                if (si > 0.0 && sj > 0.0) r = ((double)(int)v[b] - (double)N * mi * mj) / ((double)N * si * sj);
                if (d == 0) r = 1.0;
                if (taper) r *= 1.0 - (double)d / (double)(w + 1);
                r = (1.0 - s) * r + (d == 0 ? s : 0.0);
                if (d == 0) r *= 0.5;
                U[sgv_dsym_index(t, d, ngr)] = (float)r;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_d) : "memory");
}

// r_j = (sum_n g_nj y_n - mu_j sum_n y_n) / (sd_j sqrt(N)) for the own markers
__global__ void __launch_bounds__(256)
k_ld_xty(const int8_t* __restrict__ G, int64_t ldg, int64_t g0, int64_t N, const double* __restrict__ mu, const double* __restrict__ sd,
         const double* __restrict__ y, double ysum, int64_t row_lo, int64_t rows, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= rows) return;
    const int64_t jb = row_lo + t - g0;
    const int8_t* row = G + jb * ldg;
    double acc = 0.0;
    for (int64_t n = lane; n < N; n += 32) acc += (double)row[n] * y[n];
    acc = warp_sum(acc);
    if (lane == 0) out[t] = sd[jb] > 0.0 ? (acc - mu[jb] * ysum) / (sd[jb] * sqrt((double)N)) : 0.0;
}

extern "C" int sgv_ld_build_banded(sgv_handle c, int cohort, const int8_t* G, int on_device, int64_t g0, int64_t nmark,
                                   int64_t N, int64_t ldg, int64_t w, double s, int taper, const double* y, double* xty_out) {
    SGV_CHECK(c != nullptr && c->M > 0, "handle not configured");
    SGV_CHECK(cohort >= 0 && cohort < c->K, "cohort index %d out of range [0,%d)", cohort, c->K);
    SGV_CUDA(cudaSetDevice(c->device));
    SGV_CHECK(G != nullptr && N > 0 && ldg >= N && ldg % 16 == 0, "genotype rows must be padded to a multiple of 16 bytes (ldg >= N)");
    SGV_CHECK(w >= 0 && w < c->M && sgv_dsym_feasible(w), "half-bandwidth %lld not supported by the half-band layout", (long long)w);
    SGV_CHECK(N < (1LL << 29), "N too large for exact int32 Gram sums of genotypes in {0,1,2}");
    SGV_CHECK(c->world == 1 || c->halo, "banded LD construction needs a halo partition when sharded");
    const int64_t E = sgv_dsym_ext(c, w), Ml = c->Ml, row_lo = c->row_lo, M = c->M;
    const int64_t need_lo = std::max<int64_t>(0, row_lo - E), need_hi = std::min<int64_t>(M, row_lo + Ml + w);
    SGV_CHECK(g0 <= need_lo && g0 + nmark >= need_hi, "genotype window [%lld,%lld) must cover markers [%lld,%lld) (own rows, "
              "the extension rows before them and w markers after them)", (long long)g0, (long long)(g0 + nmark),
              (long long)need_lo, (long long)need_hi);
    const int8_t* dG = G;
    int8_t* owned = nullptr;
    if (!on_device) {
        SGV_CUDA(cudaMalloc(&owned, (size_t)nmark * ldg));
        SGV_CUDA(cudaMemcpyAsync(owned, G, (size_t)nmark * ldg, cudaMemcpyHostToDevice, c->stream));
        dG = owned;
    }
    double *mu = nullptr, *sd = nullptr, *dy = nullptr, *dout = nullptr;
    int rc = 0;
    LdMatrix& ld = c->coh[cohort].ld;
    do {
        if (cudaMalloc(&mu, 2 * (size_t)nmark * sizeof(double)) != cudaSuccess) { sgv_set_error("out of device memory"); rc = -2; break; }
        sd = mu + nmark;
        k_ld_col_stats<<<(unsigned)((nmark * 32 + 255) / 256), 256, 0, c->stream>>>(dG, ldg, nmark, N, mu, sd);
        c->launches++;
        sgv_ld_free(ld);
        const int64_t Dp = round_up(w + 1, 4), ngr = Dp / 4, ldb = round_up(Ml + E, 128);
        float* U = nullptr;
        if (cudaMalloc(&U, (size_t)Dp * ldb * sizeof(float)) != cudaSuccess) { sgv_set_error("out of device memory"); rc = -2; break; }
        ld.band = U;
        ld.owned = true;
        ld.w = w;
        ld.ldb = ldb;
        ld.ext = E;
        ld.nnz_stored = (w + 1) * Ml;
        cudaMemsetAsync(U, 0, (size_t)Dp * ldb * sizeof(float), c->stream);
        // tensor-core Gram tiles (tcgen05, kind::i8) by default; SGV_LD_DP4A=1 selects the IDP4A kernel (A/B, cross-check)
        const bool use_dp4a = getenv("SGV_LD_DP4A") != nullptr && atoi(getenv("SGV_LD_DP4A")) != 0;
        if (use_dp4a) {
            const dim3 grid((unsigned)((Ml + E + LB_T - 1) / LB_T), (unsigned)((w + LB_T - 1) / LB_T + 1));
            k_ld_band_gram<<<grid, 256, 0, c->stream>>>(dG, ldg, g0, nmark, N, mu, sd, U, ngr, w, row_lo, E, Ml + E, M, s, taper);
        } else {
            const dim3 grid((unsigned)((w + TC_T - 1) / TC_T + 1), (unsigned)((Ml + E + TC_T - 1) / TC_T));
            if (grid.y > 65535u) { sgv_set_error("too many marker tiles for one launch"); rc = -1; break; }
            if (cudaFuncSetAttribute(k_ld_band_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024) != cudaSuccess) {
                sgv_set_error("cannot raise the dynamic shared memory limit of the LD construction kernel");
                rc = -2;
                break;
            }
            k_ld_band_gram_tc<<<grid, 128, 65536 + 1024, c->stream>>>(dG, ldg, g0, nmark, N, mu, sd, U, ngr, w, row_lo, E, Ml + E, M, s,
                                                                      taper);
        }
        c->launches++;
        ld.layout = SGV_LAYOUT_DSYM;
        if ((rc = sgv_dsym_ensure_scratch(c, ld)) != 0) break;
        if (y != nullptr && xty_out != nullptr) {
            if (cudaMalloc(&dy, (size_t)(N + Ml) * sizeof(double)) != cudaSuccess) { sgv_set_error("out of device memory"); rc = -2; break; }
            dout = dy + N;
            cudaMemcpyAsync(dy, y, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, c->stream);
            double ysum = 0.0;
            for (int64_t n = 0; n < N; ++n) ysum += y[n];
            k_ld_xty<<<(unsigned)((Ml * 32 + 255) / 256), 256, 0, c->stream>>>(dG, ldg, g0, N, mu, sd, dy, ysum, row_lo, Ml, dout);
            c->launches++;
            cudaMemcpyAsync(xty_out, dout, (size_t)Ml * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            sgv_set_error("LD construction kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = -2;
        }
    } while (0);
    cudaFree(mu);
    cudaFree(dy);
    cudaFree(owned);
    if (rc != 0) sgv_ld_free(ld);
    return rc;
}

// Copy a cohort's half band (tiled layout, Dp x ldb floats) into a caller's device buffer: lets one constructed matrix
// be adopted by several handles (benchmarks build once and run several solvers).
extern "C" int sgv_ld_copy_band(sgv_handle c, int cohort, float* dst_dev, int64_t nfloats, int64_t* w, int64_t* ldb, int64_t* ext) {
    SGV_CHECK(c != nullptr && cohort >= 0 && cohort < c->K, "bad handle / cohort");
    SGV_CUDA(cudaSetDevice(c->device));
    const LdMatrix& ld = c->coh[cohort].ld;
    SGV_CHECK(ld.layout == SGV_LAYOUT_DSYM, "cohort %d does not hold a half band", cohort);
    const int64_t need = round_up(ld.w + 1, 4) * ld.ldb;
    if (w) *w = ld.w;
    if (ldb) *ldb = ld.ldb;
    if (ext) *ext = ld.ext;
    if (dst_dev == nullptr) return 0;                             // size query
    SGV_CHECK(nfloats >= need, "destination holds %lld floats, %lld needed", (long long)nfloats, (long long)need);
    SGV_CUDA(cudaMemcpyAsync(dst_dev, ld.band, (size_t)need * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
