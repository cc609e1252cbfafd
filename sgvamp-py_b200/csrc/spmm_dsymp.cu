// Symmetric half-band SpMM, PERSISTENT form: one kernel per pass, no finish kernel, no ypart / tails round trip.
//
// Same arithmetic per stored value as spmm_dsym.cu (forward use from a sliding register window, transposed use as a
// systolic pipeline over the lanes of a warp, matrix stream through a per-warp bulk-copy ring).  What changes is the
// ownership of rows: a CTA does not own one tile but a contiguous RANGE of 128-row units (ticket order, balanced to
// one unit) and walks down it tile by tile:
//   * the transposed sums that reach past a tile (the `tails` of the one-tile kernel: Dp rows, 2/3 of a tile at w = 500)
//     stay in shared memory as a carry and are consumed by the CTA's own next tiles; a row is complete as soon as its
//     tile is combined, so the epilogue (q = gamw y + gam2 p, the four dot products of the fused CG step) runs right
//     there and nothing but q is written;
//   * the x window slides: only the TR new entries are staged per tile (the one-tile kernel re-stages TR + Dp entries
//     per tile, in CG mode three vector loads each);
//   * the bulk-copy ring runs ahead across tile boundaries, so the matrix stream never drains while a tile is being
//     combined (one stage per warp is lent to hold the forward sums during the combine and re-armed right after).
// Only the first Dp rows of a range need sums from the previous range: their partial sums are parked in `yhead`, the
// range's own carry-out goes to `tails[range]` with a release flag, and each CTA finishes its head rows at the very
// end, after an (almost always already satisfied) acquire on its predecessor's flag.  Ranges are handed out by an
// atomic ticket, so a CTA's predecessor is always running or finished: no co-residency assumption, no deadlock.
// Per-range dot-product partials are added in range order by the last CTA: results are bit-reproducible.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "dsym_common.cuh"

#ifndef DSP_UNROLL
#define DSP_UNROLL 1     // unroll factor of the group loop (tuning knob)
#endif
#ifndef DSP_L2HINT
#define DSP_L2HINT 1     // evict-first L2 policy on the matrix stream
#endif
// Measured and dropped (B200, M = 1M, w = 500, ms per pass isolated / in the CG loop; kept 0.383 / 0.416):
//   cp.async.bulk.prefetch.L2 4-16 groups ahead of the ring        0.426-0.444 / 0.464-0.515
//   group loop unrolled x2 / x4 (stage index still a run-time value) 0.391-0.408 / 0.448-0.450
//   fence.proxy.async before every refill (see DSP_ISSUE)           0.392-0.401 / 0.432-0.433
//   refill after the group's FMAs instead of right after its loads  0.383 / 0.425
#ifndef DSP_DEBUG
#define DSP_DEBUG 0      // 1: compile the phase clock of the whole-solve kernel in (development builds; costs registers)
#endif
#define DSP_STR2(x) #x
#define DSP_STR(x) DSP_STR2(x)

struct DsPersist {
    const float*        U;
    int                 Dp;       // stored diagonals (multiple of 4)
    int                 units;    // ldb / 128
    int                 upr;      // > 0: fixed units per range, the last range takes what is left; 0: even split over the grid
    int64_t             E;        // extension rows stored before the first own row
    double2*            yhead;    // [ranges][Dp] partial sums of the first Dp rows of a range
    double2*            tails;    // [ranges][Dp] carry-out of a range (sums for the Dp rows after it)
    unsigned long long* flags;    // [ranges] == epoch once tails[range] is complete
    unsigned*           ticket;   // range ticket (reset by the last CTA)
    unsigned long long  epoch;
    int                 epi;      // non-CG instantiation: EPI_Q or EPI_PLAIN
    int                 nph;      // phases of the drain step (1 when a segment spans >= 128 diagonals)
};

// Whole-solve mode (SOLVE): one cooperative launch runs every CG step of an LMMSE solve; step n reads the vector
// buffers (n+1)&1 - the neighbours' too - and writes n&1, exactly as the one-step launches do.
struct DsSolve {
    double2 *pp[2], *rr[2], *qq[2];                       // own arena buffers
    const double2 *ppL[2], *rrL[2], *qqL[2];              // left neighbour's (null: none)
    const double2 *ppR[2], *rrR[2], *qqR[2];              // right neighbour's
    unsigned long long* gen;                              // step barrier: == epoch + n + 1 once step n's state transition is done
    unsigned*           exit_ticket;
    int                 max_steps;
    unsigned long long* dbg;                              // SGV_DS_DEBUG: accumulated phase durations in ns (see DSP_T)
};

__device__ __forceinline__ unsigned long long dsp_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

static inline int dsp_per(int Dp, int S) { return (((Dp + S - 1) / S) + 3) & ~3; }

size_t sgv_dsymp_smem_bytes(int64_t w, int rw, int s, int nst) {
    const int Dp = (int)round_up(w + 1, 4);
    const int TR = 128 * rw, NW = rw * s;
    size_t b = (size_t)NW * nst * DS_STAGE_FLOATS * sizeof(float);         // per-warp rings
    b += (size_t)4 * dia_plane_len(TR + Dp) * sizeof(double2);              // x window
    b += (size_t)rw * (Dp + 128) * sizeof(double2);                         // transposed sums per row-warp
    b += (size_t)Dp * sizeof(double2);                                      // carry
    b += (size_t)NW * nst * 8;                                              // mbarriers
    b += (size_t)NW * 8 * sizeof(double) + (size_t)NW * sizeof(int) + 16;   // dot slots, lent stages, misc
    return b;
}

template <int RW, int S, int NST, bool CG, bool SOLVE>
__global__ void __launch_bounds__(32 * RW * S, 2)
k_dsym_persist(SpmmArgs a, DsPersist g, DsSolve sv) {
    SGV_LOAD_DEV_SCALARS(a);
    static_assert(S >= 4 && (NST & (NST - 1)) == 0, "one new window entry per thread; ring depth a power of two");
    static_assert(CG || !SOLVE, "whole-solve mode is the CG instantiation");
    if (!SOLVE && a.check_done && a.rc.st->done[0] && a.rc.st->done[1]) return;
    constexpr int TR = 128 * RW, NT = 32 * RW * S, NW = RW * S;
    constexpr int NV = CG ? 8 : 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Dp = g.Dp;
    const int W = TR + Dp;
    const int PL = dia_plane_len(W);
    const int AL = Dp + 128;
    float* ring = reinterpret_cast<float*>(smem_raw);
    double2* xw = reinterpret_cast<double2*>(smem_raw + (size_t)NW * NST * DS_STAGE_FLOATS * sizeof(float));
    double2* Aall = xw + 4 * PL;
    double2* carry = Aall + RW * AL;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(carry + Dp);
    double* sdot = reinterpret_cast<double*>(bars + NW * NST);
    int* s_lent = reinterpret_cast<int*>(sdot + NW * 8);
    int* s_misc = s_lent + NW;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rw = wid % RW, s = wid / RW;
    const int G = gridDim.x;
    if (tid == 0) s_misc[0] = (int)atomicAdd(g.ticket, 1u);
    const unsigned bar0 = smem_u32(bars + wid * NST);
    const unsigned ring0 = smem_u32(ring + (size_t)wid * NST * DS_STAGE_FLOATS);
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < NST; ++t) mbar_init(bar0 + 8 * t, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int j = tid; j < NW * 8; j += NT) sdot[j] = 0.0;
    __syncthreads();
    const int c = s_misc[0];                                   // this CTA's range (ticket order)
    // every range but the last spans >= Dp rows (the host chooses the split, see dsp_units_per_range)
    const int u0 = g.upr > 0 ? c * g.upr : (int)((int64_t)c * g.units / G);
    const int u1 = g.upr > 0 ? min(g.units, (c + 1) * g.upr) : (int)((int64_t)(c + 1) * g.units / G);
    const int nunits = u1 - u0;
    const int ntiles = (nunits + RW - 1) / RW;
    const int64_t rb = (int64_t)u0 * 128, re = (int64_t)u1 * 128;   // storage rows of the range

    // this warp's diagonals [d0, d1): segments are laid out from the TOP of the band, so that every d1 is a multiple of
    // `per` below Dp and the drain ranges [d1, d1 + 128) of different segments do not overlap when per >= 128
    const int per = (((Dp + S - 1) / S) + 3) & ~3;
    int d1 = Dp - (S - 1 - s) * per;
    if (d1 < 0) d1 = 0;
    const int d0 = max(0, d1 - per);
    const int ngw = (d1 - d0) >> 2;                             // groups of 4 diagonals per tile
    const int ntw = (ngw > 0 && nunits > rw) ? (nunits - rw + RW - 1) / RW : 0;   // tiles in which this warp has rows

    // ring walker (warp-uniform): next group to request
#if DSP_L2HINT
    const unsigned long long pol = l2_policy_evict_first();
#define DSP_COPY(DST, SRC, BYTES, BAR) bulk_g2s_hint(DST, SRC, BYTES, BAR, pol)
#else
#define DSP_COPY(DST, SRC, BYTES, BAR) bulk_g2s(DST, SRC, BYTES, BAR)
#endif
    int pf_left = 0, pf_gi = 0;
    const float* const pf_start = g.U + ((int64_t)(u0 + rw) * (Dp >> 2) + (d0 >> 2)) * DS_STAGE_FLOATS;
    const float* pf_ptr = pf_start;
    const int64_t unit_stride = (int64_t)RW * (Dp >> 2) * DS_STAGE_FLOATS;
#define DSP_ISSUE(STAGE, FENCE)                                                                                       \
    do {                                                                                                         \
        if (lane == 0) {                                                                                         \
            if (FENCE) fence_proxy_async_smem();                                                                 \
            mbar_expect_tx(bar0 + 8 * (STAGE), DS_STAGE_FLOATS * 4);                                             \
            DSP_COPY(ring0 + (STAGE) * DS_STAGE_FLOATS * 4, pf_ptr + (int64_t)pf_gi * DS_STAGE_FLOATS,           \
                     DS_STAGE_FLOATS * 4, bar0 + 8 * (STAGE));                                                   \
        }                                                                                                        \
        --pf_left;                                                                                               \
        if (++pf_gi == ngw) {                                                                                    \
            pf_gi = 0;                                                                                           \
            pf_ptr += unit_stride;                                                                               \
        }                                                                                                        \
    } while (0)
    int q = 0;                                                  // groups consumed so far (stage = q % NST)
    // start of a pass: the walker goes back to the range's first group; the first NST groups are requested at once
#define DSP_ARM_PASS()                                                                                           \
    do {                                                                                                         \
        pf_left = ntw * ngw;                                                                                     \
        pf_gi = 0;                                                                                               \
        pf_ptr = pf_start;                                                                                       \
        _Pragma("unroll") for (int t = 0; t < NST; ++t)                                                          \
            if (pf_left > 0) DSP_ISSUE((q + t) & (NST - 1), false);                                              \
    } while (0)
    bool prearmed = false;

    // CG scalars of the step (see spmm_dsym.cu)
    double al0 = 0.0, al1 = 0.0, beta0 = 0.0, beta1 = 0.0;
    bool first = true, fz0 = false, fz1 = false;
    const volatile CgState* vst = a.rc.st;                      // SOLVE: rewritten between steps by the last CTA
    unsigned long long epoch = g.epoch;
    // phase clock of the whole-solve kernel (SGV_DS_DEBUG=1; thread 0 of range 0 and of the last-arriving CTA):
    // dbg[0] stage window, [1] tiles, [2] head fix-up + partials, [3] wait for the slowest CTA, [4] local reduction,
    // [5] cross-rank exchange, [6] barrier release seen by range 0, [7] steps
    unsigned long long tp = 0;
#define DSP_T(SLOT)                                                                  \
    do {                                                                             \
        if (DSP_DEBUG && SOLVE && sv.dbg != nullptr && tid == 0 && c == 0) {                      \
            const unsigned long long tn = dsp_now();                                 \
            atomicAdd(sv.dbg + (SLOT), tn - tp);                                     \
            tp = tn;                                                                 \
        }                                                                            \
    } while (0)
  for (int step_i = 0;; ++step_i, ++epoch) {
    if (SOLVE) {                                                // the set-up kernel leaves step = 0: step n of the solve is
        const int n = step_i, prev = (n + 1) & 1, cur = n & 1;  // this loop's n-th turn, known without reading the state
        a.v = sv.pp[prev]; a.r = sv.rr[prev]; a.q = sv.qq[prev];
        a.p_new = sv.pp[cur]; a.r_new = sv.rr[cur]; a.out = sv.qq[cur];
        a.v_left = sv.ppL[prev]; a.r_left = sv.rrL[prev]; a.q_left = sv.qqL[prev];
        a.v_right = sv.ppR[prev]; a.r_right = sv.rrR[prev]; a.q_right = sv.qqR[prev];
        first = n == 0;
    }
    if (!prearmed) DSP_ARM_PASS();
    prearmed = false;
    if (DSP_DEBUG && SOLVE && sv.dbg != nullptr && tid == 0 && c == 0) tp = dsp_now();
    // Value of the input vector at local column `col` (0 = first own row of this rank): own memory, the left / right
    // neighbour's arena, or zero outside the matrix.  CG mode: the new direction p = r + beta p_old with the pending
    // update r -= alpha q applied on the fly; the rows this CTA owns get r, p and x += alpha p_old written.
    // Split into a load half and a compute / store half, so that a thread can put the loads of several entries in
    // flight before it touches any of them (the stores in between would otherwise serialise the round trips).
    struct Stg {
        double2 rv, po, qo, xv;
        int64_t col;
        int     kind;     // 0: outside the matrix, 1: loaded, 2: loaded + a row of this range (to be written)
    };
    auto stage_load = [&](int64_t col, Stg& t) {
        t.col = col;
        t.kind = 0;
        t.rv = t.po = t.qo = t.xv = make_double2(0.0, 0.0);
        const double2 *src = nullptr, *rsrc = nullptr, *qsrc = nullptr;
        int64_t idx = col;
        if (col >= 0 && col < a.M) {
            src = a.v;
            rsrc = a.r;
            qsrc = a.q;
        } else if (col < 0 && a.v_left != nullptr && a.n_left + col >= 0) {
            src = a.v_left;
            rsrc = a.r_left;
            qsrc = a.q_left;
            idx = a.n_left + col;
        } else if (col >= a.M && a.v_right != nullptr) {
            src = a.v_right;
            rsrc = a.r_right;
            qsrc = a.q_right;
            idx = col - a.M;
        }
        if (src == nullptr) return;
        t.kind = 1;
        if (!CG) {
            t.po = ld_vec2(src + idx);
            return;
        }
        t.rv = ld_vec2(rsrc + idx);
        if (!first) {
            t.po = ld_vec2(src + idx);
            t.qo = ld_vec2(qsrc + idx);
        }
        const int64_t js = col + g.E;
        if (col >= 0 && col < a.M && js >= rb && js < re) {     // rows of this range: written exactly once
            t.kind = 2;
            if (!first && (SOLVE || al0 != 0.0 || al1 != 0.0)) t.xv = ld_vec2(a.x + col);   // (whole-solve: alpha not read yet)
        }
    };
    auto stage_finish = [&](const Stg& t) -> double2 {
        if (t.kind == 0) return make_double2(0.0, 0.0);
        if (!CG) return t.po;
        double2 rv = t.rv;
        if (al0 != 0.0) rv.x -= al0 * t.qo.x;                   // scipy: r -= alpha*q
        if (al1 != 0.0) rv.y -= al1 * t.qo.y;
        double2 val;
        val.x = fz0 ? t.po.x : (first ? rv.x : t.po.x * beta0 + rv.x);   // scipy: p *= beta; p += r  (first step: p = r)
        val.y = fz1 ? t.po.y : (first ? rv.y : t.po.y * beta1 + rv.y);
        if (t.kind == 2) {
            a.r_new[t.col] = rv;
            a.p_new[t.col] = val;
            if (al0 != 0.0 || al1 != 0.0) {
                double2 xv = t.xv;
                if (al0 != 0.0) xv.x += al0 * t.po.x;           // scipy: x += alpha*p
                if (al1 != 0.0) xv.y += al1 * t.po.y;
                a.x[t.col] = xv;
            }
        }
        return val;
    };
    // first window [rb - E, rb - E + W) in local coordinates: up to 4 entries per thread, their loads issued back to
    // back; in whole-solve mode BEFORE the step's CG state is read (one memory round trip less per step)
    Stg st4[4];
    auto window_loads = [&]() {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int j = tid + m * NT;
            st4[m].kind = 0;
            if (j < W) stage_load(rb - g.E + j, st4[m]);
        }
    };
    if (SOLVE) window_loads();
    // the CG state of the step: ONE thread reads it and hands it round in shared memory (every thread of every CTA
    // reading the same cache line at the same moment costs ~10 us per step at the L2 slice that holds it)
    {
        double* s_st = reinterpret_cast<double*>(Aall);        // free between passes
        if (tid == 0) {
            s_st[0] = (double)vst->step;
            s_st[1] = (double)vst->done[0];
            s_st[2] = (double)vst->done[1];
            s_st[3] = vst->alpha[0];
            s_st[4] = vst->alpha[1];
            s_st[5] = vst->rho[0];
            s_st[6] = vst->rho_prev[0];
            s_st[7] = vst->rho[1];
            s_st[8] = vst->rho_prev[1];
        }
        __syncthreads();
        const int st_step = (int)s_st[0];
        fz0 = s_st[1] != 0.0;
        fz1 = s_st[2] != 0.0;
        al0 = al1 = beta0 = beta1 = 0.0;
        if (CG && st_step != 0) {
            al0 = s_st[3];
            al1 = s_st[4];
            beta0 = s_st[5] / s_st[6];
            beta1 = s_st[7] / s_st[8];
        }
        if (!SOLVE) first = st_step == 0;
        __syncthreads();                                        // the A region is written again below
    }
    if (SOLVE && ((fz0 && fz1) || step_i >= sv.max_steps)) {
        prearmed = true;                                        // the ring holds this step's first requests: drained below
        break;
    }
    if (!SOLVE) window_loads();
    // a finished row: fused epilogue, dot products into d[]
    double d[NV];
#pragma unroll
    for (int k2 = 0; k2 < NV; ++k2) d[k2] = 0.0;
    auto finish_row = [&](int64_t i, double2 y, double2 vi, double2 rv) {   // rv: r_new[i] (CG mode)
        double2 o;
        o.x = a.gamw * y.x + a.gam2 * vi.x;
        o.y = a.gamw * y.y + a.gam2 * vi.y;
        a.out[i] = o;
        if constexpr (CG) {
            d[0] += vi.x * o.x; d[1] += vi.y * o.y;      // p.q
            d[2] += rv.x * o.x; d[3] += rv.y * o.y;      // r.q
            d[4] += o.x * o.x;  d[5] += o.y * o.y;       // q.q
            d[6] += rv.x * rv.x; d[7] += rv.y * rv.y;    // r.r
        } else {
            d[0] += vi.x * o.x;
            d[1] += vi.y * o.y;
        }
    };
    auto flush_dots = [&]() {                                  // warp sums into the warp's slot (fixed order over tiles)
#pragma unroll
        for (int k2 = 0; k2 < NV; ++k2) {
            const double v = warp_sum(d[k2]);
            if (lane == 0) sdot[wid * 8 + k2] += v;
            d[k2] = 0.0;
        }
    };

    // the window goes into shared memory; carry and the untouched tail of the A ranges start at zero
    {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int j = tid + m * NT;
            if (j < 4 * PL) xw[(j & 3) * PL + (j >> 2)] = j < W ? stage_finish(st4[m]) : make_double2(0.0, 0.0);
        }
        for (int j = tid + 4 * NT; j < 4 * PL; j += NT) xw[(j & 3) * PL + (j >> 2)] = make_double2(0.0, 0.0);
    }
    for (int j = tid; j < Dp; j += NT) carry[j] = make_double2(0.0, 0.0);
    for (int j = tid; j < RW * 128; j += NT) Aall[(j >> 7) * AL + Dp + (j & 127)] = make_double2(0.0, 0.0);
    for (int j = tid; j < NW * 8; j += NT) sdot[j] = 0.0;
    __syncthreads();
    DSP_T(0);

    const int g4 = rw * 32 + lane;                              // this thread's rows: 4*g4 .. 4*g4+3 of the tile
    double2* Arw = Aall + rw * AL;
    const double hm = lane == 31 ? 0.0 : 1.0;                   // lane 31 has no lane above: its incoming sums are zero
    for (int k = 0; k < ntiles; ++k) {
        const int64_t r0s = rb + (int64_t)k * TR;               // storage row of the tile's first row
        const bool active = ngw > 0 && k * RW + rw < nunits;
        const bool last = k == ntiles - 1;
        double2 acc0 = make_double2(0.0, 0.0), acc1 = acc0, acc2 = acc0, acc3 = acc0;
        double2 T0 = acc0, T1 = acc0, T2 = acc0, T3 = acc0;
        int lent = -1;
        if (active) {
            double2 T4 = acc0, T5 = acc0, T6 = acc0;
            const double2 O0 = xw[g4], O1 = xw[PL + g4], O2 = xw[2 * PL + g4], O3 = xw[3 * PL + g4];
            int xi = g4 + (d0 >> 2);
            double2 X0 = xw[xi], X1 = xw[PL + xi], X2 = xw[2 * PL + xi], X3 = xw[3 * PL + xi];
            _Pragma(DSP_STR(unroll DSP_UNROLL))
            for (int gi = 0; gi < ngw; ++gi) {
                const int stage = q & (NST - 1);
                mbar_wait(bar0 + 8 * stage, (unsigned)(q / NST) & 1u);
                const float4* sp = reinterpret_cast<const float4*>(ring + ((size_t)wid * NST + stage) * DS_STAGE_FLOATS) + lane;
                const float4 c0 = sp[0], c1 = sp[32], c2 = sp[64], c3 = sp[96];
                __syncwarp();
                // hand the stage straight back to the copy engine - except the tile's last one, which holds the
                // forward sums during the combine and is re-armed after it
                // (the lanes' LDS reads of the stage are ordered before lane 0's refill by __syncwarp, the same
                // consumer-release -> producer hand-off CUTLASS' TMA pipelines use; a fence.proxy.async here costs 4 %
                // of the pass - MEMBAR.ALL.CTA in SASS - and is kept only where generic WRITES precede a refill)
                if (gi + 1 < ngw) {
                    if (pf_left > 0) DSP_ISSUE(stage, false);
                } else {
                    lent = stage;
                }
                ++q;
                const double2 N0 = xw[xi + 1], N1 = xw[PL + xi + 1], N2 = xw[2 * PL + xi + 1], N3 = xw[3 * PL + xi + 1];
                DS_FWD(c0, X0, X1, X2, X3);
                DS_TRN(c0, T0, T1, T2, T3);
                DS_FWD(c1, X1, X2, X3, N0);
                DS_TRN(c1, T1, T2, T3, T4);
                DS_FWD(c2, X2, X3, N0, N1);
                DS_TRN(c2, T2, T3, T4, T5);
                DS_FWD(c3, X3, N0, N1, N2);
                DS_TRN(c3, T3, T4, T5, T6);
                X0 = N0; X1 = N1; X2 = N2; X3 = N3;
                ++xi;
                if (lane == 0) {                                 // lane 0's four sums are final for the warp
                    double2* e = Arw + d0 + 4 * gi;
                    e[0] = T0; e[1] = T1; e[2] = T2; e[3] = T3;
                }
                const double2 I0 = shfl_down1(T0), I1 = shfl_down1(T1), I2 = shfl_down1(T2), I3 = shfl_down1(T3);
                T0 = make_double2(fma(I0.x, hm, T4.x), fma(I0.y, hm, T4.y));
                T1 = make_double2(fma(I1.x, hm, T5.x), fma(I1.y, hm, T5.y));
                T2 = make_double2(fma(I2.x, hm, T6.x), fma(I2.y, hm, T6.y));
                T3 = make_double2(I3.x * hm, I3.y * hm);
                T4 = T5 = T6 = make_double2(0.0, 0.0);
            }
            // forward sums of the warp's 128 rows into the lent ring stage (2 KB = 128 x double2)
            double2* f = reinterpret_cast<double2*>(ring + ((size_t)wid * NST + lent) * DS_STAGE_FLOATS) + 4 * lane;
            f[0] = acc0; f[1] = acc1; f[2] = acc2; f[3] = acc3;
        }
        if (lane == 0) s_lent[wid] = lent;
        // the TR new window entries of the next tile: loads issued now, consumed after the combine
        double2 nx = make_double2(0.0, 0.0);
        if (!last && tid < TR) {
            Stg t1;
            stage_load(r0s + TR - g.E + Dp + tid, t1);
            nx = stage_finish(t1);
        }
        // r of the row this thread finishes in the combine (row tid of the tile): written when the row was staged, at
        // least one barrier ago; requested here so that its latency hides behind the barriers below
        double2 rpre = make_double2(0.0, 0.0);
        if (CG && tid < TR) {
            const int64_t i = r0s + tid - g.E;
            if (i >= 0 && i < a.M && r0s + tid < re) rpre = a.r_new[i];
        }
        __syncthreads();                                        // S1: forward sums and lane-0 sums of every warp are in place
        // drain: every lane holds finished sums for targets d1 + 4*lane + {0..3} (relative to the warp's first row)
        for (int ph = 0; ph < g.nph; ++ph) {
            if (active && (s % g.nph) == ph) {
                double2* e = Arw + d1 + 4 * lane;
                e[0].x += T0.x; e[0].y += T0.y;
                e[1].x += T1.x; e[1].y += T1.y;
                e[2].x += T2.x; e[2].y += T2.y;
                e[3].x += T3.x; e[3].y += T3.y;
            }
            __syncthreads();                                    // S2
        }
        // combine, in fixed order: carry-in + forward sums of the S segments + the row-warps' transposed sums
        const int t_end = last ? (int)(re - r0s) : TR;          // rows of this tile that belong to the range
        const int act_rw = min(RW, nunits - k * RW);
        double2 keep[4], xs[3];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int t = tid + m * NT;
            keep[m] = make_double2(0.0, 0.0);
            if (t < W) {
                double2 sum = t < Dp ? carry[t] : make_double2(0.0, 0.0);
                if (t < TR && (t >> 7) < act_rw) {
#pragma unroll
                    for (int s2 = 0; s2 < S; ++s2) {
                        const int w2 = s2 * RW + (t >> 7);
                        const int st2 = s_lent[w2];
                        if (st2 >= 0) {
                            const double2 v = reinterpret_cast<const double2*>(ring + ((size_t)w2 * NST + st2) * DS_STAGE_FLOATS)[t & 127];
                            sum.x += v.x;
                            sum.y += v.y;
                        }
                    }
                }
#pragma unroll
                for (int rw2 = 0; rw2 < RW; ++rw2) {
                    const int rel = t - 128 * rw2;
                    if (rw2 < act_rw && rel >= 0 && rel < AL) {
                        const double2 v = Aall[rw2 * AL + rel];
                        sum.x += v.x;
                        sum.y += v.y;
                    }
                }
                if (t < t_end) {                                // a row of this range
                    const int64_t js = r0s + t, i = js - g.E;
                    if (c > 0 && js - rb < Dp) g.yhead[(int64_t)c * Dp + (js - rb)] = sum;   // still lacks the previous range's carry
                    else if (i >= 0 && i < a.M) finish_row(i, sum, xw[(t & 3) * PL + (t >> 2)], rpre);   // t < TR: m == 0, t == tid
                } else if (!last) {
                    keep[m] = sum;                              // carry for the next tiles (index t - TR)
                } else if (t - t_end < Dp) {
                    g.tails[(int64_t)c * Dp + (t - t_end)] = sum;   // carry-out of the range
                }
            }
        }
        if (!last) {
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int j = tid + m * NT;
                xs[m] = j < Dp ? xw[((j + TR) & 3) * PL + ((j + TR) >> 2)] : make_double2(0.0, 0.0);
            }
        }
        flush_dots();
        __syncthreads();                                        // S3: carry, window, A and the lent stages have been read
        if (!last) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int t = tid + m * NT;
                if (t >= TR && t < W) carry[t - TR] = keep[m];
            }
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int j = tid + m * NT;
                if (j < Dp) xw[(j & 3) * PL + (j >> 2)] = xs[m];
            }
            if (tid < TR) {
                const int j = Dp + tid;
                xw[(j & 3) * PL + (j >> 2)] = nx;
            }
            for (int j = tid; j < RW * 128; j += NT) Aall[(j >> 7) * AL + Dp + (j & 127)] = make_double2(0.0, 0.0);
        }
        if (lent >= 0 && pf_left > 0) DSP_ISSUE(lent, true);    // re-arm the lent stage (held generic writes: proxy fence)
        __syncthreads();                                        // S4
    }

    DSP_T(1);
    // carry-out published; head rows: add the previous range's carry-out and finish them
    if (tid == 0 && c + 1 < G) {
        __threadfence();
        st_release_gpu_u64(g.flags + c, epoch);
    }
    if (c > 0) {
        if (tid == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu_u64(g.flags + (c - 1)) != epoch) {
                if (clock64() - t0 > 8000000000LL) {            // ~4 s: cannot happen unless a CTA died; never hang
                    a.rc.st->error = 1 << 30;
                    break;
                }
            }
        }
        __syncthreads();
        for (int t = tid; t < Dp; t += NT) {
            const int64_t js = rb + t, i = js - g.E;
            if (js < re && i >= 0 && i < a.M) {
                double2 y = g.yhead[(int64_t)c * Dp + t];
                const double2 tl = ld_vec2(g.tails + (int64_t)(c - 1) * Dp + t);
                y.x += tl.x;
                y.y += tl.y;
                finish_row(i, y, CG ? a.p_new[i] : ld_vec2(a.v + i), CG ? a.r_new[i] : make_double2(0.0, 0.0));
            }
        }
        flush_dots();
    }
    __syncthreads();
    if (tid < NV) {
        double tot = 0.0;
        for (int w2 = 0; w2 < NW; ++w2) tot += sdot[w2 * 8 + tid];
        a.rc.partials[(size_t)c * NV + tid] = tot;
    }
    if (SOLVE) {                                                // the matrix does not change: the next step's first
        DSP_ARM_PASS();                                         // requests go out before the step barrier
        prearmed = true;
    }
    __syncthreads();
    DSP_T(2);
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(a.rc.counter, 1u);
        s_misc[1] = (t == (unsigned)G - 1u);
    }
    __syncthreads();
    unsigned long long tl = 0;
    if (DSP_DEBUG && SOLVE && sv.dbg != nullptr && tid == 0 && s_misc[1]) {
        tl = dsp_now();
        if (c == 0) { atomicAdd(sv.dbg + 3, tl - tp); tp = tl; }
    }
    if (s_misc[1]) {
        // last CTA of the pass: add the per-range partials in range order; state transition / cross-rank exchange
        __threadfence();
        if (tid == 0) {
            *a.rc.counter = 0u;
            *g.ticket = 0u;                                     // every CTA has taken its range by now
        }
        if (CG || g.epi != EPI_PLAIN) {
            // number of the cross-rank exchange that follows: loaded now, in the shadow of the partial sums
            const unsigned long long xseq = a.rc.world > 1 ? red_next_seq(a.rc) : 0ull;
            double acc[NV];
#pragma unroll
            for (int k2 = 0; k2 < NV; ++k2) acc[k2] = 0.0;
            for (int b = tid; b < G; b += NT) {
#pragma unroll
                for (int k2 = 0; k2 < NV; ++k2) acc[k2] += __ldcg(&a.rc.partials[(size_t)b * NV + k2]);
            }
            double* red = reinterpret_cast<double*>(Aall);
            block_reduce<NV>(acc, red);
            const RedCtx& rcs = a.rc;
            if (DSP_DEBUG && SOLVE && sv.dbg != nullptr && tid == 0) {
                const unsigned long long tn = dsp_now();
                atomicAdd(sv.dbg + 4, tn - tl);
                tl = tn;
            }
            if (tid == 0 && rcs.world == 1) apply_totals(rcs.ap, rcs.st, acc);
            if (rcs.world > 1 && tid < 32) {
                publish_warp<NV>(acc, rcs, tid, xseq);
                if (rcs.inline_resolve) {
                    __syncwarp();
                    resolve_warp(rcs, tid, xseq);
                }
            }
        }
        if (SOLVE && tid == 0) {
            __threadfence();
            if (DSP_DEBUG && sv.dbg != nullptr) {
                atomicAdd(sv.dbg + 5, dsp_now() - tl);
                atomicAdd(sv.dbg + 7, 1ull);
            }
            st_release_gpu_u64(sv.gen, epoch + 1);              // step barrier: the new CG state is in place
        }
    }
    if (!SOLVE) break;
    if (!s_misc[1] && tid == 0) {
        const long long t0 = clock64();
        while (ld_acquire_gpu_u64(sv.gen) != epoch + 1) {
            if (clock64() - t0 > 40000000000LL) {               // ~20 s (a cross-rank wait inside is bounded by the same)
                a.rc.st->error = 1 << 29;
                break;
            }
        }
    }
    __syncthreads();
    if (DSP_DEBUG && SOLVE && sv.dbg != nullptr && tid == 0 && c == 0) {
        const unsigned long long tn = dsp_now();
        atomicAdd(sv.dbg + (s_misc[1] ? 6 : 3), s_misc[1] ? 0ull : tn - tp);   // range 0 waiting = slowest CTA + reduction + exchange
        tp = tn;
    }
  }
#undef DSP_T
#undef DSP_ARM_PASS
#undef DSP_ISSUE
#undef DSP_COPY
    if (SOLVE) {
        // requests made ahead for a step that does not happen must land before the CTA's shared memory is released
        if (prearmed) {
            const int n_armed = min(NST, ntw * ngw);
            for (int t = 0; t < n_armed; ++t) mbar_wait(bar0 + 8 * ((q + t) & (NST - 1)), (unsigned)((q + t) / NST) & 1u);
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned t = atomicAdd(sv.exit_ticket, 1u);
            if (t == (unsigned)G - 1u) {                        // also covers a solve that was over before its first step
                *sv.exit_ticket = 0u;
                *g.ticket = 0u;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define DSP_RW 2
#define DSP_S 4
#define DSP_NST 4
#define DSP_SMEM_LIMIT (113 * 1024)   // two CTAs per SM

static int dsp_nph(int64_t w) {
    const int Dp = (int)round_up(w + 1, 4);
    const int per = dsp_per(Dp, DSP_S);
    return per >= 128 ? 1 : (128 + per - 1) / per;
}

// The persistent kernel is used when its shared memory fits two CTAs per SM and the drain step needs at most 4
// phases (w >= ~125); narrower or much wider bands keep the one-tile kernels.  SGV_DS_PERSIST=0 disables it (A/B).
bool sgv_dsymp_feasible(int64_t w) {
    static const bool off = getenv("SGV_DS_PERSIST") != nullptr && atoi(getenv("SGV_DS_PERSIST")) == 0;
    if (off) return false;
    return sgv_dsymp_smem_bytes(w, DSP_RW, DSP_S, DSP_NST) <= DSP_SMEM_LIMIT && dsp_nph(w) <= 4;
}

static int g_dsp_coop_blocks = 0;   // co-resident CTAs per SM of the whole-solve instantiation (0: not usable)

int sgv_preload_dsymp() {
    SGV_CUDA(cudaFuncSetAttribute(k_dsym_persist<DSP_RW, DSP_S, DSP_NST, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  DSP_SMEM_LIMIT));
    SGV_CUDA(cudaFuncSetAttribute(k_dsym_persist<DSP_RW, DSP_S, DSP_NST, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  DSP_SMEM_LIMIT));
    SGV_CUDA(cudaFuncSetAttribute(k_dsym_persist<DSP_RW, DSP_S, DSP_NST, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  DSP_SMEM_LIMIT));
    int nb = 0;
    SGV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_dsym_persist<DSP_RW, DSP_S, DSP_NST, true, true>,
                                                           32 * DSP_RW * DSP_S, DSP_SMEM_LIMIT));
    g_dsp_coop_blocks = nb;
    return 0;
}

// Row ranges.  A range must span at least Dp rows (min_units 128-row units), so that its carry-out ends inside the next
// one.  Enough rows for two ranges per SM: the units are split evenly over 2 x SMs ranges (sizes differ by one unit).
// Fewer rows (a shard of a multi-GPU partition): ranges of exactly min_units units and a LAST range that takes the
// remainder - nothing follows it, so it may be shorter - which keeps the makespan at min_units units where an even split
// over fewer ranges would hand one of them an extra unit (+25 % at M = 125k, w = 500).
static int dsp_units_per_range(const sgv_ctx* c, const LdMatrix& ld) {
    const int64_t Dp = round_up(ld.w + 1, 4);
    const int64_t units = ld.ldb / 128, min_units = (Dp + 127) / 128, slots = (int64_t)c->sm_count * 2;
    return units >= slots * min_units ? 0 : (int)min_units;
}
static int dsp_ranges(const sgv_ctx* c, const LdMatrix& ld) {
    const int64_t units = ld.ldb / 128, upr = dsp_units_per_range(c, ld);
    if (upr == 0) return c->sm_count * 2;
    return (int)std::max<int64_t>(1, (units + upr - 1) / upr);
}

int sgv_dsymp_ensure_scratch(sgv_ctx* c, const LdMatrix& ld) {
    const int64_t Dp = round_up(ld.w + 1, 4);
    const int64_t need = (int64_t)c->sm_count * 2 * Dp;
    if (c->dsp_cap < need) {
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        if (c->dsp_yhead) cudaFree(c->dsp_yhead);
        if (c->dsp_tails) cudaFree(c->dsp_tails);
        c->dsp_yhead = c->dsp_tails = nullptr;
        c->dsp_cap = 0;
        SGV_CUDA(cudaMalloc(&c->dsp_yhead, need * sizeof(double2)));
        SGV_CUDA(cudaMalloc(&c->dsp_tails, need * sizeof(double2)));
        c->dsp_cap = need;
    }
    if (c->dsp_dbg == nullptr && getenv("SGV_DS_DEBUG") != nullptr) {
        SGV_CUDA(cudaMalloc(&c->dsp_dbg, 16 * sizeof(unsigned long long)));
        SGV_CUDA(cudaMemset(c->dsp_dbg, 0, 16 * sizeof(unsigned long long)));
    }
    if (c->dsp_flags == nullptr) {   // one hand-off flag per range + the step barrier word of the whole-solve kernel
        SGV_CUDA(cudaMalloc(&c->dsp_flags, ((size_t)c->sm_count * 2 + 2) * sizeof(unsigned long long)));
        SGV_CUDA(cudaMemset(c->dsp_flags, 0, ((size_t)c->sm_count * 2 + 2) * sizeof(unsigned long long)));
    }
    return sgv_ensure_partials(c, (int64_t)c->sm_count * 2);
}

static void dsp_fill(sgv_ctx* c, const LdMatrix& ld, int epi, DsPersist& g) {
    g.U = ld.band;
    g.Dp = (int)round_up(ld.w + 1, 4);
    g.units = (int)(ld.ldb / 128);
    g.upr = dsp_units_per_range(c, ld);
    g.E = ld.ext;
    g.yhead = c->dsp_yhead;
    g.tails = c->dsp_tails;
    g.flags = c->dsp_flags;
    g.ticket = c->counter + 12;
    g.epi = epi;
    g.nph = dsp_nph(ld.w);
}

int sgv_launch_dsymp(sgv_ctx* c, const LdMatrix& ld, int epi, SpmmArgs& a) {
    SGV_CHECK(c->dsp_yhead != nullptr && c->dsp_flags != nullptr, "DSYM persistent scratch not allocated");
    SGV_CHECK(epi == EPI_CG || epi == EPI_Q || epi == EPI_PLAIN, "epilogue %d not available in the persistent kernel", epi);
    DsPersist g;
    dsp_fill(c, ld, epi, g);
    g.epoch = ++c->dsp_epoch;
    DsSolve sv;
    memset(&sv, 0, sizeof(sv));
    const int G = dsp_ranges(c, ld);
    SGV_TRY(sgv_ensure_partials(c, G));
    a.rc.partials = c->partials;
    a.rc.counter = c->counter;
    const size_t smem = sgv_dsymp_smem_bytes(ld.w, DSP_RW, DSP_S, DSP_NST);
    if (epi == EPI_CG)
        k_dsym_persist<DSP_RW, DSP_S, DSP_NST, true, false><<<G, 32 * DSP_RW * DSP_S, smem, c->stream>>>(a, g, sv);
    else
        k_dsym_persist<DSP_RW, DSP_S, DSP_NST, false, false><<<G, 32 * DSP_RW * DSP_S, smem, c->stream>>>(a, g, sv);
    c->launches++;
    return 0;
}

// Whole CG solve in one cooperative launch (the state must have been initialised by the set-up kernel).  Usable when
// every range CTA is co-resident and no rank shares its GPU with another (the step barrier spans the cross-rank
// exchange).  SGV_DS_SOLVE=0 falls back to one launch per step.
bool sgv_dsymp_solve_usable(const sgv_ctx* c, const LdMatrix& ld) {
    static const bool off = getenv("SGV_DS_SOLVE") != nullptr && atoi(getenv("SGV_DS_SOLVE")) == 0;
    if (off || !c->coop_ok || (c->world > 1 && c->host_barrier)) return false;
    return ld.layout == SGV_LAYOUT_DSYM && sgv_dsymp_feasible(ld.w) &&
           dsp_ranges(c, ld) <= g_dsp_coop_blocks * c->sm_count;
}

int sgv_launch_dsymp_solve(sgv_ctx* c, const LdMatrix& ld, SpmmArgs& a, int max_steps) {
    SGV_CHECK(c->dsp_yhead != nullptr && c->dsp_flags != nullptr, "DSYM persistent scratch not allocated");
    DsPersist g;
    dsp_fill(c, ld, EPI_CG, g);
    g.epoch = c->dsp_epoch + 1;
    c->dsp_epoch += (unsigned long long)max_steps + 1;           // step n uses epoch + n (flags) / epoch + n + 1 (barrier)
    DsSolve sv;
    memset(&sv, 0, sizeof(sv));
    for (int i = 0; i < 2; ++i) {
        sv.pp[i] = c->pp[i];
        sv.rr[i] = c->rr2[i];
        sv.qq[i] = c->qq2[i];
    }
    if (c->world > 1 && c->halo) {
        for (int side = 0; side < 2; ++side) {
            const int q = side == 0 ? c->rank - 1 : c->rank + 1;
            if (q < 0 || q >= c->world) continue;
            const PeerView& pv = c->peer[q];
            SGV_CHECK(pv.base != nullptr && pv.Ml >= ld.w, "neighbour %d not attached or shorter than the half-bandwidth", q);
            for (int i = 0; i < 2; ++i) {
                const double2* p = reinterpret_cast<double2*>(pv.base + arena_off_pp(pv.Ml, i));
                const double2* r = reinterpret_cast<double2*>(pv.base + arena_off_rr(pv.Ml, i));
                const double2* qv = reinterpret_cast<double2*>(pv.base + arena_off_qq(pv.Ml, i));
                if (side == 0) { sv.ppL[i] = p; sv.rrL[i] = r; sv.qqL[i] = qv; }
                else { sv.ppR[i] = p; sv.rrR[i] = r; sv.qqR[i] = qv; }
            }
            if (side == 0) a.n_left = pv.Ml;
        }
    }
    sv.gen = c->dsp_flags + (size_t)c->sm_count * 2;
    sv.exit_ticket = c->counter + 13;
    sv.max_steps = max_steps;
    sv.dbg = c->dsp_dbg;
    const int G = dsp_ranges(c, ld);
    SGV_TRY(sgv_ensure_partials(c, G));
    a.rc.partials = c->partials;
    a.rc.counter = c->counter;
    size_t smem = sgv_dsymp_smem_bytes(ld.w, DSP_RW, DSP_S, DSP_NST);
    void* args[] = {&a, &g, &sv};
    SGV_CUDA(cudaLaunchCooperativeKernel((const void*)k_dsym_persist<DSP_RW, DSP_S, DSP_NST, true, true>, dim3(G),
                                         dim3(32 * DSP_RW * DSP_S), args, smem, c->stream));
    c->launches++;
    return 0;
}
