// Device-side helpers: streaming loads, deterministic block / grid / cross-rank reductions.
#pragma once
#include "sgv_internal.cuh"

__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldg_stream_i1(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
// vectors that a peer GPU (or another CTA's earlier kernel) wrote: bypass the non-coherent path
__device__ __forceinline__ double2 ld_vec2(const double2* p) {
    double2 r;
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

#define SGV_INF __longlong_as_double(0x7ff0000000000000LL)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum NV values over the block in a fixed order.  Result valid in warp 0.
// `red` is shared scratch of at least NV*32 doubles.  Contains __syncthreads().
template <int NV, bool MIN = false>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = MIN ? warp_min(v[k]) : warp_sum(v[k]);
    __syncthreads();   // protect `red` from a previous use
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[k * 32 + wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = (lane < nw) ? red[k * 32 + lane] : (MIN ? SGV_INF : 0.0);
            v[k] = MIN ? warp_min(x) : warp_sum(x);
        }
    }
}

// scipy.sparse.linalg.cg loop-top test for one column, applied right after rho = r.r is known.
// (scipy 1.18.1 _isolve/iterative.py: `if np.linalg.norm(r) < atol: return x, 0`, atol = 1e-5*|b|;
//  loop exhaustion returns info = maxiter without a final test.)
__device__ __forceinline__ void cg_top_test(CgState* s, int c) {
    if (s->done[c]) return;
    if (s->iters[c] >= s->maxit) {
        s->done[c] = 1;
        s->info[c] = s->maxit;
        return;
    }
    if (sqrt(s->rho[c]) < 1e-5 * sqrt(s->bnorm2[c])) {
        s->done[c] = 1;
        s->info[c] = 0;
    }
}

// gamw / gam2 of an SpMM launched from the fused VAMP iteration live on the device (first statement of every kernel
// that applies them)
#define SGV_LOAD_DEV_SCALARS(a)                                   \
    do {                                                          \
        if ((a).vs_cohort >= 0) {                                 \
            (a).gamw = (a).rc.st->vs.gamw[(a).vs_cohort];         \
            (a).gam2 = (a).rc.st->vs.gam2[(a).vs_cohort];         \
        }                                                         \
    } while (0)

// The state transition that follows a completed reduction.  Executed by exactly one thread per
// rank: the finaliser of the reducing kernel (world == 1) or the resolve kernel (world > 1).
__device__ __forceinline__ void apply_totals(const ApplyArgs& ap, CgState* s, const double* t) {
    switch (ap.kind) {
        case AP_STATS:
            for (int k = 0; k < ap.nv; ++k) s->stats[ap.off + k] = t[k];
            break;
        case AP_PQ:
            s->pq[0] = t[0];
            s->pq[1] = t[1];
            break;
        case AP_RESID:   // r = b - A x0 done: rho and the loop-top test of iteration 0
            for (int c = 0; c < 2; ++c) {
                if (!s->done[c]) {
                    s->rho[c] = t[c];
                    cg_top_test(s, c);
                }
            }
            break;
        case AP_SETUP:   // |b|^2 known: initialise the CG state of both columns
            s->maxit = ap.maxit;
            s->step = 0;
            s->alpha[0] = s->alpha[1] = 0.0;
            s->ambig[0] = s->ambig[1] = 0;
            for (int c = 0; c < 2; ++c) {
                s->bnorm2[c] = t[c];
                s->rho[c] = 0.0;
                s->rho_prev[c] = 0.0;
                s->pq[c] = 0.0;
                s->iters[c] = 0;
                s->info[c] = 0;
                s->done[c] = 0;
                s->zero_b[c] = 0;
                if (t[c] == 0.0) {          // scipy: `if bnrm2 == 0: return b, 0`
                    s->done[c] = 1;
                    s->zero_b[c] = 1;
                } else {                    // r = b.copy() or b - A x0 (formed by the set-up kernel); loop-top test of iteration 0
                    s->rho[c] = ap.x0_zero ? t[c] : t[2 + c];
                    cg_top_test(s, c);
                }
            }
            break;
        case AP_EM: {   // one EM pass done (src/sgvamp.py:134-136) + the convergence test of the driver loop (:254-256)
            EmState& e = s->em;
            if (e.done) break;
            double wsum = 0.0;
            for (int q = 0; q < e.K; ++q) wsum += e.a[q] * t[q];
            const double lam_new = wsum / e.asum / e.Mtot;
            double om_new[SGV_MAX_L], dn = 0.0, on = 0.0;
            for (int l = 0; l < e.Lm1; ++l) {
                om_new[l] = t[8 + l] / t[15];
                dn += (om_new[l] - e.omegas[l]) * (om_new[l] - e.omegas[l]);
                on += e.omegas[l] * e.omegas[l];
            }
            const double om_err = sqrt(dn) / sqrt(on);
            const double lam_err = fabs(lam_new - e.lam) / lam_new;
            e.lam = lam_new;
            for (int l = 0; l < e.Lm1; ++l) e.omegas[l] = om_new[l];
            e.steps += 1;
            e.relerr = fmax(om_err, lam_err);
            if ((om_err < e.tol && lam_err < e.tol) || e.steps >= e.maxit) e.done = 1;
            break;
        }
        case AP_CGFUSED:   // t = [p.q, r.q, q.q, r.r] x 2 columns of the step just multiplied
            // alpha = rho/(p.q); the update x += alpha p, r -= alpha q is applied when the next kernel stages
            // its window (or by the post kernel), but its effect on rho is known now:
            //     |r - alpha q|^2 = r.r - 2 alpha r.q + alpha^2 q.q
            // with r.r the exactly summed norm of the residual this step used (no drift accumulates).
            // scipy sums r.r of the updated residual; the extrapolation cancels, so its error is bounded by
            // ~eps (|r| + alpha |q|)^2.  When the extrapolated value lies within that band of the stopping threshold the
            // decision is postponed (ambig): the step is taken tentatively and the NEXT pass - which sums r.r of the
            // updated residual exactly, as scipy does - decides; if the column had in fact converged, the tentative
            // step is discarded (alpha = 0, no count).  Iteration counts therefore follow the summed norm always.
            for (int c = 0; c < 2; ++c) {
                if (s->done[c]) {
                    s->alpha[c] = 0.0;
                    continue;
                }
                const double rr = t[6 + c];
                if (s->ambig[c]) {          // postponed loop-top test of this step, now with the exact sum
                    s->ambig[c] = 0;
                    s->rho[c] = rr;
                    if (sqrt(rr) < 1e-5 * sqrt(s->bnorm2[c])) {
                        s->done[c] = 1;
                        s->info[c] = 0;
                        s->alpha[c] = 0.0;  // the tentative step is void: nothing pending
                        continue;
                    }
                }
                const double al = rr / t[c];
                double rn = rr - 2.0 * al * t[2 + c] + al * al * t[4 + c];
                if (rn < 0.0) rn = 0.0;
                s->alpha[c] = al;
                s->rho_prev[c] = rr;
                s->rho[c] = rn;
                s->iters[c] += 1;
                cg_top_test(s, c);
                if (s->iters[c] < s->maxit) {
                    const double sr = sqrt(rr) + fabs(al) * sqrt(t[4 + c]);
                    const double band = s->band_eps * sr * sr;                     // ~1e3 eps (|r| + alpha |q|)^2
                    const double thr = 1e-10 * s->bnorm2[c];
                    if (fabs(rn - thr) <= band) {
                        s->ambig[c] = 1;
                        s->done[c] = 0;
                    }
                }
            }
            s->step += 1;
            break;
        // ---- device-resident scalar chain of the VAMP loop.  Every operation is a separately rounded IEEE operation in the
        // reference's order (no FMA contraction), so the values equal what the Python floats of src/sgvamp.py would hold.
        case AP_DENOISE: {   // t[0] = sum_j d(j): alpha1 (:285-291) and gam2 (:305) of every cohort
            VampScal& v = s->vs;
            const double dmean = __ddiv_rn(t[0], ap.Mtot);
            s->log.dmean = dmean;
            for (int k = 0; k < v.K; ++k) {
                double a1 = __dmul_rn(__dmul_rn(v.a[k], v.gam1[k]), dmean);
                if (ap.it > 0) a1 = __dadd_rn(__dmul_rn(ap.rho, a1), __dmul_rn(__dsub_rn(1.0, ap.rho), v.alpha1[k]));
                v.alpha1[k] = a1;
                v.gam2[k] = __ddiv_rn(__dmul_rn(v.gam1[k], __dsub_rn(1.0, a1)), a1);
            }
            s->log.lam = s->em.lam;
            for (int l = 0; l < v.Lm1; ++l) s->log.omegas[l] = s->em.omegas[l];
            s->log.em_steps = s->em.steps;
            s->log.em_relerr = s->em.relerr;
            break;
        }
        case AP_POST: {      // t = [u.Sigma2u, xhat2.r, xhat2^T R xhat2, u^T R Sigma2u] of cohort ap.cohort
            VampScal& v = s->vs;
            const int k = ap.cohort;
            const double gam2 = v.gam2[k];
            double a2 = __ddiv_rn(__dmul_rn(gam2, t[0]), ap.Mtot);                                       // :338-340
            if (ap.lmmse_damp) a2 = __dadd_rn(__dmul_rn(ap.rho, a2), __dmul_rn(__dsub_rn(1.0, ap.rho), v.alpha2[k]));   // :345-346
            v.alpha2[k] = a2;
            const double gam1 = __ddiv_rn(__dmul_rn(gam2, __dsub_rn(1.0, a2)), a2);                      // :347
            v.gam1[k] = gam1;
            double gw = v.gamw[k];
            if (ap.learn_gamw) {                                                                         // :350-364
                const double N = v.N[k];
                double z = __dadd_rn(__dsub_rn(N, __dmul_rn(2.0, t[1])), t[2]);
                if (z < 0.0) z = 0.0;
                gw = __ddiv_rn(1.0, __dadd_rn(__ddiv_rn(z, N), __ddiv_rn(t[3], N)));
            }
            gw = fmax(gw, 1.0);                                                                          // :374
            v.gamw[k] = gw;
            IterLog& L = s->log;
            L.row[k][0] = (double)ap.it; L.row[k][1] = gw; L.row[k][2] = gam1; L.row[k][3] = gam2;
            L.row[k][4] = v.alpha1[k]; L.row[k][5] = a2; L.row[k][6] = s->em.lam;
            for (int cidx = 0; cidx < 2; ++cidx) {
                L.cg_iters[k][cidx] = s->iters[cidx];
                L.cg_info[k][cidx] = s->done[cidx] ? s->info[cidx] : s->maxit;   // never passed the test: scipy reports maxiter
            }
            L.passes[k] = s->iters[0] > s->iters[1] ? s->iters[0] : s->iters[1];
            L.error = s->error;
            break;
        }
        case AP_METRICS:
            for (int k = 0; k < 4; ++k) s->log.metrics[k] = t[k];
            s->log.error = s->error;
            break;
        case AP_CGUPDATE:   // x, r updated: rho_prev <- rho, rho <- r.r, count, loop-top test
            for (int c = 0; c < 2; ++c) {
                if (s->done[c]) continue;
                s->rho_prev[c] = s->rho[c];
                s->rho[c] = t[c];
                s->iters[c] += 1;
                cg_top_test(s, c);
            }
            s->step += 1;
            break;
    }
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_entry(InboxEntry* p, double v, unsigned long long seq) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned f = (unsigned)seq;
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)bits), "r"(f),
                 "r"((unsigned)(bits >> 32)), "r"(f)
                 : "memory");
}
__device__ __forceinline__ bool ld_entry(const InboxEntry* p, unsigned long long seq, double& v) {
    unsigned lo, f1, hi, f2;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(p) : "memory");
    v = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    const unsigned f = (unsigned)seq;
    return f1 == f && f2 == f;
}

// Sequence number of the cross-rank exchange in flight: one more than the number of exchanges this rank has
// COMPLETED, counted on the device (RedCtx::pubseq).  Launches that exit early (CG / EM loop already converged) do
// not exchange and do not advance it, so the inbox slot index (seq % SGV_INBOX_SLOTS) only moves with real
// exchanges: every rank completes exchange s before any rank can publish s + 1, hence no rank is ever more than one
// slot ahead of a peer that is still reading.  All ranks run the same sequence of exchanges (the skip decisions
// depend on flags that are bit-identical everywhere), so the counters agree without being communicated.
__device__ __forceinline__ unsigned long long red_next_seq(const RedCtx& rc) {
    return *reinterpret_cast<volatile unsigned long long*>(rc.pubseq) + 1ull;
}

// Wait (bounded) until every rank's partial sums of the exchange in flight are in this rank's inbox, add the
// rows in rank order (bit-identical on all ranks) and apply the state transition.  Called by one full
// warp: lane q collects rank q's row.  On time-out the error flag is raised and the CG / EM loops are
// marked done, so that nothing hangs.
__device__ __forceinline__ void resolve_warp(const RedCtx& rc, int lane, unsigned long long seq) {
    Inbox* me = rc.inbox[rc.rank];
    const int slot = (int)(seq % SGV_INBOX_SLOTS);
    const int nv = rc.ap.nv;
    double mine[SGV_MAX_PARTIAL_VALUES];
    bool good = true;
    // phase 1: all world*nv entries are polled in parallel, 32 at a time (a serial poll costs one L2 round
    // trip per entry: ~12 us for the 16 sums of an EM pass)
    {
        const int total = rc.world * nv;
        const long long t0 = clock64();
        for (int e = lane; e < total; e += 32) {
            const InboxEntry* p = &me->e[slot][e / nv][e % nv];
            double dummy;
            while (!ld_entry(p, seq, dummy)) {
                if (clock64() - t0 > 40000000000LL) {   // ~20 s
                    good = false;
                    break;
                }
            }
        }
    }
    __syncwarp();
    // phase 2: lane q reads rank q's row (all present now; independent loads)
    if (lane < rc.world) {
#pragma unroll
        for (int k = 0; k < SGV_MAX_PARTIAL_VALUES; ++k) {
            mine[k] = 0.0;
            if (k < nv) ld_entry(&me->e[slot][lane][k], seq, mine[k]);
        }
    }
    const unsigned bad = __ballot_sync(0xffffffffu, !good);
    double t[SGV_MAX_PARTIAL_VALUES];
    for (int k = 0; k < nv; ++k) {
        double acc = rc.ap.is_min ? SGV_INF : 0.0;
        const double v = lane < rc.world ? mine[k] : 0.0;
        for (int q = 0; q < rc.world; ++q) {           // rank order
            const double x = __shfl_sync(0xffffffffu, v, q);
            acc = rc.ap.is_min ? fmin(acc, x) : acc + x;
        }
        t[k] = acc;
    }
    if (lane == 0) {
        *rc.pubseq = seq;                               // this exchange is complete on this rank
        if (bad) {
            if (!rc.st->error) {
                rc.st->error = (int)bad;
                rc.st->err_seq = seq;
            }
            rc.st->done[0] = rc.st->done[1] = 1;
            rc.st->em.done = 1;
            return;
        }
        apply_totals(rc.ap, rc.st, t);
    }
}

// Write this rank's NV partial sums of reduction `seq` into every rank's inbox; called by a full warp, lane e
// handles entry e = (peer, value) so that the peer stores are issued in parallel.
template <int NV>
__device__ __forceinline__ void publish_warp(const double (&acc)[NV], const RedCtx& rc, int lane, unsigned long long seq) {
    const int slot = (int)(seq % SGV_INBOX_SLOTS);
    __threadfence();
    for (int e = lane; e < rc.world * NV; e += 32) {
        const int q = e / NV, k = e % NV;
        double val = 0.0;
#pragma unroll
        for (int kk = 0; kk < NV; ++kk)
            if (kk == k) val = acc[kk];
        st_entry(&rc.inbox[q]->e[slot][rc.rank][k], val, seq);
    }
}

// Grid-wide deterministic reduction: every block contributes NV values; the block that takes the
// last ticket combines the per-block partials in index order.  world == 1: it applies the state
// transition directly.  world > 1: it publishes this rank's totals into every rank's inbox (peer
// stores) followed by the sequence flag; the resolve kernel launched next combines the rows.
// `partials` must hold NV * (number of blocks) doubles and *counter must be 0 on entry.
template <int NV, bool MIN = false>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], const RedCtx& rc, double* red) {
    __shared__ int s_last;
    const unsigned nblk = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    block_reduce<NV, MIN>(v, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) rc.partials[(size_t)bid * NV + k] = v[k];
        __threadfence();
        unsigned t = atomicAdd(rc.counter, 1u);
        s_last = (t == nblk - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = MIN ? SGV_INF : 0.0;
    for (unsigned b = threadIdx.x; b < nblk; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = __ldcg(&rc.partials[(size_t)b * NV + k]);
            acc[k] = MIN ? fmin(acc[k], x) : acc[k] + x;
        }
    }
    block_reduce<NV, MIN>(acc, red);          // totals valid in every lane of warp 0
    if (threadIdx.x == 0) {
        *rc.counter = 0u;
        if (rc.world == 1) apply_totals(rc.ap, rc.st, acc);
    }
    if (rc.world > 1 && threadIdx.x < 32) {
        // This rank's vector writes of the kernel (read by the neighbours as halos once they have seen these
        // entries) are already ordered: every block fenced at GPU scope before taking its ticket, and peers read
        // this memory through this GPU's L2.  The world x NV entries are written by the 32 lanes in parallel: a
        // single thread issuing them one after the other costs ~0.25 us per peer store (16 us at 8 ranks x 8 sums).
        const unsigned long long seq = red_next_seq(rc);     // read by every lane before lane 0 advances the counter
        publish_warp<NV>(acc, rc, threadIdx.x, seq);
        // one GPU per rank: this block completes the cross-rank reduction itself (the kernel ends when every
        // rank has published, which is also the ordering point for the halo reads of the next kernel)
        if (rc.inline_resolve) {
            __syncwarp();
            resolve_warp(rc, threadIdx.x, seq);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epi_row(const SpmmArgs& a, int64_t i, double2 acc, double2 vi, double (&dots)[2]) {
    double2 o;
    o.x = a.gamw * acc.x + a.gam2 * vi.x;
    o.y = a.gamw * acc.y + a.gam2 * vi.y;
    if (EPI == EPI_Q) {
        a.out[i] = o;
        dots[0] += vi.x * o.x;
        dots[1] += vi.y * o.y;
    } else if (EPI == EPI_RESID) {
        double2 b = a.bb[i];
        double2 r = make_double2(b.x - o.x, b.y - o.y);
        a.out[i] = r;
        dots[0] += r.x * r.x;
        dots[1] += r.y * r.y;
    } else if (EPI == EPI_STATS) {
        double2 b = a.bb[i];
        dots[0] += vi.x * o.x;   // xhat2^T R xhat2
        dots[1] += b.y * o.y;    // u^T R Sigma2_u
    } else {
        a.out[i] = o;
    }
}


// length of one of the 4 (index mod 4) planes of a shared-memory vector window of W entries; odd
// multiple-of-8 padding keeps the planes on different banks
__host__ __device__ inline int dia_plane_len(int W) {
    int pl = (W + 3) / 4 + 1;
    while ((pl & 7) != 1) ++pl;
    return pl;
}

