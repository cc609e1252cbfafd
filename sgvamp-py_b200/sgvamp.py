"""sgVAMP solver, B200-native.  Drop-in for the reference module ``src/sgvamp.py``.

Same class name, constructor and ``infer`` signature, same output files
(``{out}_xhat_it_{it}.bin``, ``{out}_r1_cohort_{k}_it_{it}.bin``, ``{out}_cohort_{k}.csv``,
``{out}_metrics.csv``; reference src/sgvamp.py:33-76).  The numeric hot path - denoiser,
EM/MLE reductions, the two conjugate-gradient solves of the LMMSE step, the Hutchinson probe and
the gamw statistics - runs in hand-written sm_100a CUDA kernels behind the C ABI of
``libsgvamp_b200.so`` (``include/sgvamp_b200.h``).  Scalars (gam1, gam2, alpha1, alpha2, gamw)
stay on the host in fp64 exactly where the reference keeps them; M-vectors never leave the GPU
except for the per-iteration output dumps.  There is no CPU fallback.

Differences from the reference that a caller can see:
  * one process drives all K cohorts (``R``/``r``/``N`` may be length-K lists, ``comm`` may be
    None); the reference's rank-per-cohort mode is kept when ``comm.Get_size() == K > 1``;
  * extra keyword-only arguments of ``infer`` (``s``, ``probes``, ``layout``, ``write_outputs``).
"""
from __future__ import annotations

import csv
import logging
import os
import queue
import threading
import time

import numpy as np
import scipy.optimize
import scipy.sparse

import sgv_native as nat
import shard as shd


class DeviceDIA:
    """LD already resident in HBM in diagonal-major band layout (see sgv_ld_adopt_dia)."""

    def __init__(self, ptr, w, ldb, keepalive=None):
        self.ptr, self.w, self.ldb, self.keepalive = int(ptr), int(w), int(ldb), keepalive


class DeviceDSYM:
    """Symmetric LD already resident in HBM as a half band (upper diagonals 0..w; see sgv_ld_adopt_dsym)."""

    def __init__(self, ptr, w, ldb, ext=0, keepalive=None):
        self.ptr, self.w, self.ldb, self.ext, self.keepalive = int(ptr), int(w), int(ldb), int(ext), keepalive


class DiaWindow:
    """Column window of a matrix in scipy's DIA format: data[k, j - col0] = R[j - offsets[k], j] for the
    columns j in [col0, col0 + data.shape[1]) of an M x M matrix (what one rank of a row partition needs)."""

    def __init__(self, data, offsets, col0, M):
        self.data, self.offsets, self.col0, self.M = data, np.asarray(offsets), int(col0), int(M)


class DeviceDense:
    """Dense fp32 row-major LD already resident in HBM (see sgv_ld_adopt_dense)."""

    def __init__(self, ptr, ld, keepalive=None):
        self.ptr, self.ld, self.keepalive = int(ptr), int(ld), keepalive


class DeviceDenseCols:
    """This rank's slice of a dense LD partitioned by rows, already resident in HBM as the fp32 column panel
    P[j][i] = R[row_lo + i][j] (M rows of `ld` floats; see sgv_ld_adopt_dense_colpanel)."""

    def __init__(self, ptr, ld, keepalive=None):
        self.ptr, self.ld, self.keepalive = int(ptr), int(ld), keepalive


class DeviceBlockDiag:
    """Block-diagonal LD already resident in HBM: dense row-major panels (see sgv_ld_adopt_blockdiag); `starts` are the
    local block boundaries of this rank (0 ... local rows)."""

    def __init__(self, ptr, starts, offs, lds, keepalive=None):
        self.ptr, self.starts, self.offs, self.lds, self.keepalive = int(ptr), starts, offs, lds, keepalive


class _SoloComm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def bcast(self, obj, root=0):
        return obj


class _Worker:
    """Ordered background worker: output dumps and host copies leave the solver loop."""

    def __init__(self):
        self.q = queue.Queue()
        self.err = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            fn, done = job
            try:
                if self.err is None:
                    fn()
            except Exception as e:  # surfaced at close()
                self.err = e
            finally:
                done.set()

    def submit(self, fn):
        done = threading.Event()
        self.q.put((fn, done))
        return done

    def close(self):
        self.q.put(None)
        self.t.join()
        if self.err is not None:
            raise self.err


class _ProbeSource:
    """Rademacher probes exactly as the reference draws them (src/sgvamp.py:326: numpy legacy global
    RNG, one call per cohort per iteration, in iteration order), generated one step ahead on a
    helper thread so that the ~15 ns/sample host RNG overlaps the GPU work of the previous step."""

    def __init__(self, order, M):
        self.q = queue.Queue(maxsize=2)
        self.states = {}            # RNG state before the draw of (it, k): what a checkpoint needs to continue the sequence
        self.end_state = None
        self.t = threading.Thread(target=self._run, args=(order, M), daemon=True)
        self.t.start()

    def _run(self, order, M):
        for key in order:
            self.states[key] = np.random.get_state()
            u = (np.random.binomial(p=1 / 2, n=1, size=M) * 2 - 1).astype(np.int8)
            self.q.put((key, u))
        self.end_state = np.random.get_state()

    def get(self, key):
        k, u = self.q.get()
        assert k == key
        return u


class VAMP:
    def __init__(self, N, Nt, M, K, rho, gamw, gam1, a, prior_vars, prior_probs, out_dir, out_name, comm=None,
                 device=0, stream=None, shard=None, shard_rows=None, halo=True):
        self.eps = 1e-32
        self.N = N
        self.Nt = Nt
        self.M = int(M)
        self.K = int(K)
        self.L = len(prior_probs)
        self.rho = rho
        self.gamw = gamw
        self.gam1 = gam1
        self.a = np.asarray(a, dtype=np.float64)
        self.lam = 1 - prior_probs[0]
        self.sigmas = np.array(prior_vars[1:], dtype=np.float64) * Nt       # src/sgvamp.py:27
        self.omegas = np.array([p / sum(prior_probs[1:]) for p in prior_probs[1:]])
        self.comm = comm if comm is not None else _SoloComm()
        self.gam = None
        size = self.comm.Get_size() if hasattr(self.comm, "Get_size") else 1
        self.rank_mode = size > 1
        if self.rank_mode and size != self.K:
            raise Exception("communicator size must equal the number of cohorts K")
        self.rank = self.comm.Get_rank() if self.rank_mode else 0
        self.my_cohorts = [self.rank] if self.rank_mode else list(range(self.K))
        # marker-row sharding over GPUs (no reference counterpart): this process owns rows [lo, hi)
        self.shard = shard if shard is not None else shd.SoloShard()
        if self.shard.world > 1 and self.rank_mode:
            raise Exception("row sharding and rank-per-cohort mode cannot be combined")
        self.bounds = list(shard_rows) if shard_rows is not None else shd.partition_rows(self.M, self.shard.world)
        self.lo, self.hi = self.bounds[self.shard.rank]
        self.Ml = self.hi - self.lo
        self.root = self.shard.rank == 0
        # halo: True - banded LD, neighbours' halos; False - block-diagonal LD sharded at block boundaries;
        # "rows" - dense LD, every rank holds its rows and gathers the vector pair of all ranks before a product
        self.rows = isinstance(halo, str) and halo == "rows"
        self.halo = (not self.rows) and bool(halo)
        if self.rows and self.shard.world == 1:
            raise Exception('halo="rows" partitions a dense LD over the ranks of a shard (world > 1)')
        self.device = device
        self.out_dir, self.out_name = out_dir, out_name
        if out_dir is not None and self.root:
            # the reference truncates every cohort's CSV from every rank (src/sgvamp.py:38-43), racing with ranks
            # that already append; here rank 0 creates the files and the others wait for it
            if not self.rank_mode or self.rank == 0:
                self.setup_io(out_dir, out_name)
            if self.rank_mode:
                self.comm.bcast(0, root=0)
        self.handle = nat.Handle(device=device, stream=stream)
        if self.shard.world == 1:
            self.handle.configure(self.M, self.K)
        else:
            self.handle.configure_part(self.M, self.K, self.shard.rank, self.shard.world, self.lo, self.hi,
                                       2 if self.rows else int(self.halo))
            shd.attach_peers(self.handle, self.shard)
        self.handle.set_weights(self.a)
        # the pinned read-back ring of the fused loop (page-locking 4 x 8 MB at M = 1M takes 20-130 ms depending on the
        # host): allocated in the background from here on, so that it overlaps the LD upload instead of delaying infer()
        self._pinned_cache = {}
        self._prefetch = threading.Thread(target=self._prefetch_pinned, daemon=True)
        self._prefetch.start()
        self._ld_loaded = [False] * self.K
        self._keep = []
        self.stats = {}

    # ------------------------------------------------------------------------------------------
    # output files (byte-compatible with src/sgvamp.py:33-76)
    # ------------------------------------------------------------------------------------------
    def setup_io(self, out_dir, out_name):
        self.out_dir = out_dir
        self.out_name = out_name
        for i in range(self.K):
            with open(os.path.join(self.out_dir, "%s_cohort_%d.csv" % (self.out_name, i + 1)), "w", newline="") as f:
                csv.writer(f, delimiter="\t").writerow(["it", "gamw", "gam1", "gam2", "alpha1", "alpha2", "lam"])
        with open(os.path.join(self.out_dir, "%s_metrics.csv" % self.out_name), "w", newline="") as f:
            csv.writer(f, delimiter="\t").writerow(["it", "alignment", "l2"])

    def write_params_to_file(self, params, cohort_idx):
        with open(os.path.join(self.out_dir, "%s_cohort_%d.csv" % (self.out_name, cohort_idx + 1)), "a", newline="") as f:
            csv.writer(f, delimiter="\t").writerow(params)

    def write_metrics_to_file(self, metrics):
        with open(os.path.join(self.out_dir, "%s_metrics.csv" % self.out_name), "a", newline="") as f:
            csv.writer(f, delimiter="\t").writerow(metrics)

    def write_xhat_to_file(self, it, xhat):
        np.ascontiguousarray(xhat, dtype=np.float64).ravel().tofile(
            os.path.join(self.out_dir, "%s_xhat_it_%d.bin" % (self.out_name, it)))

    def write_r1_to_file(self, it, r1, k):
        np.ascontiguousarray(r1, dtype=np.float64).ravel().tofile(
            os.path.join(self.out_dir, "%s_r1_cohort_%d_it_%d.bin" % (self.out_name, k, it)))

    # ------------------------------------------------------------------------------------------
    # LD / XTy ingestion
    # ------------------------------------------------------------------------------------------
    def load_ld(self, cohort, R, s=0.0, layout="auto", assume_symmetric=False):
        """Upload one cohort's LD matrix; Rused = (1-s) R + s I is applied on the device
        (src/main.py:265).  R: scipy sparse, ndarray / np.matrix, DeviceDIA or DeviceDense."""
        h = self.handle
        lay = {"auto": nat.LAYOUT_AUTO, "dense": nat.LAYOUT_DENSE, "dia": nat.LAYOUT_DIA,
               "blockdiag": nat.LAYOUT_BLOCKDIAG, "csr": nat.LAYOUT_CSR, "dsym": nat.LAYOUT_DSYM}[layout]
        if isinstance(R, DeviceDIA):
            assert s == 0.0, "device-resident LD must already be regularised"
            h.adopt_dia(cohort, R.ptr, R.w, R.ldb)
            self._keep.append(R)
        elif isinstance(R, DeviceDSYM):
            assert s == 0.0, "device-resident LD must already be regularised"
            h.adopt_dsym(cohort, R.ptr, R.w, R.ldb, R.ext)
            self._keep.append(R)
        elif isinstance(R, DeviceDense):
            assert s == 0.0, "device-resident LD must already be regularised"
            h.adopt_dense(cohort, R.ptr, R.ld)
            self._keep.append(R)
        elif isinstance(R, DeviceBlockDiag):
            assert s == 0.0, "device-resident LD must already be regularised"
            h.adopt_blockdiag(cohort, R.ptr, R.starts, R.offs, R.lds)
            self._keep.append(R)
        elif isinstance(R, DeviceDenseCols):
            assert s == 0.0, "device-resident LD must already be regularised"
            h.adopt_dense_colpanel(cohort, R.ptr, R.ld)
            self._keep.append(R)
        elif self.rows and scipy.sparse.issparse(R) and (lay == nat.LAYOUT_CSR or (
                lay == nat.LAYOUT_AUTO and R.nnz < 0.25 * R.shape[0] * R.shape[1])):
            # general sparse LD partitioned by rows: CSR rows with global column indices
            R = R.tocsr()
            if R.shape == (self.M, self.M):
                R = R[self.lo:self.hi]
            if R.shape != (self.Ml, self.M):
                raise Exception("LD rows of shape %s do not match rows [%d,%d) of M=%d" % (R.shape, self.lo, self.hi, self.M))
            if not R.has_canonical_format:
                R = R.copy()
                R.sum_duplicates()
            rc = h.upload_csr(cohort, R.indptr, R.indices, R.data, s=s, layout=nat.LAYOUT_CSR)
            if rc == -3:   # the CSR layout needs an explicitly stored diagonal when s != 0
                R = R.copy()
                R.setdiag(R.diagonal(k=self.lo), k=self.lo)     # inserts the missing entries (as explicit zeros)
                R.sort_indices()
                rc = h.upload_csr(cohort, R.indptr, R.indices, R.data, s=s, layout=nat.LAYOUT_CSR)
            h._ck(rc)
        elif self.rows:
            # dense LD partitioned by rows: the whole matrix or this rank's rows [lo, hi)
            Rd = R.toarray() if scipy.sparse.issparse(R) else np.asarray(R)
            if Rd.shape == (self.M, self.M):
                Rd = Rd[self.lo:self.hi]
            if Rd.shape != (self.Ml, self.M):
                raise Exception("LD rows of shape %s do not match rows [%d,%d) of M=%d" % (Rd.shape, self.lo, self.hi, self.M))
            h.upload_dense_rows(cohort, Rd, s=s)
        elif isinstance(R, DiaWindow) or (scipy.sparse.issparse(R) and R.format == "dia"):
            # banded LD in DIA form: diagonals travel as they are (no index arrays; half of them for symmetric LD)
            if isinstance(R, DiaWindow):
                data, offsets, col0, Mr = R.data, R.offsets, R.col0, R.M
            else:
                data, offsets, col0, Mr = R.data, R.offsets, 0, R.shape[0]
                if R.shape != (self.M, self.M):
                    raise Exception("LD matrix shape %s does not match M=%d" % (R.shape, self.M))
            if Mr != self.M:
                raise Exception("LD matrix size %d does not match M=%d" % (Mr, self.M))
            if lay not in (nat.LAYOUT_AUTO, nat.LAYOUT_DIA, nat.LAYOUT_DSYM):
                raise Exception("LD in DIA format maps to the band layouts (auto / dia / dsym)")
            if self.shard.world > 1:
                wmax = int(np.max(np.abs(offsets))) if len(offsets) else 0
                if not self.halo:
                    raise Exception("banded LD in DIA format needs halo=True when sharded")
                if min(hi_ - lo_ for lo_, hi_ in self.bounds) < wmax:
                    raise Exception("row shards are shorter than the LD half-bandwidth %d" % wmax)
            h._ck(h.upload_dia(cohort, data, offsets, s=s, layout=lay, assume_symmetric=assume_symmetric, col0=col0))
        elif scipy.sparse.issparse(R) and self.shard.world > 1:
            R = R.tocsr()
            if R.shape == (self.M, self.M):
                R = R[self.lo:self.hi]                 # global matrix given: keep this rank's rows
            if R.shape != (self.Ml, self.M):
                raise Exception("LD shard shape %s does not match rows [%d,%d) of M=%d" % (R.shape, self.lo, self.hi, self.M))
            if not R.has_canonical_format:
                R = R.copy()
                R.sum_duplicates()
            indptr, indices = R.indptr, R.indices
            if self.halo:                              # banded: neighbours supply w halo entries of the input vector
                wmax = max(self.shard.allgather(shd.local_bandwidth(indptr, indices, self.lo,
                                                                    sorted_indices=bool(R.has_sorted_indices))))
                if min(hi_ - lo_ for lo_, hi_ in self.bounds) < wmax:
                    raise Exception("row shards are shorter than the LD half-bandwidth %d" % wmax)
                h.set_bandwidth_hint(wmax)
                if lay not in (nat.LAYOUT_AUTO, nat.LAYOUT_DIA, nat.LAYOUT_DSYM):
                    raise Exception("a row partition with halos needs a band layout (auto / dia / dsym)")
                h._ck(h.upload_csr(cohort, indptr, indices, R.data, s=s, layout=lay))
            else:                                      # block-diagonal sharded by block: no coupling across shards
                Rl = R[:, self.lo:self.hi].tocsr()
                if Rl.nnz != R.nnz:
                    raise Exception("LD couples markers across the shard boundary of rank %d; shard block-diagonal LD "
                                    "at block boundaries (shard.partition_blocks) or use halo=True" % self.shard.rank)
                Rl.sort_indices()
                rc = h.upload_csr(cohort, Rl.indptr, Rl.indices, Rl.data, s=s, layout=lay)
                if rc == -3:
                    Rl = Rl.tolil()
                    Rl.setdiag(Rl.diagonal())
                    Rl = Rl.tocsr()
                    Rl.sort_indices()
                    rc = h.upload_csr(cohort, Rl.indptr, Rl.indices, Rl.data, s=s, layout=lay)
                h._ck(rc)
        elif scipy.sparse.issparse(R):
            R = R.tocsr()
            if R.shape != (self.M, self.M):
                raise Exception("LD matrix shape %s does not match M=%d" % (R.shape, self.M))
            if not R.has_canonical_format:
                R = R.copy()
                R.sum_duplicates()
            rc = h.upload_csr(cohort, R.indptr, R.indices, R.data, s=s, layout=lay)
            if rc == -3:   # CSR layout needs an explicitly stored diagonal when s != 0
                R = (R + scipy.sparse.diags(np.zeros(self.M)).tocsr()).tocsr()
                R.setdiag(R.diagonal())
                R.sort_indices()
                rc = h.upload_csr(cohort, R.indptr, R.indices, R.data, s=s, layout=lay)
            h._ck(rc)
        else:
            Rd = np.asarray(R)
            if Rd.shape != (self.M, self.M):
                raise Exception("LD matrix shape %s does not match M=%d" % (Rd.shape, self.M))
            h.upload_dense(cohort, Rd, s=s)
        self._ld_loaded[cohort] = True
        return h.ld_info(cohort)

    def build_ld_from_genotypes(self, cohort, G, N, w, s=0.0, taper=True, y=None, g0=0, device_ptr=None, nmark=None, ldg=None):
        """Banded LD of one cohort (and r = X^T y when y is given) computed on the GPU from int8 genotypes in {0,1,2},
        marker-major, following simulation/sim_gen_phen_mult.py:39-55 restricted to |i-j| <= w (see sgv_ld_build_banded).
        Returns r (this rank's rows) or None."""
        r = self.handle.build_banded(cohort, G, N, w, s=s, taper=taper, y=y, g0=g0, device_ptr=device_ptr, nmark=nmark, ldg=ldg)
        self._ld_loaded[cohort] = True
        return r

    # ------------------------------------------------------------------------------------------
    # per-step methods kept for API compatibility (reference src/sgvamp.py:93-194)
    # ------------------------------------------------------------------------------------------
    def _push_prior(self, handle=None):
        (handle or self.handle).set_prior(float(self.lam), self.omegas, self.sigmas)

    def _one_marker(self, rs, gam1s):
        aux = getattr(self, "_aux", None)
        if aux is None:
            aux = self._aux = nat.Handle(device=self.device)
            aux.configure(1, self.K)
        aux.set_weights(self.a)
        self._push_prior(aux)
        for k in range(self.K):
            aux.set_vec(k, nat.VEC_R1, np.array([rs[k]], dtype=np.float64))
        dfac = aux.denoise(np.asarray(gam1s, dtype=np.float64), 1.0, False)
        return aux.get_vec(0, nat.VEC_XHAT1)[0], dfac

    def denoiser_meta(self, rs, gam1s):
        return self._one_marker(rs, gam1s)[0]

    def der_denoiser_meta(self, rs, gam1s):
        k = self.comm.Get_rank()
        return self.a[k] * gam1s[k] * self._one_marker(rs, gam1s)[1]

    def _set_r1s(self, r1s):
        r1s = np.asarray(r1s, dtype=np.float64).reshape(self.K, self.M)
        for k in range(self.K):
            self.handle.set_vec(k, nat.VEC_R1, r1s[k])

    def prior_update_em(self, r1s, gam1s):
        if r1s is not None:
            self._set_r1s(r1s)
        self._push_prior()
        lam, om, _, _ = self.handle.prior_em(gam1s, 1, 0.0, self.L - 1)
        self.lam, self.omegas = lam, om

    def Lagrangian_der(self, x, omega0, sigma2, r1s, gam1s):
        if r1s is not None:
            self._set_r1s(r1s)
        self._push_prior()
        return self.handle.lagrangian(gam1s, x, omega0, sigma2)

    def prior_update_mle(self, r1s, gam1s):
        if r1s is not None:
            self._set_r1s(r1s)
        rank = self.comm.Get_rank()
        omega0 = np.zeros(self.L)
        omega0[0] = 1 - self.lam
        omega0[1:] = self.lam * self.omegas
        sigma2 = np.zeros(self.L)
        sigma2[0] = 1e-16
        sigma2[1:] = self.sigmas
        x0 = np.zeros(self.L + 1)
        x0[:-1] = omega0
        x0[-1] = 1 if self.gam is None else self.gam
        self._push_prior()
        nev = [0]

        def func(x):
            nev[0] += 1
            return self.handle.lagrangian(gam1s, x, omega0, sigma2)

        x, _, ier, _ = scipy.optimize.fsolve(func=func, x0=x0, full_output=True)
        self.stats["mle_evals"] = self.stats.get("mle_evals", 0) + nev[0]
        if ier != 1:
            if rank == 0:
                logging.info("WARNING: fsolve not converged. No prior update!")
            return "not_converged"
        elif any(s_ <= 0 for s_ in x[:-1]):
            if rank == 0:
                logging.info("WARNING: Negative values in MLE. No prior update!")
            return "negative"
        x[:-1] /= sum(x[:-1])
        self.lam = 1 - x[0]
        self.omegas = np.array([w / sum(x[1:-1]) for w in x[1:-1]])
        self.gam = x[self.L]
        return "ok"

    # ------------------------------------------------------------------------------------------
    # the solver loop (reference src/sgvamp.py:196-389)
    # ------------------------------------------------------------------------------------------
    def infer(self, R, r, iterations, x0=None, cg_maxit=500, em_prior_maxit=100, learn_gamw=True, lmmse_damp=True,
              prior_update=None, update_prior_from=1, *, s=0.0, probes=None, layout="auto", write_outputs=True,
              iter_hook=None, gather_outputs=True, checkpoint_path=None, checkpoint_every=0, resume_from=None, n_probes=1):
        """checkpoint_path / checkpoint_every / resume_from (no reference counterpart: the reference cannot restart,
        SURVEY 5.4): every `checkpoint_every` iterations the state needed to continue (r1, xhat1, xhat2, Sigma2_u, the
        scalar chain, the prior, the legacy RNG state of the probe sequence) is written to `checkpoint_path` (one .npz per
        row shard); `resume_from` continues such a run at the iteration after the checkpoint - the entries of the returned
        list before it are None.

        n_probes (no reference counterpart; 1 = the reference): number of Rademacher probes averaged in the Hutchinson
        estimates of alpha2 (src/sgvamp.py:338) and of the gamw update's trace (:359).  Probe 0 is the reference's probe
        (solved together with xhat2, warm-started); probes 1.. are solved from zero, two per extra 2-RHS solve
        (sgv_probe_pair).  Injected probes: an array of shape (K, iterations, n_probes, M) or a callable
        probes(k, it, M, p); the default draws them from numpy's legacy global RNG right after probe 0."""
        n_probes = int(n_probes)
        if n_probes < 1:
            raise Exception("n_probes must be >= 1")
        M, K, Nt, rho = self.M, self.K, self.Nt, self.rho
        h = self.handle
        rank = self.rank if self.shard.world == 1 else self.shard.rank   # only gates logging / rank-per-cohort mode
        mine = self.my_cohorts
        Rs = list(R) if isinstance(R, (list, tuple)) else [R]
        rs = list(r) if isinstance(r, (list, tuple)) else [r]
        if isinstance(self.N, (list, tuple, np.ndarray)):
            Ns = [float(n) for n in np.ravel(self.N)]
            Ns = Ns if len(Ns) == K else [Ns[0]] * K
        else:
            Ns = [float(self.N)] * K
        if len(Rs) != len(mine) or len(rs) != len(mine):
            raise Exception("expected %d LD matrices / XTy vectors, got %d / %d" % (len(mine), len(Rs), len(rs)))
        for idx, k in enumerate(mine):
            if Rs[idx] is not None:
                self.load_ld(k, Rs[idx], s=s, layout=layout)
            elif not self._ld_loaded[k]:
                raise Exception("no LD matrix for cohort %d" % k)
            h.set_xty(k, self._local(rs[idx]))
        h.reset_state()                                                 # :199-217
        ck = dict(path=checkpoint_path, every=int(checkpoint_every or 0))
        st0 = self._restore(resume_from) if resume_from is not None else None
        if (not self.rank_mode and prior_update != "mle" and h.iteration_supported() and n_probes == 1
                and os.environ.get("SGV_STEPWISE", "0") != "1"):
            return self._infer_fused(rs, Ns, iterations, x0, cg_maxit, em_prior_maxit, learn_gamw, lmmse_damp,
                                     prior_update, update_prior_from, probes, write_outputs, iter_hook, gather_outputs, ck, st0)
        write = write_outputs and self.out_dir is not None
        sharded = self.shard.world > 1
        worker = _Worker()
        tm = self.timers = dict(prior=0.0, denoise=0.0, lmmse=0.0, host_tail=0.0)
        pc = time.perf_counter
        probe_src = None
        if probes is None:
            if st0 is not None and st0.get("rng_state") is not None:
                np.random.set_state(st0["rng_state"])
            probe_src = _ProbeSource([key for it in range(st0["it_next"] if st0 else 0, iterations) for k in mine
                                      for key in [(it, k)] + [(it, k, p) for p in range(1, n_probes)]], M)
        sqrtNt = np.sqrt(Nt)
        truth = None
        if x0 is not None:
            truth = self._local(x0)
        gam1 = [np.float64(self.gam1)] * K
        gamw = [self.gamw] * K
        alpha1 = [np.float64(0.0)] * K
        alpha2 = [np.float64(0.0)] * K
        it0 = 0
        if st0 is not None:
            it0 = st0["it_next"]
            gam1, gamw = [np.float64(v) for v in st0["gam1"]], [float(v) for v in st0["gamw"]]
            alpha1, alpha2 = [np.float64(v) for v in st0["alpha1"]], [np.float64(v) for v in st0["alpha2"]]
        xhat1s = [None] * iterations
        self.history = dict(rows=[], cg_iters=[], cg_info=[], em_steps=[], spmm_passes=[], lam=[], omegas=[])
        NS = 3                                                          # host-copy ring depth
        pin_x = self._pinned_ring("x", NS, 1)
        pin_r = self._pinned_ring("r", NS, len(mine)) if write else None
        slot_done = [None] * NS

        if rank == 0:
            logging.debug(f"a = {self.a}")
        h.sync()
        self.shard.barrier()       # every rank's state is initialised before any kernel touches peer memory
        for it in range(it0, iterations):
            if iter_hook is not None:
                iter_hook(it)
            if rank == 0:
                logging.info(f"\n -----ITERATION {it} -----")
            t_it = pc()
            gam1s = np.array(gam1, dtype=np.float64)
            if self.rank_mode:                                          # :228-233
                got = self.comm.allgather_r1(h, gam1s[rank]) if hasattr(self.comm, "allgather_r1") else None
                if got is not None:                                     # device-side: one collective into the r1 block
                    gam1s = np.array(got, dtype=np.float64)
                else:                                                   # the reference's 2K broadcasts through the host
                    mine_r1 = h.get_vec(rank, nat.VEC_R1)
                    for i in range(K):
                        gam1s[i] = self.comm.bcast(gam1s[i], root=i)
                        got_i = self.comm.bcast(mine_r1 if i == rank else None, root=i)
                        if i != rank:
                            h.set_vec(i, nat.VEC_R1, got_i)
            # prior update :242-259
            em_steps = 0
            if it >= update_prior_from:
                if prior_update == "mle":
                    if rank == 0:
                        logging.info("...Updating prior parameters using MLE")
                    self.prior_update_mle(None, gam1s)
                elif prior_update == "em":
                    if rank == 0:
                        logging.info("...Updating prior parameters using EM")
                    self._push_prior()
                    self.lam, self.omegas, em_steps, rel = h.prior_em(gam1s, em_prior_maxit, 1e-6, self.L - 1)
                    if rank == 0:
                        logging.info(f"... prior-learning EM algorithm performed {em_steps} steps and had final relative error = {rel:0.9f}")
            self._push_prior()
            t_pr = pc()
            tm["prior"] += t_pr - t_it
            # denoising :270-293
            if rank == 0:
                logging.info("...Denoising")
            dmean = np.float64(h.denoise(gam1s, rho, it > 0))
            t_a = pc()
            slot = it % NS
            if slot_done[slot] is not None:
                slot_done[slot].wait()                                   # the worker has drained this slot
            t_b = pc()
            tm["slot_wait"] = tm.get("slot_wait", 0.0) + (t_b - t_a)
            tm["denoise_call"] = tm.get("denoise_call", 0.0) + (t_a - t_pr)
            h.get_vec_async(0, nat.VEC_XHAT1, 1.0, pin_x[slot][0])
            if write:
                for idx, k in enumerate(mine):
                    h.get_vec_async(k, nat.VEC_R1, 1.0, pin_r[slot][idx])
            t_dn = pc()
            tm["denoise"] += t_dn - t_pr
            rows_it, iters_it, info_it, passes_it = {}, {}, {}, 0
            for k in mine:
                a1 = self.a[k] * gam1s[k] * dmean                        # :285
                if it > 0:
                    a1 = rho * a1 + (1 - rho) * alpha1[k]                # :290-291
                alpha1[k] = a1
                logging.info(f"...LMMSE cohort {k}")
                alpha2_prev = alpha2[k]
                gam2 = gam1[k] * (1 - a1) / a1                           # :305
                def probe(p):
                    if probes is None:
                        u_ = probe_src.get((it, k) if p == 0 else (it, k, p))    # :326 (same RNG calls, same order)
                    elif callable(probes):
                        u_ = probes(k, it, M) if p == 0 else probes(k, it, M, p)
                    else:
                        u_ = np.asarray(probes)[k, it] if n_probes == 1 else np.asarray(probes)[k, it, p]
                    if sharded and len(u_) == M:
                        u_ = u_[self.lo:self.hi]                          # every rank draws the same global probe
                    return u_

                u = probe(0)
                out = h.lmmse(k, float(gamw[k]), float(gam2), float(a1), float(rho), cg_maxit, lmmse_damp, learn_gamw,
                              it == 0, u)
                for c in range(2):
                    if out.cg_info[c] > 0:
                        logging.info(f"Rank {k} WARNING: CG {c + 1} convergence after {out.cg_info[c]} iterations not achieved!")
                u_sigma2u, u_R_sigma2u = np.float64(out.u_sigma2u), out.u_R_sigma2u
                if n_probes > 1:                                         # further probes, two per extra solve
                    extra = [probe(p) for p in range(1, n_probes)]
                    for i in range(0, len(extra), 2):
                        ub = extra[i + 1] if i + 1 < len(extra) else None
                        po = h.probe_pair(k, float(gamw[k]), float(gam2), cg_maxit, extra[i], ub)
                        for c in range(2 if ub is not None else 1):
                            u_sigma2u += np.float64(po.u_s[c])
                            u_R_sigma2u += po.u_R_s[c]
                        out.spmm_passes += po.spmm_passes
                    u_sigma2u /= n_probes
                    u_R_sigma2u /= n_probes
                a2 = gam2 * u_sigma2u / M                                # :338-340
                if lmmse_damp:
                    a2 = rho * a2 + (1 - rho) * alpha2_prev              # :345-346
                alpha2[k] = a2
                gam1[k] = gam2 * (1 - a2) / a2                           # :347
                h.update_r1(k, float(a2))                                # :348
                gw = gamw[k]
                if learn_gamw:                                           # :350-364
                    N = Ns[k]
                    z = N - 2 * out.xhat2_r + out.xhat2_R_xhat2
                    if z < 0:
                        z = 0
                    gw = float(1 / (z / N + u_R_sigma2u / N))
                gw = max(gw, 1.0)                                        # :374
                gamw[k] = gw
                row = [it, gw, gam1[k], gam2, a1, a2, self.lam]
                rows_it[k] = row
                iters_it[k] = (out.cg_iters[0], out.cg_iters[1])
                info_it[k] = (out.cg_info[0], out.cg_info[1])
                passes_it += out.spmm_passes
                if self.out_dir is not None and write_outputs and self.root:
                    self.write_params_to_file(row, k)                    # :377
            t_lm = pc()
            tm["lmmse"] += t_lm - t_dn

            def drain(it=it, slot=slot):
                h.wait_copies()
                xh = pin_x[slot][0].copy()
                xhat1s[it] = xh.reshape(-1, 1)
                if write and sharded:                                    # this rank's slices; assembled after the loop
                    self._write_part("%s_xhat_it_%d.bin" % (self.out_name, it), xh / sqrtNt)
                    for idx, k in enumerate(mine):
                        self._write_part("%s_r1_cohort_%d_it_%d.bin" % (self.out_name, k + 1, it), pin_r[slot][idx] / sqrtNt)
                elif write:
                    if rank == 0:
                        (xh / sqrtNt).tofile(os.path.join(self.out_dir, "%s_xhat_it_%d.bin" % (self.out_name, it)))
                    for idx, k in enumerate(mine):
                        (pin_r[slot][idx] / sqrtNt).tofile(
                            os.path.join(self.out_dir, "%s_r1_cohort_%d_it_%d.bin" % (self.out_name, k + 1, it)))

            slot_done[slot] = worker.submit(drain)
            if truth is not None:                                        # :379-387
                d = h.metrics(truth if it == it0 else None)
                alignment = d[0] / np.sqrt(d[1]) / np.sqrt(d[2])
                l2 = np.sqrt(d[3]) / np.sqrt(d[2])
                if rank == 0:
                    logging.debug(f"Alignment(xhat1, x0) = {alignment:0.9f} \n")
                    logging.debug(f"L2_error(xhat1, x0) = {l2:0.9f} \n")
                    if self.out_dir is not None and write_outputs and self.root:
                        self.write_metrics_to_file([it, alignment, l2])
                self.history.setdefault("metrics", []).append((it, alignment, l2))
            self.history["rows"].append(rows_it)
            self.history["cg_iters"].append(iters_it)
            self.history["cg_info"].append(info_it)
            self.history["em_steps"].append(em_steps)
            self.history["spmm_passes"].append(passes_it)
            self.history["lam"].append(float(self.lam))
            self.history["omegas"].append(np.array(self.omegas, dtype=np.float64).copy())
            if ck["path"] and ck["every"] > 0 and (it + 1) % ck["every"] == 0:
                self._checkpoint(ck["path"], it + 1, gam1, gamw, alpha1, alpha2, probe_src, mine)
            tm["host_tail"] += pc() - t_lm
        if iter_hook is not None:
            iter_hook(iterations)
        h.sync()
        worker.close()
        self.gam1_final, self.gamw_final = gam1, gamw
        if sharded and gather_outputs:
            # the data path has no host collective; the per-rank output slices are assembled once, here
            have = [i for i in range(iterations) if xhat1s[i] is not None]
            full = shd.gather_rows(self.shard, np.stack([xhat1s[i].ravel() for i in have]) if have else np.zeros((0, self.Ml)),
                                   self.bounds)
            for j, i in enumerate(have):
                xhat1s[i] = full[j].reshape(M, 1)
        if sharded and write:
            self._assemble_parts(["%s_xhat_it_%d.bin" % (self.out_name, it) for it in range(it0, iterations)] +
                                 ["%s_r1_cohort_%d_it_%d.bin" % (self.out_name, k + 1, it) for it in range(it0, iterations) for k in mine])
        return xhat1s

    def _infer_fused(self, rs, Ns, iterations, x0, cg_maxit, em_prior_maxit, learn_gamw, lmmse_damp, prior_update,
                     update_prior_from, probes, write_outputs, iter_hook, gather_outputs, ck=None, st0=None):
        """The same loop with the scalar chain on the device (sgv_iteration_*): one VAMP iteration is enqueued without a
        host round trip and the host reads one small record per iteration, one iteration behind the GPU.  Used whenever
        the prior update does not need the host (EM or none) and this process drives all cohorts."""
        M, K, Nt, rho = self.M, self.K, self.Nt, self.rho
        h = self.handle
        rank = self.shard.rank
        mine = self.my_cohorts
        write = write_outputs and self.out_dir is not None
        sharded = self.shard.world > 1
        sqrtNt = np.sqrt(Nt)
        NS = nat.ITER_SLOTS
        worker = _Worker()
        tm = self.timers = dict(enqueue=0.0, wait=0.0, host_tail=0.0)
        pc = time.perf_counter
        t_enter = pc()
        ck = ck or dict(path=None, every=0)
        it0 = st0["it_next"] if st0 is not None else 0
        probe_src = None
        if probes is None:
            if st0 is not None and st0.get("rng_state") is not None:
                np.random.set_state(st0["rng_state"])
            probe_src = _ProbeSource([(it, k) for it in range(it0, iterations) for k in mine], M)
        self._push_prior()
        if st0 is None:
            h.vamp_begin([self.gam1] * K, [self.gamw] * K, Ns)
        else:
            h.vamp_begin(st0["gam1"], st0["gamw"], Ns)
            h.vamp_set_alphas(st0["alpha1"], st0["alpha2"])
        truth = None
        if x0 is not None:
            truth = self._local(x0)
            h.set_truth(truth)
        xhat1s = [None] * iterations
        self.history = dict(rows=[], cg_iters=[], cg_info=[], em_steps=[], spmm_passes=[], lam=[], omegas=[])
        pin_x = self._pinned_ring("x", NS, 1)
        pin_r = self._pinned_ring("r", NS, K) if write else None
        slot_done = [None] * NS
        if rank == 0:
            logging.debug(f"a = {self.a}")
        h.sync()
        self.shard.barrier()       # every rank's state is initialised before any kernel touches peer memory
        tm["setup"] = pc() - t_enter
        t_loop = pc()

        def finish(it):
            slot = it % NS
            t0 = pc()
            out = h.iteration_wait(slot)
            tm["wait"] += pc() - t0
            self.lam = out.lam
            self.omegas = np.array(out.omegas[: self.L - 1], dtype=np.float64)
            if rank == 0:
                logging.info(f"\n -----ITERATION {it} -----")
                if prior_update == "em" and it >= update_prior_from:
                    logging.info(f"... prior-learning EM algorithm performed {out.em_steps} steps and had final relative error = {out.em_relerr:0.9f}")
            rows_it, iters_it, info_it, passes_it = {}, {}, {}, 0
            for k in mine:
                ck = out.coh[k]
                row = [it] + [np.float64(v) for v in ck.row[1:6]] + [float(ck.row[6])]
                row[1] = float(row[1])
                for c_ in range(2):
                    if ck.cg_info[c_] > 0:
                        logging.info(f"Rank {k} WARNING: CG {c_ + 1} convergence after {ck.cg_info[c_]} iterations not achieved!")
                rows_it[k] = row
                iters_it[k] = (ck.cg_iters[0], ck.cg_iters[1])
                info_it[k] = (ck.cg_info[0], ck.cg_info[1])
                passes_it += ck.spmm_passes
                if write and self.root:
                    self.write_params_to_file(row, k)                    # :377
            if truth is not None:                                        # :379-387
                d = out.metrics
                alignment = d[0] / np.sqrt(d[1]) / np.sqrt(d[2])
                l2 = np.sqrt(d[3]) / np.sqrt(d[2])
                if rank == 0:
                    logging.debug(f"Alignment(xhat1, x0) = {alignment:0.9f} \n")
                    logging.debug(f"L2_error(xhat1, x0) = {l2:0.9f} \n")
                    if write and self.root:
                        self.write_metrics_to_file([it, alignment, l2])
                self.history.setdefault("metrics", []).append((it, alignment, l2))
            self.history["rows"].append(rows_it)
            self.history["cg_iters"].append(iters_it)
            self.history["cg_info"].append(info_it)
            self.history["em_steps"].append(int(out.em_steps) if (prior_update == "em" and it >= update_prior_from) else 0)
            self.history["spmm_passes"].append(passes_it)
            self.history["lam"].append(float(self.lam))
            self.history["omegas"].append(np.array(self.omegas, dtype=np.float64).copy())

            def drain(it=it, slot=slot):
                h.wait_copies()
                xh = pin_x[slot][0].copy()
                xhat1s[it] = xh.reshape(-1, 1)
                if write and sharded:                                    # this rank's slices; assembled after the loop
                    self._write_part("%s_xhat_it_%d.bin" % (self.out_name, it), xh / sqrtNt)
                    for k in mine:
                        self._write_part("%s_r1_cohort_%d_it_%d.bin" % (self.out_name, k + 1, it), pin_r[slot][k] / sqrtNt)
                elif write:
                    (xh / sqrtNt).tofile(os.path.join(self.out_dir, "%s_xhat_it_%d.bin" % (self.out_name, it)))
                    for k in mine:
                        (pin_r[slot][k] / sqrtNt).tofile(
                            os.path.join(self.out_dir, "%s_r1_cohort_%d_it_%d.bin" % (self.out_name, k + 1, it)))

            slot_done[slot] = worker.submit(drain)

        finished = it0 - 1
        for it in range(it0, iterations):
            if iter_hook is not None:
                iter_hook(it)
            t0 = pc()
            slot = it % NS
            if slot_done[slot] is not None:
                slot_done[slot].wait()                                   # the worker has drained this slot's host buffers
            pb = h.iteration_probe_buffer(slot)
            for k in mine:
                if probes is None:
                    u = probe_src.get((it, k))                           # :326 (same RNG calls, same order)
                elif callable(probes):
                    u = probes(k, it, M)
                else:
                    u = np.asarray(probes)[k, it]
                if sharded and len(u) == M:
                    u = u[self.lo:self.hi]                               # every rank draws the same global probe
                pb[k, :] = u
            h.iteration_enqueue(it, prior_update == "em" and it >= update_prior_from, em_prior_maxit, 1e-6, rho, cg_maxit,
                                lmmse_damp, learn_gamw, truth is not None, pin_x[slot][0],
                                pin_r[slot] if write else None, slot)
            tm["enqueue"] += pc() - t0
            t1 = pc()
            if finished < it - 1:
                finish(it - 1)                                           # one iteration behind the GPU
                finished = it - 1
            if ck["path"] and ck["every"] > 0 and (it + 1) % ck["every"] == 0:
                finish(it)                                               # a checkpoint needs this iteration's results now
                finished = it
                row = self.history["rows"][-1]
                self._checkpoint(ck["path"], it + 1, [row[k][2] for k in range(K)], [row[k][1] for k in range(K)],
                                 [row[k][4] for k in range(K)], [row[k][5] for k in range(K)], probe_src, mine)
            tm["host_tail"] += pc() - t1
        if finished < iterations - 1:
            finish(iterations - 1)
        if iter_hook is not None:
            iter_hook(iterations)
        tm["loop"] = pc() - t_loop
        t_end = pc()
        h.sync()
        worker.close()
        tm["drain"] = pc() - t_end
        if iterations > 0:
            last = self.history["rows"][-1]
            self.gam1_final = [last[k][2] for k in mine]
            self.gamw_final = [last[k][1] for k in mine]
        if sharded and gather_outputs:
            have = [i for i in range(iterations) if xhat1s[i] is not None]
            full = shd.gather_rows(self.shard, np.stack([xhat1s[i].ravel() for i in have]) if have else np.zeros((0, self.Ml)),
                                   self.bounds)
            for j, i in enumerate(have):
                xhat1s[i] = full[j].reshape(M, 1)
        if sharded and write:
            self._assemble_parts(["%s_xhat_it_%d.bin" % (self.out_name, it) for it in range(it0, iterations)] +
                                 ["%s_r1_cohort_%d_it_%d.bin" % (self.out_name, k + 1, it) for it in range(it0, iterations) for k in mine])
        return xhat1s

    # checkpoint / resume
    def _ckpt_file(self, path):
        return path if self.shard.world == 1 else "%s.rank%d" % (path, self.shard.rank)

    def _checkpoint(self, path, it_next, gam1, gamw, alpha1, alpha2, probe_src, mine):
        h = self.handle
        h.sync()
        d = dict(it_next=it_next, M=self.M, lo=self.lo, hi=self.hi, K=self.K, gam1=np.array(gam1, dtype=np.float64),
                 gamw=np.array(gamw, dtype=np.float64), alpha1=np.array(alpha1, dtype=np.float64),
                 alpha2=np.array(alpha2, dtype=np.float64), lam=float(self.lam), omegas=np.array(self.omegas, dtype=np.float64),
                 gam=np.float64(np.nan if self.gam is None else self.gam), xhat1=h.get_vec(0, nat.VEC_XHAT1))
        for k in mine:
            d["r1_%d" % k] = h.get_vec(k, nat.VEC_R1)
            d["xhat2_%d" % k] = h.get_vec(k, nat.VEC_XHAT2)
            d["sig_%d" % k] = h.get_vec(k, nat.VEC_SIGMA2U)
        if probe_src is not None:
            t0 = time.time()
            st = None
            while st is None and time.time() - t0 < 30:     # the producer thread runs at most two draws ahead
                st = probe_src.states.get((it_next, mine[0])) or probe_src.end_state
                if st is None:
                    time.sleep(0.001)
            if st is not None:
                d["rng_name"], d["rng_keys"], d["rng_pos"], d["rng_has_gauss"], d["rng_gauss"] = st[0], st[1], st[2], st[3], st[4]
        f = self._ckpt_file(path)
        with open(f + ".tmp", "wb") as fh:
            np.savez(fh, **d)
        os.replace(f + ".tmp", f)

    def _restore(self, path):
        z = np.load(self._ckpt_file(path), allow_pickle=False)
        if int(z["M"]) != self.M or int(z["lo"]) != self.lo or int(z["hi"]) != self.hi or int(z["K"]) != self.K:
            raise Exception("checkpoint %s does not match this problem / row partition" % path)
        h = self.handle
        h.set_vec(0, nat.VEC_XHAT1, z["xhat1"])
        for k in self.my_cohorts:
            h.set_vec(k, nat.VEC_R1, z["r1_%d" % k])
            h.set_vec(k, nat.VEC_XHAT2, z["xhat2_%d" % k])       # injected warm starts: R x0 is re-formed by one real pass
            h.set_vec(k, nat.VEC_SIGMA2U, z["sig_%d" % k])
        self.lam = float(z["lam"])
        self.omegas = np.array(z["omegas"], dtype=np.float64)
        self.gam = None if np.isnan(z["gam"]) else float(z["gam"])
        st = dict(it_next=int(z["it_next"]), gam1=z["gam1"], gamw=z["gamw"], alpha1=z["alpha1"], alpha2=z["alpha2"], rng_state=None)
        if "rng_keys" in z.files:
            st["rng_state"] = (str(z["rng_name"]), z["rng_keys"], int(z["rng_pos"]), int(z["rng_has_gauss"]), float(z["rng_gauss"]))
        return st

    # sharded output dumps: every rank streams its row slice of a dump into a part file while the loop runs; after the
    # loop rank 0 concatenates the parts in rank order (nothing of size iterations x M is ever held in host memory)
    def _part_path(self, name, rank):
        return os.path.join(self.out_dir, ".%s.part%d" % (name, rank))

    def _write_part(self, name, vec):
        np.ascontiguousarray(vec, dtype=np.float64).ravel().tofile(self._part_path(name, self.shard.rank))

    def _assemble_parts(self, names):
        self.shard.barrier()                                    # every rank's parts are on disk
        if self.root:
            for name in names:
                with open(os.path.join(self.out_dir, name), "wb") as out:
                    for q in range(self.shard.world):
                        pp = self._part_path(name, q)
                        with open(pp, "rb") as f:
                            while True:
                                buf = f.read(1 << 24)
                                if not buf:
                                    break
                                out.write(buf)
                        os.remove(pp)
        self.shard.barrier()

    def _local(self, v):
        """This rank's rows of a marker vector given either globally (length M) or already sliced."""
        v = np.asarray(v, dtype=np.float64).ravel()
        if v.shape[0] == self.M and self.shard.world > 1:
            return np.ascontiguousarray(v[self.lo:self.hi])
        if v.shape[0] != self.Ml:
            raise Exception("vector of length %d does not match M=%d (local rows %d)" % (v.shape[0], self.M, self.Ml))
        return v

    def _prefetch_pinned(self):
        try:
            key = ("x", nat.ITER_SLOTS, 1)
            self._pinned_cache[key] = [[self.handle.pinned_array(self.Ml)] for _ in range(nat.ITER_SLOTS)]
        except Exception:                                               # allocated on demand instead
            pass

    def _pinned_ring(self, tag, depth, width):
        if self._prefetch is not None:
            self._prefetch.join()
            self._prefetch = None
        key = (tag, depth, width)
        cache = self._pinned_cache
        if key not in cache:
            cache[key] = [[self.handle.pinned_array(self.Ml) for _ in range(width)] for _ in range(depth)]
        return cache[key]

    def close(self):
        if self._prefetch is not None:
            self._prefetch.join()
            self._prefetch = None
        if self.shard.world > 1 and self.handle.h:
            self.handle.sync()
            self.shard.barrier()   # peers may still be reading this rank's arena
        self.handle.close()
        if getattr(self, "_aux", None) is not None:
            self._aux.close()
