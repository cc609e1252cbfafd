"""Generate golden vectors by running the UNMODIFIED reference in this container.

Usage (build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py

Imports /root/reference/src/sgvamp.py by path, drives ``VAMP.infer`` with a stub ``comm``
(K=1) or one thread per cohort sharing a barrier-based ``bcast`` (K>1), injects the probe
vectors by rebinding the module-level name ``sgvamp.binomial`` and records CG iteration counts
by wrapping ``sgvamp.con_grad``.  Inputs and per-iteration outputs are written to
``tests/golden/<case>.npz``.  Nothing from the reference is copied into the repo.
"""
import csv
import os
import sys
import tempfile
import threading

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
sys.path.insert(0, "/root/reference/src")
import ldgen  # noqa: E402
import sgvamp as ref  # noqa: E402  (the reference module)

assert ref.__file__.startswith("/root/reference/"), ref.__file__

_tls = threading.local()


class SoloComm:
    def Get_rank(self):
        return 0

    def bcast(self, obj, root=0):
        return obj


class ThreadComm:
    """K threads, thread-local rank, barrier-based bcast (root publishes a copy)."""

    def __init__(self, K):
        self.K = K
        self.bar = threading.Barrier(K)
        self.slot = None

    def Get_rank(self):
        return _tls.rank

    def bcast(self, obj, root=0):
        if _tls.rank == root:
            self.slot = np.array(obj, copy=True) if isinstance(obj, np.ndarray) else obj
        self.bar.wait()
        out = self.slot
        out = out.copy() if isinstance(out, np.ndarray) else out
        self.bar.wait()
        return out


def _probe_table(K, iterations, M, seed):
    rng = np.random.RandomState(seed)
    return (rng.binomial(p=0.5, n=1, size=(K, iterations, M)) * 2 - 1).astype(np.int64)


def run_reference(case):
    K, M, iters = case["K"], case["M"], case["iterations"]
    probes = _probe_table(K, iters, M, case["probe_seed"])
    Rused = [ldgen_regularise(R, case["s"], dense_as_matrix=case.get("dense_as_matrix", True)) for R in case["R"]]
    N_list = case["N_list"]
    Nt = sum(N_list)
    a = np.array(N_list) / Nt
    outdir = tempfile.mkdtemp(prefix="golden_")
    comm = SoloComm() if K == 1 else ThreadComm(K)
    counters = [dict(it=-1, n=[]) for _ in range(K)]
    results = [None] * K

    orig_cg = ref.con_grad
    orig_binomial = ref.binomial

    def cg_wrapped(A, b, maxiter=None, x0=None):
        cnt = [0]

        def cb(_x):
            cnt[0] += 1

        x, info = orig_cg(A, b, maxiter=maxiter, x0=x0, callback=cb)
        counters[_tls.rank]["n"].append((cnt[0], info))
        return x, info

    def binomial_injected(p=None, n=None, size=None):
        c = counters[_tls.rank]
        c["it"] += 1
        return ((probes[_tls.rank, c["it"]] + 1) // 2).astype(np.int64)

    def binomial_recorded(p=None, n=None, size=None):
        # default-RNG case: the reference's own call (src/sgvamp.py:326) on numpy's legacy global state, only recorded
        c = counters[_tls.rank]
        c["it"] += 1
        b = orig_binomial(p=p, n=n, size=size)
        probes[_tls.rank, c["it"]] = b * 2 - 1
        return b

    ref.con_grad = cg_wrapped
    if case.get("rng_seed") is not None:
        assert K == 1, "one global RNG stream: single cohort only"
        np.random.seed(case["rng_seed"])
        ref.binomial = binomial_recorded
    else:
        ref.binomial = binomial_injected

    def worker(k):
        _tls.rank = k
        x0 = None
        if case.get("x0") is not None:
            x0 = case["x0"] * np.sqrt(N_list[k])       # src/main.py:276
        v = ref.VAMP(N=N_list[k], Nt=Nt, M=M, K=K, rho=case["rho"], gamw=case["gamw"], gam1=case["gam1"],
                     a=a, prior_vars=case["prior_vars"], prior_probs=case["prior_probs"],
                     out_dir=outdir, out_name="g", comm=comm)
        xs = v.infer(Rused[k], case["r"][k].reshape(M, 1), iters, x0=x0, cg_maxit=case["cg_maxit"],
                     em_prior_maxit=case["em_prior_maxit"], learn_gamw=case["learn_gamw"],
                     lmmse_damp=case["lmmse_damp"], prior_update=case["prior_update"],
                     update_prior_from=case["update_prior_from"])
        results[k] = (np.stack([x.ravel() for x in xs]), v)

    try:
        if K == 1:
            worker(0)
        else:
            ths = [threading.Thread(target=worker, args=(k,)) for k in range(K)]
            [t.start() for t in ths]
            [t.join() for t in ths]
    finally:
        ref.con_grad = orig_cg
        ref.binomial = orig_binomial

    xhat = results[0][0]                                 # identical on all ranks
    for k in range(1, K):
        assert np.array_equal(results[k][0], xhat), "xhat1 differs across ranks"
    rows = np.zeros((iters, K, 7))
    for k in range(K):
        with open(os.path.join(outdir, "g_cohort_%d.csv" % (k + 1)), newline="") as f:
            rd = list(csv.reader(f, delimiter="\t"))
        assert rd[0] == ["it", "gamw", "gam1", "gam2", "alpha1", "alpha2", "lam"]
        for it in range(iters):
            rows[it, k] = [float(v) for v in rd[1 + it]]
    r1_dump = np.zeros((iters, K, M))
    xhat_dump = np.zeros((iters, M))
    for it in range(iters):
        xhat_dump[it] = np.fromfile(os.path.join(outdir, "g_xhat_it_%d.bin" % it))
        for k in range(K):
            r1_dump[it, k] = np.fromfile(os.path.join(outdir, "g_r1_cohort_%d_it_%d.bin" % (k + 1, it)))
    cg_iters = np.zeros((iters, K, 2), dtype=np.int64)
    cg_info = np.zeros((iters, K, 2), dtype=np.int64)
    for k in range(K):
        n = counters[k]["n"]
        assert len(n) == 2 * iters
        for it in range(iters):
            cg_iters[it, k] = [n[2 * it][0], n[2 * it + 1][0]]
            cg_info[it, k] = [n[2 * it][1], n[2 * it + 1][1]]
    metrics = None
    if case.get("x0") is not None:
        with open(os.path.join(outdir, "g_metrics.csv"), newline="") as f:
            rd = list(csv.reader(f, delimiter="\t"))
        metrics = np.array([[float(v) for v in row] for row in rd[1:]])
    with open(os.path.join(outdir, "g_cohort_1.csv"), "rb") as f:
        csv_bytes = f.read()
    return dict(xhat=xhat, xhat_dump=xhat_dump, r1_dump=r1_dump, rows=rows, cg_iters=cg_iters,
                cg_info=cg_info, probes=probes, metrics=metrics, csv_bytes=csv_bytes,
                final_lam=results[0][1].lam, final_omegas=np.asarray(results[0][1].omegas))


def ldgen_regularise(R, s, dense_as_matrix=True):
    """Exactly src/main.py:265 (dense input becomes np.matrix there)."""
    M = R.shape[0]
    out = (1 - s) * R + s * scipy.sparse.identity(M)
    if scipy.sparse.issparse(out):
        return out.tocsr()
    return out if dense_as_matrix else np.asarray(out)


def save_case(name, case, out):
    d = dict(
        K=case["K"], M=case["M"], iterations=case["iterations"], s=case["s"], rho=case["rho"],
        gamw=case["gamw"], gam1=case["gam1"], N_list=np.array(case["N_list"], dtype=np.float64),
        prior_vars=np.array(case["prior_vars"]), prior_probs=np.array(case["prior_probs"]),
        cg_maxit=case["cg_maxit"], em_prior_maxit=case["em_prior_maxit"],
        learn_gamw=case["learn_gamw"], lmmse_damp=case["lmmse_damp"],
        prior_update=str(case["prior_update"]), update_prior_from=case["update_prior_from"],
        r=np.stack(case["r"]), probes=out["probes"].astype(np.int8),
        xhat=out["xhat"], xhat_dump=out["xhat_dump"], r1_dump=out["r1_dump"], rows=out["rows"],
        cg_iters=out["cg_iters"], cg_info=out["cg_info"],
        csv_bytes=np.frombuffer(out["csv_bytes"], dtype=np.uint8),
        final_lam=out["final_lam"], final_omegas=out["final_omegas"],
        layout=case["layout"],
    )
    if case.get("rng_seed") is not None:
        d["rng_seed"] = case["rng_seed"]
    if case.get("x0") is not None:
        d["x0"] = case["x0"]
        d["metrics"] = out["metrics"]
    for k, R in enumerate(case["R"]):
        if scipy.sparse.issparse(R):
            R = R.tocsr()
            d["R%d_indptr" % k] = R.indptr.astype(np.int64)
            d["R%d_indices" % k] = R.indices.astype(np.int32)
            # all cases use fp32-representable values, so storing fp32 is lossless
            assert np.array_equal(R.data.astype(np.float32).astype(np.float64), R.data)
            d["R%d_data" % k] = R.data.astype(np.float32)
        else:
            assert np.array_equal(np.asarray(R).astype(np.float32).astype(np.float64), np.asarray(R))
            d["R%d_dense" % k] = np.asarray(R).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print("%-28s M=%d K=%d it=%d  cg_iters[0..2]=%s  alpha1[-1]=%.4g gamw[-1]=%.4g lam=%.4g  max|xhat|=%.3g" % (
        name, case["M"], case["K"], case["iterations"], out["cg_iters"][:3, 0].tolist(),
        out["rows"][-1, 0, 4], out["rows"][-1, 0, 1], out["final_lam"], np.abs(out["xhat"]).max()))


def base(**kw):
    c = dict(K=1, iterations=6, s=0.0, rho=0.5, gamw=2.0, gam1=1e-6, cg_maxit=500, em_prior_maxit=100,
             learn_gamw=True, lmmse_damp=False, prior_update="em", update_prior_from=1, probe_seed=1234,
             x0=None)
    c.update(kw)
    return c


def main():
    f32 = ldgen.round_to_f32
    cases = {}

    # dense K=1 L=2 (config-1 regime, reduced): N = 4M
    Rs, rs, beta, Nl = ldgen.sim_dense(M=320, N=1280, lam=0.02, h2=0.5, seed=1)
    cm = max(1, int(320 * 0.02))
    cases["dense_L2_em"] = base(M=320, R=[f32(Rs[0])], r=rs, N_list=Nl, prior_vars=[0, 0.5 / cm],
                                prior_probs=[0.98, 0.02], layout="dense", x0=beta, iterations=8)
    # dense K=1 L=4, cg_maxit=50, s=0.1 (config-2 regime, reduced)
    Rs, rs, beta, Nl = ldgen.sim_dense(M=384, N=768, lam=0.02, h2=0.5, seed=2)
    v = 0.5 / max(1, int(384 * 0.02))
    cases["dense_L4_em_s01"] = base(M=384, R=[f32(Rs[0])], r=rs, N_list=Nl, s=0.1,
                                    prior_vars=[0, 0.1 * v, v, 10 * v], prior_probs=[0.97, 0.01, 0.01, 0.01],
                                    cg_maxit=50, layout="dense", iterations=7)
    # dense, no prior update, fixed gamw, lmmse damping on (the infer() signature default)
    Rs, rs, beta, Nl = ldgen.sim_dense(M=256, N=2048, lam=0.03, h2=0.4, seed=3)
    cm = max(1, int(256 * 0.03))
    cases["dense_noprior_fixedgamw_damp"] = base(M=256, R=[f32(Rs[0])], r=rs, N_list=Nl,
                                                 prior_vars=[0, 0.4 / cm], prior_probs=[0.97, 0.03],
                                                 prior_update="none", learn_gamw=False, lmmse_damp=True,
                                                 layout="dense", iterations=6)
    # dense MLE prior update, L=2 and L=3
    Rs, rs, beta, Nl = ldgen.sim_dense(M=320, N=1600, lam=0.03, h2=0.5, seed=4)
    cm = max(1, int(320 * 0.03))
    cases["dense_L2_mle"] = base(M=320, R=[f32(Rs[0])], r=rs, N_list=Nl, prior_vars=[0, 0.5 / cm],
                                 prior_probs=[0.97, 0.03], prior_update="mle", layout="dense", iterations=6)
    cases["dense_L3_mle"] = base(M=320, R=[f32(Rs[0])], r=rs, N_list=Nl, prior_vars=[0, 0.2 / cm, 2.0 / cm],
                                 prior_probs=[0.96, 0.02, 0.02], prior_update="mle", layout="dense", iterations=6)
    # banded CSR (config-5 regime, reduced)
    R, r, x0, N = ldgen.sim_banded(M=2000, w=40, N_ld=512, N=2000, lam=0.01, h2=0.5, seed=5)
    cm = max(1, int(2000 * 0.01))
    cases["banded_L2_em_s01"] = base(M=2000, R=[f32(R)], r=[r], N_list=[N], s=0.1, prior_vars=[0, 0.5 / cm],
                                     prior_probs=[0.99, 0.01], layout="banded", iterations=8,
                                     x0=x0 / np.sqrt(N))
    # block-diagonal CSR (config-3 regime, reduced), L=3
    R, r, x0, N, starts = ldgen.sim_blockdiag(M=1500, block_lo=60, block_hi=260, N_ld=512, N=3000,
                                              lam=0.01, h2=0.5, seed=6)
    cm = max(1, int(1500 * 0.01))
    cases["blockdiag_L3_em_s01"] = base(M=1500, R=[f32(R)], r=[r], N_list=[N], s=0.1,
                                        prior_vars=[0, 0.1 / cm, 1.0 / cm], prior_probs=[0.99, 0.005, 0.005],
                                        layout="blockdiag", iterations=7)
    # general sparse CSR (irregular pattern: banded minus random entries) -> exercises the CSR kernel
    R, r, x0, N = ldgen.sim_banded(M=1200, w=30, N_ld=512, N=1200, lam=0.02, h2=0.5, seed=7)
    rng = np.random.default_rng(70)
    Rc = R.tocoo()
    keep = (rng.random(Rc.nnz) < 0.35) | (Rc.row == Rc.col)
    Rk = scipy.sparse.csr_matrix((Rc.data[keep], (Rc.row[keep], Rc.col[keep])), shape=R.shape)
    Rk = ((Rk + Rk.T) * 0.5).tocsr()
    Rk.setdiag(1.0)
    Rk = (0.5 * Rk + 0.5 * scipy.sparse.identity(1200)).tocsr()   # keep it PD after thinning
    Rk.sort_indices()
    cm = max(1, int(1200 * 0.02))
    cases["csr_irregular_L2_em"] = base(M=1200, R=[f32(Rk)], r=[r], N_list=[N], s=0.0,
                                        prior_vars=[0, 0.5 / cm], prior_probs=[0.98, 0.02], layout="csr",
                                        iterations=6)
    # K=3 cohorts, dense, shared beta, EM (config-4 regime, reduced)
    Rs, rs, beta, Nl = ldgen.sim_dense(M=256, N=0, lam=0.03, h2=0.5, seed=8, N_list=[800, 1200, 1000])
    cm = max(1, int(256 * 0.03))
    cases["dense_K3_L2_em"] = base(M=256, K=3, R=[f32(R_) for R_ in Rs], r=rs, N_list=Nl,
                                   prior_vars=[0, 0.5 / cm], prior_probs=[0.97, 0.03], layout="dense",
                                   iterations=6, x0=beta)
    # K=2 cohorts with banded LD (different genotype samples and N, shared effects), L=3: the multi-cohort path on the
    # sparse layouts
    R1, r1, x01, N1 = ldgen.sim_banded(M=1000, w=30, N_ld=512, N=1500, lam=0.02, h2=0.5, seed=11)
    R2, r2o, x02, N2 = ldgen.sim_banded(M=1000, w=30, N_ld=512, N=2500, lam=0.02, h2=0.5, seed=12)
    beta = x01 / np.sqrt(N1)
    r2 = R2 @ (beta * np.sqrt(N2)) + (r2o - R2 @ x02)
    cm = max(1, int(1000 * 0.02))
    cases["banded_K2_L3_em_s01"] = base(M=1000, K=2, R=[f32(R1), f32(R2)], r=[r1, r2], N_list=[N1, N2], s=0.1,
                                        prior_vars=[0, 0.1 / cm, 1.0 / cm], prior_probs=[0.98, 0.01, 0.01],
                                        layout="banded", iterations=6, x0=beta)
    # CG that never reaches its tolerance (cg_maxit=4: every solve ends with info = maxiter), warm starts carry on
    R, r, x0, N = ldgen.sim_banded(M=1200, w=25, N_ld=512, N=1200, lam=0.01, h2=0.5, seed=13)
    cm = max(1, int(1200 * 0.01))
    cases["banded_L2_em_cgmaxit4"] = base(M=1200, R=[f32(R)], r=[r], N_list=[N], s=0.1, prior_vars=[0, 0.5 / cm],
                                          prior_probs=[0.99, 0.01], layout="banded", iterations=6, cg_maxit=4)
    # probes NOT injected: np.random.seed + the reference's own binomial call (src/sgvamp.py:326) on the legacy global RNG
    R, r, x0, N = ldgen.sim_banded(M=1500, w=20, N_ld=512, N=1500, lam=0.01, h2=0.5, seed=14)
    cm = max(1, int(1500 * 0.01))
    cases["banded_L2_em_defaultrng"] = base(M=1500, R=[f32(R)], r=[r], N_list=[N], s=0.1, prior_vars=[0, 0.5 / cm],
                                            prior_probs=[0.99, 0.01], layout="banded", iterations=6, rng_seed=20261018)
    # general sparse CSR in a stable regime (all iterations and the CG counts are compared): irregular symmetric
    # pattern = banded sample LD with 60 % of the off-diagonal pairs dropped, diagonally loaded to stay PD
    R, r, x0, N = ldgen.sim_banded(M=1600, w=24, N_ld=512, N=3200, lam=0.01, h2=0.5, seed=15)
    rng = np.random.default_rng(150)
    Ru = scipy.sparse.triu(R, 1).tocoo()
    keep = rng.random(Ru.nnz) < 0.4
    Ru = scipy.sparse.csr_matrix((Ru.data[keep], (Ru.row[keep], Ru.col[keep])), shape=R.shape)
    Rk = (Ru + Ru.T).tocsr()
    load = 0.3 - min(0.0, float(np.linalg.eigvalsh(Rk.toarray())[0]))      # diagonal load: smallest eigenvalue 0.3/load
    Rk = ((Rk + load * scipy.sparse.identity(1600)) * (1.0 / load)).tocsr()  # unit diagonal again
    Rk = f32(Rk)
    Rk.sort_indices()
    # r ~ N(Rk x0, (1-h2) Rk): the summary-statistic form of the reference recipe for this LD
    Lc = np.linalg.cholesky(Rk.toarray())
    r = Rk @ x0 + np.sqrt(0.5) * (Lc @ rng.standard_normal(1600))
    cm = max(1, int(1600 * 0.01))
    cases["csr_irregular_stable_L2_em"] = base(M=1600, R=[f32(Rk)], r=[r], N_list=[N], s=0.0,
                                               prior_vars=[0, 0.5 / cm], prior_probs=[0.99, 0.01], layout="csr",
                                               iterations=7, x0=x0 / np.sqrt(N))
    # a fixture whose LD is NOT fp32-representable is produced at test time from dense_L2_em by
    # perturbing R; the reference perturbation study bounds that effect separately.

    only = sys.argv[1:]
    for name, case in cases.items():
        if only and name not in only:
            continue
        out = run_reference(case)
        save_case(name, case, out)


if __name__ == "__main__":
    main()
