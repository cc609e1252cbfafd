"""CPU oracle for the sgVAMP hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product path (``sgvamp-py_b200/``) never
imports it and has no CPU fallback.

This is a numpy restatement of the reference algorithm (pure Python, so there is no C to
compile).  Every function cites the reference lines it follows; citations are relative to
``/root/reference``.  Third-party arithmetic the reference delegates to is restated from the
versions installed in this image (the reference pins none): ``scipy.sparse.linalg.cg`` as in
scipy 1.18.1 ``_isolve/iterative.py`` (see :func:`cg`), ``scipy.optimize.fsolve`` (called, not
restated - MINPACK hybrd stays third-party on both sides), ``numpy.random.binomial``.

Pinning: the reference has no tests or golden vectors.  The oracle is pinned against outputs
of the *unmodified reference executed in the build container* (``tests/golden/make_golden.py``
imports ``/root/reference/src/sgvamp.py`` and stores its per-iteration outputs under
``tests/golden/*.npz``); ``tests/test_oracle_vs_golden.py`` checks this file against them.

Differences from the reference that are deliberate and do not change results beyond rounding:
  * all K cohorts run in one process (the reference runs one MPI rank per cohort and
    broadcasts ``gam1`` / ``r1`` so that every rank holds identical copies, src/sgvamp.py:228-233);
  * the per-marker Python loops (src/sgvamp.py:273,285) are evaluated for all markers at once;
    ``per_marker=True`` keeps the literal one-marker-at-a-time evaluation for cross-checking;
  * ``A = gamw*R + gam2*I`` (src/sgvamp.py:312) is applied as an operator instead of being
    materialised (``materialise_A=True`` restores the reference behaviour for timing).
"""
from __future__ import annotations

import csv
import os
import time

import numpy as np
import scipy.optimize
import scipy.sparse
import scipy.sparse.linalg


# --------------------------------------------------------------------------------------
# scipy.sparse.linalg.cg, restated (scipy 1.18.1, _isolve/iterative.py: cg)
# --------------------------------------------------------------------------------------
def cg(matvec, b, x0, maxiter, rtol=1e-5, atol=0.0):
    """Conjugate gradients with scipy's exact control flow.

    Follows the call sites src/sgvamp.py:316,332 (no rtol/atol passed -> rtol=1e-5, atol=0).
    Returns (x, info, n_updates) where info is 0 (converged test hit at loop top) or
    ``maxiter`` (loop exhausted, no final test), and n_updates counts matvecs in the loop.
    """
    b = np.asarray(b, dtype=np.float64).ravel()
    x = np.array(x0, dtype=np.float64).ravel().copy()
    bnrm2 = np.linalg.norm(b)
    atol_eff = max(float(atol), float(rtol) * float(bnrm2))
    if bnrm2 == 0:
        return b.copy(), 0, 0
    r = b - matvec(x) if x.any() else b.copy()
    rho_prev, p = None, None
    n_updates = 0
    for iteration in range(maxiter):
        if np.linalg.norm(r) < atol_eff:
            return x, 0, n_updates
        rho_cur = np.dot(r, r)
        if iteration > 0:
            beta = rho_cur / rho_prev
            p *= beta
            p += r
        else:
            p = r.copy()
        q = matvec(p)
        alpha = rho_cur / np.dot(p, q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho_cur
        n_updates += 1
    return x, maxiter, n_updates


# --------------------------------------------------------------------------------------
# prior / denoiser pieces
# --------------------------------------------------------------------------------------
class Prior:
    """Prior parametrisation of src/sgvamp.py:21-28."""

    def __init__(self, prior_vars, prior_probs, Nt):
        self.L = len(prior_probs)
        self.lam = 1 - prior_probs[0]
        self.sigmas = np.array(prior_vars[1:], dtype=np.float64) * Nt
        self.omegas = np.array([p / sum(prior_probs[1:]) for p in prior_probs[1:]], dtype=np.float64)
        self.gam = None  # MLE Lagrange multiplier memo, src/sgvamp.py:31


def denoise_all(r1s, gam1s, a, prior, per_marker=False):
    """Posterior mean and rank-independent derivative factor for every marker.

    Follows denoiser_meta / der_denoiser_meta, src/sgvamp.py:93-114.
    Returns (xhat (M,), dfac (M,)) with  der_k[j] = a[k]*gam1s[k]*dfac[j]  (the rank enters
    src/sgvamp.py:112-113 only through that scalar factor).
    """
    r1s = np.asarray(r1s, dtype=np.float64)
    K, M = r1s.shape
    lam, omegas, sigmas = prior.lam, prior.omegas, prior.sigmas
    w = a * gam1s
    if per_marker:
        xhat = np.empty(M)
        dfac = np.empty(M)
        for j in range(M):
            xhat[j], dfac[j] = _denoise_one(r1s[:, j], w, lam, omegas, sigmas)
        return xhat, dfac
    sigma2 = 1.0 / (np.sum(w) + 1.0 / sigmas)                      # (L-1,)   :95
    mu = (w @ r1s)[:, None] * sigma2[None, :]                       # (M,L-1)  :96
    score = mu * mu / sigma2[None, :]                               #          :97
    mi = score.argmax(axis=1)
    rows = np.arange(M)
    mu_max = mu[rows, mi][:, None]
    s2_max = sigma2[mi][:, None]
    EXP = np.exp(0.5 * (mu * mu * s2_max - mu_max * mu_max * sigma2[None, :])
                 / (sigma2[None, :] * s2_max))                       #          :98
    sq = np.sqrt(sigma2 / sigmas)[None, :]
    Num = lam * np.sum(omegas[None, :] * EXP * mu * sq, axis=1)      #          :99
    EXP2 = np.exp(-0.5 * (mu_max[:, 0] ** 2 / s2_max[:, 0]))         #          :100
    Den = (1 - lam) * EXP2 + lam * np.sum(omegas[None, :] * EXP * sq, axis=1)  # :101
    DerNum = lam * np.sum(omegas[None, :] * EXP * (mu * mu + sigma2[None, :]) * sq, axis=1)  # :112 / w_rank
    DerDen = lam * np.sum(omegas[None, :] * mu * EXP * sq, axis=1)                             # :113 / w_rank
    return Num / Den, (DerNum * Den - DerDen * Num) / (Den * Den)


def _denoise_one(rs, w, lam, omegas, sigmas, w_rank=None):
    """Literal single-marker evaluation of src/sgvamp.py:93-114.  Without ``w_rank`` the derivative is returned
    without the factor a[rank]*gam1s[rank]; with ``w_rank = (a[rank], gam1s[rank])`` the two factors sit
    inside the sums exactly where the reference puts them (:112-113): bit-identical to der_denoiser_meta."""
    sigma2_meta = 1.0 / (sum(w) + 1.0 / sigmas)
    mu_meta = np.inner(rs, w) * sigma2_meta
    max_ind = (np.array(mu_meta * mu_meta / sigma2_meta)).argmax()
    EXP = np.exp(0.5 * (mu_meta * mu_meta * sigma2_meta[max_ind]
                        - mu_meta[max_ind] * mu_meta[max_ind] * sigma2_meta)
                 / (sigma2_meta * sigma2_meta[max_ind]))
    Num = lam * sum(omegas * EXP * mu_meta * np.sqrt(sigma2_meta / sigmas))
    EXP2 = np.exp(-0.5 * ((mu_meta[max_ind]) ** 2 / sigma2_meta[max_ind]))
    Den = (1 - lam) * EXP2 + lam * sum(omegas * EXP * np.sqrt(sigma2_meta / sigmas))
    if w_rank is not None:
        DerNum = lam * sum(omegas * EXP * (mu_meta * mu_meta + sigma2_meta) * w_rank[0] * w_rank[1] * np.sqrt(sigma2_meta / sigmas))
        DerDen = lam * sum(omegas * mu_meta * EXP * w_rank[0] * w_rank[1] * np.sqrt(sigma2_meta / sigmas))
        return Num / Den, (DerNum * Den - DerDen * Num) / (Den * Den)
    DerNum = lam * sum(omegas * EXP * (mu_meta * mu_meta + sigma2_meta) * np.sqrt(sigma2_meta / sigmas))
    DerDen = lam * sum(omegas * mu_meta * EXP * np.sqrt(sigma2_meta / sigmas))
    return Num / Den, (DerNum * Den - DerDen * Num) / (Den * Den)


def prior_update_em(r1s, gam1s, a, prior):
    """One EM pass, src/sgvamp.py:116-136.  Mutates ``prior.lam`` / ``prior.omegas``."""
    K, M = r1s.shape
    Lm1 = prior.L - 1
    pv = prior.sigmas.reshape(1, 1, Lm1)
    g = gam1s.reshape(K, 1, 1)
    ginv = 1.0 / g
    r2 = np.power(r1s.reshape(K, M, 1), 2)
    e = -r2 / 2 / (pv + ginv)
    exp_max = e.max(axis=2).reshape(K, M, 1)                                            # :127
    xi = prior.lam * prior.omegas.reshape(1, 1, Lm1) * np.exp(e - exp_max) / np.sqrt(ginv + pv)  # :128
    sum_xi = xi.sum(axis=2).reshape(K, M, 1)
    xi_tilde = xi / sum_xi
    pi = 1.0 / (1.0 + (1 - prior.lam) * np.exp(-r2 / 2 * g - exp_max) / np.sqrt(ginv) / sum_xi)  # :131
    prior.lam = np.mean(np.average(pi, axis=0, weights=a))                              # :134
    aw = a.reshape(K, 1, 1)
    prior.omegas = np.sum(pi * xi_tilde * aw, axis=(0, 1)) / np.sum(pi * aw, axis=(0, 1))  # :136


def em_loop(r1s, gam1s, a, prior, em_prior_maxit):
    """The EM driver loop of src/sgvamp.py:250-257.  Returns (steps, final relative error)."""
    steps, rel = 0, 0.0
    for em_it in range(em_prior_maxit):
        old_omegas, old_lam = prior.omegas, prior.lam
        prior_update_em(r1s, gam1s, a, prior)
        om_err = np.linalg.norm(prior.omegas - old_omegas) / np.linalg.norm(old_omegas)
        lam_err = np.abs(prior.lam - old_lam) / prior.lam
        steps, rel = em_it + 1, max(om_err, lam_err)
        if om_err < 1e-6 and lam_err < 1e-6:
            break
    return steps, rel


def lagrangian_der(x, omega0, sigma2, r1s, gam1s, a, L):
    """Residual of the MLE stationarity system, src/sgvamp.py:139-160."""
    K, M = r1s.shape
    y = np.zeros(L + 1)
    omega = x[:L]
    gam = x[L]
    pv = sigma2.reshape(1, 1, L)
    ginv = 1.0 / gam1s.reshape(K, 1, 1)
    r2 = np.power(r1s.reshape(K, M, 1), 2)
    e = -r2 / 2 / (pv + ginv)
    exp_max = e.max()                                               # :153 (global shift)
    probs = np.exp(e - exp_max) / np.sqrt(pv + ginv)                # :154
    Num = a.reshape(K, 1, 1) * probs
    Den = np.sum(probs * omega.reshape(1, 1, L), axis=2).reshape(K, M, 1)
    y[:L] = np.sum(Num / Den, axis=(0, 1)) + (omega0 - 1) / omega + gam
    y[L] = sum(omega) - 1.0
    return y


def prior_update_mle(r1s, gam1s, a, prior):
    """MLE prior update via fsolve with accept/reject, src/sgvamp.py:162-194.

    Returns one of "ok", "not_converged", "negative".
    """
    L = prior.L
    omega0 = np.zeros(L)
    omega0[0] = 1 - prior.lam
    omega0[1:] = prior.lam * prior.omegas
    sigma2 = np.zeros(L)
    sigma2[0] = 1e-16
    sigma2[1:] = prior.sigmas
    x0 = np.zeros(L + 1)
    x0[:-1] = omega0
    x0[-1] = 1 if prior.gam is None else prior.gam
    x, _, ier, _ = scipy.optimize.fsolve(
        func=lambda x_: lagrangian_der(x_, omega0, sigma2, r1s, gam1s, a, L), x0=x0, full_output=True)
    if ier != 1:
        return "not_converged"
    if any(s <= 0 for s in x[:-1]):
        return "negative"
    x[:-1] /= sum(x[:-1])
    prior.lam = 1 - x[0]
    prior.omegas = np.array([w / sum(x[1:-1]) for w in x[1:-1]])
    prior.gam = x[L]
    return "ok"


# --------------------------------------------------------------------------------------
# the VAMP loop
# --------------------------------------------------------------------------------------
def default_probe(cohort, it, M):
    """Rademacher probe exactly as src/sgvamp.py:326 (numpy legacy global RNG)."""
    return np.random.binomial(p=1 / 2, n=1, size=M) * 2 - 1


class VAMPOracle:
    """All-cohorts-in-one-process restatement of class VAMP (src/sgvamp.py:14-389)."""

    def __init__(self, N_list, M, rho, gamw, gam1, prior_vars, prior_probs,
                 out_dir=None, out_name=None):
        self.N_list = [float(n) for n in np.atleast_1d(N_list)]
        self.K = len(self.N_list)
        self.Nt = sum(self.N_list)
        self.M = M
        self.rho = rho
        self.gamw0 = gamw
        self.gam10 = gam1
        self.a = np.array(self.N_list) / self.Nt                    # src/main.py:287
        self.prior = Prior(prior_vars, prior_probs, self.Nt)
        self.out_dir, self.out_name = out_dir, out_name
        if out_dir is not None:
            self._setup_io()

    # -- output files, byte-compatible with src/sgvamp.py:33-76 --------------------------
    def _setup_io(self):
        for i in range(self.K):
            with open(os.path.join(self.out_dir, "%s_cohort_%d.csv" % (self.out_name, i + 1)), "w", newline="") as f:
                csv.writer(f, delimiter="\t").writerow(["it", "gamw", "gam1", "gam2", "alpha1", "alpha2", "lam"])
        with open(os.path.join(self.out_dir, "%s_metrics.csv" % self.out_name), "w", newline="") as f:
            csv.writer(f, delimiter="\t").writerow(["it", "alignment", "l2"])

    def _append(self, fname, row):
        with open(os.path.join(self.out_dir, fname), "a", newline="") as f:
            csv.writer(f, delimiter="\t").writerow(row)

    def infer(self, R_list, r_list, iterations, x0=None, cg_maxit=500, em_prior_maxit=100,
              learn_gamw=True, lmmse_damp=True, prior_update=None, update_prior_from=1,
              probe_fn=default_probe, per_marker=False, materialise_A=False, timers=None, threads=1, n_probes=1):
        """Follows VAMP.infer, src/sgvamp.py:196-389, for all K cohorts at once.

        ``R_list[k]`` is Rused of cohort k (already regularised as in src/main.py:265), as a
        scipy sparse matrix or a dense ndarray.  Returns a dict of per-iteration records.
        """
        K, M, Nt, rho, a, prior = self.K, self.M, self.Nt, self.rho, self.a, self.prior
        R_list = [R if scipy.sparse.issparse(R) else np.asarray(R) for R in R_list]
        r = [np.asarray(rk, dtype=np.float64).reshape(M) for rk in r_list]
        r1 = [rk.copy() for rk in r]
        xhat1 = np.zeros(M)
        xhat2 = [np.zeros(M) for _ in range(K)]
        Sig_prev = [np.zeros(M) for _ in range(K)]
        gam1 = [self.gam10] * K
        gamw = [self.gamw0] * K
        alpha1 = [0.0] * K
        alpha2 = [0.0] * K
        I = scipy.sparse.identity(M)
        out = dict(xhat1=[], r1_in=[], rows=[], cg_iters=[], cg_info=[], lam=[], omegas=[],
                   em_steps=[], mle_status=[], metrics=[], gamw_raw=[])
        tm = timers if timers is not None else {}
        pool = None
        if threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            pool = ThreadPoolExecutor(threads)

        def split_rows(A):
            # row-partitioned operator for the thread pool: every row's sum is formed exactly as in
            # the single-threaded csr_matvec, so results are bit-identical
            if pool is None or not scipy.sparse.issparse(A):
                return None
            b = np.linspace(0, A.shape[0], threads + 1).astype(int)
            return [A[b[i]:b[i + 1]] for i in range(threads)]

        def apply(A, parts, v):
            if parts is None:
                return np.asarray(A @ v).ravel()
            return np.concatenate(list(pool.map(lambda P: P @ v, parts)))

        R_parts = [split_rows(R) for R in R_list]

        def tick(name, t0):
            tm[name] = tm.get(name, 0.0) + (time.perf_counter() - t0)

        for it in range(iterations):
            gam1s = np.array(gam1, dtype=np.float64)                 # :228-233
            r1s = np.stack(r1)
            # prior update :242-259
            t0 = time.perf_counter()
            em_steps, mle_status = 0, None
            if it >= update_prior_from:
                if prior_update == "mle":
                    mle_status = prior_update_mle(r1s, gam1s, a, prior)
                elif prior_update == "em":
                    em_steps, _ = em_loop(r1s, gam1s, a, prior, em_prior_maxit)
            tick("prior", t0)
            out["em_steps"].append(em_steps)
            out["mle_status"].append(mle_status)
            out["lam"].append(float(prior.lam))
            out["omegas"].append(np.array(prior.omegas, dtype=np.float64).copy())

            # denoising :270-293
            t0 = time.perf_counter()
            xhat1_prev = xhat1
            alpha1_prev = list(alpha1)
            xhat1, dfac = denoise_all(r1s, gam1s, a, prior, per_marker=per_marker)
            if it > 0:
                xhat1 = rho * xhat1 + (1 - rho) * xhat1_prev          # :275-276
            tick("denoise", t0)
            out["xhat1"].append(xhat1.copy())
            out["r1_in"].append(r1s.copy())
            if self.out_dir is not None:
                (xhat1 / np.sqrt(Nt)).tofile(os.path.join(self.out_dir, "%s_xhat_it_%d.bin" % (self.out_name, it)))
                for k in range(K):
                    (r1[k] / np.sqrt(Nt)).tofile(
                        os.path.join(self.out_dir, "%s_r1_cohort_%d_it_%d.bin" % (self.out_name, k + 1, it)))
            dmean = np.mean(dfac)
            rows_it, iters_it, info_it, gamw_raw_it = [], [], [], []
            for k in range(K):
                a1 = a[k] * gam1s[k] * dmean                         # :285
                if it > 0:
                    a1 = rho * a1 + (1 - rho) * alpha1_prev[k]       # :290-291  (clip at :293 is a no-op)
                alpha1[k] = a1
                # LMMSE :301-323
                t0 = time.perf_counter()
                R = R_list[k]
                N = self.N_list[k]
                xhat2_prev, alpha2_prev = xhat2[k], alpha2[k]
                gam2 = gam1[k] * (1 - a1) / a1                       # :305
                r2 = (xhat1 - a1 * r1[k]) / (1 - a1)                 # :310
                gw = gamw[k]
                if materialise_A:
                    A = gw * R + gam2 * I                            # :312
                    A = np.asarray(A) if not scipy.sparse.issparse(A) else A.tocsr()
                    A_parts = split_rows(A)
                    mv = lambda v, A=A, A_parts=A_parts: apply(A, A_parts, v)
                else:
                    mv = lambda v, R=R, gw=gw, gam2=gam2, P=R_parts[k]: gw * apply(R, P, v) + gam2 * v
                mu2 = gw * r[k] + gam2 * r2                          # :313
                tick("lmmse_setup", t0)
                t0 = time.perf_counter()
                x2, info1, n1 = cg(mv, mu2, xhat2_prev, cg_maxit)    # :316
                if lmmse_damp:
                    x2 = rho * x2 + (1 - rho) * xhat2_prev           # :322-323 (no it>0 guard)
                u = np.asarray(probe_fn(k, it, M))                   # :326
                Sig, info2, n2 = cg(mv, u, Sig_prev[k], cg_maxit)    # :332
                tick("cg", t0)
                t0 = time.perf_counter()
                Sig_prev[k] = Sig
                uSu = u @ Sig
                # n_probes > 1 (an extension the reference does not have; n_probes = 1 is the reference):
                # further probes probe_fn(k, it, M, p), each solved from x0 = 0, averaged into both trace estimates
                extra = []
                for p_ in range(1, n_probes):
                    up = np.asarray(probe_fn(k, it, M, p_))
                    sp, _, _ = cg(mv, up, np.zeros(M), cg_maxit)
                    extra.append((up, sp))
                    uSu = uSu + up @ sp
                uSu = uSu / n_probes
                a2 = gam2 * uSu / M                                  # :338-340
                if lmmse_damp:
                    a2 = rho * a2 + (1 - rho) * alpha2_prev          # :345-346
                gam1_new = gam2 * (1 - a2) / a2                      # :347
                r1[k] = (x2 - a2 * r2) / (1 - a2)                    # :348
                gw_new = gw
                if learn_gamw:                                       # :350-364
                    z = N - 2 * (x2 @ r[k]) + x2 @ apply(R, R_parts[k], x2)
                    if z < 0:
                        z = 0
                    TrRS = u @ apply(R, R_parts[k], Sig)
                    for up, sp in extra:
                        TrRS = TrRS + up @ apply(R, R_parts[k], sp)
                    TrRS = TrRS / n_probes
                    gw_new = float(1 / (z / N + TrRS / N))
                gamw_raw_it.append(gw_new)
                gw_new = max(gw_new, 1.0)                            # :374
                tick("gamw", t0)
                xhat2[k], alpha2[k], gam1[k], gamw[k] = x2, a2, gam1_new, gw_new
                row = [it, gw_new, gam1_new, gam2, a1, a2, float(prior.lam)]
                rows_it.append(row)
                iters_it.append((n1, n2))
                info_it.append((info1, info2))
                if self.out_dir is not None:
                    self._append("%s_cohort_%d.csv" % (self.out_name, k + 1), row)
            out["rows"].append(rows_it)
            out["cg_iters"].append(iters_it)
            out["cg_info"].append(info_it)
            out["gamw_raw"].append(gamw_raw_it)
            if x0 is not None:                                       # :379-387
                t = np.asarray(x0, dtype=np.float64).ravel()
                alignment = np.inner(xhat1, t) / np.linalg.norm(xhat1) / np.linalg.norm(t)
                l2 = np.linalg.norm(xhat1 - t) / np.linalg.norm(t)
                out["metrics"].append((it, alignment, l2))
                if self.out_dir is not None:
                    self._append("%s_metrics.csv" % self.out_name, [it, alignment, l2])
        if pool is not None:
            pool.shutdown()
        return out


def regularise(R, s):
    """Rused = (1-s) R + s I, src/main.py:265 (dense input stays an ndarray here)."""
    M = R.shape[0]
    if scipy.sparse.issparse(R):
        return ((1 - s) * R + s * scipy.sparse.identity(M)).tocsr()
    return np.asarray((1 - s) * R + s * scipy.sparse.identity(M))
