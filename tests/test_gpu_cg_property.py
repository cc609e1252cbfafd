"""Property test of the fused conjugate-gradient step (symmetric half-band layout) against scipy.sparse.linalg.cg.

The fused step decides convergence from |r - alpha q|^2 = r.r - 2 alpha r.q + alpha^2 q.q one pass ahead, where scipy sums
r.r of the updated residual; when the extrapolated value is too close to the threshold to trust, the decision is
postponed to the next pass's exact sum (sgv_device.cuh, AP_CGFUSED).  Parity is defined on the iteration counts
(one flipped count costs a third of the 1e-4 budget), so the counts and `info` of both solves must equal
scipy's on many random SPD bands: smooth spectra, a few distinct eigenvalues (the residual collapses in one step),
warm starts, tight maxiter - with the normal band and with the postponed path forced on every step.
"""
import numpy as np
import pytest
import scipy.sparse
import scipy.sparse.linalg
from hypothesis import given, settings, strategies as st

from golden_util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import sgv_native
    return sgv_native


def _spd_band(M, w, kind, rng):
    """Symmetric positive semi-definite banded matrix with fp32-representable values."""
    if kind == 0:      # B B^T of a random lower band: smooth spectrum, half-bandwidth 2*(w//2)
        hw = max(1, w // 2)
        diags = [rng.standard_normal(M - o) for o in range(hw + 1)]
        B = scipy.sparse.diags(diags, [-o for o in range(hw + 1)], shape=(M, M), format="csr")
        R = (B @ B.T) * (1.0 / (hw + 1))
    elif kind == 1:    # identical small blocks: only a handful of distinct eigenvalues, CG collapses after that many steps
        b = int(rng.integers(2, 6))
        Q = rng.standard_normal((b, b))
        blk = Q @ Q.T / b
        nb = M // b + 1
        R = scipy.sparse.kron(scipy.sparse.identity(nb), blk, format="csr")[:M, :M]
    else:              # AR(1)-like correlation band with a Bartlett taper: the shape of real LD
        rho = rng.uniform(0.3, 0.95)
        offs = np.arange(-w, w + 1)
        diags = [np.full(M - abs(o), rho ** abs(o) * (1.0 - abs(o) / (w + 1.0))) for o in offs]
        R = scipy.sparse.diags(diags, offs, shape=(M, M), format="csr")
    R = R.tocsr()
    R.data = R.data.astype(np.float32).astype(np.float64)
    R = ((R + R.T) * 0.5).tocsr()          # exact symmetry after rounding
    R.sort_indices()
    return R


def _scipy_cg(A, b, x0, maxit):
    n = [0]

    def cb(_x):
        n[0] += 1

    x, info = scipy.sparse.linalg.cg(A, b, maxiter=maxit, x0=x0, callback=cb)
    return x, info, n[0]


def _count_is_stable(A, b, x0, maxit, n_ref):
    """True when scipy's iteration count does not hinge on the last bit: the same recursion with q perturbed at
    rounding level (3 eps, three different patterns) stops after the same number of steps.  Where it does hinge on
    it (|r| within rounding-amplified reach of the threshold), no two correct implementations can be expected to
    agree - scipy built against another BLAS would not agree with itself."""
    M = b.shape[0]
    atol = 1e-5 * np.linalg.norm(b)
    if atol == 0.0:
        return True
    for trial in range(3):
        prng = np.random.default_rng(1000 + trial)
        x = np.array(x0, dtype=np.float64)
        r = b - A @ x if x.any() else b.copy()
        rho_prev, p, n = None, None, maxit
        for it in range(maxit):
            if np.linalg.norm(r) < atol:
                n = it
                break
            rho = r @ r
            p = r.copy() if it == 0 else r + (rho / rho_prev) * p
            q = (A @ p) * (1.0 + 6.6e-16 * prng.standard_normal(M))
            al = rho / (p @ q)
            x += al * p
            r -= al * q
            rho_prev = rho
        if n != n_ref:
            return False
    return True


def _check(nat, seed, kind, M, w, warm_scale, maxit, loggam2):
    rng = np.random.default_rng(seed)
    R = _spd_band(M, w, kind, rng)
    gamw = float(rng.uniform(0.5, 5.0))
    gam2 = float(10.0 ** loggam2)
    xty, xhat1, r1 = rng.standard_normal(M), rng.standard_normal(M), rng.standard_normal(M)
    x2p = rng.standard_normal(M) * warm_scale
    sgp = rng.standard_normal(M) * warm_scale
    u = rng.integers(0, 2, M) * 2 - 1
    alpha1 = 0.3
    h = nat.Handle()
    h.configure(M, 1)
    h._ck(h.upload_csr(0, R.indptr, R.indices, R.data, layout=nat.LAYOUT_DSYM))
    assert h.ld_info(0)["layout"] == "dsym"
    h.set_xty(0, xty)
    h.set_vec(0, nat.VEC_XHAT1, xhat1)
    h.set_vec(0, nat.VEC_R1, r1)
    h.set_vec(0, nat.VEC_XHAT2, x2p)
    h.set_vec(0, nat.VEC_SIGMA2U, sgp)
    out = h.lmmse(0, gamw, gam2, alpha1, 0.5, maxit, False, True, warm_scale == 0.0, u)
    A = (gamw * R + gam2 * scipy.sparse.identity(M)).tocsr()
    mu2 = gamw * xty + gam2 * (xhat1 - alpha1 * r1) / (1 - alpha1)
    x2, i1, n1 = _scipy_cg(A, mu2, x2p, maxit)
    sg, i2, n2 = _scipy_cg(A, u.astype(np.float64), sgp, maxit)
    ctx = (seed, kind, M, w, warm_scale, maxit, loggam2)
    checked = 0
    for col, (b, x0, xs, info, n) in enumerate([(mu2, x2p, x2, i1, n1), (u.astype(np.float64), sgp, sg, i2, n2)]):
        if not _count_is_stable(A, b, x0, maxit, n):
            continue                       # scipy's own count hinges on the last bit here: nothing to compare
        checked += 1
        assert (out.cg_iters[col], out.cg_info[col]) == (n, info), (ctx, col, out.cg_iters[col], out.cg_info[col], n, info)
        if info == 0 and n > 0:
            got = h.get_vec(0, nat.VEC_XHAT2 if col == 0 else nat.VEC_SIGMA2U)
            assert rel_l2(got, xs) < 1e-5 * max(1.0, warm_scale), (ctx, col)
    h.close()
    return checked


_CHECKED = [0]
CASES = dict(seed=st.integers(0, 10**6), kind=st.integers(0, 2), M=st.integers(200, 6000), w=st.integers(1, 48),
             warm_scale=st.sampled_from([0.0, 1.0, 30.0]), maxit=st.sampled_from([2, 7, 500]),
             loggam2=st.floats(-3.0, 1.0))


@settings(max_examples=80, deadline=None, derandomize=True)
@given(**CASES)
def test_fused_cg_counts_equal_scipy(nat, seed, kind, M, w, warm_scale, maxit, loggam2):
    _CHECKED[0] += _check(nat, seed, kind, M, w, warm_scale, maxit, loggam2)


def test_fused_cg_enough_solves_were_compared():
    """(runs after the property test) at least 100 of its 160 solves had a well-defined count and were compared"""
    assert _CHECKED[0] >= 100, _CHECKED[0]


@pytest.mark.parametrize("band", ["1e-3", "10"])
def test_fused_cg_counts_with_postponed_decisions(nat, monkeypatch, band):
    """SGV_CG_BAND widens the 'too close to call' band: 1e-3 postpones every decision near convergence, 10 postpones
    every single one - the exact-sum path alone must reproduce scipy's counts too."""
    monkeypatch.setenv("SGV_CG_BAND", band)
    rng = np.random.default_rng(11)
    for i in range(12):
        _check(nat, int(rng.integers(0, 10**6)), i % 3, int(rng.integers(300, 4000)), int(rng.integers(2, 40)),
               [0.0, 1.0, 30.0][i % 3], [500, 500, 3][(i // 3) % 3], float(rng.uniform(-3, 1)))
