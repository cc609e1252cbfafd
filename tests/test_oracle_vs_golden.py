"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz).

Bar (BASELINE.json north_star): per-iteration xhat rel-L2 <= 1e-4 and gamw/gam1/gam2/alpha1/alpha2
within 1e-4 relative.  The oracle is fp64 like the reference, so we hold it to 1e-8 here (it
only differs by summation order / operator-vs-materialised A).
"""
import numpy as np
import pytest

from golden_util import ALL_CASES, UNSTABLE, load_case, rel_err, rel_l2
from oracle import sgvamp_oracle as orc


def run_oracle(c, **kw):
    o = orc.VAMPOracle(c["N_list"], c["M"], c["rho"], c["gamw"], c["gam1"], c["prior_vars"], c["prior_probs"])
    Rused = [orc.regularise(R, c["s"]) for R in c["R"]]
    pu = c["prior_update"]
    x0 = c["x0"] * np.sqrt(c["N_list"][0]) if "x0" in c else None
    return o.infer(Rused, list(c["r"]), c["iterations"], x0=x0, cg_maxit=c["cg_maxit"],
                   em_prior_maxit=c["em_prior_maxit"], learn_gamw=c["learn_gamw"], lmmse_damp=c["lmmse_damp"],
                   prior_update=pu, update_prior_from=c["update_prior_from"],
                   probe_fn=lambda k, it, M: c["probes"][k, it], **kw)


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_matches_reference(name):
    c = load_case(name)
    out = run_oracle(c)
    tol = 1e-6 if name in UNSTABLE else 1e-8
    Nt = sum(c["N_list"])
    for it in range(c["iterations"]):
        assert rel_l2(out["xhat1"][it], c["xhat"][it]) < tol, (name, it)
        assert rel_l2(out["xhat1"][it] / np.sqrt(Nt), c["xhat_dump"][it]) < tol
        for k in range(c["K"]):
            assert rel_l2(out["r1_in"][it][k] / np.sqrt(Nt), c["r1_dump"][it, k]) < tol
            assert rel_err(out["rows"][it][k][1:], c["rows"][it, k, 1:]) < tol, (name, it, k)
            assert tuple(out["cg_iters"][it][k]) == tuple(c["cg_iters"][it, k]), (name, it, k)
            assert tuple(out["cg_info"][it][k]) == tuple(c["cg_info"][it, k])
    assert rel_err(out["lam"][-1], c["final_lam"]) < tol
    assert rel_err(out["omegas"][-1], c["final_omegas"]) < tol
    if "metrics" in c:
        assert rel_err(np.array(out["metrics"])[:, 1:], c["metrics"][:, 1:]) < 1e-8


@pytest.mark.parametrize("name", ["dense_L2_em", "dense_K3_L2_em"])
def test_per_marker_and_materialised_modes(name):
    c = load_case(name)
    out = run_oracle(c, per_marker=True, materialise_A=True)
    for it in range(c["iterations"]):
        assert rel_l2(out["xhat1"][it], c["xhat"][it]) < 1e-9


def test_cg_restatement_matches_scipy():
    import scipy.sparse.linalg as sla
    rng = np.random.default_rng(0)
    B = rng.standard_normal((300, 200))
    A = B.T @ B / 300 + 0.05 * np.eye(200)
    for trial in range(5):
        b = rng.standard_normal(200)
        x0 = rng.standard_normal(200) * (trial % 2)
        for maxiter in (3, 50):
            cnt = [0]
            xs, info_s = sla.cg(A, b, x0=x0, maxiter=maxiter, callback=lambda _x: cnt.__setitem__(0, cnt[0] + 1))
            xo, info_o, n = orc.cg(lambda v: A @ v, b, x0, maxiter)
            assert info_s == info_o and n == cnt[0]
            assert rel_l2(xo, xs) < 1e-12
    x, info, n = orc.cg(lambda v: A @ v, np.zeros(200), np.ones(200), 10)
    assert info == 0 and n == 0 and not x.any()
