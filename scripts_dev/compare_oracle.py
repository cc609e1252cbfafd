import sys, os, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import torch, bench, sgvamp
from oracle import sgvamp_oracle as orc
M = int(sys.argv[1]); w = int(sys.argv[2]); iters = int(sys.argv[3]); rho = float(sys.argv[4])
dev = torch.device("cuda", 0)
band, ldb, r, x0, tg = bench.build_problem(torch, M, w, 5, dev)
Rh, keep = bench.band_to_host_csr(torch, band, M, w, pinned=False)
p = bench.vamp_params(M); N = bench.n_gwas(M)
probes = bench.make_probes(iters, M, 5)
v = sgvamp.VAMP(N=N, Nt=N, M=M, K=1, rho=rho, gamw=p["gamw"], gam1=p["gam1"], a=np.array([1.0]),
                prior_vars=p["prior_vars"], prior_probs=p["prior_probs"], out_dir=None, out_name="d")
xs = v.infer(sgvamp.DeviceDIA(band.data_ptr(), w, ldb), r, iters, cg_maxit=500, lmmse_damp=False, prior_update="em", probes=probes)
o = orc.VAMPOracle([N], M, rho, p["gamw"], p["gam1"], p["prior_vars"], p["prior_probs"])
t0 = time.time()
R64 = Rh.astype(np.float64)
ref = o.infer([R64], [r], iters, cg_maxit=500, lmmse_damp=False, prior_update="em", probe_fn=lambda k, it, M_: probes[k, it], threads=16)
print("oracle %.1fs" % (time.time() - t0))
for it in range(iters):
    g, rr = v.history["rows"][it][0], ref["rows"][it][0]
    e = np.linalg.norm(xs[it].ravel() - ref["xhat1"][it]) / np.linalg.norm(ref["xhat1"][it])
    se = max(abs(g[i] - rr[i]) / abs(rr[i]) for i in range(1, 7))
    al = float(np.dot(ref["xhat1"][it], x0) / np.linalg.norm(ref["xhat1"][it]) / np.linalg.norm(x0))
    print("it %2d xhat relL2 %.2e scal %.2e | gpu cg %s ref cg %s | ref gamw %.4g lam %.4g align %.4f" % (
        it, e, se, v.history["cg_iters"][it][0], ref["cg_iters"][it][0], rr[1], rr[6], al))
