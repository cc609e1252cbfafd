import sys, os, time, json
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import torch, bench, sgvamp, ldgen
M = int(sys.argv[1]); w = int(sys.argv[2]); iters = int(sys.argv[3])
pu = sys.argv[4] if len(sys.argv) > 4 else "em"
rho = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
s_reg = float(sys.argv[6]) if len(sys.argv) > 6 else 0.1
ldrho = float(sys.argv[7]) if len(sys.argv) > 7 else 0.97
nfac = float(sys.argv[8]) if len(sys.argv) > 8 else 2.0
bench.S_REG = s_reg
bench.n_gwas = lambda M_: int(nfac * M_)
_g = ldgen.genotypes_device
ldgen.genotypes_device = lambda torch_, n, lo, hi, seed, dev, rho=ldrho, B=256: _g(torch_, n, lo, hi, seed, dev, rho=ldrho, B=B)
dev = torch.device("cuda", 0)
band, ldb, r, x0, tg = bench.build_problem(torch, M, w, 5, dev)
p = bench.vamp_params(M)
probes = bench.make_probes(iters, M, 5)
N = bench.n_gwas(M)
v = sgvamp.VAMP(N=N, Nt=N, M=M, K=1, rho=rho, gamw=p["gamw"], gam1=p["gam1"], a=np.array([1.0]),
                prior_vars=p["prior_vars"], prior_probs=p["prior_probs"], out_dir=None, out_name="d")
t0 = time.time()
xs = v.infer(sgvamp.DeviceDIA(band.data_ptr(), w, ldb), r, iters, cg_maxit=500, lmmse_damp=False, prior_update=pu, probes=probes)
print("== M=%d w=%d pu=%s rho=%g s=%g ldrho=%g N/M=%g  offdiag1=%.3f infer %.3fs" % (M, w, pu, rho, s_reg, ldrho, nfac, band[w+1, :M-1].mean().item()/(1-s_reg), time.time() - t0))
for it in range(iters):
    row = v.history["rows"][it][0]
    al = float(np.dot(xs[it].ravel(), x0) / np.linalg.norm(xs[it]) / np.linalg.norm(x0))
    print("it %2d gamw %.4g gam1 %.4g gam2 %.4g a1 %.4g a2 %.4g lam %.4g | cg %s em %d | align %.4f |xhat| %.3g" % (
        it, row[1], row[2], row[3], row[4], row[5], row[6], v.history["cg_iters"][it][0], v.history["em_steps"][it], al, np.linalg.norm(xs[it])))
