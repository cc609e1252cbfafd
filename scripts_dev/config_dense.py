"""BASELINE.json configs[1]-like run: M=50k dense LD, K=1, L=4, cg-maxit=50, learned gamw, one B200.
Reference recipe (simulation/sim_gen_phen_mult.py:28-55) with the genotype matrix generated in row chunks on the
GPU (torch as a data-generation utility): X ~ Binomial(2, 0.4) standardised, R = X^T X / N, r = X^T y / sqrt(N)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import numpy as np, torch
import sgvamp
M = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * M
its = 10
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(2)
lam, h2 = 0.01, 0.5
cm = int(M * lam)
beta = torch.zeros(M, device=dev, dtype=torch.float64)
idx = torch.randperm(M, generator=g, device=dev)[:cm]
beta[idx] = torch.randn(cm, generator=g, device=dev, dtype=torch.float64) * np.sqrt(h2 / cm)
ld = (M + 3) // 4 * 4
R = torch.zeros((M, ld), device=dev, dtype=torch.float32)
r = torch.zeros(M, device=dev, dtype=torch.float64)
t0 = time.time()
p = 0.4
mu, sd = 2 * p, np.sqrt(2 * p * (1 - p))
for c0 in range(0, N, 8192):
    n = min(8192, N - c0)
    X = ((torch.rand((n, M), generator=g, device=dev) < p).float() + (torch.rand((n, M), generator=g, device=dev) < p).float() - mu) / sd
    y = X.double() @ beta + torch.randn(n, generator=g, device=dev, dtype=torch.float64) * np.sqrt(1 - h2)
    R[:, :M] += X.T @ X
    r += X.double().T @ y
R /= N
R = (R + R.T.contiguous()[:, :ld] if ld == M else R)   # exact symmetry of the fp32 accumulation
if ld == M:
    R *= 0.5
r /= np.sqrt(N)
torch.cuda.synchronize()
print("generated M=%d N=%d in %.1f s" % (M, N, time.time() - t0), flush=True)
v0 = h2 / cm
v = sgvamp.VAMP(N=N, Nt=N, M=M, K=1, rho=0.5, gamw=2.0, gam1=1e-6, a=np.array([1.0]), prior_vars=[0.0, 0.1 * v0, v0, 10 * v0],
                prior_probs=[0.97, 0.01, 0.01, 0.01], out_dir=None, out_name="c2")
x0 = (beta * np.sqrt(N)).cpu().numpy()
probes = (np.random.RandomState(1).binomial(1, 0.5, size=(1, its, M)) * 2 - 1).astype(np.int8)
R_dev = sgvamp.DeviceDense(R.data_ptr(), ld, keepalive=R)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    xs = v.infer(R_dev if rep == 0 else None, r.cpu().numpy(), its, cg_maxit=50, learn_gamw=True, lmmse_damp=False, prior_update="em",
                 probes=probes, write_outputs=False)
    torch.cuda.synchronize(); dt = time.time() - t0
    al = float(np.dot(xs[-1].ravel(), x0) / np.linalg.norm(xs[-1]) / np.linalg.norm(x0))
    print("run %d: %d iterations in %.3f s = %.1f it/s; cg iters %s; alignment %.4f; layout %s" % (
        rep, its, dt, its / dt, [tuple(v.history["cg_iters"][i][0]) for i in range(its)], al, v.handle.ld_info(0)["layout"]), flush=True)
v.close()
