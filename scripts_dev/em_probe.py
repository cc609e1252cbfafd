import sys, os, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import torch, sgv_native as nat
M = 1000000
torch.zeros(1, device="cuda")
h = nat.Handle(); h.configure(M, 1)
rng = np.random.default_rng(0)
r1 = rng.standard_normal(M) * 3
h.set_vec(0, nat.VEC_R1, r1); h.set_weights(np.array([1.0]))
def em(n=1, maxit=15):
    out = []
    for _ in range(n):
        h.set_prior(0.01, np.array([1.0]), np.array([25.0]))
        t = time.perf_counter(); lam, om, steps, rel = h.prior_em(np.array([1.0]), maxit, 0.0, 1); out.append((time.perf_counter() - t) * 1e3)
    return out, steps
def dn(n=1):
    out = []
    for _ in range(n):
        t = time.perf_counter(); h.denoise(np.array([1.0]), 0.5, False); out.append((time.perf_counter() - t) * 1e3)
    return out
print("em x6 (15 passes each) ms:", [round(x, 2) for x in em(6)[0]])
print("denoise x6 ms:", [round(x, 3) for x in dn(6)])
band = torch.randn((1001, M), device="cuda", dtype=torch.float32)
h.adopt_dia(0, band.data_ptr(), 500, M)
print("spmm_bench 40 reps ms/launch:", h.spmm_bench(0, 40))
print("em x6 after heavy:", [round(x, 2) for x in em(6)[0]])
time.sleep(1.0)
print("em x6 after 1s idle:", [round(x, 2) for x in em(6)[0]])
print("denoise x6 ms:", [round(x, 3) for x in dn(6)])
p = [h.pinned_array(M) for _ in range(3)]
t = time.perf_counter(); h.get_vec_async(0, nat.VEC_XHAT1, 1.0, p[0]); h.wait_copies(); print("get_vec_async+wait ms", (time.perf_counter() - t) * 1e3)
t = time.perf_counter(); h.get_vec_async(0, nat.VEC_XHAT1, 1.0, p[1]); h.wait_copies(); print("get_vec_async+wait ms", (time.perf_counter() - t) * 1e3)
print("em x6 after async copies:", [round(x, 2) for x in em(6)[0]])
import threading
def bg():
    for _ in range(50):
        h.wait_copies(); x = p[0].copy(); time.sleep(0.002)
th = threading.Thread(target=bg); th.start()
print("em x6 with bg thread:", [round(x, 2) for x in em(6)[0]])
th.join()
