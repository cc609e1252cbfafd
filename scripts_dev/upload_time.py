import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200")); sys.path.insert(0, REPO)
import numpy as np, torch, scipy.sparse
import sgv_native as nat
M, w = 400_000, 500
offs = np.arange(-w, w + 1)
rng = np.random.default_rng(0)
t0 = time.time()
nnz = M * (2 * w + 1) - w * (w + 1)
# banded symmetric CSR built directly
rows = np.arange(M)
lo = np.maximum(rows - w, 0); hi = np.minimum(rows + w, M - 1)
cnt = hi - lo + 1
indptr = np.zeros(M + 1, dtype=np.int64); indptr[1:] = np.cumsum(cnt)
indices = (np.repeat(lo - indptr[:-1], cnt) + np.arange(nnz)).astype(np.int32)
r = np.repeat(rows, cnt)
d = np.abs(indices - r)
mn = np.minimum(indices, r)
data = (np.sin(mn * 0.37 + d * 1.3) * 0.01).astype(np.float32)   # symmetric by construction
data[d == 0] = 1.0
print("built", time.time() - t0, nnz)
ti = torch.from_numpy(indices).pin_memory(); td = torch.from_numpy(data).pin_memory()
h = nat.Handle(); h.configure(M, 1)
for lay, name in [(nat.LAYOUT_DIA, "dia"), (nat.LAYOUT_DSYM, "dsym"), (nat.LAYOUT_DIA, "dia"), (nat.LAYOUT_DSYM, "dsym")]:
    torch.cuda.synchronize(); t0 = time.time()
    h._ck(h.upload_csr(0, indptr, ti.numpy(), td.numpy(), s=0.1, layout=lay))
    torch.cuda.synchronize()
    print(name, "upload %.3f s" % (time.time() - t0), h.ld_info(0)["layout"])
