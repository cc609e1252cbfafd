"""Isolated SpMM timing of the dense-panel / block-diagonal / CSR kernels (GB/s vs algorithmic bytes)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import numpy as np, torch, scipy.sparse
import sgv_native as nat
which = sys.argv[1]
h = nat.Handle()
if which == "dense":
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
    ld = (M + 3) // 4 * 4
    R = torch.randn((M, ld), device="cuda", dtype=torch.float32) * 0.01
    h.configure(M, 1)
    h.adopt_dense(0, R.data_ptr(), ld)
elif which == "blockdiag":
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
    rng = np.random.default_rng(3)
    sizes = []
    left = M
    while left > 0:
        b = int(min(left, rng.integers(500, 3500))); sizes.append(b); left -= b
    # upload through CSR would need 5.7 GB host arrays; build per block on host in fp32
    blocks = [scipy.sparse.csr_matrix(np.full((b, b), 0.01, dtype=np.float32) + np.eye(b, dtype=np.float32)) for b in sizes]
    R = scipy.sparse.block_diag(blocks, format="csr")
    h.configure(M, 1)
    h._ck(h.upload_csr(0, R.indptr, R.indices, R.data))
elif which == "csr":
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
    R = scipy.sparse.random(M, M, density=300.0 / M, format="csr", dtype=np.float32, random_state=1)
    R = (R + R.T).tocsr(); R.sort_indices()
    h.configure(M, 1)
    h._ck(h.upload_csr(0, R.indptr, R.indices, R.data, layout=nat.LAYOUT_CSR))
info = h.ld_info(0)
x = np.random.default_rng(0).standard_normal((h.M, 2))
h.spmm(0, x)
for rep in range(2):
    ms = h.spmm_bench(0, 20)
    print("%s M=%d layout=%s nnz_stored=%.3g: %.4f ms/pass  %.0f GB/s algorithmic" % (which, h.M, info["layout"], info["nnz_stored"], ms, info["bytes_per_pass"] / ms / 1e6))
h.close()
