"""Timing of the LD construction kernels (tcgen05 int8 Gram tiles vs IDP4A) at the benchmark's shape."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import numpy as np, torch
import sgv_native as nat
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 500
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
g = torch.Generator(device="cuda"); g.manual_seed(1)
Gt = (torch.rand((M, N), generator=g, device="cuda") < 0.3).to(torch.int8) + (torch.rand((M, N), generator=g, device="cuda") < 0.3).to(torch.int8)
for flag in ("0", "1", "0", "1"):
    os.environ["SGV_LD_DP4A"] = flag
    h = nat.Handle(); h.configure(M, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.build_banded(0, None, N, w, s=0.1, taper=True, device_ptr=Gt.data_ptr(), nmark=M, ldg=N)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    macs = M * (w + 1.0) * N
    print("%s: M=%d w=%d N=%d: %.1f ms for the whole construction (stats + Gram + epilogue), %.1f TMAC/s on the %.2e band MACs" % (
        "IDP4A  " if flag == "1" else "tcgen05", M, w, N, dt * 1e3, macs / dt / 1e12, macs))
    h.close()
