import sys, os, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import torch, bench, sgvamp
M, w, iters = int(sys.argv[1]), 500, 8
dev = torch.device("cuda", 0)
band, ldb, r, x0, tg = bench.build_problem(torch, M, w, 5, dev)
p = bench.vamp_params(M); N = bench.n_gwas(M)
probes = bench.make_probes(iters, M, 5)
def solver():
    return sgvamp.VAMP(N=N, Nt=N, M=M, K=1, rho=p["rho"], gamw=p["gamw"], gam1=p["gam1"], a=np.array([1.0]),
                       prior_vars=p["prior_vars"], prior_probs=p["prior_probs"], out_dir=None, out_name="d")
def run(tag, sampler=False, profile=False, events=False, adopt=True):
    smp = bench.ClockSampler(0) if sampler else None
    v = solver()
    evs = []
    def hook(it):
        if it == 2 and profile:
            v.handle.profile(True)
        if events:
            e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    v.infer(sgvamp.DeviceDIA(band.data_ptr(), w, ldb), r, iters, cg_maxit=500, lmmse_damp=False, prior_update="em", probes=probes, iter_hook=hook)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-28s %.1f ms/it  timers %s em_steps %s" % (tag, dt / iters * 1e3, {k: round(x, 3) for k, x in v.timers.items()}, v.history["em_steps"]))
    v.close()
    if smp: smp.close()
run("warm")
run("baseline")
run("sampler", sampler=True)
run("profile", profile=True)
run("events", events=True)
run("all", sampler=True, profile=True, events=True)
run("baseline again")
