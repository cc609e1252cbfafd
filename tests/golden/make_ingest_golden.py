"""Golden vectors for the ingestion path (src/main.py:126-265), produced by the UNMODIFIED reference
driver run in this container.

src/main.py is a script: it is executed with runpy once per cohort (one thread per MPI rank) against
 - a stub `mpi4py.MPI` whose COMM_WORLD has a thread-local rank and queue-based send/recv (tags as in
   src/main.py:214-247), and
 - a stub `sgvamp` module whose VAMP records what the driver passes to `infer` (R = Rused, r, x0) and
   returns zeros, so that only the ingestion code of the reference runs.
Inputs (tiny .bim / .ld / .assoc.linear / .npy files) are written next to the goldens under
tests/golden/ingest/ and are what tests/test_ingest_cpu.py feeds to sgvamp-py_b200/ingest.py.
Run:  python tests/golden/make_ingest_golden.py
"""
import os
import queue
import runpy
import sys
import threading
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ingest")
REF_MAIN = "/root/reference/src/main.py"


def write_inputs():
    """Two cohorts with overlapping but different SNP sets; cohort 1 lacks rs3 and rs7, cohort 2 lacks rs5."""
    rng = np.random.default_rng(42)
    snps = ["rs%d" % i for i in range(1, 11)]
    coord = {rs: 1000 + 37 * i for i, rs in enumerate(snps)}
    sets = [[rs for rs in snps if rs not in ("rs3", "rs7")], [rs for rs in snps if rs != "rs5"]]
    sets[1] = sets[1][::-1][2:] + sets[1][::-1][:2]          # cohort 2 lists its SNPs in a different order
    for k, ss in enumerate(sets):
        with open(os.path.join(OUT, "c%d.bim" % (k + 1)), "w") as f:
            for rs in ss:
                f.write("1\t%s\t0\t%d\tA\tG\n" % (rs, coord[rs]))
        with open(os.path.join(OUT, "c%d.ld" % (k + 1)), "w") as f:
            f.write(" CHR_A BP_A SNP_A CHR_B BP_B SNP_B R\n")
            order = sorted(ss, key=lambda r: coord[r])
            for i in range(len(order)):
                for j in range(i + 1, min(i + 4, len(order))):
                    f.write(" 1 %d %s 1 %d %s %.6f\n" % (coord[order[i]], order[i], coord[order[j]], order[j],
                                                        rng.uniform(-0.6, 0.6)))
        with open(os.path.join(OUT, "c%d.assoc.linear" % (k + 1)), "w") as f:
            f.write(" CHR SNP BP A1 TEST NMISS BETA STAT P\n")
            for i, rs in enumerate(ss):
                beta = "NA" if (k == 0 and i == 2) else "%.6f" % rng.normal(0, 0.05)
                f.write(" 1 %s %d A ADD 100 %s 0.1 0.5\n" % (rs, coord[rs], beta))
    x0 = rng.normal(0, 1, 10)
    np.save(os.path.join(OUT, "x0.npy"), x0)


class FakeComm:
    def __init__(self, K):
        self.K = K
        self.local = threading.local()
        self.q = {}
        self.lock = threading.Lock()
        self.bar = threading.Barrier(K)
        self.slot = None

    def _chan(self, src, dst, tag):
        with self.lock:
            return self.q.setdefault((src, dst, tag), queue.Queue())

    def Get_rank(self):
        return self.local.rank

    def Get_size(self):
        return self.K

    def send(self, obj, dest, tag=0):
        self._chan(self.local.rank, dest, tag).put(obj)

    def recv(self, source, tag=0):
        return self._chan(source, self.local.rank, tag).get(timeout=60)

    def bcast(self, obj, root=0):
        if self.K == 1:
            return obj
        if self.local.rank == root:
            self.slot = np.array(obj, copy=True) if isinstance(obj, np.ndarray) else obj
        self.bar.wait()
        out = self.slot
        out = out.copy() if isinstance(out, np.ndarray) else out
        self.bar.wait()
        return out


def run_reference_solver(argv, K, probe_seed, iterations, M):
    """The unmodified reference driver AND the unmodified reference solver (src/sgvamp.py), one thread per rank, with the
    probe of (rank k, iteration it) injected as the it-th draw of RandomState(probe_seed + k) by rebinding the solver
    module's `binomial` (src/sgvamp.py:5,326) - what `main.py --probe-seed` reproduces."""
    import importlib.util
    comm = FakeComm(K)
    mpi = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")
    MPI.COMM_WORLD = comm
    MPI.Finalize = lambda: None
    mpi.MPI = MPI
    spec = importlib.util.spec_from_file_location("sgvamp", "/root/reference/src/sgvamp.py")
    sg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sg)
    streams = [np.random.RandomState(probe_seed + k) for k in range(K)]
    sg.binomial = lambda p=None, n=None, size=None: streams[comm.local.rank].binomial(p=p, n=n, size=size)
    saved = {n: sys.modules.get(n) for n in ("mpi4py", "mpi4py.MPI", "sgvamp")}
    sys.modules.update({"mpi4py": mpi, "mpi4py.MPI": MPI, "sgvamp": sg})
    old_argv = sys.argv
    sys.argv = ["main.py"] + argv
    errs = []

    def body(k):
        comm.local.rank = k
        try:
            runpy.run_path(REF_MAIN, run_name="__main__")
        except BaseException as e:   # noqa: BLE001
            errs.append((k, e))
            comm.bar.abort()

    ts = [threading.Thread(target=body, args=(k,)) for k in range(K)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    sys.argv = old_argv
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m
    if errs:
        raise errs[0][1]
    out_dir = argv[argv.index("--out-dir") + 1]
    name = argv[argv.index("--out-name") + 1]
    res = dict(xhat=np.stack([np.fromfile(os.path.join(out_dir, "%s_xhat_it_%d.bin" % (name, it))) for it in range(iterations)]))
    import csv
    rows = np.zeros((iterations, K, 7))
    for k in range(K):
        with open(os.path.join(out_dir, "%s_cohort_%d.csv" % (name, k + 1)), newline="") as f:
            rd = list(csv.reader(f, delimiter="\t"))
        for it in range(iterations):
            rows[it, k] = [float(v) for v in rd[1 + it]]
    res["rows"] = rows
    res["r1"] = np.stack([[np.fromfile(os.path.join(out_dir, "%s_r1_cohort_%d_it_%d.bin" % (name, k + 1, it))) for k in range(K)]
                          for it in range(iterations)])
    return res


def run_reference(argv, K):
    comm = FakeComm(K)
    captured = {}
    mpi = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")
    MPI.COMM_WORLD = comm
    MPI.Finalize = lambda: None
    mpi.MPI = MPI
    sg = types.ModuleType("sgvamp")

    class VAMP:
        def __init__(self, **kw):
            self.kw = kw

        def infer(self, R, r, iterations, x0=None, **kw):
            k = comm.Get_rank()
            captured[k] = dict(R=np.asarray(R.todense()) if hasattr(R, "todense") else np.asarray(R), r=np.asarray(r).ravel(),
                               x0=None if x0 is None else np.asarray(x0).ravel(), M=self.kw["M"], a=np.asarray(self.kw["a"]))
            return [np.zeros((self.kw["M"], 1)) for _ in range(iterations)]

    sg.VAMP = VAMP
    saved = {n: sys.modules.get(n) for n in ("mpi4py", "mpi4py.MPI", "sgvamp")}
    sys.modules.update({"mpi4py": mpi, "mpi4py.MPI": MPI, "sgvamp": sg})
    old_argv = sys.argv
    sys.argv = ["main.py"] + argv
    errs = []

    def body(k):
        comm.local.rank = k
        try:
            runpy.run_path(REF_MAIN, run_name="__main__")
        except BaseException as e:   # noqa: BLE001
            errs.append((k, e))

    ts = [threading.Thread(target=body, args=(k,)) for k in range(K)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    sys.argv = old_argv
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m
    if errs:
        raise errs[0][1]
    return captured


def main():
    os.makedirs(OUT, exist_ok=True)
    write_inputs()
    p = lambda n: os.path.join(OUT, n)
    out = {}
    # K = 2, PLINK .ld + .assoc.linear + .bim with missing SNPs in both cohorts
    cap = run_reference(["--ld-files", p("c1.ld") + "," + p("c2.ld"), "--r-files", p("c1.assoc.linear") + "," + p("c2.assoc.linear"),
                         "--bim-files", p("c1.bim") + "," + p("c2.bim"), "--true-signal-file", p("x0.npy"),
                         "--out-dir", OUT, "--out-name", "ref_k2", "--N", "400,900", "--M", "8,9", "--K", "2",
                         "--iterations", "1", "--s", "0.2"], 2)
    for k in range(2):
        for key in ("R", "r", "x0", "a"):
            out["k2_%s_%d" % (key, k)] = cap[k][key]
        out["k2_M"] = cap[k]["M"]
    out["k2_bim"] = open(p("ref_k2.bim")).read()
    # K = 1, .ld of cohort 2 alone
    cap = run_reference(["--ld-files", p("c2.ld"), "--r-files", p("c2.assoc.linear"), "--bim-files", p("c2.bim"),
                         "--out-dir", OUT, "--out-name", "ref_k1", "--N", "900", "--M", "9", "--K", "1",
                         "--iterations", "1", "--s", "0.0"], 1)
    out["k1_R_0"], out["k1_r_0"], out["k1_M"] = cap[0]["R"], cap[0]["r"], cap[0]["M"]
    out["k1_bim"] = open(p("ref_k1.bim")).read()
    # K = 2 end to end: reference driver + reference solver, 4 iterations, probes from RandomState(77 + k)
    run = run_reference_solver(["--ld-files", p("c1.ld") + "," + p("c2.ld"), "--r-files", p("c1.assoc.linear") + "," + p("c2.assoc.linear"),
                                "--bim-files", p("c1.bim") + "," + p("c2.bim"), "--true-signal-file", p("x0.npy"),
                                "--out-dir", OUT, "--out-name", "ref_run", "--N", "400,900", "--M", "8,9", "--K", "2",
                                "--iterations", "4", "--s", "0.3", "--prior-vars", "0,0.001", "--prior-probs", "0.7,0.3",
                                "--gamw", "2"], 2, 77, 4, 10)
    out["k2run_xhat"], out["k2run_rows"], out["k2run_r1"] = run["xhat"], run["rows"], run["r1"]
    np.savez_compressed(os.path.join(OUT, "reference.npz"), **out)
    for n in os.listdir(OUT):
        if n.startswith("ref_k") or n.startswith("ref_run"):
            os.remove(p(n))
    print("wrote", os.path.join(OUT, "reference.npz"), {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
