"""Command-line driver, drop-in for the reference's ``src/main.py``.

Same flags (both the code spellings of src/main.py:28-50 and the README spellings, e.g.
``--mle-prior-update``), same input formats for r (.txt / .npy / .linear) and R (.npz / .npy),
same output files.  One process drives all K cohorts on the GPU (the reference starts one MPI
rank per cohort); ``Rused = (1-s) R + s I`` (src/main.py:265) is applied on the device at upload.

The `.bim` reference-order merge with missing SNPs and the PLINK `.ld` triple loader
(src/main.py:126-165, 203-257) live in ``ingest.py``; the reference's MPI exchange of missing rows
becomes an in-memory lookup because one process holds all cohorts.
"""
import argparse
import logging
import os
import struct
import sys
import time

import numpy as np
import scipy.sparse

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ingest  # noqa: E402
from sgvamp import VAMP  # noqa: E402


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("-ld_files", "--ld-files", help="Path to LD matrices in .npz/.npy files, separated by comma")
    p.add_argument("-r_files", "--r-files", help="Path to XTy .npy/.txt/.linear files separated by comma")
    p.add_argument("-true_signal_file", "--true-signal-file", help="Path to true signal .npy/.bin file", default=None)
    p.add_argument("-out_dir", "--out-dir", help="Output directory")
    p.add_argument("-out_name", "--out-name", help="Output file name")
    p.add_argument("-N", "--N", help="Number of samples in each cohort, separated by comma")
    p.add_argument("-M", "--M", help="Number of markers in each cohort, separated by comma")
    p.add_argument("-K", "--K", help="Number of cohorts", default=1)
    p.add_argument("-L", "--L", help="Number of prior mixture components", default=2)
    p.add_argument("-iterations", "--iterations", help="Number of iterations", default=10)
    p.add_argument("-prior_vars", "--prior-vars", help="Prior mixture variances", default="0,1")
    p.add_argument("-prior_probs", "--prior-probs", help="Prior mixture probabilities", default="0.99,0.01")
    p.add_argument("-gamw", "--gamw", help="Initial noise precision", default=5)
    p.add_argument("-gam1", "--gam1", help="Initial signal precision", default=0.000001)
    p.add_argument("-lmmse_damp", "--lmmse-damp", help="Use LMMSE damping", default=False)
    p.add_argument("-learn_gamw", "--learn-gamw", help="Learn or fix gamw", default=True)
    p.add_argument("-rho", "--rho", help="Damping factor rho", default=0.5)
    p.add_argument("-cg_maxit", "--cg-maxit", help="CG max iterations", default=500)
    p.add_argument("-s", "--s", help="Rused = (1-s) * R + s * Id", default=0.0)
    p.add_argument("-prior_update", "--prior-update", "--mle-prior-update", dest="prior_update",
                   help="Learning prior probabilities: 'em' or 'mle'", default="em")
    p.add_argument("-update_prior_from", "--update-prior-from", help="Learn prior from this iteration onwards", default=1)
    p.add_argument("-em_prior_maxit", "--em-prior-maxit", help="Max EM prior-learning iterations", default=100)
    p.add_argument("-bim_files", "--bim-files", help="Path to files containing list of snps", default=None)
    p.add_argument("--device", help="CUDA device index", default=0)
    p.add_argument("--bim-source-quirk", help="1 (default): choose the cohort that supplies a missing SNP exactly as "
                   "src/main.py:162 does (argmax position); 0: the candidate cohort with the largest N", default=1)
    p.add_argument("--layout", help="LD layout in HBM: auto|dense|dia|dsym|blockdiag|csr", default="auto")
    p.add_argument("--row-partition", help="under torchrun with N ranks: 1 = split the marker rows of every cohort over the "
                   "N GPUs (banded LD; halos and reductions exchanged inside the kernels) instead of one rank per cohort; "
                   "default: 1 when K == 1", default=None)
    p.add_argument("--probe-seed", help="deterministic Hutchinson probes: the probe of cohort k in iteration it is the it-th draw "
                   "of numpy RandomState(seed + k).binomial(p=1/2, n=1, size=M) (default: the reference's draws from numpy's "
                   "global RNG, src/sgvamp.py:326)", default=None)
    p.add_argument("--n-probes", help="Rademacher probes averaged in the Hutchinson estimates of alpha2 and of the gamw update "
                   "(1 = the reference, src/sgvamp.py:326-340)", default=1)
    p.add_argument("--checkpoint-path", help="write a restart file here every --checkpoint-every iterations", default=None)
    p.add_argument("--checkpoint-every", help="iterations between restart files (0: never)", default=0)
    p.add_argument("--resume-from", help="continue the run stored in this restart file", default=None)
    return p


class _SeededProbes:
    """probes(k, it, M[, p]): draw number it * n_probes + p of RandomState(seed + k), drawn in order and cached."""

    def __init__(self, seed, K, n_probes=1):
        self.rs = [np.random.RandomState(seed + k) for k in range(K)]
        self.have = [[] for _ in range(K)]
        self.n_probes = n_probes

    def __call__(self, k, it, M, p=0):
        idx = it * self.n_probes + p
        while len(self.have[k]) <= idx:
            self.have[k].append((self.rs[k].binomial(p=1 / 2, n=1, size=M) * 2 - 1).astype(np.int8))
        return self.have[k][idx]


def main(argv=None):
    logging.basicConfig(format="%(message)s", level=logging.INFO)
    a = build_parser().parse_args(argv)
    logging.info(" ### VAMP for summary statistics (B200) ###\n")
    K, L = int(a.K), int(a.L)
    ld_list, r_list = a.ld_files.split(","), a.r_files.split(",")
    N_list = [int(n) for n in a.N.split(",")]
    M_list = [int(m) for m in a.M.split(",")]
    prior_vars = [float(x) for x in a.prior_vars.split(",")]
    prior_probs = [float(x) for x in a.prior_probs.split(",")]
    if len(ld_list) != K:
        raise Exception("Specified number of cohorts is not equal to number of LD matrices provided!")
    if len(r_list) != K:
        raise Exception("Specified number of cohorts is not equal to number of marginal estimates provided!")
    if len(prior_vars) != L:
        raise Exception("Number of prior variances must be L!")
    if len(prior_probs) != L:
        raise Exception("Number of prior mixture probabilites must be L!")
    Nt = sum(N_list)
    lmmse_damp, learn_gamw = bool(int(a.lmmse_damp)), bool(int(a.learn_gamw))
    s = float(a.s)

    ts = time.time()
    bim_list = a.bim_files.split(",") if a.bim_files is not None else None
    if bim_list is not None and len(bim_list) != K:
        raise Exception("Specified number of cohorts is not equal to number of .bim files provided!")
    M, Rs, rs, merged = ingest.load_all(ld_list, r_list, bim_list, N_list, M_list,
                                        source_quirk=bool(int(a.bim_source_quirk)))
    if merged is not None:
        logging.info(f"Total number of markers in reference is {M} \n")
        if int(os.environ.get("RANK", "0")) == 0:                                               # rank 0 saves it, src/main.py:148-150
            ingest.write_ref_bim(merged["ref_df"], os.path.join(a.out_dir, a.out_name + ".bim"))
    logging.info(f"Loading R and r took {time.time() - ts:0.2f} seconds\n")
    x0 = None
    if a.true_signal_file is not None:
        if a.true_signal_file.endswith(".bin"):
            with open(a.true_signal_file, "rb") as f:
                x0 = np.array(struct.unpack(str(M) + "d", f.read(M * 8)))
        elif a.true_signal_file.endswith(".npy"):
            x0 = np.load(a.true_signal_file).astype(np.float64).ravel()
        else:
            raise Exception("Unsupported true signal format!")
        x0 = x0 * np.sqrt(N_list[0])                                # src/main.py:276 (rank 0 writes the metrics)
    avec = np.array(N_list) / sum(N_list)                           # src/main.py:287
    # Deployment shapes: (default) one process drives all K cohorts on one GPU; or the reference's own shape,
    # one rank per cohort (src/main.py:16-18,85): `torchrun --nproc-per-node K main.py ...` puts cohort k on GPU k
    # and the per-iteration exchange of r1 / gam1 (src/sgvamp.py:228-233) goes over NCCL.
    world = int(os.environ.get("WORLD_SIZE", "1"))
    comm, device = None, int(a.device)
    row_part = world > 1 and (bool(int(a.row_partition)) if a.row_partition is not None else K == 1)
    shard_obj = None
    if row_part:
        # one cohort (or all of them) row-partitioned over the N GPUs of the box: every rank reads the inputs and
        # keeps its own rows; rank 0 writes the outputs
        import torch
        import torch.distributed as dist
        import shard
        device = int(os.environ.get("LOCAL_RANK", "0"))
        if not dist.is_initialized():
            torch.cuda.set_device(device)
            dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        shard_obj = shard.TorchShard()
    elif world > 1:
        import torch
        import torch.distributed as dist
        import shard
        if world != K:
            raise Exception("one rank per cohort: WORLD_SIZE (%d) must equal K (%d)" % (world, K))
        device = int(os.environ.get("LOCAL_RANK", "0"))
        if not dist.is_initialized():
            torch.cuda.set_device(device)
            dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        comm = shard.TorchComm()
        me = comm.Get_rank()
        Rs, rs = [Rs[me]], [rs[me]]
        if x0 is not None:
            x0 = x0 / np.sqrt(N_list[0]) * np.sqrt(N_list[me])     # src/main.py:276 scales by the rank's own N
    solver = VAMP(N=(N_list[comm.Get_rank()] if comm else (N_list if K > 1 else N_list[0])), Nt=Nt, M=M, K=K,
                  rho=float(a.rho), gam1=float(a.gam1), gamw=float(a.gamw), a=avec, prior_vars=prior_vars,
                  prior_probs=prior_probs, out_dir=a.out_dir, out_name=a.out_name, comm=comm, device=device,
                  shard=shard_obj, halo=True)
    logging.info("...Running sgVAMP\n")
    ts = time.time()
    one = comm is not None or K == 1
    xhat1 = solver.infer(Rs[0] if one else Rs, rs[0] if one else rs, int(a.iterations), x0=x0,
                         cg_maxit=int(a.cg_maxit), em_prior_maxit=int(a.em_prior_maxit), learn_gamw=learn_gamw,
                         lmmse_damp=lmmse_damp, prior_update=a.prior_update,
                         update_prior_from=int(a.update_prior_from), s=s, layout=a.layout,
                         probes=_SeededProbes(int(a.probe_seed), K, int(a.n_probes)) if a.probe_seed is not None else None,
                         n_probes=int(a.n_probes),
                         checkpoint_path=a.checkpoint_path, checkpoint_every=int(a.checkpoint_every), resume_from=a.resume_from)
    logging.info(f"sgVAMP inference running time: {(time.time() - ts):0.4f}s\n")
    # README names the dump {out}__xhat_it_{it}.bin, the code writes {out}_xhat_it_{it}.bin: provide both
    is_root = (comm is None or comm.Get_rank() == 0) and (shard_obj is None or shard_obj.rank == 0)
    for it in range(int(a.iterations) if is_root else 0):
        src = os.path.join(a.out_dir, "%s_xhat_it_%d.bin" % (a.out_name, it))
        dst = os.path.join(a.out_dir, "%s__xhat_it_%d.bin" % (a.out_name, it))
        if os.path.exists(src) and not os.path.exists(dst):
            os.symlink(os.path.basename(src), dst)
    if x0 is not None:
        al = [float(np.inner(x.squeeze(), x0) / np.linalg.norm(x) / np.linalg.norm(x0)) for x in xhat1]
        l2 = [float(np.linalg.norm(x.squeeze() - x0) / np.linalg.norm(x0)) for x in xhat1]
        logging.info(f"Alignment(x1hat, x0) over iterations: \n {al}\n")
        logging.info(f"L2 error(x1hat, x0) over iterations: \n {l2}\n")
    solver.close()
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    return xhat1


if __name__ == "__main__":
    main()
