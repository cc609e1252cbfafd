"""Worker of test_rank_per_cohort_under_torchrun: the reference's deployment shape (one rank per cohort,
src/main.py:16-18,85) on GPUs - `torchrun --nproc-per-node K` - with the r1 / gam1 exchange of src/sgvamp.py:228-233
carried by shard.TorchComm (one NCCL all-gather into the library's r1 block).  Compares against the reference golden
and exits non-zero on any mismatch."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "sgvamp-py_b200"))


def main():
    import torch
    import torch.distributed as dist
    from golden_util import load_case, rel_err, rel_l2
    import sgvamp
    import shard as shd
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = load_case(sys.argv[1])
    K, M = c["K"], c["M"]
    assert dist.get_world_size() == K
    comm = shd.TorchComm()
    Nt = sum(c["N_list"])
    v = sgvamp.VAMP(N=c["N_list"][rank], Nt=Nt, M=M, K=K, rho=c["rho"], gamw=c["gamw"], gam1=c["gam1"],
                    a=np.array(c["N_list"]) / Nt, prior_vars=c["prior_vars"], prior_probs=c["prior_probs"],
                    out_dir=None, out_name="g", comm=comm, device=local)
    x0 = c["x0"] * np.sqrt(c["N_list"][0]) if "x0" in c else None
    xs = v.infer(c["R"][rank], c["r"][rank], c["iterations"], x0=x0, cg_maxit=c["cg_maxit"], em_prior_maxit=c["em_prior_maxit"],
                 learn_gamw=c["learn_gamw"], lmmse_damp=c["lmmse_damp"], prior_update=c["prior_update"],
                 update_prior_from=c["update_prior_from"], s=c["s"], probes=lambda kk, it, M_: c["probes"][kk, it])
    used_device_exchange = comm.allgather_r1(v.handle, 1.0) is not None
    for it in range(c["iterations"]):
        assert rel_l2(xs[it], c["xhat"][it]) <= 1e-4, (it, rel_l2(xs[it], c["xhat"][it]))
        assert rel_err(v.history["rows"][it][rank][1:6], c["rows"][it, rank, 1:6]) <= 1e-4
        assert tuple(v.history["cg_iters"][it][rank]) == tuple(c["cg_iters"][it, rank])
    assert used_device_exchange
    v.close()
    dist.barrier()
    dist.destroy_process_group()
    print("RANK_MODE_OK %d" % rank)


if __name__ == "__main__":
    main()
