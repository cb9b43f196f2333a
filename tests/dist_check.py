"""Multi-GPU check (run under torchrun on a >= 2 GPU box: `gpurun --gpus 2 -- torchrun --standalone --nproc-per-node 2
tests/dist_check.py`): the time-partitioned chain (one segment per rank, boundary all-gather inside the library) must
reproduce the single-GPU run of the merged problem."""
import os
import sys
import pathlib

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch
import torch.distributed as dist

import gaussianvi_b200 as gv
from gaussianvi_b200 import problems
from gaussianvi_b200.dist import attach_nccl


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    niters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gv.Context(local)
    attach_nccl(ctx, rank, world)
    seg = problems.make_cfg3_segment(rank, world, N=N)
    p = problems.build_device_problem(ctx, seg)
    opts = gv.Problem.default_opts()
    opts.reuse_accepted_sweep = 1
    stats = [p.iterate(opts) for _ in range(niters)]
    mu = p.mean()
    cD, cO = p.covariance()
    ok = True
    if rank == 0:
        # single-GPU reference on a second, communicator-free context
        ctx1 = gv.Context(local)
        merged = problems.merge_segments([problems.make_cfg3_segment(r, world, N=N) for r in range(world)])
        q = problems.build_device_problem(ctx1, merged)
        ref = [q.iterate(opts) for _ in range(niters)]
        mu1 = q.mean()
        cD1, cO1 = q.covariance()
        m = seg.S - 1
        e_cost = max(abs(a.cost - b.cost) / abs(b.cost) for a, b in zip(stats, ref))
        e_mu = np.abs(mu - mu1[:(m + 1) * 4]).max() / np.abs(mu1).max()
        e_cov = np.abs(cD - cD1[:m + 1]).max() / np.abs(cD1).max()
        nb = [s.n_backtrack for s in stats], [s.n_backtrack for s in ref]
        print(f"dist_check world={world} N={N}: rel err cost {e_cost:.2e} mu {e_mu:.2e} cov {e_cov:.2e} backtracks {nb}")
        ok = e_cost < 1e-10 and e_mu < 1e-9 and e_cov < 1e-9 and nb[0] == nb[1]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    p.close()
    dist.destroy_process_group()
    if not int(flag.item()):
        sys.exit(1)
    if rank == 0:
        print("dist_check OK")


if __name__ == "__main__":
    main()
