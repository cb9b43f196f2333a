"""World-size-2 test of the multi-GPU host logic on CPU (gloo): each process builds ITS time segment
(problems.make_cfg3_segment), assembles its share of Vddmu / Vdmu with the oracle's factor arithmetic, reduces its
segment to the boundary record [Dfirst | Dlast | CL | CR | O | gfirst | glast | gl | gr] of gaussianvi_b200/csrc/bt_cr.h,
all-gathers the records, solves the chain of rank boundaries and back-substitutes.  The result must equal the oracle's
dmu on the merged single-process problem -- this pins segment ownership, the splitting of shared blocks and the
boundary-record algebra the CUDA path uses (the kernels themselves are replayed in tests/test_capi_host.py)."""
import os
import sys
import pathlib

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _local_system(seg):
    """This rank's share of (Vddmu, -Vdmu) as dense local arrays, from the oracle's per-factor arithmetic."""
    sys.path[:0] = [str(ROOT), str(ROOT / "tests"), str(ROOT / "oracle")]
    import oracle_bridge as ob
    factors = ob.build_factors(seg)
    # shares of the precision are not SPD on their own: take the covariance from the MERGED problem (passed in meta)
    cov = seg.meta["cov_blocks"]
    mu = np.asarray(seg.mu0, float).reshape(-1)
    for f in factors:
        f.update_mu_from_joint(mu)
        f.update_precision_from_joint(cov.block(f.start_index, f.dim // seg.d))
    S, d = seg.S, seg.d
    Vdmu = np.zeros(S * d)
    import gvi_oracle as o
    V = o.BlockTri.identity(S, d, 0.0)
    for f in factors:
        f.calculate_partial_V()
        Vdmu[f.offset:f.offset + f.dim] += f.Vdmu
        V.add_block(f.start_index, f.dim // d, f.Vddmu)
    return V, -Vdmu


def _worker(rank, world, port, N, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [str(ROOT), str(ROOT / "tests"), str(ROOT / "oracle")]
    import torch
    import torch.distributed as dist
    import gvi_oracle as o
    import oracle_bridge as ob
    from gaussianvi_b200 import problems
    dist.init_process_group("gloo", rank=rank, world_size=world)
    segs = [problems.make_cfg3_segment(r, world, N=N) for r in range(world)]
    merged = problems.merge_segments(segs)
    ref = ob.build_oracle(merged, niters=1)
    seg = segs[rank]
    m, d = seg.S - 1, seg.d
    # local slice of the merged covariance (both neighbours hold the same blocks at a shared state)
    seg.meta["cov_blocks"] = o.BlockTri(ref.cov.D[rank * m:rank * m + m + 1].copy(), ref.cov.O[rank * m:(rank + 1) * m].copy())
    V, g = _local_system(seg)
    A = V.dense()
    n = (m + 1) * d
    I = np.arange(d, n - d)                     # interior
    B = np.r_[np.arange(d), np.arange(n - d, n)]  # first, last
    Aii, Aib, Abb = A[np.ix_(I, I)], A[np.ix_(I, B)], A[np.ix_(B, B)]
    sol = np.linalg.solve(Aii, np.c_[Aib, g[I]])
    Sb = Abb - Aib.T @ sol[:, :2 * d]           # [[Dfirst + CL, O], [O^T, Dlast + CR]]
    gb = g[B] - Aib.T @ sol[:, 2 * d]
    rec = np.r_[Sb[:d, :d].ravel(), Sb[d:, d:].ravel(), Sb[:d, d:].ravel(), gb]
    gathered = [torch.zeros(len(rec), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(rec))
    # chain of rank boundaries, solved redundantly
    P = world
    Dt = np.zeros((P + 1, d, d))
    Ot = np.zeros((P, d, d))
    gt = np.zeros((P + 1, d))
    for r, t in enumerate(gathered):
        t = t.numpy()
        Dt[r] += t[:d * d].reshape(d, d)
        Dt[r + 1] += t[d * d:2 * d * d].reshape(d, d)
        Ot[r] = t[2 * d * d:3 * d * d].reshape(d, d)
        gt[r] += t[3 * d * d:3 * d * d + d]
        gt[r + 1] += t[3 * d * d + d:]
    xt = o.block_solve(o.BlockTri(Dt, Ot), gt.reshape(-1)).reshape(P + 1, d)
    xb = np.r_[xt[rank], xt[rank + 1]]
    xi = sol[:, 2 * d] - sol[:, :2 * d] @ xb
    x = np.zeros(n)
    x[B] = xb
    x[I] = xi
    dmu_ref, _ = ref.compute_gradients()
    err = np.abs(x - dmu_ref[rank * m * d:rank * m * d + n]).max() / np.abs(dmu_ref).max()
    out.put((rank, float(err)))
    dist.destroy_process_group()


def test_time_partition_two_processes_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 14, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-9, (rank, err)


def test_bench_reference_arm_under_two_ranks():
    """bench.py --impl reference under torchrun semantics: rank 0 prints the line, the other rank exits 0 silently."""
    import json
    import subprocess
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", GVIB200_BENCH_CPU_FACTORS="2000")
    outs = []
    for rank in (0, 1):
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                           env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    line = json.loads(outs[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert outs[1] == ""


def test_bench_reference_arm_at_the_drivers_arguments():
    """The driver runs `bench.py --impl reference --steps 20 --warmup 5`: 25 iterations of the reference's own schedule
    and arithmetic.  Un-rewound, the O(dim^4) Vddmu loop of the linear factors loses positive definiteness at iteration 6
    of this workload (round 1's crash); the arm rewinds a host-side snapshot like the GPU arm rewinds on the device."""
    import json
    import subprocess
    env = dict(os.environ, GVIB200_BENCH_CPU_FACTORS="10000")
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "20", "--warmup", "5"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip())
    assert line["impl"] == "reference" and line["steps"] == 20 and line["warmup"] == 5 and line["value"] > 0
    assert line["cpu_baseline"]["sample_factors"] == 10000 and line["e2e"]["value"] == line["value"]


def test_c_oracle_faithful_linear_form_loses_spd_where_the_closed_form_converges():
    """Documents WHY the rewind is needed: with the reference's fourth-moment form of the linear
    factors' Vddmu (ngd/NGDFactorizedLinear.h:107-119) cfg3 at N = 2 000 stops being SPD within a dozen iterations, the
    closed form 2CA/T (same algebra) keeps going."""
    import gvi_oracle as o
    import gvi_oracle_c as oc
    from gaussianvi_b200 import problems
    spec = problems.make_cfg3(N=2000)
    lean = oc.COracle(spec, o.table)
    for _ in range(12):
        st = lean.iterate(schedule=1)
        assert st.status == 0 and st.accepted
    faithful = oc.COracle(spec, o.table)
    status = [faithful.iterate(schedule=0).status for _ in range(12)]
    assert status[0] == 0 and status[1] == 0 and status[2] == 0
    assert any(s != 0 for s in status), status


def _cfg5_worker(rank, world, port, n_problems, N, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gvi_oracle as o
    import gvi_oracle_c as oc
    from gaussianvi_b200 import problems
    from gaussianvi_b200.dist import gather_problem_results, shard_problems
    first, count = shard_problems(n_problems, rank, world)
    spec = problems.make_cfg5(n_problems=count, N=N, first_seed=1000 + first)
    c = oc.COracle(spec, o.table)
    for _ in range(3):
        st = c.iterate(schedule=1)
        assert st.status == 0 and st.accepted and st.n_backtrack == 0
    parts = gather_problem_results((first, count, c.mean().reshape(count, -1)))
    if rank == 0:
        out.put(parts)
    dist.barrier()
    dist.destroy_process_group()


def test_cfg5_problem_batch_shards_over_two_ranks_gloo():
    """SURVEY 8(e), independent problems: contiguous ranges of the problem index per rank, no collective on the iteration
    path, one gather at the end -- the gathered means equal the single-rank batch of all problems (CPU oracle on both
    sides: this pins the host-side sharding / seeding logic that bench.py --config cfg5 uses)."""
    import numpy as np
    import torch.multiprocessing as mp
    import gvi_oracle as o
    import gvi_oracle_c as oc
    from gaussianvi_b200 import problems
    from gaussianvi_b200.dist import shard_problems
    n_problems, N = 5, 30
    assert [shard_problems(5, r, 2) for r in range(2)] == [(0, 3), (3, 2)]
    assert [shard_problems(4096, r, 8) for r in (0, 7)] == [(0, 512), (3584, 512)]
    assert sum(shard_problems(10, r, 4)[1] for r in range(4)) == 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    procs = [ctx.Process(target=_cfg5_worker, args=(r, 2, port, n_problems, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    parts = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = oc.COracle(problems.make_cfg5(n_problems=n_problems, N=N), o.table)
    for _ in range(3):
        full.iterate(schedule=1)
    want = full.mean().reshape(n_problems, -1)
    got = np.zeros_like(want)
    for first, count, mu in parts:
        got[first:first + count] = mu
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
