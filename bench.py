#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json: NGD iterations / s (and sigma-point evaluations / s) of the
factorized NGD-GVI hot path at N = 100 000 nonlinear factors, d = 4, sparse Gauss-Hermite degree 6 (953 nodes).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl gvib200|reference] [--config cfg3|cfg5]

One "step" = one full NGD iteration (quadrature sweep(s) + precision assembly + block-tridiagonal solve + candidate +
selected inverse / log det + line-search cost) of the synthetic cfg3 trajectory.

  value     whole-job NGD iterations / s with all state resident in HBM (gvib200_ngd_iterate), CUDA-event timed on the
            problem's stream, max over ranks.  At N > 1 every rank owns its own contiguous 100k-factor time segment
            of one long chain (weak scaling; the boundary records and the cost travel through peer-mapped mailboxes as
            NVLink stores issued by the kernels); value = N_ranks * iterations / s, i.e. iterations / s normalised to
            100k factors.  `strong_scaling` (N > 1): one 100k-factor chain cut over the ranks.
  e2e       the same iteration through the C-ABI with HOST buffers: per step set_state(mu, Lambda) from pinned host
            memory, one iteration, mean + covariance blocks read back.
  roofline  dominant kernel (the fused sigma-point / cost / moment kernel K1): algorithmic FP64 flops per launch
            (89 per sigma point) / average launch duration from per-launch CUDA events, against the FP64
            FMA peak measured live by a DFMA micro-benchmark (MEASURED_PEAKS.json carries no FP64 figure).
  schedule_faithful / culling   the same steps with the reference's sweep schedule / with every factor evaluated.
  cpu_baseline  the oracle's C restatement (oracle/gvi_oracle_c.c, OpenMP) on a bounded sample of the same workload.

--impl reference times the CPU path alone (rank 0): the reference's own schedule and arithmetic of one iteration, state
rewound from a host-side snapshot like the GPU arm, at the full 100k factors when that fits the time budget.
--config cfg5: BASELINE configs[4], 4096 independent N=1k problems sharded over the GPUs (an extra bench line).
"""
import argparse
import json
import os
import pathlib
import statistics
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_FACTORS = 100_000
DEG = 6
N_NODES = 953
FLOPS_FULL = 89     # per sigma point, full-moment sweep, d = 4, planar hinge
FLOPS_COST = 61     # per sigma point, cost-only sweep
K1_DRAM_BYTES_NCU = 14230784  # dram read + write bytes of one K1S launch: STATIC figure from the ncu --set full capture profiles/r2_k1s_ncu.txt
METRIC = "NGD iters/sec & sigma-pt evals/sec at N=100k factors, d=4, SpGH deg 6"
UNIT = "NGD iters/s"
CPU_SAMPLE_FACTORS = 10_000
WORKLOAD = "cfg3: NGD-GH hinge-SDF factors + LTV GP prior, d=4, N=100k factors, sparse-GH degree 6 (953 nodes)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="gvib200", choices=["gvib200", "reference"])
    ap.add_argument("--schedule", default="reuse", choices=["reuse", "faithful"],
                    help="reuse: the accepted trial's full-moment sweep seeds the next iteration (identical results); "
                         "faithful: 1 moment sweep + T_ls cost sweeps per iteration")
    ap.add_argument("--factors", type=int, default=N_FACTORS, help="(development only) factors per GPU")
    ap.add_argument("--config", default="cfg3", choices=["cfg3", "cfg5"],
                    help="cfg3: the headline chain (BASELINE configs[2]); cfg5: 4096 independent N=1k problems sharded "
                         "over the GPUs (configs[4]) -- an extra bench line, the driver runs the default")
    ap.add_argument("--problems", type=int, default=4096, help="cfg5: total number of independent problems")
    ap.add_argument("--rewind-every", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-factors", type=int, default=int(os.environ.get("GVIB200_BENCH_CPU_FACTORS", "0")),
                    help="(development / tests) factors of the CPU sample; default: chosen from a probe iteration")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_arm(n_factors, steps, warmup, schedule, rewind_every=4):
    """Time the oracle's C restatement on a bounded sample of cfg3 (n_factors factors); returns
    (iterations / s normalised to 100k factors, seconds per sample iteration, threads, T_ls, psi sweeps / iteration).

    Like the GPU arm the state is rewound to a snapshot taken after one set-up iteration, every `rewind_every` steps,
    so that every timed step does the same work.  With the reference's own arithmetic (schedule 0) the rewind is also
    what keeps the run alive: NGDFactorizedLinear's O(dim^4) fourth-moment form of Vddmu (ngd/NGDFactorizedLinear.h:
    107-119) feeds the rounding error of the pairwise marginal's inverse back into the next precision, the error grows
    ~1000x per iteration and Vddmu stops being positive definite after 6 un-rewound iterations of this workload (the
    algebraically identical closed form 2CA/T, which the GPU path and the lean schedule use, converges in 22)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import gvi_oracle as o          # table generator of the oracle
    import gvi_oracle_c as oc       # the C restatement (CPU baseline; never part of the product path)
    from gaussianvi_b200 import problems
    spec = problems.make_cfg3(N=n_factors)
    c = oc.COracle(spec, o.table)
    st = c.iterate(schedule=schedule)  # set-up iteration (also the snapshot the GPU arm rewinds to)
    if st.status != 0 or not st.accepted:
        raise RuntimeError(f"CPU oracle set-up iteration failed: status {st.status}, accepted {st.accepted}")
    snap = c.snapshot()
    rewind_every = max(1, int(rewind_every))
    it = 0

    def step():
        nonlocal it
        if it and it % rewind_every == 0:
            c.restore(snap)
        it += 1
        return c.iterate(schedule=schedule)

    for _ in range(warmup):
        step()
    c.restore(snap)
    it = 0
    times, nb, sweeps = [], [], []
    for k in range(steps):
        restore = (it and it % rewind_every == 0)
        t = time.perf_counter()
        st = step()
        times.append(time.perf_counter() - t)  # includes the (memcpy) rewind, as on the GPU
        nb.append(st.n_backtrack + 1)
        sweeps.append(st.n_psi_sweeps)
        if st.status != 0 or not st.accepted:
            raise RuntimeError(f"CPU oracle iteration failed at timed step {k} ({(it - 1) % rewind_every + 1} iterations "
                               f"after the snapshot{', rewound' if restore else ''}): status {st.status}, "
                               f"accepted {st.accepted}, n_backtrack {st.n_backtrack}")
    t_iter = sum(times) / len(times)
    scale = n_factors / N_FACTORS
    return scale / t_iter, t_iter, oc.num_threads(), statistics.mean(nb), statistics.mean(sweeps)


def cpu_sample_factors(steps, warmup, budget_s=150.0):
    """The reference arm runs the TRUE 100k-factor workload when (steps + warmup) iterations fit `budget_s` on this
    host (probed with one iteration at the 10k-factor sample: every stage is linear in the factor count), else the
    10k-factor sample."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import gvi_oracle as o
    import gvi_oracle_c as oc
    from gaussianvi_b200 import problems
    c = oc.COracle(problems.make_cfg3(N=CPU_SAMPLE_FACTORS), o.table)
    c.iterate(schedule=0)
    t = time.perf_counter()
    c.iterate(schedule=0)
    t10k = time.perf_counter() - t
    est = t10k * (N_FACTORS / CPU_SAMPLE_FACTORS) * (steps + warmup + 1)
    return (N_FACTORS if est <= budget_s else CPU_SAMPLE_FACTORS), t10k


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    n_sample, t_probe = cpu_sample_factors(steps, warmup)
    if args.cpu_factors:
        n_sample = args.cpu_factors
    rewind = min(args.rewind_every, 4)
    value, t_iter, threads, tls, sweeps = cpu_arm(n_sample, steps, warmup, schedule=0, rewind_every=rewind)
    if n_sample == N_FACTORS:
        sample = (f"one NGD iteration per step of the cfg3 generator at the full {N_FACTORS} factors ({N_FACTORS + 2} states)")
    else:
        sample = (f"one NGD iteration per step of the cfg3 generator at {n_sample} factors ({n_sample + 2} states; the full "
                  f"size would need ~{t_probe * 10 * (steps + warmup + 1):.0f} s on this host); work is linear in the factor "
                  f"count, value scaled by {n_sample}/{N_FACTORS}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * t_iter * (N_FACTORS / n_sample), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "factors": N_FACTORS, "states": N_FACTORS + 2, "dim_state": 4, "gh_degree": DEG,
                   "nodes": N_NODES, "rewind": f"host-side snapshot restore every {rewind} steps (inside the timed region)"},
        "sigma_pt_evals_per_s": value * N_FACTORS * N_NODES * sweeps,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + f"; the reference's own schedule ({sweeps:.0f} psi sweeps, (3+T_ls) chain "
                                            f"inversions per iteration, T_ls={tls:.2f}) and arithmetic (three separate "
                                            f"integrals per factor, O(dim^4) Vddmu loop of the linear factors)",
                         "sample_factors": n_sample, "seconds_per_sample_iteration": t_iter,
                         "why_port": "the reference is header-only C++ on Eigen 3.4 + GSL + a MATLAB-generated table; none "
                                     "is in this image, so oracle/_ref cannot be built"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rank, world, local_rank):
    import numpy as np
    import torch
    import gaussianvi_b200 as gv
    from gaussianvi_b200 import problems

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    K, W = max(1, args.steps), max(3, args.warmup)
    N = args.factors
    ctx = gv.Context(local_rank)
    if world > 1:
        # ONE chain of world * (N + 1) + 1 states cut along the time axis: every rank owns a contiguous segment of N (+1)
        # hinge factors; per block-tridiagonal pass the ranks exchange their boundary records (96 doubles each) and per
        # cost evaluation (cost, flags) through each other's peer-mapped mailboxes (NVLink stores issued by the kernels;
        # NCCL all-gathers when no mailbox could be mapped)
        from gaussianvi_b200.dist import attach_nccl
        attach_nccl(ctx, rank, world)
        spec = problems.make_cfg3_segment(rank, world, N=N, deg=DEG)
    else:
        spec = problems.make_cfg3(N=N, deg=DEG)
    prob = problems.build_device_problem(ctx, spec)
    info = prob.info()
    pts = int(info.sigma_points_per_sweep)
    opts = gv.Problem.default_opts()
    opts.niters_lowtemp = 1 << 30          # no temperature switch inside the timed run
    opts.reuse_accepted_sweep = 1 if args.schedule == "reuse" else 0
    fp64_peak = ctx.fp64_peak_tflops()

    # ---- set-up: one iteration, then snapshot the steady state the timed blocks rewind to.  N > 1: a few more iterations
    # from that snapshot (rewound afterwards, fewer than the rewind interval) so that both NCCL communicators have their
    # connections and buffers established before anything is measured
    prob.iterate(opts)
    prob.snapshot_save()
    if world > 1:
        for _ in range(min(5, max(1, args.rewind_every - 1))):
            prob.iterate(opts)
        prob.snapshot_restore()
    # clocks / throttle reasons are sampled from the warm-up on through every measured region of this run (the K timed
    # steps alone last a few milliseconds: too short for more than one NVML sample)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(W):
        prob.iterate(opts)
    prob.snapshot_restore()

    def run_steps(k, collect=None, o=None):
        # gvib200_optimize (the reference user's call, GVIGH::optimize) in blocks of --rewind-every iterations
        o = opts if o is None else o
        i = 0
        while i < k:
            if i:
                prob.snapshot_restore()     # device-to-device rewind, keeps every step's work identical
            n = min(args.rewind_every, k - i)
            sts = prob.optimize(n, o)
            if len(sts) != n:
                raise SystemExit("bench.py: gvib200_optimize stopped early (converged) inside a timed block")
            if collect is not None:
                collect.extend(sts)
            i += n

    # ---- device-resident timed region
    stats = []
    barrier()
    prob.evaluated_factors(reset=True)
    l0 = ctx.launch_count()
    t_wall = time.perf_counter()
    prob.timer_start()
    run_steps(K, stats)
    ms = prob.timer_stop()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = ctx.launch_count() - l0
    ms = max_over_ranks(ms)
    tls = statistics.mean(s.n_backtrack + 1 for s in stats)
    n_full = sum(s.n_moment_sweeps for s in stats)
    n_cost = sum(s.n_cost_sweeps for s in stats)
    if not all(s.accepted for s in stats):
        raise SystemExit("bench.py: a timed NGD iteration was not accepted")
    ms_per_step = ms / K
    value = world * (N / N_FACTORS) * 1e3 / ms_per_step
    # free-space culling (library default, bit-identical results): only part of the factors is evaluated per sweep; every
    # throughput / roofline figure below counts the sigma points that were actually evaluated
    n_eval = prob.evaluated_factors(reset=True)                   # factor evaluations over the K timed steps
    eval_frac = n_eval / float(info.n_gh_factors * (n_full + n_cost)) if (n_full + n_cost) else 1.0
    evals_nominal = sum_over_ranks(pts * (n_full + n_cost)) / (ms * 1e-3)
    evals = sum_over_ranks(n_eval * N_NODES) / (ms * 1e-3)

    # ---- per-launch profile of the same K steps (event pair per launch; separate pass, not the timed one)
    prob.snapshot_restore()
    prob.profile_begin()
    run_steps(K)
    prof = prob.profile_end()
    k1 = prof.get("k_moments<full>", (0, 0.0))
    k1_ms = k1[1] / max(k1[0], 1)
    n_eval_prof = prob.evaluated_factors(reset=True)
    pts_launch = n_eval_prof * N_NODES / max(k1[0], 1)            # sigma points evaluated per launch of K1
    k1_tflops = pts_launch * FLOPS_FULL / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else 0.0

    # ---- K1 timed alone (nothing else on the GPU: inside the iteration the HBM-bound linear factors share the SMs with it):
    # the culling + prologue pass and K1 launched back to back ten times, K1's own launches from their event pairs
    prob.snapshot_restore()
    prob.iterate(opts)
    prob.evaluated_factors(reset=True)
    prob.profile_begin()
    k1_stage_ms, _ = prob.time_stage(5, 10, opts)
    prof_alone = prob.profile_end()
    n_eval_alone = prob.evaluated_factors(reset=True) / 10.0
    ka = prof_alone.get("k_moments<full>", (0, 0.0))
    k1_alone_ms = ka[1] / max(ka[0], 1)
    kc = prof_alone.get("k_cull", (0, 0.0))
    cull_alone_ms = kc[1] / max(kc[0], 1)
    k1_alone_tflops = n_eval_alone * N_NODES * FLOPS_FULL / (k1_alone_ms * 1e-3) / 1e12 if k1_alone_ms > 0 else 0.0

    # ---- the same K steps with the culling switched off (every sigma point of every factor evaluated)
    prob.snapshot_restore()
    prob.set_option("cull", 0)
    for _ in range(3):
        prob.iterate(opts)
    prob.snapshot_restore()
    barrier()
    prob.timer_start()
    run_steps(K)
    ms_all = max_over_ranks(prob.timer_stop())
    barrier()
    value_all = world * (N / N_FACTORS) * 1e3 / (ms_all / K)
    prob.set_option("cull", 1)
    prob.snapshot_restore()
    prob.iterate(opts)
    prob.snapshot_restore()
    # ---- the same K steps with the reference's sweep schedule (1 moment sweep + T_ls cost sweeps per iteration, no reuse of
    # the accepted trial's sweep): identical iterates, one more quadrature sweep per iteration
    opts_f = gv.Problem.default_opts()
    opts_f.niters_lowtemp = 1 << 30
    opts_f.reuse_accepted_sweep = 0
    for _ in range(3):
        prob.iterate(opts_f)
    prob.snapshot_restore()
    barrier()
    prob.timer_start()
    run_steps(K, None, opts_f)
    ms_faithful = max_over_ranks(prob.timer_stop())
    barrier()
    value_faithful = world * (N / N_FACTORS) * 1e3 / (ms_faithful / K)
    prob.snapshot_restore()
    prob.iterate(opts)
    prob.snapshot_restore()
    # ---- strong scaling (N > 1): ONE chain of ~100k factors cut over the ranks, N_FACTORS / world hinge factors each
    strong = None
    if world > 1:
        Ns = N // world
        spec_s = problems.make_cfg3_segment(rank, world, N=Ns, deg=DEG)
        prob_s = problems.build_device_problem(ctx, spec_s)
        prob_s.iterate(opts)
        prob_s.snapshot_save()
        for _ in range(W):
            prob_s.iterate(opts)
        prob_s.snapshot_restore()
        barrier()
        prob_s.timer_start()
        for i in range(K):
            if i and i % args.rewind_every == 0:
                prob_s.snapshot_restore()
            st = prob_s.iterate(opts)
        ms_s = max_over_ranks(prob_s.timer_stop())
        barrier()
        strong = {"value": 1e3 / (ms_s / K), "unit": UNIT, "ms_per_step": ms_s / K, "factors_total": Ns * world,
                  "factors_per_gpu": Ns, "note": f"one chain of {world * (Ns + 1) + 1} states cut into {world} time segments"}
        prob_s.close()
    clocks = sampler.stop()
    prof_total = sum(v[1] for v in prof.values())
    # chain engine: algorithmic HBM bytes per block-tridiagonal pass (~8*8*d^2 per state per inversion)
    S, d = info.num_states, info.dim_state
    chain_ms = sum(prof.get(k, (0, 0.0))[1] for k in ("k_bt_forward", "k_bt_top", "k_bt_back"))
    chain_passes = 2 * K  # one dmu solve + one selected inverse per iteration (T_ls = 1)
    chain_bytes = 8 * 8 * d * d * S
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "of measured (MEASURED_PEAKS.json)" if peaks else "of fallback (6650 GB/s)"
    flops_iter = eval_frac * pts * (FLOPS_FULL * n_full + FLOPS_COST * n_cost) / K   # executed (culled factors cost nothing)

    # ---- end to end through the C-ABI with host buffers (pinned), every step: H2D state, iterate, D2H result
    # host buffers in the C-ABI layout (column-major d x d blocks), pinned
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        t.numpy()[...] = a
        return t.numpy()
    mu_h = pinned(np.ascontiguousarray(spec.mu0, dtype=np.float64))
    pD_h = pinned(np.ascontiguousarray(np.transpose(spec.prec0_D, (0, 2, 1))))
    pO_h = pinned(np.ascontiguousarray(np.transpose(spec.prec0_O, (0, 2, 1))))
    out_mu = pinned(np.zeros_like(mu_h))
    out_cD = pinned(np.zeros_like(pD_h))
    out_cO = pinned(np.zeros((max(S - 1, 1), d, d)))
    e2e_steps = max(3, min(K, 12))
    h2d = mu_h.nbytes + pD_h.nbytes + pO_h.nbytes
    d2h = out_mu.nbytes + out_cD.nbytes + pO_h.nbytes
    # (a) one handle, synchronous calls: upload, iterate, download one after the other
    prob.set_state_raw(mu_h, pD_h, pO_h); prob.iterate(opts); prob.get_mean_into(out_mu); prob.get_cov_blocks_into(out_cD, out_cO)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prob.set_state_raw(mu_h, pD_h, pO_h)      # H2D + selected inverse + factor marginals
        st = prob.iterate(opts)                    # one NGD iteration
        prob.get_mean_into(out_mu)                 # D2H
        prob.get_cov_blocks_into(out_cD, out_cO)   # D2H
    torch.cuda.synchronize()
    e2e_sync_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_sync_value = world * (N / N_FACTORS) * e2e_steps / e2e_sync_s
    # (b) the same calls' asynchronous variants over several handles of the same problem (what a caller with a stream of
    # independent planning problems does): every step still uploads ITS state from pinned host memory, runs one iteration
    # and downloads ITS mean + covariance blocks, but the upload / download of one handle overlap the iteration of the
    # other (both PCIe directions and the SMs busy at once).  Every result is checked against the synchronous one.
    e2e_value, pipe_steps = e2e_sync_value, e2e_steps
    # (N > 1: one handle only -- two handles of a distributed problem would interleave their peer-mailbox exchanges on
    #  different streams, and the epoch order of those exchanges must be the same on every rank)
    if world == 1:
        # handles in flight: one uploading, one iterating, one downloading (a handle's stream keeps its own download -> next
        # upload -> selected inverse -> iteration in order).  Measured on the B200 box: 1 handle synchronous 545 it/s, 2
        # handles 941, 3 handles 1149, 4 handles 1146 (compute bound: a step from a host-provided state costs two quadrature
        # sweeps, a selected inverse and the factor marginals on top of the iteration's chain passes).
        NH = int(os.environ.get("GVIB200_BENCH_E2E_HANDLES", "3"))
        extra = [problems.build_device_problem(ctx, spec) for _ in range(NH - 1)]
        ref_mu, ref_cD = out_mu.copy(), out_cD.copy()
        handles = [(prob, (out_mu, out_cD, out_cO))]
        for h in extra:
            handles.append((h, (pinned(np.zeros_like(mu_h)), pinned(np.zeros_like(pD_h)), pinned(np.zeros((max(S - 1, 1), d, d))))))
        for h, o in handles:                            # warm every handle up (first-use set-up of the new ones)
            h.set_state_raw_async(mu_h, pD_h, pO_h); h.iterate(opts); h.get_mean_into_async(o[0]); h.get_cov_blocks_into_async(o[1], o[2])
            h.sync()
        pipe_steps = NH * e2e_steps
        barrier()
        t0 = time.perf_counter()
        handles[0][0].set_state_raw_async(mu_h, pD_h, pO_h)
        for i in range(pipe_steps):
            h, o = handles[i % NH]
            if i + 1 < pipe_steps:
                handles[(i + 1) % NH][0].set_state_raw_async(mu_h, pD_h, pO_h)  # next step's H2D: overlaps this iteration
            st = h.iterate(opts)                                                 # blocks on THIS handle's cost only
            h.get_mean_into_async(o[0])                                          # D2H: overlaps the next iterations
            h.get_cov_blocks_into_async(o[1], o[2])
        for h, _ in handles:
            h.sync()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e_value = world * (N / N_FACTORS) * pipe_steps / e2e_s
        for _, o in handles:
            if not (np.array_equal(o[0], ref_mu) and np.array_equal(o[1], ref_cD)):
                raise SystemExit("bench.py: a pipelined end-to-end step does not reproduce the synchronous result bit for bit")
        for h in extra:
            h.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample: ~10-30 s of CPU work)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = args.cpu_factors or CPU_SAMPLE_FACTORS
        v_ref, t_ref, threads, tls_c, sw_ref = cpu_arm(n_cpu, 8, 1, schedule=0)
        v_lean, t_lean, _, _, sw_lean = cpu_arm(n_cpu, 8, 1, schedule=1)
        cpu = {"value": v_ref, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"8 NGD iterations of the cfg3 generator at {n_cpu} factors after a set-up iteration and 1 warm-up "
                         f"(state rewound every 4), reference schedule and arithmetic ({sw_ref:.0f} psi sweeps / iteration); "
                         f"scaled by {n_cpu}/{N_FACTORS} (work is linear in the factor count)",
               "sample_factors": n_cpu, "seconds_per_sample_iteration": t_ref,
               "lean_schedule_value": v_lean, "lean_psi_sweeps": sw_lean,
               "lean_note": "same C code with the GPU path's schedule: one fused moment sweep + T_ls cost sweeps, closed-form "
                            "linear factors, (1+T_ls) chain inversions -- the hardware-only comparison"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "factors_per_gpu": N, "states_per_gpu": S, "dim_state": d, "gh_degree": DEG, "nodes": N_NODES,
                       "linear_factors_per_gpu": int(info.n_linear_factors), "schedule": args.schedule,
                       "T_ls_mean": tls, "parallelism": "1 GPU" if world == 1 else
                       f"{world} ranks: ONE chain of {world * (N + 1) + 1} states cut along the time axis, one contiguous segment "
                       f"of ~{N} hinge factors per rank (weak scaling); per iteration 2 boundary exchanges (96 doubles per rank) + 1 cost / flag "
                       f"exchange (4 doubles per rank) as NVLink peer stores issued by the kernels (mailboxes mapped over CUDA IPC)",
                       "l2": "working set per iteration (state, factor marginals, chain workspace, SDF: ~0.25 GB) exceeds the "
                             "126 MB L2; no flush",
                       "rewind": f"device-side snapshot restore every {args.rewind_every} steps (inside the timed region)",
                       "culling": "library default: a hinge factor whose whole sigma-point box lies provably in free space (bound from "
                                  "the distance field) is not evaluated -- its moments are exactly zero either way, results are "
                                  "bit-identical (tests/test_gpu_parity.py::test_free_space_culling_is_bit_identical)"},
            "culling": {"evaluated_factor_fraction": eval_frac, "value_all_factors_evaluated": value_all,
                        "ms_per_step_all_factors_evaluated": ms_all / K},
            "schedule_faithful": {"value": value_faithful, "ms_per_step": ms_faithful / K,
                                  "note": "reuse_accepted_sweep = 0: the reference's 1 moment sweep + T_ls cost sweeps per "
                                          "iteration (identical iterates); the headline value reuses the accepted trial's "
                                          "full-moment sweep as the next iteration's gradient sweep"},
            "strong_scaling": strong,
            "sigma_pt_evals_per_s": evals, "sigma_pt_evals_per_s_nominal": evals_nominal,
            "wall_ms_per_step": 1e3 * t_wall / K,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": pipe_steps,
                    "call": ("gvib200_set_state_async + gvib200_ngd_iterate + gvib200_get_mean_async + gvib200_get_cov_blocks_async "
                             "per step, three handles in rotation (one uploading, one iterating, one downloading); every "
                             "step uploads its state from pinned host memory and downloads its mean + covariance blocks")
                            if world == 1 else "gvib200_set_state + gvib200_ngd_iterate + gvib200_get_mean + gvib200_get_cov_blocks",
                    "single_handle_synchronous": {"value": e2e_sync_value, "steps": e2e_steps,
                                                  "call": "gvib200_set_state + gvib200_ngd_iterate + gvib200_get_mean + "
                                                          "gvib200_get_cov_blocks, one after the other"}},
            "roofline": {"bound": "fp64", "kernel": "k_moments<4, CostPlanarHinge, full> (K1)",
                         "achieved": k1_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": k1_tflops / fp64_peak if fp64_peak else None, "traffic": K1_DRAM_BYTES_NCU,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of k_moments_sym per launch, ncu --set full "
                                           "capture profiles/r2_k1s_ncu.txt (14.23 MB; the capture of the final code, profiles/r2_final_iteration_ncu.txt, shows the same) -- a static figure, not measured in this run (bytes; the kernel is FP64 bound, not HBM bound)",
                         "peak_source": "DFMA micro-benchmark run in this process (gvib200_fp64_peak); MEASURED_PEAKS.json "
                                        "has no FP64 figure",
                         "algorithmic_flops_per_launch": pts_launch * FLOPS_FULL, "avg_launch_ms": k1_ms, "launches": k1[0],
                         "sigma_points_evaluated_per_launch": pts_launch, "sigma_points_nominal_per_launch": pts,
                         "kernel_alone": {"ms": k1_alone_ms, "achieved": k1_alone_tflops,
                                          "frac": k1_alone_tflops / fp64_peak if fp64_peak else None,
                                          "culling_prologue_pass_ms": cull_alone_ms, "stage_ms": k1_stage_ms,
                                          "note": "K1 with nothing else on the GPU (gvib200_time_stage 5: the fused culling + "
                                                  "prologue pass and K1 back to back, each launch timed by its own event pair); "
                                                  "inside the iteration the closed-form linear factors (HBM bound) run "
                                                  "underneath K1 on the side stream"},
                         "share_of_step": k1[1] / prof_total if prof_total else None,
                         "whole_iteration": {"flops": flops_iter, "achieved": flops_iter / (ms_per_step * 1e-3) / 1e12,
                                             "frac": flops_iter / (ms_per_step * 1e-3) / 1e12 / fp64_peak if fp64_peak else None},
                         "chain_engine_hbm": {"bound": "hbm", "achieved": chain_bytes * chain_passes / (chain_ms * 1e-3) / 1e9 if chain_ms else None,
                                              "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src,
                                              "frac": chain_bytes * chain_passes / (chain_ms * 1e-3) / 1e9 / hbm_peak if chain_ms else None,
                                              "algorithmic_bytes_per_pass": chain_bytes}},
            "kernel_ms_per_step": {k: v[1] / K for k, v in sorted(prof.items())},
            "kernel_launches_per_step": {k: v[0] / K for k, v in sorted(prof.items())},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    prob.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ cfg5 (extra bench line)
CFG5_N = 1000
CFG5_METRIC = "independent-problem NGD iterations/s: 4096 problems of N=1k factors, d=4, SpGH deg 6, sharded over the GPUs"
CFG5_UNIT = "problem-iterations/s"


def cfg5_cpu(n_problems, steps, warmup, total):
    """C oracle on a block-diagonal batch of `n_problems` problems (reference schedule and arithmetic, state rewound every
    4 steps like cpu_arm); returns problem-iterations / s and a description."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import gvi_oracle as o
    import gvi_oracle_c as oc
    from gaussianvi_b200 import problems
    c = oc.COracle(problems.make_cfg5(n_problems=n_problems, N=CFG5_N), o.table)
    st = c.iterate(schedule=0)
    if st.status != 0 or not st.accepted:
        raise RuntimeError(f"cfg5 CPU oracle set-up iteration failed: status {st.status}")
    snap = c.snapshot()
    times = []
    for k in range(warmup + steps):
        if k and k % 4 == 0:
            c.restore(snap)
        t = time.perf_counter()
        st = c.iterate(schedule=0)
        if k >= warmup:
            times.append(time.perf_counter() - t)
        if st.status != 0 or not st.accepted:
            raise RuntimeError(f"cfg5 CPU oracle iteration {k} failed: status {st.status}, accepted {st.accepted}")
    t_iter = sum(times) / len(times)
    return n_problems / t_iter, oc.num_threads(), (
        f"{steps} NGD iterations of a batch of {n_problems} of the {total} problems after a set-up iteration and {warmup} "
        f"warm-up (reference schedule and arithmetic, state rewound every 4); problems are independent, so "
        f"problem-iterations/s does not depend on the batch size")


def run_cfg5(args, rank, world, local_rank):
    """BASELINE configs[4] / SURVEY 8(e) first bullet: `--problems` independent copies of the cfg3 generator at N = 1000
    (seeds 1000 ...), sharded over the ranks as contiguous ranges; every rank runs its range as ONE block-diagonal batch
    on its GPU.  No collective on the iteration path (strong scaling: the total is fixed)."""
    K, W = max(1, args.steps), max(3, args.warmup)
    if args.impl == "reference":
        if rank != 0:
            return
        v, threads, sample = cfg5_cpu(8, K, max(0, args.warmup), args.problems)
        print(json.dumps({
            "impl": "reference", "metric": CFG5_METRIC, "value": v, "unit": CFG5_UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": max(0, args.warmup), "ms_per_step": 1e3 * args.problems / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg5: {args.problems} independent cfg3 problems of N={CFG5_N} factors"},
            "cpu_baseline": {"value": v, "unit": CFG5_UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": CFG5_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    import numpy as np
    import torch
    import gaussianvi_b200 as gv
    from gaussianvi_b200 import problems
    from gaussianvi_b200.dist import shard_problems
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    first, count = shard_problems(args.problems, rank, world)
    ctx = gv.Context(local_rank)
    t_build = time.perf_counter()
    spec = problems.make_cfg5(n_problems=count, N=CFG5_N, first_seed=1000 + first, ctx=ctx)   # LTV set-up on the device
    prob = problems.build_device_problem(ctx, spec)
    t_build = time.perf_counter() - t_build
    info = prob.info()
    opts = gv.Problem.default_opts()
    opts.niters_lowtemp = 1 << 30
    opts.reuse_accepted_sweep = 1 if args.schedule == "reuse" else 0
    prob.iterate(opts)
    prob.snapshot_save()
    for _ in range(W):
        prob.iterate(opts)
    prob.snapshot_restore()

    def run_steps(k, collect=None):
        for i in range(k):
            if i and i % args.rewind_every == 0:
                prob.snapshot_restore()
            st = prob.iterate(opts)
            if collect is not None:
                collect.append(st)

    sampler = ClockSampler(local_rank)
    stats = []
    barrier()
    prob.evaluated_factors(reset=True)
    l0 = ctx.launch_count()
    sampler.start()
    prob.timer_start()
    run_steps(K, stats)
    ms = prob.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - l0
    ms = max_over_ranks(ms)
    if not all(s.accepted and s.n_backtrack == 0 for s in stats):
        raise SystemExit("bench.py: a timed cfg5 iteration was not accepted at the first trial")
    n_eval = prob.evaluated_factors(reset=True)
    value = args.problems * K / (ms * 1e-3)
    # per-launch profile of the same steps (separate pass)
    prob.snapshot_restore()
    prob.profile_begin()
    run_steps(K)
    prof = prob.profile_end()
    k1 = prof.get("k_moments<full>", (0, 0.0))
    n_eval_prof = prob.evaluated_factors(reset=True)
    k1_ms = k1[1] / max(k1[0], 1)
    k1_tflops = n_eval_prof * N_NODES * FLOPS_FULL / max(k1[0], 1) / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else 0.0
    fp64_peak = ctx.fp64_peak_tflops()
    # end to end with host buffers
    S, d = info.num_states, info.dim_state

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        t.numpy()[...] = a
        return t.numpy()
    mu_h = pinned(np.ascontiguousarray(spec.mu0, dtype=np.float64))
    pD_h = pinned(np.ascontiguousarray(np.transpose(spec.prec0_D, (0, 2, 1))))
    pO_h = pinned(np.ascontiguousarray(np.transpose(spec.prec0_O, (0, 2, 1))))
    out_mu, out_cD, out_cO = pinned(np.zeros_like(mu_h)), pinned(np.zeros_like(pD_h)), pinned(np.zeros((max(S - 1, 1), d, d)))
    e2e_steps = max(3, min(K, 6))
    prob.set_state_raw(mu_h, pD_h, pO_h); prob.iterate(opts); prob.get_mean_into(out_mu); prob.get_cov_blocks_into(out_cD, out_cO)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prob.set_state_raw(mu_h, pD_h, pO_h)
        prob.iterate(opts)
        prob.get_mean_into(out_mu)
        prob.get_cov_blocks_into(out_cD, out_cO)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    # ---- the same batch with a line search PER PROBLEM (gvib200_set_batch / gvib200_batch_iterate: per-problem cost from the
    # per-node log pivots, per-problem step sizes; exactly the independent optimizer objects of the reference also when
    # the problems disagree about a trial)
    Sb = spec.meta["states_per_problem"]
    prob.snapshot_restore()
    prob.set_batch(np.arange(count + 1, dtype=np.int32) * Sb)
    for _ in range(3):
        prob.batch_iterate(opts)
    prob.snapshot_restore()
    barrier()
    prob.timer_start()
    trials = 0
    for i in range(K):
        if i and i % args.rewind_every == 0:
            prob.snapshot_restore()
        sts, ntr = prob.batch_iterate(opts)
        trials += ntr
    ms_pp = max_over_ranks(prob.timer_stop())
    barrier()
    if not all(s_.accepted for s_ in sts):
        raise SystemExit("bench.py: a timed per-problem iteration left a problem without an accepted trial")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, threads, sample = cfg5_cpu(8, 6, 1, args.problems)
        cpu = {"value": v, "unit": CFG5_UNIT, "cores": threads, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps({
            "metric": CFG5_METRIC, "value": value, "unit": CFG5_UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"cfg5: {args.problems} independent cfg3 problems (seeds 1000..) of N={CFG5_N} hinge factors, "
                                   f"{CFG5_N + 2} states, d=4, sparse-GH degree 6",
                       "problems_total": args.problems, "problems_per_gpu": count, "states_per_gpu": S,
                       "factors_per_gpu": int(info.n_factors), "schedule": args.schedule,
                       "parallelism": f"{world} rank(s): contiguous ranges of the problem index, one block-diagonal batch per GPU, "
                                      f"no collective on the iteration path",
                       "line_search": "one step size and the summed cost per batch (every timed iteration accepts the first trial, "
                                      "where this equals the independent iterations)",
                       "l2": "working set per iteration exceeds the 126 MB L2; no flush",
                       "rewind": f"device-side snapshot restore every {args.rewind_every} steps (inside the timed region)",
                       "setup_s": t_build},
            "iterations_per_s_per_batch": K / (ms * 1e-3),
            "per_problem_line_search": {"value": args.problems * K / (ms_pp * 1e-3), "unit": CFG5_UNIT, "ms_per_step": ms_pp / K,
                                        "trial_sweeps_per_iteration": trials / K,
                                        "call": "gvib200_set_batch + gvib200_batch_iterate (tests/test_gpu_batch.py: every "
                                                "problem walks the path of its own independent oracle run)"},
            "evaluated_factor_fraction": n_eval / float(info.n_gh_factors * sum(s.n_moment_sweeps + s.n_cost_sweeps for s in stats)),
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": args.problems * e2e_steps / e2e_s, "unit": CFG5_UNIT,
                    "h2d_bytes_per_step": int(mu_h.nbytes + pD_h.nbytes + pO_h.nbytes),
                    "d2h_bytes_per_step": int(out_mu.nbytes + out_cD.nbytes + pO_h.nbytes), "steps": e2e_steps,
                    "call": "gvib200_set_state + gvib200_ngd_iterate + gvib200_get_mean + gvib200_get_cov_blocks"},
            "roofline": {"bound": "fp64", "kernel": "k_moments_sym<4, CostPlanarHinge, full> (K1)", "achieved": k1_tflops,
                         "peak": fp64_peak, "unit": "TFLOP/s", "frac": k1_tflops / fp64_peak if fp64_peak else None,
                         "traffic": None, "avg_launch_ms": k1_ms, "launches": k1[0],
                         "peak_source": "DFMA micro-benchmark run in this process (gvib200_fp64_peak)"},
            "kernel_ms_per_step": {k: v[1] / K for k, v in sorted(prof.items())},
            "cpu_baseline": cpu}), flush=True)
    prob.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    # stdout carries ONE JSON line: libraries that write to fd 1 (NCCL prints its version banner there) are sent to
    # stderr for the whole run; the JSON line goes to the saved descriptor
    global print
    real_out = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    _print = print

    def print(*a, **k):  # noqa: A001 -- only the final JSON line is printed through this
        k.setdefault("file", real_out)
        _print(*a, **k)
        real_out.flush()
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config == "cfg5":
        run_cfg5(args, rank, world, local_rank)
        return
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
