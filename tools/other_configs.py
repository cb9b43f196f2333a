"""dev: device time per iteration of the other BASELINE.json configs (parity-test cases, not bench lines): cfg1 (1-D example),
cfg2 (all-linear, S = 1000), cfg4 (Prox-GVI, dim-12 factors, sparse-GH degree 4) and cfg5 (independent N = 1000 problems
batched block-diagonally).  Usage: python tools/other_configs.py [n_cfg4_states] [n_cfg5_problems]"""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import gaussianvi_b200 as gv
from gaussianvi_b200 import problems

S4 = int(sys.argv[1]) if len(sys.argv) > 1 else 10_001
B5 = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = gv.Context(0)


def run(name, spec, iters, prox=False, reuse=True):
    t0 = time.perf_counter()
    p = problems.build_device_problem(ctx, spec, prox=prox)
    opts = gv.Problem.default_opts()
    opts.step_size_base = spec.meta.get("step_size_base", 0.55)
    opts.niters_lowtemp = 1 << 30
    opts.reuse_accepted_sweep = 0 if prox else (1 if reuse else 0)
    step = p.prox_iterate if prox else p.iterate
    step(opts)
    p.snapshot_save()
    for _ in range(2):
        step(opts)
    p.snapshot_restore()
    p.timer_start()
    acc = 0
    for i in range(iters):
        if i and i % 4 == 0:
            p.snapshot_restore()
        acc += step(opts).accepted
    ms = p.timer_stop() / iters
    info = p.info()
    p.snapshot_restore()
    p.profile_begin()
    for i in range(4):
        step(opts)
    prof = p.profile_end()
    print("    per-kernel ms / iteration:", {k: round(v[1] / 4, 4) for k, v in sorted(prof.items())})
    print(f"{name}: states {info.num_states} d {info.dim_state} GH factors {info.n_gh_factors} linear {info.n_linear_factors} "
          f"sigma points / sweep {info.sigma_points_per_sweep}: {ms:.4f} ms / iteration ({1e3 / ms:.0f} it/s), accepted {acc}/{iters}, "
          f"set-up {time.perf_counter() - t0:.1f} s", flush=True)
    p.close()


run("cfg1 1-D example (deg 10)", problems.make_cfg1(), 8)
run("cfg2 all-linear S=1000", problems.make_cfg2(), 8)
run(f"cfg4 Prox-GVI dim-12 GH factors deg 4, S={S4}", problems.make_cfg4(S=S4), 8, prox=True)
run(f"cfg5 {B5} x (N=1000) batched", problems.make_cfg5(n_problems=B5), 8)
