"""Quick stage timing at the headline shape (dev tool; bench.py is the contract)."""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import gaussianvi_b200 as gv
from gaussianvi_b200 import problems

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
t = time.time(); spec = problems.make_cfg3(N=N); print(f"gen {time.time()-t:.2f}s")
ctx = gv.Context(0)
print("fp64 peak TFLOP/s", ctx.fp64_peak_tflops())
t = time.time(); p = problems.build_device_problem(ctx, spec); print(f"build+set_state {time.time()-t:.2f}s")
opts = gv.Problem.default_opts()
n_nodes = 953
for stage, name in [(0, "moment sweep (K2+K1+linear)"), (1, "cost sweep"), (2, "assemble+solve"), (3, "candidate+selinv")]:
    p.time_stage(stage, 2, opts)
    ms, nl = p.time_stage(stage, 10, opts)
    extra = ""
    if stage == 0:
        extra = f"  -> {N*n_nodes*89/ms/1e9:.2f} TFLOP/s (89 flop/pt)"
    if stage == 1:
        extra = f"  -> {N*n_nodes*61/ms/1e9:.2f} TFLOP/s (61 flop/pt)"
    print(f"stage {stage} {name}: {ms:.4f} ms/rep, {nl} launches{extra}")
for reuse in (0, 1):
    p.set_state(spec.mu0, spec.prec0_D, spec.prec0_O); p.reset_schedule()
    opts.reuse_accepted_sweep = reuse
    p.iterate(opts)
    t = time.time(); sts = [p.iterate(opts) for _ in range(8)]; dt = (time.time() - t) / 8
    print(f"reuse={reuse}: {dt*1e3:.3f} ms/iter wall, backtracks {[s.n_backtrack for s in sts]}, cost {sts[-1].cost:.6f}")
