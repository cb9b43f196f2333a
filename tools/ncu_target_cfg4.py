"""Short, deterministic workload for ncu captures of the dim-12 path: cfg4 (Prox-GVI, sparse-GH degree 4) at S states,
two Prox iterations (dev tool)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import gaussianvi_b200 as gv
from gaussianvi_b200 import problems
S = int(sys.argv[1]) if len(sys.argv) > 1 else 10_001
spec = problems.make_cfg4(S=S)
ctx = gv.Context(0)
p = problems.build_device_problem(ctx, spec, prox=True)
opts = gv.Problem.default_opts()
opts.step_size_base = spec.meta["step_size_base"]
opts.niters_lowtemp = 1 << 30
for _ in range(2):
    st = p.prox_iterate(opts)
print("cost", st.cost, "launches", ctx.launch_count())
