"""Short, deterministic workload for ncu captures: cfg3 at N factors, two NGD iterations (dev tool)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import gaussianvi_b200 as gv
from gaussianvi_b200 import problems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reuse = int(sys.argv[2]) if len(sys.argv) > 2 else 0
spec = problems.make_cfg3(N=N)
ctx = gv.Context(0)
p = problems.build_device_problem(ctx, spec)
opts = gv.Problem.default_opts()
opts.reuse_accepted_sweep = reuse
for _ in range(2):
    st = p.iterate(opts)
print("cost", st.cost, "launches", ctx.launch_count())
