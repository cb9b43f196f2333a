"""dev tool: snapshot/restore behaviour at full size"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import gaussianvi_b200 as gv
from gaussianvi_b200 import problems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reuse = int(sys.argv[2]) if len(sys.argv) > 2 else 1
spec = problems.make_cfg3(N=N)
ctx = gv.Context(0)
p = problems.build_device_problem(ctx, spec)
opts = gv.Problem.default_opts()
opts.reuse_accepted_sweep = reuse
opts.niters_lowtemp = 1 << 30
def show(tag, st):
    print(tag, "cost %.9f new %.9f acc %d nb %d sweeps %d/%d" % (st.cost, st.new_cost, st.accepted, st.n_backtrack, st.n_moment_sweeps, st.n_cost_sweeps))
show("it0", p.iterate(opts))
p.snapshot_save()
for i in range(14):
    show("a%d" % i, p.iterate(opts))
p.snapshot_restore()
for i in range(4):
    show("b%d" % i, p.iterate(opts))
