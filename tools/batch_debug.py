"""dev: per-problem paths of a heterogeneous batch (tests/test_gpu_batch.py) next to independent oracle runs."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
for q in (ROOT, ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(q))
import numpy as np
import gaussianvi_b200 as gv
from gaussianvi_b200 import capi, problems
import oracle_bridge as ob
import test_gpu_batch as tb
base, lowtemp, mb, niters = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
subs = tb.make_subs()
refs = []
for s in subs:
    s.meta["step_size_base"] = base
    s.meta["niters_lowtemp"] = lowtemp
    r = ob.build_oracle(s, niters=niters)
    r.set_max_iter_backtrack(mb)
    refs.append((r, r.optimize()))
spec, off = tb.concat(subs)
ctx = gv.Context(0)
p = problems.build_device_problem(ctx, spec)
p.set_batch(off)
opts = capi.Problem.default_opts()
opts.step_size_base, opts.niters_lowtemp, opts.max_backtrack, opts.reuse_accepted_sweep = base, lowtemp, mb, 1
print("cost0 gpu", p.batch_costs())
for it in range(niters):
    stats, ntr = p.batch_iterate(opts)
    print("iteration", it, "trials", ntr)
    for q, (ref, recs) in enumerate(refs):
        s = stats[q]
        o = recs[it] if it < len(recs) else None
        print(f"   q{q} gpu nb {s.n_backtrack} acc {s.accepted} cost {s.cost:.12g} new {s.new_cost:.12g} step {s.step:.4g} sw {s.switched_high_T} cv {s.converged} st {s.status} | "
              + (f"oracle nb {o.n_backtrack} acc {int(o.accepted)} cost {o.cost:.12g} step {o.step:.4g}" if o else "oracle stopped"))
