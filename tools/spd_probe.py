"""Find (sigma, clearance) for which cfg3 stays SPD through 10 iterations at full size (dev tool)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import gaussianvi_b200 as gv
from gaussianvi_b200 import problems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
ctx = gv.Context(0)
for sigma in (0.1, 0.05):
    for clearance in (None, 0.3, 0.6, 1.0):
        spec = problems.make_cfg3(N=N, sigma=sigma, clearance=clearance)
        p = problems.build_device_problem(ctx, spec)
        opts = gv.Problem.default_opts()
        res = []
        try:
            for it in range(12):
                st = p.iterate(opts)
                res.append((st.n_backtrack, round(st.cost, 3)))
            (E0, _, _), = p.moments()
            print(sigma, clearance, "OK", res[:3], res[-1], "active factors", float((E0 > 0).mean()))
        except gv.GviError as e:
            print(sigma, clearance, "FAIL at iter", len(res), e)
        p.close()
