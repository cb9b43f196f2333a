"""Dev tool: walks the oracle through the line-search regimes of tests/test_gpu_branches.py and prints, per regime, the
accept / reject path and the smallest relative margin |new_cost - cost_iter| / |cost_iter| over all decisions (a NaN cost
of a not-SPD candidate counts as an infinite margin).  A regime is usable as a GPU parity test when the margin is far
above the GPU-vs-oracle cost agreement (~1e-12)."""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
for q in (ROOT, ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(q))
import numpy as np  # noqa: E402
import oracle_bridge as ob  # noqa: E402
from gaussianvi_b200 import problems  # noqa: E402


def run(name, spec, niters, base, lowtemp, maxbt):
    spec.meta["step_size_base"] = base
    spec.meta["niters_lowtemp"] = lowtemp
    ref = ob.build_oracle(spec, niters=niters)
    ref.set_max_iter_backtrack(maxbt)
    margins = []
    orig = ref.onestep_linesearch

    def wrapped(step, dmu, dprec):
        c, a, b = orig(step, dmu, dprec)
        ci = ref.records[-1].cost
        margins.append(abs(c - ci) / abs(ci) if np.isfinite(c) else np.inf)
        return c, a, b
    ref.onestep_linesearch = wrapped
    recs = ref.optimize()
    print(f"{name}: base {base} niters_lowtemp {lowtemp} max_backtrack {maxbt} niters {niters} -> {len(recs)} iterations, "
          f"T {ref.T}, min margin {min(margins):.2e}, {sum(np.isinf(margins))} not-SPD candidates")
    print("    path (n_backtrack, accepted):", [(r.n_backtrack, int(r.accepted)) for r in recs])


if __name__ == "__main__":
    for base, lt, mb, n in [(2.5, 100, 1, 40), (2.0, 3, 10, 40), (1.6, 3, 10, 12)]:
        run("cfg1", problems.make_cfg1(), n, base, lt, mb)
    for base, lt, mb, n in [(1.8, 100, 1, 16), (2.2, 4, 5, 12), (1.3, 100, 6, 8), (2.2, 100, 2, 16), (2.6, 100, 3, 12), (3.0, 100, 3, 12)]:
        run("cfg3 N=40", problems.make_cfg3(N=40), n, base, lt, mb)
