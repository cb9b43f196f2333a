"""GPU parity on the branches of the optimizer that an accept-first-trial run never takes:
rejected trials (`cnt > 0`: separate candidate kernel + unfused selected inverse), a candidate precision that is not
positive definite treated as a rejection, back-tracking exhaustion -> switch to the high temperature
(gvibase/GVI-GH-GBP-impl.h:104-119), second exhaustion -> `converged` (:117-118), the temperature switch at iteration
`niters_lowtemp` (:49-58), and an indefinite Vddmu reported as GVIB200_ENOTSPD with the state left untouched.

The regimes were picked with the oracle so that EVERY accept / reject decision has a relative cost margin >= 1e-9
(printed by tools/regime_margins.py): GPU and oracle costs agree to ~1e-12, so both must walk the same path."""
import numpy as np
import pytest

import gvi_oracle as o
import oracle_bridge as ob
from gaussianvi_b200 import capi, problems

pytestmark = pytest.mark.gpu

FINAL_TOL = 1e-7


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def run_regime(gpu_ctx, spec, niters, base, lowtemp, max_backtrack, reuse, alpha=1.0):
    spec.meta["step_size_base"] = base
    spec.meta["niters_lowtemp"] = lowtemp
    ref = ob.build_oracle(spec, niters=niters)
    ref.set_max_iter_backtrack(max_backtrack)
    ref.alpha = alpha
    recs = ref.optimize()
    p = problems.build_device_problem(gpu_ctx, spec)
    opts = capi.Problem.default_opts()
    opts.step_size_base = base
    opts.niters_lowtemp = lowtemp
    opts.max_backtrack = max_backtrack
    opts.reuse_accepted_sweep = reuse
    opts.ema_alpha = alpha
    stats = p.optimize(niters, opts)
    return p, stats, ref, recs


def check_path(p, stats, ref, recs, lowtemp, max_backtrack):
    """Per iteration: same accept / reject decisions, same step, same cost; the temperature switch and `converged`
    where GVIGH::optimize puts them; final mean / covariance blocks within 1e-7."""
    assert len(stats) == len(recs), (len(stats), len(recs))
    is_low = True
    for it, (s, r) in enumerate(zip(stats, recs)):
        switched = False
        if it == lowtemp and is_low:
            is_low, switched = False, True
        assert s.status == 0
        assert bool(s.accepted) == r.accepted, it
        assert s.n_backtrack == r.n_backtrack, (it, s.n_backtrack, r.n_backtrack)
        assert abs(s.cost - r.cost) < 1e-9 * max(1.0, abs(r.cost)), (it, s.cost, r.cost)
        if r.accepted:
            assert abs(s.step - r.step) < 1e-15
        converged = False
        if not r.accepted:  # back-tracking exhausted
            assert r.n_backtrack == max_backtrack + 1
            if is_low:
                is_low, switched = False, True
            else:
                converged = True
        assert bool(s.switched_high_T) == switched, it
        assert bool(s.converged) == converged, it
    cD, cO = p.covariance()
    e_mu = rel(p.mean(), ref.mean())
    e_cov = rel(np.concatenate([cD.reshape(-1), cO.reshape(-1)]),
                np.concatenate([ref.cov.D.reshape(-1), ref.cov.O.reshape(-1)]))
    pD, pO = p.precision()
    e_prec = rel(np.concatenate([pD.reshape(-1), pO.reshape(-1)]),
                 np.concatenate([ref.prec.D.reshape(-1), ref.prec.O.reshape(-1)]))
    assert e_mu < FINAL_TOL and e_cov < FINAL_TOL and e_prec < FINAL_TOL, (e_mu, e_cov, e_prec)
    return e_mu, e_cov


@pytest.mark.parametrize("reuse", [0, 1])
@pytest.mark.parametrize("base,lowtemp,max_backtrack,niters,expect", [
    # (0,acc) (2,exhausted -> T_high) (2,exhausted -> converged); every rejected candidate but one is not SPD
    (2.5, 100, 1, 40, dict(iters=3, rejected=True, exhausted=2, converged=True)),
    # switch at iteration 3 by count, rejections of 1..8 trials, converged by exhaustion at iteration 15
    (2.0, 3, 10, 40, dict(iters=16, rejected=True, exhausted=1, converged=True)),
    # a run that ends by its iteration count in the high-temperature phase, rejections on the way
    (1.6, 3, 10, 12, dict(iters=12, rejected=True, exhausted=0, converged=False)),
])
def test_cfg1_rejections_temperature_switch_and_convergence(gpu_ctx, base, lowtemp, max_backtrack, niters, expect, reuse):
    p, stats, ref, recs = run_regime(gpu_ctx, problems.make_cfg1(), niters, base, lowtemp, max_backtrack, reuse)
    assert len(recs) == expect["iters"]
    assert any(r.accepted and r.n_backtrack > 0 for r in recs) or not expect["rejected"] or expect["exhausted"]
    assert sum(1 for r in recs if not r.accepted) == expect["exhausted"]
    e = check_path(p, stats, ref, recs, lowtemp, max_backtrack)
    # after `converged` the optimizer does nothing more (GVI-GH-GBP-impl.h:44-47)
    if expect["converged"]:
        st = p.iterate(capi.Problem.default_opts())
        assert st.converged == 1 and st.accepted == 0 and st.n_moment_sweeps == 0 and st.n_cost_sweeps == 0
    print("cfg1 regime", base, lowtemp, max_backtrack, "reuse", reuse, "path", [(r.n_backtrack, int(r.accepted)) for r in recs], "err", e)


@pytest.mark.parametrize("reuse", [0, 1])
@pytest.mark.parametrize("base,lowtemp,max_backtrack,niters,expect", [
    # every candidate of the first two iterations is not SPD: exhausted -> T_high, exhausted -> converged
    (1.8, 100, 1, 16, dict(iters=2, exhausted=2)),
    # multi-state chain with rejections (2, 2, 1, 0, 2, 1 ...), switch by count at iteration 4, 12 iterations
    (2.2, 4, 5, 12, dict(iters=12, exhausted=0)),
    # rejections early, exhaustion in the low-temperature phase at the last iteration (-> switch to T_high)
    (2.2, 100, 2, 16, dict(iters=16, exhausted=1)),
])
def test_cfg3_chain_rejections_temperature_switch_and_convergence(gpu_ctx, base, lowtemp, max_backtrack, niters, expect, reuse):
    p, stats, ref, recs = run_regime(gpu_ctx, problems.make_cfg3(N=40), niters, base, lowtemp, max_backtrack, reuse)
    assert len(recs) == expect["iters"]
    assert sum(1 for r in recs if not r.accepted) == expect["exhausted"]
    e = check_path(p, stats, ref, recs, lowtemp, max_backtrack)
    print("cfg3 N=40 regime", base, lowtemp, max_backtrack, "reuse", reuse, "path", [(r.n_backtrack, int(r.accepted)) for r in recs], "err", e)


@pytest.mark.parametrize("reuse", [0, 1])
@pytest.mark.parametrize("alpha,base,lowtemp,max_backtrack,niters", [
    (0.5, 0.55, 100, 10, 8),   # every first trial accepted, half of each step taken
    (0.7, 2.2, 4, 5, 12),      # with rejections, the count switch and a final exhaustion (-> converged)
])
def test_ema_update_of_the_cuda_path(gpu_ctx, alpha, base, lowtemp, max_backtrack, niters, reuse):
    """EMA of an accepted proposal, GVIGH::set_alpha of the reference's GPU path (gvibase/GVI-GH-Cuda.h:223,
    GVI-GH-Cuda-impl.h:112-114): alpha * new + (1 - alpha) * current for mean and precision."""
    p, stats, ref, recs = run_regime(gpu_ctx, problems.make_cfg3(N=40), niters, base, lowtemp, max_backtrack, reuse, alpha)
    e = check_path(p, stats, ref, recs, lowtemp, max_backtrack)
    # the EMA really is in effect: the same regime without it ends elsewhere
    p1, _, _, _ = run_regime(gpu_ctx, problems.make_cfg3(N=40), niters, base, lowtemp, max_backtrack, reuse, 1.0)
    assert rel(p1.mean(), p.mean()) > 1e-6
    print("EMA alpha", alpha, "path", [(r.n_backtrack, int(r.accepted)) for r in recs], "err", e)


def test_rejected_trial_with_culling_and_speculation_is_bit_identical(gpu_ctx):
    """The speculative second V-buffer set, the zero-copy cost hand-over and the free-space culling under rejected trials:
    with and without reuse / culling the iterates agree bit for bit (same arithmetic, different schedule)."""
    out = []
    for reuse, cull in ((0, 1), (1, 1), (1, 0)):
        spec = problems.make_cfg3(N=9000)  # > 8192 factors: the one-launch total with the mapped-memory hand-over
        p = problems.build_device_problem(gpu_ctx, spec)
        p.set_option("cull", cull)
        opts = capi.Problem.default_opts()
        opts.step_size_base = 2.2
        opts.niters_lowtemp = 1 << 30
        opts.max_backtrack = 5
        opts.reuse_accepted_sweep = reuse
        stats = p.optimize(7, opts)
        cD, cO = p.covariance()
        out.append((p.mean(), cD, cO, [(s.n_backtrack, s.accepted, s.cost, s.switched_high_T) for s in stats]))
    assert any(s[0] > 0 for s in out[0][3]), out[0][3]   # the regime does reject trials
    for other in out[1:]:
        assert out[0][3] == other[3]
        assert np.array_equal(out[0][0], other[0]) and np.array_equal(out[0][1], other[1]) and np.array_equal(out[0][2], other[2])


def test_cfg3_10k_sixteen_iterations_match_c_oracle(gpu_ctx):
    """Round 1's reference arm died at iteration 6 of this very input: that was the reference's O(dim^4) form of the
    linear factors' Vddmu (tests/test_dist_gloo.py), not the workload -- with the closed form both the C oracle and the
    GPU run on, same decisions, same iterates."""
    import gvi_oracle_c as oc
    N, niters = 10_000, 16
    spec = problems.make_cfg3(N=N)
    c = oc.COracle(spec, o.table)
    p = problems.build_device_problem(gpu_ctx, spec)
    opts = capi.Problem.default_opts()
    opts.niters_lowtemp = 1 << 30
    for it in range(niters):
        r = c.iterate(schedule=1)
        s = p.iterate(opts)
        assert r.status == 0 and s.status == 0, (it, r.status, s.status)
        assert bool(s.accepted) == bool(r.accepted) and s.n_backtrack == r.n_backtrack, it
        assert abs(s.cost - r.cost) < 1e-10 * abs(r.cost), (it, s.cost, r.cost)
    cD, cO = p.covariance()
    rD, rO = c.cov_blocks()
    assert rel(p.mean(), c.mean()) < FINAL_TOL
    assert rel(cD, rD) < FINAL_TOL and rel(cO, rO) < FINAL_TOL


def test_indefinite_vddmu_is_reported_and_leaves_the_state_untouched(gpu_ctx):
    """With the class default sigma = 15.5 (helpers/CudaOperation.h:456) the hinge curvature makes Vddmu indefinite on
    this prior at the very first iteration.  The reference hands such a matrix to CG (ngd/NGD-GH-impl.h:59-60) and gets an
    arbitrary vector; this library (and the oracle's direct solve) report it: GVIB200_ENOTSPD, iteration not counted,
    mean / precision / covariance unchanged, the handle stays usable."""
    import gvi_oracle_c as oc
    spec = problems.make_cfg3(N=300, sigma=15.5, clearance=None)
    c = oc.COracle(spec, o.table)
    assert c.iterate(schedule=1).status != 0
    for reuse in (0, 1):
        p = problems.build_device_problem(gpu_ctx, spec)
        before = (p.mean(), p.precision(), p.covariance())
        opts = capi.Problem.default_opts()
        opts.reuse_accepted_sweep = reuse
        opts.niters_lowtemp = 1  # a counted iteration would switch the temperature on the next call
        for _ in range(2):
            with pytest.raises(capi.GviError) as e:
                p.iterate(opts)
            assert e.value.code == capi.E_NOTSPD
        after = (p.mean(), p.precision(), p.covariance())
        assert np.array_equal(before[0], after[0])
        for b, a in zip(before[1] + before[2], after[1] + after[2]):
            assert np.array_equal(b, a)
        c0, _ = p.cost()  # the handle is still usable
        assert np.isfinite(c0)
