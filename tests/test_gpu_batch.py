"""Batches of independent problems with a line search PER PROBLEM (gvib200_set_batch / gvib200_batch_iterate): every problem
of a block-diagonal batch must walk exactly the path of its own optimizer object in the reference (GVIGH::optimize,
gvibase/GVI-GH-GBP-impl.h:33-130) -- here: of an independent oracle run -- also when the problems DISAGREE about a trial
(some accept, some back-track, some exhaust their back-tracking and switch temperature or converge).

The regimes were picked with the oracle so that the problems do disagree and every accept / reject decision has a
relative cost margin >= 2e-7 (or is a not-SPD candidate, whose cost is NaN on both sides)."""
import numpy as np
import pytest

import oracle_bridge as ob
from gaussianvi_b200 import capi, problems
from gaussianvi_b200.problems import GhGroupSpec, LinGroupSpec, ProblemSpec

pytestmark = pytest.mark.gpu

SUBS = [dict(N=30, clearance=0.6, seed=1, prec0=100.0), dict(N=40, clearance=0.3, seed=2, prec0=10.0),
        dict(N=25, clearance=0.15, seed=3, prec0=1000.0), dict(N=35, clearance=0.45, seed=4, prec0=3.0),
        dict(N=28, clearance=0.05, seed=5, prec0=30.0)]


def make_subs():
    subs = []
    for i, kw in enumerate(SUBS):
        s = problems.make_cfg3(**kw)
        rng = np.random.default_rng(100 + i)
        s.mu0 = s.mu0 + 0.05 * i * rng.standard_normal(s.mu0.shape)
        subs.append(s)
    return subs


def concat(subs):
    """Block-diagonal batch of heterogeneous sub-problems (same group structure): what problems.make_cfg5 does for equal ones."""
    d = subs[0].d
    off = np.concatenate([[0], np.cumsum([s.S for s in subs])]).astype(np.int32)
    out = ProblemSpec(S=int(off[-1]), d=d)
    out.sdf = subs[0].sdf
    out.mu0 = np.concatenate([s.mu0 for s in subs])
    out.prec0_D = np.concatenate([s.prec0_D for s in subs])
    out.prec0_O = np.zeros((out.S - 1, d, d))
    for b, s in enumerate(subs):
        out.prec0_O[off[b]:off[b] + s.S - 1] = s.prec0_O
    for gi in range(len(subs[0].groups)):
        g0 = subs[0].groups[gi]
        starts = np.concatenate([s.groups[gi].start + off[b] for b, s in enumerate(subs)]).astype(np.int32)
        if isinstance(g0, GhGroupSpec):
            out.groups.append(GhGroupSpec(g0.kind, g0.dim, g0.deg, starts, g0.params, g0.T, g0.T_high))
        else:
            cat = lambda name: np.concatenate([getattr(s.groups[gi], name) for s in subs])
            out.groups.append(LinGroupSpec(
                start=starts, Lambda=cat("Lambda"), Psi=cat("Psi"), mu_t=cat("mu_t"), Kinv=cat("Kinv"),
                C=np.concatenate([np.broadcast_to(np.asarray(s.groups[gi].C, float), (len(s.groups[gi].start),)) for s in subs]),
                T=g0.T, T_high=g0.T_high))
    out.meta = dict(subs[0].meta, name="batch")
    return out, off


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


@pytest.mark.parametrize("base,lowtemp,max_backtrack,niters", [(1.7, 100, 5, 6), (1.9, 100, 1, 7), (1.7, 3, 6, 7)])
def test_per_problem_line_search_matches_independent_oracle_runs(gpu_ctx, base, lowtemp, max_backtrack, niters):
    subs = make_subs()
    refs = []
    for s in subs:
        s.meta["step_size_base"] = base
        s.meta["niters_lowtemp"] = lowtemp
        r = ob.build_oracle(s, niters=niters)
        r.set_max_iter_backtrack(max_backtrack)
        refs.append((r, r.optimize()))
    paths = [[(x.n_backtrack, int(x.accepted)) for x in recs] for _, recs in refs]
    assert len(set(map(tuple, paths))) > 1, "the regime must make the problems disagree"

    spec, off = concat(subs)
    p = problems.build_device_problem(gpu_ctx, spec)
    p.set_batch(off)
    opts = capi.Problem.default_opts()
    opts.step_size_base = base
    opts.niters_lowtemp = lowtemp
    opts.max_backtrack = max_backtrack
    opts.reuse_accepted_sweep = 1
    hist = []
    for it in range(niters):
        if it == niters - 2:   # the last two iterations through gvib200_batch_optimize (GVIGH::optimize of every problem)
            two = p.batch_optimize(2, opts)
            hist.extend(two + [two[-1]] * (2 - len(two)))   # (stops early only when every problem has converged)
            break
        stats, ntr = p.batch_iterate(opts)
        hist.append(stats)
        # trial sweeps of the batch: the largest number any problem needed (a problem that exhausts its back-tracking takes
        # max_backtrack + 1 trials and one more, with step 0, that parks it at its current state)
        need = 1
        for _, recs in refs:
            if it < len(recs):
                need = max(need, recs[it].n_backtrack + 1 if recs[it].accepted else recs[it].n_backtrack + 1)
        assert ntr >= 1 and ntr <= need + 1, (it, ntr, need)
    for q, (ref, recs) in enumerate(refs):
        for it in range(niters):
            s = hist[it][q]
            if it < len(recs):
                r = recs[it]
                assert s.status == 0
                assert bool(s.accepted) == r.accepted, (q, it)
                assert s.n_backtrack == r.n_backtrack, (q, it, s.n_backtrack, r.n_backtrack)
                assert abs(s.cost - r.cost) < 1e-9 * max(1.0, abs(r.cost)), (q, it, s.cost, r.cost)
                if r.accepted:
                    assert abs(s.step - r.step) < 1e-15
            else:  # the reference's optimizer object has stopped (converged): the problem stays where it is
                assert s.converged == 1 and not s.accepted
    mu = p.mean()
    cD, cO = p.covariance()
    for q, (ref, recs) in enumerate(refs):
        a, b = off[q], off[q + 1]
        d = spec.d
        assert rel(mu[a * d:b * d], ref.mean()) < 1e-7, q
        assert rel(cD[a:b], ref.cov.D) < 1e-7, q
        assert rel(cO[a:b - 1], ref.cov.O) < 1e-7, q
    p.close()


def test_agreeing_batch_equals_joint_line_search(gpu_ctx):
    """When every problem accepts the same trial the per-problem path and the joint path (gvib200_ngd_iterate on the batch)
    produce the same iterates."""
    spec = problems.make_cfg5(n_problems=4, N=40)
    Sb = spec.meta["states_per_problem"]
    off = np.arange(5, dtype=np.int32) * Sb
    opts = capi.Problem.default_opts()
    opts.reuse_accepted_sweep = 1
    pj = problems.build_device_problem(gpu_ctx, spec)
    for _ in range(4):
        st = pj.iterate(opts)
        assert st.accepted and st.n_backtrack == 0
    pb = problems.build_device_problem(gpu_ctx, spec)
    pb.set_batch(off)
    for _ in range(4):
        stats, ntr = pb.batch_iterate(opts)
        assert ntr == 1 and all(s.accepted and s.n_backtrack == 0 for s in stats)
    assert rel(pb.mean(), pj.mean()) < 1e-13
    assert rel(pb.covariance()[0], pj.covariance()[0]) < 1e-12
    assert abs(pb.batch_costs().sum() - st.new_cost) < 1e-10 * abs(st.new_cost)
    pj.close()
    pb.close()


@pytest.mark.parametrize("n_problems", [8, 300])
def test_per_problem_costs_on_tiled_and_three_level_chains(gpu_ctx, n_problems):
    """The per-node log pivots come out of every level of the chain engine (tiles, the separator chain's own tiles on a
    three-level plan, the top): per-problem costs of a long batch equal the costs of the same problems run one by one,
    and two per-problem iterations reproduce the single-problem iterates (8 x 1002 states: tiled; 300 x 1002: three levels)."""
    N = 1000
    spec = problems.make_cfg5(n_problems=n_problems, N=N, ctx=gpu_ctx)
    Sb = spec.meta["states_per_problem"]
    off = np.arange(n_problems + 1, dtype=np.int32) * Sb
    opts = capi.Problem.default_opts()
    opts.reuse_accepted_sweep = 1
    pb = problems.build_device_problem(gpu_ctx, spec)
    info = pb.info()
    assert info.chain_tiles > 0 and info.chain_levels == (3 if n_problems == 300 else 2)
    pb.set_batch(off)
    c0 = pb.batch_costs()
    sample = sorted({0, 1, n_problems // 2, n_problems - 1})
    singles = {}
    for q in sample:
        ps = problems.build_device_problem(gpu_ctx, problems.make_cfg3(N=N, seed=1000 + q, ctx=gpu_ctx))
        st = ps.iterate(opts)
        assert abs(c0[q] - st.cost) < 1e-11 * abs(st.cost), (q, c0[q], st.cost)
        ps.iterate(opts)
        singles[q] = (ps.mean(), ps.covariance()[0])
        ps.close()
    for _ in range(2):
        stats, ntr = pb.batch_iterate(opts)
        assert ntr == 1 and all(s.accepted for s in stats)
    mu, cD = pb.mean(), pb.covariance()[0]
    d = spec.d
    for q, (m1, c1) in singles.items():
        a, b = off[q], off[q + 1]
        assert rel(mu[a * d:b * d], m1) < 1e-12, q
        assert rel(cD[a:b], c1) < 1e-11, q
    assert abs(pb.batch_costs().sum() - sum(s.new_cost for s in stats)) < 1e-12 * abs(sum(s.new_cost for s in stats))
    pb.close()


def test_problem_with_indefinite_vddmu_is_reported_alone(gpu_ctx):
    """One problem of the batch has a Vddmu without a Cholesky factor (sigma = 15.5 on a path through the obstacles, as in
    test_gpu_branches.test_indefinite_vddmu_...): ITS status is GVIB200_ENOTSPD and its state stays where it is, every
    iteration; its neighbours -- across whose boundary the garbage of its solve must not leak -- follow their own oracle."""
    kws = [dict(N=25, sigma=15.5, clearance=1.2, seed=4), dict(N=30, sigma=15.5, clearance=None),
           dict(N=35, sigma=15.5, clearance=1.5, seed=5)]
    subs = [problems.make_cfg3(**kw) for kw in kws]
    niters = 3
    refs = {}
    for q in (0, 2):
        r = ob.build_oracle(subs[q], niters=niters)
        refs[q] = (r, r.optimize())
        assert all(x.accepted for x in refs[q][1])
    spec, off = concat(subs)
    p = problems.build_device_problem(gpu_ctx, spec)
    p.set_batch(off)
    opts = capi.Problem.default_opts()
    opts.reuse_accepted_sweep = 1
    mu0, cov0 = p.mean(), p.covariance()[0]
    for it in range(niters):
        stats, _ = p.batch_iterate(opts)
        assert stats[1].status == capi.E_NOTSPD and not stats[1].accepted
        for q in (0, 2):
            r = refs[q][1][it]
            assert stats[q].status == 0 and stats[q].accepted and stats[q].n_backtrack == r.n_backtrack
            assert abs(stats[q].cost - r.cost) < 1e-9 * abs(r.cost), (q, it)
    mu, cD = p.mean(), p.covariance()[0]
    d = spec.d
    a, b = off[1], off[2]
    assert np.array_equal(mu[a * d:b * d], mu0[a * d:b * d])
    assert rel(cD[a:b], cov0[a:b]) < 1e-13   # recomputed from the same precision, not copied
    for q in (0, 2):
        a, b = off[q], off[q + 1]
        assert rel(mu[a * d:b * d], refs[q][0].mean()) < 1e-7, q
        assert rel(cD[a:b], refs[q][0].cov.D) < 1e-7, q
    p.close()
