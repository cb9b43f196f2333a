"""GPU tests at BASELINE.json's full sizes (N = 100 000 factors, d = 4, degree 6) through size-independent properties:
the two moment kernels agree, free space gives exact zeros, the C oracle agrees on a random sample of factors, the chain
engine agrees with the C oracle's inverse_GBP, the iteration is monotone, SPD and bit-reproducible, and the all-linear
chain moves the precision by (1 - a) Lambda + a Lambda* towards the exact Gaussian posterior."""
import ctypes as C

import numpy as np
import pytest

import gvi_oracle as o
import gvi_oracle_c as oc
import oracle_bridge as ob
from gaussianvi_b200 import capi, problems

pytestmark = pytest.mark.gpu
N_FULL = 100_000


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


@pytest.fixture(scope="module")
def batch(gpu_ctx):
    spec = problems.make_factor_batch(N=N_FULL)
    return spec, problems.build_device_problem(gpu_ctx, spec)


def test_full_size_moment_kernels_agree_and_match_c_oracle(batch):
    spec, p = batch
    (a0, a1, a2), = p.moments()
    p.set_option("generic_k1", 1)
    (b0, b1, b2), = p.moments()
    p.set_option("generic_k1", 0)
    assert np.array_equal(a0 == 0, b0 == 0)                       # free space: exact zeros in both kernels
    free = a0 == 0
    assert free.sum() > 1000 and not a1[free].any() and not a2[free].any()
    nz = ~free
    # signed weights (sum |w| = 231) cancel: a sum may be orders of magnitude smaller than its terms, so next to the
    # relative tolerance an absolute floor of ~231 * max psi * eps is allowed
    FLOOR = 1e-12
    s0 = np.abs(b0[nz])
    assert (np.abs(a0[nz] - b0[nz]) / (s0 + FLOOR / 1e-11)).max() < 1e-11
    s2 = np.abs(b2[nz]).reshape(nz.sum(), -1).max(1)
    assert (np.abs(a2[nz] - b2[nz]).reshape(nz.sum(), -1).max(1) / (s2 + FLOOR / 1e-11)).max() < 1e-11
    s1 = np.maximum(np.abs(b1[nz]).max(1), np.sqrt(s2 * s0))
    assert (np.abs(a1[nz] - b1[nz]).max(1) / (s1 + FLOOR / 1e-11)).max() < 1e-11
    # C oracle (x-space sums in node order, the reference's arithmetic) on a random sample of 2000 factors
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(N_FULL, 2000, replace=False)).astype(np.int32)
    sub = problems.ProblemSpec(S=len(idx), d=4)
    sub.sdf = spec.sdf
    sub.groups.append(problems.GhGroupSpec(capi.COST_PLANAR_HINGE, 4, 6, np.arange(len(idx), dtype=np.int32), spec.groups[0].params))
    sub.mu0 = spec.mu0.reshape(-1, 4)[idx].reshape(-1)
    sub.prec0_D = spec.prec0_D[idx]
    sub.prec0_O = np.zeros((len(idx) - 1, 4, 4))
    c = oc.COracle(sub, o.table)
    r0, r1, r2 = c.moments(faithful=False)
    g0, g2 = a0[idx], a2[idx]
    m = r0 != 0
    assert np.array_equal(m, g0 != 0)
    assert (np.abs(g0[m] - r0[m]) / (np.abs(r0[m]) + 1e-2)).max() < 1e-10     # floor: 1e-12 absolute
    d2 = np.abs(g2[m] - r2[m]).reshape(m.sum(), -1).max(1) / (np.abs(r2[m]).reshape(m.sum(), -1).max(1) + 1e-2)
    assert d2.max() < 1e-10


def test_full_size_iteration_properties(gpu_ctx):
    spec = problems.make_cfg3(N=N_FULL)
    opts = capi.Problem.default_opts()
    opts.reuse_accepted_sweep = 1
    runs = []
    for _ in range(2):
        p = problems.build_device_problem(gpu_ctx, spec)
        stats = [p.iterate(opts) for _ in range(6)]
        runs.append((np.array([s.cost for s in stats]), p.mean(), p.covariance()[0]))
        assert all(s.accepted and s.n_backtrack == 0 and s.status == 0 for s in stats)   # SPD throughout, T_ls = 1
        p.close()
    costs = runs[0][0]
    assert np.all(np.diff(costs) < 0)                                   # monotone decrease
    assert np.array_equal(runs[0][0], runs[1][0])                       # bit-reproducible: no atomics, fixed trees
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    # the covariance blocks the optimizer holds are the selected inverse of its precision (C oracle inverse_GBP)
    p = problems.build_device_problem(gpu_ctx, spec)
    for _ in range(2):
        p.iterate(opts)
    pD, pO = p.precision()
    cD, cO = p.covariance()
    S, d = spec.S, spec.d
    Dc = np.ascontiguousarray(np.transpose(pD, (0, 2, 1)))
    Oc = np.ascontiguousarray(np.transpose(pO, (0, 2, 1)))
    rD, rO = np.zeros_like(Dc), np.zeros((S, d, d))
    dp = C.POINTER(C.c_double)
    assert oc.lib().orc_inverse_gbp(S, d, Dc.ctypes.data_as(dp), Oc.ctypes.data_as(dp), rD.ctypes.data_as(dp), rO.ctypes.data_as(dp)) == 0
    assert rel(cD, np.transpose(rD, (0, 2, 1))) < 1e-9
    assert rel(cO, np.transpose(rO[:S - 1], (0, 2, 1))) < 1e-9


def test_full_size_all_linear_chain_moves_towards_exact_posterior(gpu_ctx):
    """cfg2 generator at S = 100 000: every accepted step is Lambda <- (1 - a) Lambda + a Lambda*, and dmu solves
    Lambda* dmu = -Vdmu (checked by the block-tridiagonal residual)."""
    spec = problems.make_cfg2(S=N_FULL)
    p = problems.build_device_problem(gpu_ctx, spec)
    dmu, dD, dO = p.gradients()
    Vd, VD, VO = p.get_V()
    x = dmu.reshape(-1, 4)
    r = np.einsum("sij,sj->si", VD, x)
    r[:-1] += np.einsum("sij,sj->si", VO, x[1:])
    r[1:] += np.einsum("sji,sj->si", VO, x[:-1])
    assert np.abs(r.reshape(-1) + Vd).max() < 1e-6 * np.abs(Vd).max()   # kappa(Vddmu) ~ 3e5 for the anchored chain
    opts = capi.Problem.default_opts()
    st = p.iterate(opts)
    assert st.accepted and st.n_backtrack == 0
    a = opts.step_size_base * opts.backtrack_ratio
    pD, pO = p.precision()
    assert rel(pD, (1 - a) * spec.prec0_D + a * VD) < 1e-13
    assert rel(pO, (1 - a) * spec.prec0_O + a * VO) < 1e-13


# ---------------------------------------------------------------- end-to-end parity at BASELINE.json's sizes
def _run_vs_c_oracle(gpu_ctx, spec, niters, reuse, step_size_base=None):
    """`niters` NGD iterations on the GPU and in the C oracle (lean schedule = the GPU path's arithmetic: closed-form
    linear factors, direct block solve): per iteration the same decision and cost, at the end mu and the covariance
    blocks within the north star's 1e-7."""
    c = oc.COracle(spec, o.table)
    p = problems.build_device_problem(gpu_ctx, spec)
    opts = capi.Problem.default_opts()
    opts.niters_lowtemp = 1 << 30
    opts.reuse_accepted_sweep = reuse
    base = spec.meta.get("step_size_base", 0.55) if step_size_base is None else step_size_base
    opts.step_size_base = base
    for it in range(niters):
        r = c.iterate(step_size_base=base, schedule=1)
        s = p.iterate(opts)
        assert r.status == 0 and s.status == 0, (it, r.status, s.status)
        assert bool(s.accepted) == bool(r.accepted) and s.n_backtrack == r.n_backtrack, (it, s.n_backtrack, r.n_backtrack)
        assert abs(s.cost - r.cost) < 1e-10 * max(1.0, abs(r.cost)), (it, s.cost, r.cost)
    cD, cO = p.covariance()
    rD, rO = c.cov_blocks()
    e = (rel(p.mean(), c.mean()), rel(cD, rD), rel(cO, rO))
    p.close()
    return e


@pytest.mark.parametrize("reuse", [1, 0])
def test_cfg3_headline_size_ten_iterations_match_c_oracle(gpu_ctx, reuse):
    """BASELINE config 3 at its full size: N = 100 000 hinge factors + 100 001 LTV factors, d = 4, degree 6."""
    e = _run_vs_c_oracle(gpu_ctx, problems.make_cfg3(N=N_FULL), 10, reuse)
    print("cfg3 N=100k, 10 iterations, reuse", reuse, ": rel err mu %.2e cov diag %.2e cov off %.2e" % e)
    assert max(e) < 1e-7


def test_cfg2_s1000_ten_iterations_match_c_oracle(gpu_ctx):
    """BASELINE config 2 at its full size: all-linear 2-D point robot, S = 1000 (999 GP + 2 fixed + anchors)."""
    spec = problems.make_cfg2(S=1000)
    e = _run_vs_c_oracle(gpu_ctx, spec, 10, 0)
    print("cfg2 S=1000, 10 iterations: rel err mu %.2e cov diag %.2e cov off %.2e" % e)
    assert max(e) < 1e-7


@pytest.mark.parametrize("closed_form", [False, True])
def test_cfg4_s1001_prox_iterations_match_oracle(gpu_ctx, closed_form):
    """BASELINE config 4 (Prox-GVI, dim-12 two-state factors, sparse GH degree 4 = 2649 nodes) at S = 1001: three
    iterations against the NumPy oracle (proxgd/ProxGVI-GH-impl.h:124-205)."""
    spec = problems.make_cfg4(S=1001, closed_form=closed_form)
    p = problems.build_device_problem(gpu_ctx, spec, prox=True)
    opts = capi.Problem.default_opts()
    opts.step_size_base = spec.meta["step_size_base"]
    opts.niters_lowtemp = spec.meta["niters_lowtemp"]
    ref = ob.build_oracle_prox(spec, niters=3)
    recs = ref.optimize()
    stats = [p.prox_iterate(opts) for _ in range(3)]
    for s, r in zip(stats, recs):
        assert s.n_backtrack == r.n_backtrack and bool(s.accepted) == r.accepted
        assert abs(s.cost - r.cost) < 1e-9 * max(1.0, abs(r.cost))
    cD, cO = p.covariance()
    e_mu = rel(p.mean(), ref.mean())
    e_cov = rel(np.concatenate([cD.reshape(-1), cO.reshape(-1)]), np.concatenate([ref.cov.D.reshape(-1), ref.cov.O.reshape(-1)]))
    print("prox cfg4 S=1001 closed_form", closed_form, "rel err mu", e_mu, "cov", e_cov)
    assert e_mu < 1e-7 and e_cov < 1e-7
