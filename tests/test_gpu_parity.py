"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded
inputs.  Tolerances are BASELINE.json's: per-factor quadrature moments 1e-10 relative (tensor-wise
max norm), final mu / Sigma after the reference's iteration count 1e-7 relative."""
import numpy as np
import pytest

import gvi_oracle as o
import oracle_bridge as ob
from gaussianvi_b200 import capi, problems

pytestmark = pytest.mark.gpu

MOMENT_TOL = 1e-10
FINAL_TOL = 1e-7
GOLDEN = ob.ROOT / "tests" / "golden"


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def rand_spd_chain(rng, S, d):
    D = np.zeros((S, d, d))
    O = np.zeros((max(S - 1, 0), d, d))
    for i in range(S):
        A = rng.standard_normal((d, d))
        D[i] += 0.1 * A @ A.T + np.eye(d)
    for i in range(S - 1):
        B = rng.standard_normal((2 * d, 2 * d))
        M = B @ B.T
        D[i] += M[:d, :d]
        D[i + 1] += M[d:, d:]
        O[i] += M[:d, d:]
    return D, O


# ---------------------------------------------------------------- block-tridiagonal engine (a9, a10)
@pytest.mark.parametrize("S,d", [(1, 4), (2, 4), (3, 1), (9, 4), (33, 2), (100, 4), (257, 3), (1000, 4), (40, 6)])
def test_selected_inverse_matches_dense(gpu_ctx, S, d):
    rng = np.random.default_rng(S * 10 + d)
    D, O = rand_spd_chain(rng, S, d)
    cD, cO, ld = gpu_ctx.selected_inverse(D, O)
    A = o.BlockTri(D, O).dense()
    Ai = np.linalg.inv(A)
    eD = np.array([Ai[i * d:(i + 1) * d, i * d:(i + 1) * d] for i in range(S)])
    assert rel(cD, eD) < 1e-12
    if S > 1:
        eO = np.array([Ai[i * d:(i + 1) * d, (i + 1) * d:(i + 2) * d] for i in range(S - 1)])
        assert rel(cO, eO) < 1e-12
    assert abs(ld - np.linalg.slogdet(A)[1]) < 1e-10 * max(1.0, abs(ld))


def test_selected_inverse_matches_gbp_oracle_long_chain(gpu_ctx):
    """src/GBP.cpp:133-158 property at a size the dense inverse cannot reach: GBP marginals == selected inverse."""
    rng = np.random.default_rng(5)
    S, d = 20000, 4
    D, O = rand_spd_chain(rng, S, d)
    cD, cO, ld = gpu_ctx.selected_inverse(D, O)
    ref = o.inverse_gbp(o.BlockTri(D, O))
    assert rel(cD, ref.D) < 1e-11
    assert rel(cO, ref.O) < 1e-11
    assert abs(ld - o.logdet(o.BlockTri(D, O))) < 1e-9 * abs(ld)


@pytest.mark.parametrize("S,d", [(1, 2), (7, 4), (300, 4), (5000, 4)])
def test_blocktri_solve(gpu_ctx, S, d):
    rng = np.random.default_rng(S + d)
    D, O = rand_spd_chain(rng, S, d)
    rhs = rng.standard_normal(S * d)
    x, ld = gpu_ctx.blocktri_solve(D, O, rhs)
    xe = o.block_solve(o.BlockTri(D, O), rhs)
    assert rel(x, xe) < 1e-11


def test_not_spd_is_reported(gpu_ctx):
    D = np.tile(np.eye(2), (5, 1, 1))
    D[3] = -np.eye(2)
    O = np.zeros((4, 2, 2))
    with pytest.raises(capi.GviError) as e:
        gpu_ctx.selected_inverse(D, O)
    assert e.value.code == capi.E_NOTSPD


# ---------------------------------------------------------------- quadrature known answers (a1, a2)
def single_factor_problem(ctx, kind, dim, deg, params, mean, cov):
    spec = problems.ProblemSpec(S=1, d=dim)
    spec.groups.append(problems.GhGroupSpec(kind, dim, deg, np.zeros(1, np.int32), params, 1.0, 10.0))
    spec.mu0 = np.asarray(mean, float)
    spec.prec0_D = np.linalg.inv(np.atleast_2d(np.asarray(cov, float)))[None]
    spec.prec0_O = np.zeros((0, dim, dim))
    return spec, problems.build_device_problem(ctx, spec)


def test_kat_stereo_1d_deg6(gpu_ctx):
    """tests/test_GH.cpp:134-161: E[phi] = 1.1129, E[(x-mu) phi] = -1.2144 (1e-4), sparse deg 6, mu=20, Sigma=9."""
    prm = capi.Stereo1DParams(20.0, 400.0, 0.1, 0.09, 9.0, 0.05)
    spec, p = single_factor_problem(gpu_ctx, capi.COST_STEREO_1D, 1, 6, prm, [20.0], [[9.0]])
    (E0, E1, E2), = p.moments()
    assert abs(E0[0] - 1.1129) < 1e-4 and abs(E1[0, 0] + 1.2144) < 1e-4
    f = ob.build_factors(spec, fast=False)[0]
    r0, r1, r2 = o.moments(f.psi, [20.0], [[9.0]], f.Z, f.w)
    assert rel(E0[0], r0) < MOMENT_TOL and rel(E1[0], r1) < MOMENT_TOL and rel(E2[0], r2) < MOMENT_TOL


@pytest.mark.parametrize("dim,deg,mean,cov,c,expected,tol", [
    (4, 3, np.zeros(4), 1e-4 * np.eye(4), 1e4, 4.0, 1e-10),                                  # test_gh_spgh.cpp:76-90
    (3, 8, np.ones(3), np.eye(3), 1e4, 6.00e4, 1e-7),                                         # :194-220
    (2, 10, np.ones(2), np.linalg.inv(np.array([[1, -0.74], [-0.74, 1.0]])), 1e4, 6.420866489831914e4, 1e-5),  # :92-124
])
def test_kat_quadratic(gpu_ctx, dim, deg, mean, cov, c, expected, tol):
    spec, p = single_factor_problem(gpu_ctx, capi.COST_QUADRATIC, dim, deg, np.array([c]), mean, cov)
    (E0, E1, E2), = p.moments()
    assert abs(E0[0] - expected) < tol
    f = ob.build_factors(spec, fast=False)[0]
    r0, r1, r2 = o.moments(f.psi, mean, cov, f.Z, f.w)
    # E1 vanishes by symmetry for a centred quadratic: compare against the natural scale sqrt(|Sigma|) |E0|
    scale1 = max(np.abs(r1).max(), np.sqrt(np.abs(cov).max()) * abs(r0))
    assert rel(E0[0], r0) < MOMENT_TOL and np.abs(E1[0] - r1).max() < MOMENT_TOL * scale1 and rel(E2[0], r2) < MOMENT_TOL


def test_moments_hinge_factor_batch(gpu_ctx):
    """SURVEY 8(d) factor-batch micro-input (seed 11) at a size the NumPy oracle finishes in seconds."""
    N = 1500
    spec = problems.make_factor_batch(N=N)
    p = problems.build_device_problem(gpu_ctx, spec)
    (E0, E1, E2), = p.moments()
    covD = spec.meta["Sigma"]     # the generator's covariances (prec0_D is their inverse): nothing read back from the GPU
    g = spec.groups[0]
    psi = ob.psi_for_group(spec, g, 0)
    Z, w = o.table(4, 6)
    mu = spec.mu0.reshape(N, 4)
    worst = [0.0, 0.0, 0.0]
    nz = 0
    for k in range(N):
        r0, r1, r2 = o.moments_fast(psi, mu[k], covD[k], Z, w)
        if r0 == 0.0:  # free space: the hinge is 0 at every node -> exactly 0 on the GPU too
            assert E0[k] == 0.0 and not E1[k].any() and not E2[k].any()
            continue
        nz += 1
        worst[0] = max(worst[0], rel(E0[k], r0))
        worst[1] = max(worst[1], rel(E1[k], r1))
        worst[2] = max(worst[2], rel(E2[k], r2))
    print("hinge factor batch: non-zero factors", nz, "worst rel err", worst)
    assert nz > N // 10
    assert max(worst) < MOMENT_TOL


def test_linear_gp_quadrature_equals_closed_form(gpu_ctx):
    """Closed form == quadrature for linear factors (gp/factorized_opts_linear.h:12-14 'for comparison'):
    GH of a quadratic is exact, so a LINEAR_GP GH factor and the closed-form factor must agree to rounding."""
    S, d, dt = 6, 4, 0.1
    lin = problems.minacc_group(S, 0.8 * np.eye(2), dt)
    rng = np.random.default_rng(1)
    D, O = rand_spd_chain(rng, S, d)
    mu = rng.standard_normal(S * d)
    Phi = -lin.Lambda[0][:, :d]
    rec = np.concatenate([np.tile(Phi.T.reshape(-1), (S - 1, 1)), np.tile(lin.Kinv[0].T.reshape(-1), (S - 1, 1))], axis=1)
    anchor = problems.fixed_prior_group([0, S - 1], np.zeros((2, d)), np.eye(d), d)  # makes Vddmu SPD
    a = problems.ProblemSpec(S=S, d=d, groups=[anchor, lin], mu0=mu, prec0_D=D, prec0_O=O)
    b = problems.ProblemSpec(S=S, d=d, groups=[anchor, problems.GhGroupSpec(capi.COST_LINEAR_GP, 2 * d, 4, lin.start, rec)],
                             mu0=mu, prec0_D=D, prec0_O=O)
    pa, pb = problems.build_device_problem(gpu_ctx, a), problems.build_device_problem(gpu_ctx, b)
    ca, fa = pa.cost()
    cb, fb = pb.cost()
    assert rel(fb, fa) < 1e-10
    pa.gradients(), pb.gradients()
    va, vb = pa.get_V(), pb.get_V()
    for x, y in zip(va, vb):
        assert rel(y, x) < 1e-9


# ---------------------------------------------------------------- end-to-end traces (a3..a12)
def test_cfg1_golden_trace(gpu_ctx):
    """src/1d_example.cpp against the reference's committed outputs data/1d/*.csv (10 iterations)."""
    spec = problems.make_cfg1()
    p = problems.build_device_problem(gpu_ctx, spec)
    opts = capi.Problem.default_opts()
    opts.step_size_base = 0.75
    opts.niters_lowtemp = 10
    g = lambda n: np.loadtxt(GOLDEN / "ref_1d" / f"{n}.csv", delimiter=",").reshape(-1)
    means, covs, precs, costs, fcs = [], [], [], [], []
    for it in range(10):
        means.append(p.mean()[0])
        covs.append(p.covariance()[0][0, 0, 0])
        precs.append(p.precision()[0][0, 0, 0])
        st = p.iterate(opts)
        costs.append(st.cost)
        assert st.accepted == 1 and st.n_backtrack == 0
    assert rel(means, g("mean")) < 1e-10
    assert rel(covs, g("cov")) < 1e-10
    assert rel(precs, g("precision")) < 1e-10
    assert rel(costs, g("cost")) < 1e-10


def test_cfg1_factor_costs_and_costmap(gpu_ctx):
    spec = problems.make_cfg1()
    p = problems.build_device_problem(gpu_ctx, spec)
    opts = capi.Problem.default_opts()
    opts.step_size_base = 0.75
    stats, fc, mt = p.optimize(10, opts, want_traces=True)
    g = np.loadtxt(GOLDEN / "ref_1d" / "factor_costs.csv", delimiter=",").reshape(-1)
    assert rel(fc[:, 0], g) < 1e-10
    # data/1d/costmap.csv: cost_value over mu in [18,25), precision in [0.05,1) (gvibase/GVI-GH.h:385-412)
    cm = np.loadtxt(GOLDEN / "ref_1d" / "costmap.csv", delimiter=",")
    nm = 40
    for i in (0, 7, 19, 39):
        for j in (0, 11, 26, 39):
            c, _ = p.cost(np.array([18 + i * 7.0 / nm]), np.array([[[0.05 + j * 0.95 / nm]]]), None)
            assert abs(c - cm[j, i]) < 1e-9 * max(1.0, abs(cm[j, i]))


def run_pair(gpu_ctx, spec, niters, reuse=0, **okw):
    p = problems.build_device_problem(gpu_ctx, spec)
    opts = capi.Problem.default_opts()
    opts.step_size_base = spec.meta.get("step_size_base", 0.55)
    opts.niters_lowtemp = spec.meta.get("niters_lowtemp", 10)
    opts.reuse_accepted_sweep = reuse
    stats = p.optimize(niters, opts)
    ref = ob.build_oracle(spec, niters=niters, **okw)
    recs = ref.optimize()
    return p, stats, ref, recs


def check_final(p, stats, ref, recs):
    assert len(stats) == len(recs)
    for s, r in zip(stats, recs):
        assert bool(s.accepted) == r.accepted and s.n_backtrack == r.n_backtrack
        assert abs(s.cost - r.cost) < 1e-8 * max(1.0, abs(r.cost))
    cD, cO = p.covariance()
    e_mu = rel(p.mean(), ref.mean())
    e_cov = rel(np.concatenate([cD.reshape(-1), cO.reshape(-1)]),
                np.concatenate([ref.cov.D.reshape(-1), ref.cov.O.reshape(-1)]))
    return e_mu, e_cov


@pytest.mark.parametrize("reuse", [0, 1])
def test_cfg2_small_all_linear(gpu_ctx, reuse):
    spec = problems.make_cfg2(S=60)
    p, stats, ref, recs = run_pair(gpu_ctx, spec, 10, reuse)
    e_mu, e_cov = check_final(p, stats, ref, recs)
    print("cfg2 S=60: rel err mu", e_mu, "cov", e_cov, "kappa(Vddmu)", np.linalg.cond(ref.Vddmu.dense()))
    assert e_mu < FINAL_TOL and e_cov < FINAL_TOL


@pytest.mark.parametrize("reuse", [0, 1])
def test_cfg3_small_hinge_ltv(gpu_ctx, reuse):
    spec = problems.make_cfg3(N=80)
    p, stats, ref, recs = run_pair(gpu_ctx, spec, 10, reuse)
    e_mu, e_cov = check_final(p, stats, ref, recs)
    print("cfg3 N=80: rel err mu", e_mu, "cov", e_cov, "kappa(Vddmu)", np.linalg.cond(ref.Vddmu.dense()))
    assert e_mu < FINAL_TOL and e_cov < FINAL_TOL


def test_cfg3_gradients_match_oracle(gpu_ctx):
    spec = problems.make_cfg3(N=200)
    p = problems.build_device_problem(gpu_ctx, spec)
    ref = ob.build_oracle(spec)
    dmu, dD, dO = p.gradients()
    Vd, VD, VO = p.get_V()
    rdmu, rdp = ref.compute_gradients()
    assert rel(Vd, ref.Vdmu) < 1e-9
    assert rel(VD, ref.Vddmu.D) < 1e-9 and rel(VO, ref.Vddmu.O) < 1e-9
    assert rel(dmu, rdmu) < 1e-8
    assert rel(dD, rdp.D) < 1e-9
    c, fc = p.cost()
    assert rel(fc, ref.factor_cost_vector()) < 1e-10
    assert abs(c - ref.cost_value()) < 1e-9 * abs(c)


# ---------------------------------------------------------------- the two moment kernels agree (K1S vs generic K1)
@pytest.mark.parametrize("dim,deg,kind", [(4, 6, "hinge"), (4, 4, "quad"), (3, 8, "quad"), (2, 10, "quad"), (1, 10, "stereo")])
def test_sign_group_kernel_matches_generic_kernel(gpu_ctx, dim, deg, kind):
    rng = np.random.default_rng(dim * 100 + deg)
    if kind == "hinge":
        spec = problems.make_factor_batch(N=999)
    else:
        N = 77
        spec = problems.ProblemSpec(S=N, d=dim)
        prm = np.array([2.5]) if kind == "quad" else capi.Stereo1DParams(20.0, 400.0, 0.1, 0.09, 9.0, -0.8)
        k = capi.COST_QUADRATIC if kind == "quad" else capi.COST_STEREO_1D
        spec.groups.append(problems.GhGroupSpec(k, dim, deg, np.arange(N, dtype=np.int32), prm, 1.0, 10.0))
        spec.mu0 = (rng.standard_normal((N, dim)) + (20.0 if kind == "stereo" else 0.0)).reshape(-1)
        A = rng.standard_normal((N, dim, dim))
        spec.prec0_D = A @ np.transpose(A, (0, 2, 1)) + 0.5 * np.eye(dim)
        spec.prec0_O = np.zeros((N - 1, dim, dim))
    p = problems.build_device_problem(gpu_ctx, spec)
    (a0, a1, a2), = p.moments()
    ca, fa = p.cost()
    p.set_option("generic_k1", 1)
    (b0, b1, b2), = p.moments()
    cb, fb = p.cost()
    assert (a0 == 0).sum() == (b0 == 0).sum()
    for f in range(len(a0)):
        assert rel(a0[f], b0[f]) < 1e-11 and rel(a2[f], b2[f]) < 1e-11
        assert np.abs(a1[f] - b1[f]).max() < 1e-11 * max(np.abs(b1[f]).max(), np.sqrt(np.abs(b2[f]).max() * abs(b0[f])), 1e-300)
    assert rel(fa, fb) < 1e-11


# ---------------------------------------------------------------- Prox-GVI (a15)
def test_prox_1d_golden_trace(gpu_ctx):
    """src/1d_example_proxGVI.cpp against the reference's committed outputs data/1d_proxgvi/*.csv (10 iterations)."""
    spec = problems.make_cfg1()
    p = problems.build_device_problem(gpu_ctx, spec, prox=True)
    opts = capi.Problem.default_opts()
    opts.step_size_base = 0.75
    opts.niters_lowtemp = 10
    g = lambda n: np.loadtxt(GOLDEN / "ref_1d_proxgvi" / f"{n}.csv", delimiter=",").reshape(-1)
    means, covs, precs, costs = [], [], [], []
    for it in range(10):
        means.append(p.mean()[0])
        covs.append(p.covariance()[0][0, 0, 0])
        precs.append(p.precision()[0][0, 0, 0])
        st = p.prox_iterate(opts)
        costs.append(st.cost)
    assert rel(means, g("mean")) < 1e-10
    assert rel(covs, g("cov")) < 1e-10
    assert rel(precs, g("precision")) < 1e-10
    assert rel(costs, g("cost")) < 1e-10


@pytest.mark.parametrize("closed_form", [False, True])
def test_prox_cfg4_small_matches_oracle(gpu_ctx, closed_form):
    """cfg4 generator at a size the oracle reaches: dim-12 two-state factors, sparse GH degree 4 (2649 nodes)."""
    spec = problems.make_cfg4(S=7, closed_form=closed_form)
    p = problems.build_device_problem(gpu_ctx, spec, prox=True)
    opts = capi.Problem.default_opts()
    opts.step_size_base = spec.meta["step_size_base"]
    opts.niters_lowtemp = spec.meta["niters_lowtemp"]
    ref = ob.build_oracle_prox(spec, niters=4)
    recs = ref.optimize()
    stats = [p.prox_iterate(opts) for _ in range(4)]
    for s, r in zip(stats, recs):
        assert s.n_backtrack == r.n_backtrack and bool(s.accepted) == r.accepted
        assert abs(s.cost - r.cost) < 1e-8 * max(1.0, abs(r.cost))
    cD, cO = p.covariance()
    e_mu = rel(p.mean(), ref.mean())
    e_cov = rel(np.concatenate([cD.reshape(-1), cO.reshape(-1)]), np.concatenate([ref.cov.D.reshape(-1), ref.cov.O.reshape(-1)]))
    print("prox cfg4 S=7 closed_form", closed_form, "rel err mu", e_mu, "cov", e_cov)
    assert e_mu < FINAL_TOL and e_cov < FINAL_TOL


def test_prox_quadrature_equals_closed_form(gpu_ctx):
    """cost_linear_gp is exactly integrable: the GH factors and ProxFactorizedLinear give the same iteration."""
    opts = capi.Problem.default_opts()
    opts.step_size_base = 0.1
    opts.niters_lowtemp = 1 << 30
    out = []
    for closed_form in (False, True):
        spec = problems.make_cfg4(S=12, closed_form=closed_form)
        p = problems.build_device_problem(gpu_ctx, spec, prox=True)
        st = [p.prox_iterate(opts) for _ in range(3)]
        out.append((p.mean(), [s.cost for s in st]))
    assert rel(out[0][0], out[1][0]) < 1e-9
    assert rel(out[0][1], out[1][1]) < 1e-9


# ---------------------------------------------------------------- cfg5: independent problems batched as one chain
def test_cfg5_batch_matches_independent_oracle_runs(gpu_ctx):
    nb, N, niters = 5, 40, 6
    spec = problems.make_cfg5(n_problems=nb, N=N)
    p = problems.build_device_problem(gpu_ctx, spec)
    opts = capi.Problem.default_opts()
    stats = p.optimize(niters, opts)
    assert all(s.accepted and s.n_backtrack == 0 for s in stats)
    mu = p.mean().reshape(nb, -1)
    cD, _ = p.covariance()
    Sb = spec.meta["states_per_problem"]
    total_cost = 0.0
    for b in range(nb):
        sub = problems.make_cfg3(N=N, seed=1000 + b)
        ref = ob.build_oracle(sub, niters=niters)
        recs = ref.optimize()
        assert all(r.accepted and r.n_backtrack == 0 for r in recs)
        total_cost += recs[-1].cost
        assert rel(mu[b], ref.mean()) < FINAL_TOL
        assert rel(cD[b * Sb:(b + 1) * Sb], ref.cov.D) < FINAL_TOL
    assert abs(stats[-1].cost - total_cost) < 1e-8 * abs(total_cost)


# ---------------------------------------------------------------- very long chains: three-level engine
def test_three_level_chain_engine_long_chain(gpu_ctx):
    """S = 700k states (beyond the two-level plan): selected inverse vs the C oracle's inverse_GBP, solve by residual."""
    import ctypes as C
    import gvi_oracle_c as oc
    rng = np.random.default_rng(9)
    S, d = 700_000, 4
    A = rng.standard_normal((S, d, d))
    D = 0.1 * A @ np.transpose(A, (0, 2, 1)) + 3.0 * np.eye(d)
    O = 0.3 * rng.standard_normal((S - 1, d, d))
    rhs = rng.standard_normal(S * d)
    cD, cO, ld = gpu_ctx.selected_inverse(D, O)
    x, _ = gpu_ctx.blocktri_solve(D, O, rhs)
    # residual of the solve: block-tridiagonal matvec
    xs = x.reshape(S, d)
    r = np.einsum("sij,sj->si", D, xs)
    r[:-1] += np.einsum("sij,sj->si", O, xs[1:])
    r[1:] += np.einsum("sji,sj->si", O, xs[:-1])
    assert np.abs(r.reshape(-1) - rhs).max() < 1e-9 * np.abs(rhs).max()
    # selected inverse against the C restatement of inverse_GBP
    Dc = np.ascontiguousarray(np.transpose(D, (0, 2, 1)))
    Oc = np.ascontiguousarray(np.transpose(O, (0, 2, 1)))
    rD, rO = np.zeros_like(Dc), np.zeros((S, d, d))
    dp = C.POINTER(C.c_double)
    assert oc.lib().orc_inverse_gbp(S, d, Dc.ctypes.data_as(dp), Oc.ctypes.data_as(dp), rD.ctypes.data_as(dp), rO.ctypes.data_as(dp)) == 0
    assert rel(cD, np.transpose(rD, (0, 2, 1))) < 1e-11
    assert rel(cO, np.transpose(rO[:S - 1], (0, 2, 1))) < 1e-11


def test_free_space_culling_is_bit_identical(gpu_ctx):
    """Option "cull" (default on): factors whose sigma-point box lies provably in free space are not evaluated.  Their
    moments are exactly zero either way, so five iterations with and without culling agree bit for bit -- and a good
    part of the cfg3 factors is in fact culled."""
    N = 3000
    spec = problems.make_cfg3(N=N)
    opts = capi.Problem.default_opts()
    opts.reuse_accepted_sweep = 1
    out = []
    for cull in (1, 0):
        p = problems.build_device_problem(gpu_ctx, spec)
        p.set_option("cull", cull)
        p.evaluated_factors(reset=True)
        stats = p.optimize(5, opts)
        n_eval = p.evaluated_factors()
        covD, covO = p.covariance()
        out.append((p.mean(), covD, covO, [s.cost for s in stats], n_eval))
    (m1, d1, o1, c1, e1), (m0, d0, o0, c0, e0) = out
    assert np.array_equal(m1, m0) and np.array_equal(d1, d0) and np.array_equal(o1, o0) and c1 == c0
    sweeps = e0 // N
    assert e0 == sweeps * N and sweeps >= 5          # without culling every sweep evaluates every factor
    print("culling: evaluated", e1, "of", e0, "factor sweeps")
    assert e1 < 0.8 * e0


@pytest.mark.parametrize("eps_radius,expect", [((0.001, 0.001), "all_culled"), ((40.0, 60.0), "none_culled")])
def test_free_space_culling_extremes(gpu_ctx, eps_radius, expect):
    """Threshold eps + r tiny: nearly every factor lies in free space (almost nothing is evaluated);
    threshold larger than the field: nothing can be culled.  Both bit-identical to the run without culling; a second
    hinge group with its own threshold rides along (two compacted lists)."""
    N = 600
    spec = problems.make_cfg3(N=N)
    g = spec.groups[2]
    g.params = capi.HingeParams(0.1 if expect == "all_culled" else 1e-4, eps_radius[0], eps_radius[1])
    half = N // 2
    start = np.asarray(g.start)
    spec.groups[2] = problems.GhGroupSpec(g.kind, g.dim, g.deg, start[:half], g.params, 1.0, 10.0)
    spec.groups.append(problems.GhGroupSpec(g.kind, g.dim, g.deg, start[half:], g.params, 1.0, 10.0))
    opts = capi.Problem.default_opts()
    opts.reuse_accepted_sweep = 1
    res = []
    for cull in (1, 0):
        p = problems.build_device_problem(gpu_ctx, spec)
        p.set_option("cull", cull)
        p.evaluated_factors(reset=True)
        stats = p.optimize(3, opts)
        n_eval = p.evaluated_factors()
        covD, covO = p.covariance()
        res.append((p.mean(), covD, covO, [s.cost for s in stats], n_eval))
    (m1, d1, o1, c1, e1), (m0, d0, o0, c0, e0) = res
    assert np.array_equal(m1, m0) and np.array_equal(d1, d0) and np.array_equal(o1, o0) and c1 == c0
    assert e0 > 0 and e0 % N == 0
    if expect == "all_culled":   # (a few factors near a disc stay: the bound is conservative by up to a block of cells)
        assert e1 < 0.3 * e0
    else:
        assert e1 == e0


# ---------------------------------------------------------------- device-side LTV prior set-up
def test_device_ltv_transition_matches_numpy(gpu_ctx):
    """gvib200_ltv_transition (one thread per link, Van Loan per quarter interval) against the NumPy set-up that stands
    for gp/LTV_prior.h:123-197 in the generators: Phi, Q, Q^-1 of 3000 random damped-oscillator links."""
    n = 3000
    hA, hB = problems.ltv_system(n, seed=17)
    idx = 4 * np.arange(n)[:, None] + np.arange(4)[None, :]
    Phi, Q, Kinv = problems.ltv_links(hA[idx], hB[idx], 0.2, ctx=None)
    dPhi, dQ, dKinv = problems.ltv_links(hA[idx], hB[idx], 0.2, ctx=gpu_ctx)
    assert rel(dPhi, Phi) < 1e-13 and rel(dQ, Q) < 1e-13
    assert max(rel(dKinv[k], Kinv[k]) for k in range(n)) < 1e-10   # kappa(Q) ~ 1e4
    # a generic (non-commuting, full B) system of state dimension 6 against scipy's expm of the whole Van Loan matrix
    from scipy.linalg import expm
    rng = np.random.default_rng(2)
    A = rng.standard_normal((5, 4, 6, 6))
    B = rng.standard_normal((5, 4, 6, 3))
    gPhi, gQ = gpu_ctx.ltv_transition(A, B, 0.4, want_inverse=False)
    for f in range(5):
        P, Qm = np.eye(6), np.zeros((6, 6))
        for k in range(4):
            M = np.zeros((12, 12))
            M[:6, :6], M[:6, 6:], M[6:, 6:] = -A[f, k], B[f, k] @ B[f, k].T, A[f, k].T
            E = expm(M * 0.1)
            Pk = E[6:, 6:].T
            Qm = Pk @ Qm @ Pk.T + Pk @ E[:6, 6:]
            P = Pk @ P
        assert rel(gPhi[f], P) < 1e-12 and rel(gQ[f], 0.5 * (Qm + Qm.T)) < 1e-12


def test_cfg5_device_generated_batch_matches_independent_oracle_runs(gpu_ctx):
    """The batch bench.py --config cfg5 builds (LTV links integrated on the device, no per-problem loop): every problem
    is bit-identical to make_cfg3(seed = 1000 + b, ctx) and iterates like an independent oracle run of it."""
    nb, N, niters = 4, 50, 5
    spec = problems.make_cfg5(n_problems=nb, N=N, first_seed=1003, ctx=gpu_ctx)
    Sb = spec.meta["states_per_problem"]
    p = problems.build_device_problem(gpu_ctx, spec)
    stats = p.optimize(niters, capi.Problem.default_opts())
    assert all(s.accepted and s.n_backtrack == 0 for s in stats)
    mu = p.mean().reshape(nb, -1)
    cD, _ = p.covariance()
    for b in range(nb):
        sub = problems.make_cfg3(N=N, seed=1003 + b, ctx=gpu_ctx)
        ltv_b = spec.groups[1]
        assert np.array_equal(ltv_b.Kinv[b * (Sb - 1):(b + 1) * (Sb - 1)], sub.groups[1].Kinv)
        assert np.array_equal(ltv_b.Lambda[b * (Sb - 1):(b + 1) * (Sb - 1)], sub.groups[1].Lambda)
        ref = ob.build_oracle(sub, niters=niters)
        ref.optimize()
        assert rel(mu[b], ref.mean()) < FINAL_TOL
        assert rel(cD[b * Sb:(b + 1) * Sb], ref.cov.D) < FINAL_TOL


# ---------------------------------------------------------------- K1G (sparse sign groups, dim > 4) vs the generic node loop
@pytest.mark.parametrize("case", ["linear_gp_12", "hinge3d_6", "arm_6", "quad_hinge_6", "linear_gp_8"])
def test_sparse_group_kernel_matches_generic_kernel(gpu_ctx, case):
    """k_moments_grp (k1_grp.cuh: reduced coordinates, Walsh sums per sign group, lane-private accumulators) against
    k_moments (node loop) on the same factors: moments, factor costs and the Vdmu / Vddmu epilogue to 1e-11."""
    rng = np.random.default_rng(len(case))
    if case.startswith("linear_gp"):
        dim = int(case.split("_")[2])
        spec = problems.make_cfg4(S=40) if dim == 12 else None
        if spec is None:   # dim 8: two 4-dimensional states under a min-acc prior evaluated by quadrature
            S, d = 30, 4
            lin = problems.minacc_group(S, 0.8 * np.eye(2), 0.1)
            Phi = -lin.Lambda[0][:, :d]
            rec = np.concatenate([np.tile(Phi.T.reshape(-1), (S - 1, 1)), np.tile(lin.Kinv[0].T.reshape(-1), (S - 1, 1))], axis=1)
            D, O = rand_spd_chain(rng, S, d)
            spec = problems.ProblemSpec(S=S, d=d, groups=[problems.fixed_prior_group([0, S - 1], np.zeros((2, d)), np.eye(d), d),
                                                          problems.GhGroupSpec(capi.COST_LINEAR_GP, 2 * d, 4, lin.start, rec)],
                                        mu0=rng.standard_normal(S * d), prec0_D=D, prec0_O=O)
        else:
            D, O = rand_spd_chain(rng, spec.S, spec.d)   # a generic (correlated) state instead of the diagonal start
            spec.prec0_D, spec.prec0_O = D, O
            spec.mu0 = spec.mu0 + 0.3 * rng.standard_normal(spec.mu0.shape)
    else:
        kind = {"hinge3d_6": capi.COST_HINGE_3D, "arm_6": capi.COST_ARM_3D, "quad_hinge_6": capi.COST_QUAD_HINGE}[case]
        spec = problems.make_factor_batch_functor(kind, N=200, d=6, deg=4)
    p = problems.build_device_problem(gpu_ctx, spec)
    with_prior = case.startswith("linear_gp")   # a factor batch without a prior has no SPD Vddmu to solve with
    mom_a = p.moments()
    ca, fa = p.cost()
    va = (p.gradients(), p.get_V())[1] if with_prior else ()
    p.set_option("generic_k1", 1)
    mom_b = p.moments()
    cb, fb = p.cost()
    vb = (p.gradients(), p.get_V())[1] if with_prior else ()
    for (a0, a1, a2), (b0, b1, b2) in zip(mom_a, mom_b):
        assert np.array_equal(a0 == 0, b0 == 0)
        for f in range(len(a0)):
            if b0[f] == 0:
                continue
            assert rel(a0[f], b0[f]) < 1e-11 and rel(a2[f], b2[f]) < 1e-11
            assert np.abs(a1[f] - b1[f]).max() < 1e-11 * max(np.abs(b1[f]).max(), np.sqrt(np.abs(b2[f]).max() * abs(b0[f])), 1e-300)
    assert rel(fa, fb) < 1e-11 and abs(ca - cb) < 1e-11 * abs(cb)
    for x, y in zip(va, vb):
        assert rel(x, y) < 1e-10
