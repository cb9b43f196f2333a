"""CPU tests: pin the oracle (oracle/gvi_oracle.py) against every golden vector / known-answer test the
reference holds for the NGD-GVI path.  The constants below are copied from the reference's
own test files / committed demo outputs (cited per test); tests/golden/ref_1d* are copies of data/1d*/*.csv."""
import numpy as np
import pytest

import gvi_oracle as o
import oracle_bridge as ob
from gaussianvi_b200 import problems

GOLDEN = ob.ROOT / "tests" / "golden"


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


# ------------------------------------------------------------------ table (nwspgr.m restatement)
def test_table_dim5_k2_golden():
    """tests/test_spgh_table_IO.cpp:68-90: nodes, weights and ROW ORDER of the (dim 5, k 2) rule."""
    Z, w = o.table(5, 2)
    pts = np.zeros((11, 5))
    for j in range(5):
        pts[j, j] = -1.0
        pts[10 - j, j] = 1.0
    wts = np.array([0.5] * 5 + [-4.0] + [0.5] * 5)
    assert Z.shape == (11, 5)
    assert np.abs(Z - pts).max() < 1e-6
    assert np.abs(w - wts).max() < 1e-6


@pytest.mark.parametrize("dim,deg,n", [(1, 10, 10), (2, 10, 381), (3, 8, 1233), (4, 2, 9), (4, 4, 137), (4, 6, 953),
                                       (8, 4, 849), (12, 4, 2649)])
def test_table_sizes_and_normalisation(dim, deg, n):
    Z, w = o.table(dim, deg)
    assert Z.shape == (n, dim) and w.shape == (n,)
    assert abs(w.sum() - 1.0) < 1e-12                       # nwspgr.m:130-133 renormalises
    assert np.abs((w[:, None] * Z).sum(0)).max() < 1e-12    # symmetric rule
    # exact for the second moment of N(0, I) (degree >= 2)
    assert np.abs((Z * w[:, None]).T @ Z - np.eye(dim)).max() < 1e-10


def test_gh_rule_10pt():
    """tests/test_GH.cpp:79-91: 10-point 1-D rule (probabilists' Hermite roots / weights) to 1e-10, as a set."""
    w_exp = np.array([4.310652630718227e-06, 4.310652630718376e-06, 7.580709343122321e-04, 7.580709343121815e-04,
                      0.344642334932012, 0.344642334932016, 0.135483702980275, 0.135483702980267,
                      0.019111580500769, 0.019111580500770])
    x_exp = np.array([4.859462828332310, -4.859462828332314, 3.581823483551924, -3.581823483551934,
                      0.484935707515505, -0.484935707515517, 1.465989094391161, -1.465989094391140,
                      2.484325841638960, -2.484325841638965])
    Z, w = o.table(1, 10)
    order_e = np.argsort(x_exp)
    order = np.argsort(Z[:, 0])
    assert np.linalg.norm(Z[order, 0] - x_exp[order_e]) < 1e-10
    assert np.linalg.norm(w[order] - w_exp[order_e]) < 1e-10


# ------------------------------------------------------------------ quadrature known answers
def _stereo(y_offset):
    def psi(X):
        x = X[:, 0]
        mu_p, f, b, sig_r_sq, sig_p_sq = 20.0, 400.0, 0.1, 0.09, 9.0
        y = f * b / mu_p + y_offset
        return (x - mu_p) ** 2 / sig_p_sq / 2 + (y - f * b / x) ** 2 / sig_r_sq / 2
    return psi


def test_kat_stereo_sparse_deg6():
    """tests/test_GH.cpp:134-161 (sparse, deg 6, mu = 20, Sigma = 9): 1.1129 / -1.2144 to 1e-4."""
    Z, w = o.table(1, 6)
    E0, E1, E2 = o.moments(_stereo(0.05), [20.0], [[9.0]], Z, w)
    assert abs(E0 - 1.1129) < 1e-4
    assert abs(E1[0] + 1.2144) < 1e-4
    # phi(mu) = 0.013888888888889 (tests/test_GH.cpp:106-108)
    assert abs(_stereo(0.05)(np.array([[20.0]]))[0] - 0.013888888888889) < 1e-5


def test_kat_multidim_vector_integrand():
    """tests/test_GH.cpp:164-183: dim 2, deg 10, correlated Sigma, E[(3 x0^2, 2 x0 x1)] = (9.63145, 5.27152) (1e-3)."""
    Z, w = o.table(2, 10)
    cov = np.array([[2.210433244916004, 1.635720601237843], [1.635720601237843, 2.210433244916004]])
    f = lambda X: np.stack([3 * X[:, 0] ** 2, 2 * X[:, 0] * X[:, 1]], axis=1)
    r = o.integrate(f, np.ones(2), cov, Z, w)
    assert np.linalg.norm(r - np.array([9.631450087970276, 5.271519032251217])) < 1e-3


@pytest.mark.parametrize("dim,deg,mean,prec_or_cov,is_prec,c,expected,tol", [
    (4, 3, np.zeros(4), 1e-4 * np.eye(4), False, 1e4, 4.0, 1e-10),                       # test_gh_spgh.cpp:76-90
    (3, 8, np.ones(3), np.eye(3), False, 1e4, 6.00e4, 1e-7),                              # :194-220
    (2, 10, np.ones(2), np.array([[1, -0.74], [-0.74, 1.0]]), True, 1e4, 6.420866489831914e4, 1e-5),  # :92-124
])
def test_kat_quadratic(dim, deg, mean, prec_or_cov, is_prec, c, expected, tol):
    Z, w = o.table(dim, deg)
    cov = np.linalg.inv(prec_or_cov) if is_prec else prec_or_cov
    r = o.integrate(lambda X: c * (X ** 2).sum(1), mean, cov, Z, w)
    assert abs(r[0] - expected) < tol * max(1.0, abs(expected)) if tol < 1e-6 else abs(r[0] - expected) < tol


# ------------------------------------------------------------------ end-to-end traces
@pytest.mark.parametrize("solver", ["direct", "cg"])
def test_cfg1_golden_trace(solver):
    """src/1d_example.cpp:38-85 against data/1d/*.csv: 10 NGD iterations, all printed digits."""
    spec = problems.make_cfg1()
    opt = ob.build_oracle(spec, fast=False, faithful_linear=True, solver=solver, niters=10)
    recs = opt.optimize()
    g = lambda n: np.loadtxt(GOLDEN / "ref_1d" / f"{n}.csv", delimiter=",").reshape(-1)
    assert len(recs) == 10
    assert rel([r.mean[0] for r in recs], g("mean")) < 1e-11
    assert rel([r.cov.D[0, 0, 0] for r in recs], g("cov")) < 1e-11
    assert rel([r.prec.D[0, 0, 0] for r in recs], g("precision")) < 1e-11
    assert rel([r.cost for r in recs], g("cost")) < 1e-11
    assert rel([r.factor_costs[0] for r in recs], g("factor_costs")) < 1e-11
    assert all(r.accepted and r.n_backtrack == 0 for r in recs)


def test_cfg1_costmap():
    """data/1d/costmap.csv: 40x40 sweep of cost_value over mu in [18,25), precision in [0.05,1)
    (gvibase/GVI-GH.h:385-412)."""
    spec = problems.make_cfg1()
    opt = ob.build_oracle(spec, fast=False)
    cm = np.loadtxt(GOLDEN / "ref_1d" / "costmap.csv", delimiter=",")
    nm = cm.shape[0]
    worst = 0.0
    for i in range(0, nm, 3):
        for j in range(0, nm, 3):
            mu = np.array([18 + i * 7.0 / nm])
            prec = o.BlockTri(np.array([[[0.05 + j * 0.95 / nm]]]), np.zeros((0, 1, 1)))
            c = opt.cost_value(mu, prec)
            worst = max(worst, abs(c - cm[j, i]) / max(1.0, abs(cm[j, i])))
    assert worst < 1e-10


# ------------------------------------------------------------------ properties the reference states
@pytest.mark.parametrize("S,d", [(20, 14), (7, 4), (2, 3), (1, 4)])
def test_gbp_equals_dense_inverse(S, d):
    """src/GBP.cpp:133-158 (d = 14, 20 states): GBP marginals == block-tridiagonal part of the dense inverse."""
    rng = np.random.default_rng(S * 100 + d)
    D = np.zeros((S, d, d))
    O = np.zeros((max(S - 1, 0), d, d))
    for i in range(S):
        A = rng.standard_normal((d, d))
        D[i] = A @ A.T + d * np.eye(d)
    for i in range(S - 1):
        O[i] = 0.3 * rng.standard_normal((d, d))
    bt = o.BlockTri(D, O)
    cov = o.inverse_gbp(bt)
    Ai = np.linalg.inv(bt.dense())
    for i in range(S):
        assert np.abs(cov.D[i] - Ai[i * d:(i + 1) * d, i * d:(i + 1) * d]).max() < 1e-12
    for i in range(S - 1):
        assert np.abs(cov.O[i] - Ai[i * d:(i + 1) * d, (i + 1) * d:(i + 2) * d]).max() < 1e-12
    assert abs(o.logdet(bt) - np.linalg.slogdet(bt.dense())[1]) < 1e-10
    rhs = rng.standard_normal(S * d)
    assert rel(o.block_solve(bt, rhs), np.linalg.solve(bt.dense(), rhs)) < 1e-12


def test_linear_closed_form_equals_quadrature():
    """gp/factorized_opts_linear.h:12-14: the ...GH aliases exist 'for comparison'.  GH of a quadratic is exact
    for deg >= 2, so NGDFactorizedLinear (closed form) and NGDFactorizedBaseGH<cost_linear_gp> must agree."""
    dt = 0.1
    lm = o.minimum_acc_gp(0.8 * np.eye(2), dt)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((8, 8))
    Sig = A @ A.T * 0.05 + 0.1 * np.eye(8)
    mu = rng.standard_normal(8)
    Phi = -lm.Lambda[:, :4]
    gh = o.GHFactor(8, 4, 4, o.make_cost_linear_gp(Phi, lm.Kinv), 0)
    cf = o.LinearFactorOpt(8, 4, lm, 0, faithful=True)
    cf2 = o.LinearFactorOpt(8, 4, lm, 0, faithful=False)
    for f in (gh, cf, cf2):
        f.mu = mu.copy()
        f.update_precision_from_joint(Sig)
        f.calculate_partial_V()
    assert rel(gh.Vdmu, cf.Vdmu) < 1e-10 and rel(gh.Vddmu, cf.Vddmu) < 1e-9
    assert rel(cf2.Vddmu, cf.Vddmu) < 1e-11      # 4-th moment loop == 2 C A / T (ngd/NGDFactorizedLinear.h:107-119)
    assert rel(gh.fact_cost_value(mu, Sig), cf.fact_cost_value(mu, Sig)) < 1e-11


def test_cfg2_fixed_point_is_exact_posterior():
    """All-linear chain: each accepted NGD step moves Lambda to (1-a) Lambda + a Lambda*, and the limit is the
    exact Gaussian posterior."""
    spec = problems.make_cfg2(S=12)
    opt = ob.build_oracle(spec, niters=1)
    opt.optimize()
    V = opt.Vddmu
    a = 0.55 * 0.75
    assert rel(opt.prec.D, (1 - a) * spec.prec0_D + a * V.D) < 1e-13
    spec.meta["niters_lowtemp"] = 1000            # no temperature switch: the limit is Vddmu at T = 1
    opt2 = ob.build_oracle(spec, niters=45)
    opt2.optimize()
    assert rel(opt2.prec.D, V.D) < 1e-8 and rel(opt2.prec.O, V.O) < 1e-8
    dmu, _ = opt2.compute_gradients()
    assert np.abs(dmu).max() < 1e-7 * np.abs(opt2.mu).max()


def test_cfg3_small_is_spd_and_accepts_first_trial():
    spec = problems.make_cfg3(N=40)
    opt = ob.build_oracle(spec, niters=4)
    recs = opt.optimize()
    assert all(r.accepted and r.n_backtrack == 0 for r in recs)
    assert np.linalg.eigvalsh(opt.Vddmu.dense()).min() > 0
    assert all(recs[i + 1].cost < recs[i].cost for i in range(len(recs) - 1))


# ------------------------------------------------------------------ Prox-GVI (a15)
def test_prox_1d_golden_trace():
    """src/1d_example_proxGVI.cpp against data/1d_proxgvi/*.csv: 10 Prox-GVI iterations, all printed digits."""
    f = o.ProxGHFactor(1, 1, 10, o.cost_1d_stereo, 0, 1.0, 10.0, fast=False)
    opt = o.ProxGVIGH([f], 1, 1, 10)
    opt.set_step_size_base(0.75)
    opt.set_niter_low_temperature(10)
    opt.set_initial_values(np.array([20.0]), o.BlockTri(np.array([[[1.0 / 9.0]]]), np.zeros((0, 1, 1))))
    recs = opt.optimize()
    g = lambda n: np.loadtxt(GOLDEN / "ref_1d_proxgvi" / f"{n}.csv", delimiter=",").reshape(-1)
    assert rel([r.mean[0] for r in recs], g("mean")) < 1e-11
    assert rel([r.cov.D[0, 0, 0] for r in recs], g("cov")) < 1e-11
    assert rel([r.prec.D[0, 0, 0] for r in recs], g("precision")) < 1e-11
    assert rel([r.cost for r in recs], g("cost")) < 1e-11
    assert rel([r.factor_costs[0] for r in recs], g("factor_costs")) < 1e-11


def test_prox_cfg4_regime_decreases_cost():
    spec = problems.make_cfg4(S=6, closed_form=True)
    recs = ob.build_oracle_prox(spec, niters=5).optimize()
    assert all(r.accepted for r in recs)
    assert all(recs[i + 1].cost < recs[i].cost for i in range(len(recs) - 1))
