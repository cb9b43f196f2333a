"""CPU tests: the C restatement (oracle/gvi_oracle_c.c, the CPU-baseline arm of bench.py) against the NumPy oracle and
the reference's golden 1-D trace, in both schedules ("reference" = what GVIGH::optimize executes, "lean")."""
import numpy as np
import pytest

import gvi_oracle as o
import gvi_oracle_c as oc
import oracle_bridge as ob
from gaussianvi_b200 import problems

GOLDEN = ob.ROOT / "tests" / "golden"


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


@pytest.mark.parametrize("schedule", [0, 1])
def test_c_oracle_cfg1_golden_trace(schedule):
    """src/1d_example.cpp against data/1d/*.csv through the C restatement."""
    spec = problems.make_cfg1()
    c = oc.COracle(spec, o.table)
    g = lambda n: np.loadtxt(GOLDEN / "ref_1d" / f"{n}.csv", delimiter=",").reshape(-1)
    means, precs, costs = [], [], []
    for it in range(10):
        means.append(c.mu[0])
        precs.append(c.LD[0, 0, 0])
        st = c.iterate(step_size_base=0.75, schedule=schedule)
        costs.append(st.cost)
        assert st.accepted == 1 and st.n_backtrack == 0 and st.status == 0
        if schedule == 0:
            assert st.n_psi_sweeps == 6 and st.n_inversions == 4   # (5 + T_ls) sweeps, (3 + T_ls) inversions
        elif it > 0:
            assert st.n_psi_sweeps == 2 and st.n_inversions == 1
    assert rel(means, g("mean")) < 1e-11
    assert rel(precs, g("precision")) < 1e-11
    assert rel(costs, g("cost")) < 1e-11


def test_c_oracle_moments_match_numpy():
    spec = problems.make_factor_batch(N=300)
    c = oc.COracle(spec, o.table)
    for faithful in (False, True):
        E0, E1, E2 = c.moments(faithful=faithful)
        cD, _ = c.covariance_blocks()
        psi = ob.psi_for_group(spec, spec.groups[0], 0)
        Z, w = o.table(4, 6)
        mu = spec.mu0.reshape(-1, 4)
        worst = 0.0
        for k in range(0, 300, 7):
            r0, r1, r2 = o.moments(psi, mu[k], cD[k], Z, w)
            if r0 == 0.0:
                assert E0[k] == 0.0
                continue
            worst = max(worst, rel(E0[k], r0), rel(E1[k], r1), rel(E2[k], r2))
        assert worst < 1e-11


@pytest.mark.parametrize("schedule", [0, 1])
def test_c_oracle_cfg3_matches_numpy(schedule):
    spec = problems.make_cfg3(N=40)
    c = oc.COracle(spec, o.table)
    ref = ob.build_oracle(spec, niters=4, faithful_linear=False)
    recs = ref.optimize()
    for it in range(4):
        st = c.iterate(schedule=schedule)
        assert st.accepted == int(recs[it].accepted) and st.n_backtrack == recs[it].n_backtrack
        assert abs(st.cost - recs[it].cost) < 1e-9 * abs(recs[it].cost)
    assert rel(c.mean(), ref.mean()) < 1e-8
    cD, cO = c.cov_blocks()
    assert rel(cD, ref.cov.D) < 1e-8 and rel(cO, ref.cov.O) < 1e-8


def test_c_oracle_cfg2_costs_and_gradients():
    spec = problems.make_cfg2(S=30)
    c = oc.COracle(spec, o.table)
    ref = ob.build_oracle(spec, niters=1)
    cost, fc = c.cost_value()
    assert abs(cost - ref.cost_value()) < 1e-10 * abs(cost)
    assert rel(c.factor_costs_in_spec_order(fc), ref.factor_cost_vector()) < 1e-11
    st = c.iterate(schedule=0)
    recs = ref.optimize()
    assert rel(c.mean(), ref.mean()) < 1e-9
    pD, pO = c.prec_blocks()
    assert rel(pD, ref.prec.D) < 1e-10 and rel(pO, ref.prec.O) < 1e-10
