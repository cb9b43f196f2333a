"""the robot cost functors beyond the planar point robot -- CudaOperation_3dpR over the 3-D
SignedDistanceField (helpers/CudaOperation.h:133-236,641-674) and CudaOperation_Quad (:565-605).

CPU: the oracle restatements against hand-computed values (trilinear interpolation reproduces an affine field exactly,
clamping outside the field, the quadrotor's five check points).  GPU: per-factor quadrature moments of both functors
against the oracle to 1e-10 (tensor-wise max norm) on a random factor batch, plus a short optimization."""
import numpy as np
import pytest

import oracle_bridge as ob
from gaussianvi_b200 import capi, problems

o = ob.o
MOMENT_TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a).max())


def test_oracle_trilinear_is_exact_on_affine_fields():
    nz, rows, cols = 5, 6, 7
    origin, cell = np.array([-1.0, 2.0, 0.5]), 0.25
    z, y, x = np.meshgrid(origin[2] + cell * np.arange(nz), origin[1] + cell * np.arange(rows),
                          origin[0] + cell * np.arange(cols), indexing="ij")
    f = lambda X, Y, Z: 0.3 * X - 1.1 * Y + 0.7 * Z + 2.0
    sdf = o.SignedDistanceField3D(origin, cell, f(x, y, z))
    rng = np.random.default_rng(0)
    ext = np.array([(cols - 1), (rows - 1), (nz - 1)]) * cell
    pts = origin + rng.uniform(0, 1, (200, 3)) * ext
    np.testing.assert_allclose(sdf.signed_distance(pts), f(pts[:, 0], pts[:, 1], pts[:, 2]), rtol=0, atol=1e-13)
    # outside the field the point is clamped to it (convertPoint3toCell, helpers/CudaOperation.h:176-205)
    out = np.array([[origin[0] - 3.0, origin[1] + 0.1, origin[2] + 10.0]])
    cl = np.array([[origin[0], origin[1] + 0.1, origin[2] + ext[2]]])
    np.testing.assert_allclose(sdf.signed_distance(out), f(cl[:, 0], cl[:, 1], cl[:, 2]), rtol=0, atol=1e-13)
    # grid nodes are reproduced bit for bit, including the last one
    corner = np.array([[origin[0] + ext[0], origin[1] + ext[1], origin[2] + ext[2]]])
    assert sdf.signed_distance(corner)[0] == sdf.data[-1, -1, -1]


def test_oracle_quadrotor_check_points():
    """vec_balls (helpers/CudaOperation.h:588-604) for phi = 0: five points from x - (L - 1.5 r)/2 in steps of L/5."""
    data = np.zeros((50, 200))
    origin, cell = (-10.0, -2.5), 0.1
    xs = origin[0] + cell * np.arange(200)
    data[:, :] = 3.0 - np.abs(xs)[None, :]          # sd depends on x only: 3 - |x|
    sdf = o.PlanarSDF(np.asarray(origin), cell, data)
    psi = o.make_quad_hinge_cost(sdf, sigma=2.0, eps=0.5, radius=1.0)
    X = np.array([[0.0, 0.0, 0.0, 0, 0, 0]])
    pts = 0.0 - (5.0 - 1.5) / 2.0 + 5.0 / 5 * np.arange(5)      # -1.75 -0.75 0.25 1.25 2.25
    want = 2.0 * np.sum((5.0 * np.maximum(0.0, 1.5 - (3.0 - np.abs(pts)))) ** 2)
    assert abs(psi(X)[0] - want) < 1e-12 * max(want, 1)
    # rotating by pi/2 moves the points along z where this field is constant: all five see sd(x = 0) = 3 -> no cost
    assert psi(np.array([[0.0, 0.0, np.pi / 2, 0, 0, 0]]))[0] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [capi.COST_HINGE_3D, capi.COST_QUAD_HINGE])
def test_moments_robot_functors(kind):
    import gaussianvi_b200 as gv
    ctx = gv.Context(0)
    N, d, deg = 96, 6, 3
    spec = problems.make_factor_batch_functor(kind, N=N, d=d, deg=deg)
    p = problems.build_device_problem(ctx, spec)
    (E0, E1, E2), = p.moments()
    covD, _ = p.covariance()
    psi = ob.psi_for_group(spec, spec.groups[0], 0)
    Z, w = o.table(d, deg)
    mu = spec.mu0.reshape(N, d)
    worst, nz = [0.0, 0.0, 0.0], 0
    for k in range(N):
        r0, r1, r2 = o.moments_fast(psi, mu[k], covD[k], Z, w)
        if r0 == 0.0:
            assert E0[k] == 0.0 and not E1[k].any() and not E2[k].any()
            continue
        nz += 1
        worst = [max(worst[0], rel(E0[k], r0)), max(worst[1], rel(E1[k], r1)), max(worst[2], rel(E2[k], r2))]
    print("functor", kind, "non-zero factors", nz, "worst rel err", worst)
    assert nz >= N // 8
    assert max(worst) < MOMENT_TOL


@pytest.mark.gpu
def test_hinge3d_chain_matches_oracle():
    """A short 3-D point-robot chain (state 6 = position + velocity): min-acc prior + fixed ends + 3-D hinge factors,
    five NGD iterations on the device against the oracle (final mean / covariance to 1e-7)."""
    import gaussianvi_b200 as gv
    ctx = gv.Context(0)
    S, d, dt = 12, 6, 0.3
    spec = problems.ProblemSpec(S=S, d=d)
    spec.sdf3d = problems.ball_sdf3d()
    data, origin, cell = spec.sdf3d
    start = np.array([origin[0] + 0.5, origin[1] + 0.5, origin[2] + 0.5, 0, 0, 0])
    goal = np.array([origin[0] + 7.0, origin[1] + 5.0, origin[2] + 3.0, 0, 0, 0])
    spec.groups.append(problems.fixed_prior_group([0, S - 1], np.stack([start, goal]), 1e-4 * np.eye(d), d))
    spec.groups.append(problems.minacc_group(S, 0.8 * np.eye(3), dt))
    spec.groups.append(problems.GhGroupSpec(capi.COST_HINGE_3D, d, 3, np.arange(1, S - 1, dtype=np.int32),
                                            capi.HingeParams(0.3, 0.5, 1.0), 1.0, 10.0))
    t = np.linspace(0, 1, S)[:, None]
    mu0 = start[None, :] * (1 - t) + goal[None, :] * t
    mu0[:, 3:] = (goal[:3] - start[:3]) / ((S - 1) * dt)
    spec.mu0 = mu0.reshape(-1)
    spec.prec0_D = np.broadcast_to(10.0 * np.eye(d), (S, d, d)).copy()
    spec.prec0_O = np.zeros((S - 1, d, d))
    spec.meta = dict(niters=5, step_size_base=0.55, niters_lowtemp=10)
    p = problems.build_device_problem(ctx, spec)
    stats = p.optimize(5, gv.Problem.default_opts())
    ref = ob.build_oracle(spec, niters=5)
    ref.optimize()
    assert rel(p.mean(), ref.mean()) < 1e-7
    covD, covO = p.covariance()
    assert rel(covD, ref.covariance().D) < 1e-7
    assert all(s.accepted for s in stats)
