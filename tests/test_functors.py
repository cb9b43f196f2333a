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


def test_oracle_arm_forward_kinematics():
    """ForwardKinematics (helpers/CudaOperation.h:366-401) on cases with a closed form: a planar two-link arm (alpha = 0,
    d = 0) puts the sphere at the classic (l1 c1 + l2 c12, l1 s1 + l2 s12, 0); n_balls is capped by the size of the state
    vector (:752); the single-precision trigonometry is visible at the 1e-8 level and nowhere above."""
    nz, rows, cols = 30, 40, 50
    origin, cell = np.array([-2.5, -2.0, -1.5]), 0.1
    z, y, x = np.meshgrid(origin[2] + cell * np.arange(nz), origin[1] + cell * np.arange(rows), origin[0] + cell * np.arange(cols),
                          indexing="ij")
    sdf = o.SignedDistanceField3D(origin, cell, 0.4 * x - 0.3 * y + 0.2 * z + 1.0)   # affine: the lookup is exact
    l1, l2, th1, th2 = 1.0, 0.7, 0.3, -0.5
    psi = o.make_arm_cost(sdf, [l1, l2], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], frames=[1], centers=[[0.0, 0.0, 0.0]], radii=[3.0],
                          sigma=2.0, eps=0.5)
    px, py = l1 * np.cos(th1) + l2 * np.cos(th1 + th2), l1 * np.sin(th1) + l2 * np.sin(th1 + th2)
    want = 2.0 * (3.5 - (0.4 * px - 0.3 * py + 1.0)) ** 2
    got = psi(np.array([[th1, th2, 9.0, 9.0]]))[0]
    assert abs(got - want) < 5e-7 * want and abs(got - want) > 0        # float trig: close, not equal
    # three spheres but a 2-dimensional state vector: only the first two are evaluated
    three = o.make_arm_cost(sdf, [l1], [0.0], [0.0], [0.0], frames=[0, 0, 0], centers=[[0, 0, 0], [-0.5, 0, 0], [-1.0, 0, 0]],
                            radii=[3.0, 3.0, 3.0], sigma=1.0, eps=0.5)
    two = o.make_arm_cost(sdf, [l1], [0.0], [0.0], [0.0], frames=[0, 0], centers=[[0, 0, 0], [-0.5, 0, 0]], radii=[3.0, 3.0],
                          sigma=1.0, eps=0.5)
    X = np.array([[0.4, 0.1]])
    assert three(X)[0] == two(X)[0] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind,d,deg", [(capi.COST_HINGE_3D, 6, 3), (capi.COST_QUAD_HINGE, 6, 3), (capi.COST_ARM_3D, 6, 3),
                                        (capi.COST_ARM_3D, 4, 4)])
def test_moments_robot_functors(kind, d, deg):
    import gaussianvi_b200 as gv
    ctx = gv.Context(0)
    N = 96
    spec = problems.make_factor_batch_functor(kind, N=N, d=d, deg=deg)
    p = problems.build_device_problem(ctx, spec)
    (E0, E1, E2), = p.moments()
    covD = spec.meta["Sigma"]   # the generator's covariances: nothing read back from the GPU feeds the oracle
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


@pytest.mark.gpu
def test_arm_chain_matches_oracle():
    """A 3-DOF arm (state 6 = joint angles + velocities) under a min-acc prior with CudaOperation_3dArm collision factors
    (helpers/CudaOperation.h:680-779): five NGD iterations on the device against the oracle."""
    import gaussianvi_b200 as gv
    ctx = gv.Context(0)
    S, d, dt = 10, 6, 0.3
    spec = problems.ProblemSpec(S=S, d=d)
    spec.sdf3d = problems.ball_sdf3d()
    start = np.array([-1.2, 0.4, 0.3, 0, 0, 0.0])
    goal = np.array([1.0, -0.6, 1.1, 0, 0, 0.0])
    spec.groups.append(problems.fixed_prior_group([0, S - 1], np.stack([start, goal]), 1e-4 * np.eye(d), d))
    spec.groups.append(problems.minacc_group(S, 0.8 * np.eye(3), dt))
    spec.groups.append(problems.GhGroupSpec(capi.COST_ARM_3D, d, 3, np.arange(1, S - 1, dtype=np.int32),
                                            problems.example_arm(3, sigma=0.2), 1.0, 10.0))
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
    recs = ref.optimize()
    assert [bool(s.accepted) for s in stats] == [r.accepted for r in recs] and [s.n_backtrack for s in stats] == [r.n_backtrack for r in recs]
    (E0, _, _), = p.moments()
    assert (E0 > 0).sum() >= 2            # the arm does touch the obstacles
    assert rel(p.mean(), ref.mean()) < 1e-7
    covD, covO = p.covariance()
    assert rel(covD, ref.covariance().D) < 1e-7
