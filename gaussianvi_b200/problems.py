"""Synthetic trajectory problems of the shapes BASELINE.json names and the glue that
turns a neutral ProblemSpec into a device problem.  Host-side set-up only (runs once per problem): the
prior models follow gp/fixed_prior.h, gp/minimum_acc_prior.h and gp/LTV_prior.h of the reference; the
numbers produced here are fed unchanged to both the GPU path and (in tests/bench) the CPU oracle.
NumPy only -- this module must not import anything under oracle/."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import functools

import numpy as np

from . import capi


@dataclass
class GhGroupSpec:
    kind: int
    dim: int
    deg: int
    start: np.ndarray                 # [n] int32
    params: object                    # ctypes struct or float64 array [n, rec]
    T: float = 1.0
    T_high: float = 10.0


@dataclass
class LinGroupSpec:
    start: np.ndarray                 # [n]
    Lambda: np.ndarray                # [n, m, dim]
    Psi: np.ndarray                   # [n, m, kdim]
    mu_t: np.ndarray                  # [n, kdim]
    Kinv: np.ndarray                  # [n, m, m]
    C: np.ndarray                     # [n]
    T: float = 1.0
    T_high: float = 10.0


@dataclass
class ProblemSpec:
    S: int
    d: int
    groups: List[object] = field(default_factory=list)   # GhGroupSpec | LinGroupSpec, id order
    sdf: Optional[Tuple[np.ndarray, Tuple[float, float], float]] = None  # (data [rows, cols], origin, cell)
    sdf3d: Optional[Tuple[np.ndarray, Tuple[float, float, float], float]] = None  # (data [nz, rows, cols], origin, cell)
    mu0: Optional[np.ndarray] = None
    prec0_D: Optional[np.ndarray] = None   # [S, d, d]
    prec0_O: Optional[np.ndarray] = None   # [S-1, d, d]
    meta: Dict[str, object] = field(default_factory=dict)

    @property
    def n_gh(self) -> int:
        return sum(len(g.start) for g in self.groups if isinstance(g, GhGroupSpec))

    @property
    def n_factors(self) -> int:
        return sum(len(g.start) for g in self.groups)


def build_device_problem(ctx: "capi.Context", spec: ProblemSpec, set_state: bool = True, prox: bool = False) -> "capi.Problem":
    p = capi.Problem(ctx, spec.S, spec.d)
    if prox:
        p.set_option("prox", 1)
    if spec.sdf is not None:
        data, origin, cell = spec.sdf
        p.set_planar_sdf(data, origin, cell)
    if spec.sdf3d is not None:
        data, origin, cell = spec.sdf3d
        p.set_sdf3d(data, origin, cell)
    for g in spec.groups:
        if isinstance(g, GhGroupSpec):
            p.add_gh_factors(g.kind, g.dim, g.deg, g.start, g.params, g.T, g.T_high)
        else:
            p.add_linear_factors(g.start, g.Lambda, g.Psi, g.mu_t, g.Kinv, g.C, g.T, g.T_high)
    p.finalize()
    if set_state and spec.mu0 is not None:
        p.set_state(spec.mu0, spec.prec0_D, spec.prec0_O)
    return p


# ----------------------------------------------------------------------------------------------
# batched matrix exponential (Taylor + scaling and squaring) for the LTV prior set-up
# ----------------------------------------------------------------------------------------------
def expm_batch(X: np.ndarray) -> np.ndarray:
    """exp of a batch of small matrices [..., n, n]."""
    X = np.asarray(X, dtype=np.float64)
    nrm = np.max(np.sum(np.abs(X), axis=-1), axis=-1)
    s = int(max(0, np.ceil(np.log2(max(float(nrm.max()), 1e-300) / 0.25))))
    Y = X / (2.0 ** s)
    n = X.shape[-1]
    E = np.broadcast_to(np.eye(n), X.shape).copy()
    term = E.copy()
    for k in range(1, 19):
        term = term @ Y / k
        E = E + term
    for _ in range(s):
        E = E @ E
    return E


def ltv_transition_batch(A: np.ndarray, B: np.ndarray, delta_t: float):
    """Phi(dt), Q(dt) of Phi' = A Phi, Q' = A Q + Q A^T + B B^T with A [n, 4, ds, ds], B [n, 4, ds, nb]
    piece-wise constant on the four quarter intervals (gp/LTV_prior.h:123-197; exact per-piece integration by
    Van Loan's block exponential instead of GSL rkf45 at tol 1e-12)."""
    n, _, ds, _ = A.shape
    h = delta_t / 4.0
    Phi = np.broadcast_to(np.eye(ds), (n, ds, ds)).copy()
    Q = np.zeros((n, ds, ds))
    for k in range(4):
        M = np.zeros((n, 2 * ds, 2 * ds))
        M[:, :ds, :ds] = -A[:, k]
        M[:, :ds, ds:] = B[:, k] @ np.transpose(B[:, k], (0, 2, 1))
        M[:, ds:, ds:] = np.transpose(A[:, k], (0, 2, 1))
        E = expm_batch(M * h)
        Phik = np.transpose(E[:, ds:, ds:], (0, 2, 1))
        Qk = Phik @ E[:, :ds, ds:]
        Qk = 0.5 * (Qk + np.transpose(Qk, (0, 2, 1)))
        Q = Phik @ Q @ np.transpose(Phik, (0, 2, 1)) + Qk
        Phi = Phik @ Phi
    return Phi, 0.5 * (Q + np.transpose(Q, (0, 2, 1)))


# ----------------------------------------------------------------------------------------------
# ingredients
# ----------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=4)
def disc_layout(rows=300, cols=400, origin=(-20.0, -10.0), cell=0.1, n_discs=12, seed=7):
    rng = np.random.default_rng(seed)
    cx = rng.uniform(origin[0], origin[0] + (cols - 1) * cell, n_discs)
    cy = rng.uniform(origin[1], origin[1] + (rows - 1) * cell, n_discs)
    rad = rng.uniform(1.0, 3.0, n_discs)
    return cx, cy, rad


@functools.lru_cache(maxsize=4)
def disc_sdf(rows=300, cols=400, origin=(-20.0, -10.0), cell=0.1, n_discs=12, seed=7):
    """Analytic signed distance to `n_discs` discs on a rows x cols grid."""
    cx, cy, rad = disc_layout(rows, cols, origin, cell, n_discs, seed)
    xs = origin[0] + cell * np.arange(cols)
    ys = origin[1] + cell * np.arange(rows)
    X, Y = np.meshgrid(xs, ys)  # [rows, cols]
    sd = np.full((rows, cols), np.inf)
    for k in range(n_discs):
        sd = np.minimum(sd, np.hypot(X - cx[k], Y - cy[k]) - rad[k])
    return sd, (float(origin[0]), float(origin[1])), float(cell)


def lissajous_nominal(S: int, delta_t: float, clearance: Optional[float] = None, discs=None) -> np.ndarray:
    """Nominal 2-D point-robot trajectory [S, 4] = (x, y, vx, vy): constant phase step per state, so every
    problem size sees the same obstacle density.  With `clearance` the path is pushed radially
    out of every disc to distance radius + clearance: a collision-free initial plan that still runs through the
    hinge band.  (A dense Lissajous at N = 100k otherwise passes through disc CENTRES, where the distance field
    has curvature -2 sigma h / rho -> -inf and Vddmu stops being positive definite, see DESIGN.md.)"""
    i = np.arange(S, dtype=np.float64)
    x = 15.0 * np.sin(14.0 * np.pi * i / 1002.0)
    y = 5.0 + 9.0 * np.sin(22.0 * np.pi * i / 1002.0)
    if clearance is not None and discs is not None:
        cx, cy, rad = discs
        for _ in range(4):
            for k in range(len(rad)):
                dx, dy = x - cx[k], y - cy[k]
                rho = np.hypot(dx, dy)
                inside = rho < rad[k] + clearance
                rho_s = np.where(rho < 1e-9, 1.0, rho)
                ux = np.where(rho < 1e-9, 1.0, dx / rho_s)
                uy = np.where(rho < 1e-9, 0.0, dy / rho_s)
                x = np.where(inside, cx[k] + (rad[k] + clearance) * ux, x)
                y = np.where(inside, cy[k] + (rad[k] + clearance) * uy, y)
    vx = np.gradient(x, delta_t) if S > 1 else np.zeros(S)
    vy = np.gradient(y, delta_t) if S > 1 else np.zeros(S)
    return np.stack([x, y, vx, vy], axis=1)


def fixed_prior_group(states, mus, K, d, T=1.0, T_high=10.0) -> LinGroupSpec:
    """FixedPriorGP(K, mu): Lambda = Psi = I, C = 1 (gp/fixed_prior.h:19-50)."""
    n = len(states)
    Kinv = np.linalg.inv(K)
    return LinGroupSpec(start=np.asarray(states, np.int32), Lambda=np.tile(np.eye(d), (n, 1, 1)),
                        Psi=np.tile(np.eye(d), (n, 1, 1)), mu_t=np.asarray(mus, float).reshape(n, d),
                        Kinv=np.tile(Kinv, (n, 1, 1)), C=np.ones(n), T=T, T_high=T_high)


def minacc_group(S: int, Qc: np.ndarray, delta_t: float, T=1.0, T_high=10.0) -> LinGroupSpec:
    """MinimumAccGP(Qc, i, dt, mu0) for i = 0..S-2 (gp/minimum_acc_prior.h:39-80,110-116)."""
    Qc = np.atleast_2d(np.asarray(Qc, float))
    dim = Qc.shape[0]
    ds = 2 * dim
    I = np.eye(dim)
    Phi = np.block([[I, delta_t * I], [np.zeros((dim, dim)), I]])
    iq = np.linalg.inv(Qc)
    invQ = np.block([[12 * iq / delta_t ** 3, -6 * iq / delta_t ** 2], [-6 * iq / delta_t ** 2, 4 * iq / delta_t]])
    Lam = np.hstack([-Phi, np.eye(ds)])
    n = S - 1
    return LinGroupSpec(start=np.arange(n, dtype=np.int32), Lambda=np.tile(Lam, (n, 1, 1)),
                        Psi=np.zeros((n, ds, 2 * ds)), mu_t=np.zeros((n, 2 * ds)), Kinv=np.tile(invQ, (n, 1, 1)),
                        C=np.full(n, 0.5), T=T, T_high=T_high)


def ltv_system(n_links: int, seed: int):
    """Damped-oscillator dynamics of SURVEY 8(d) cfg3 on the quarter intervals of `n_links` links:
    A(t) = [[0, I], [-w(t)^2 I, -c(t) I]], B = [0; I], w, c ~ U(1, 2) per quarter interval (4 n_links + 1 draws each).
    Returns hA [4 n + 1, 4, 4], hB [4 n + 1, 4, 2] (the reference's hA / hB vectors, gp/LTV_prior.h:42-60)."""
    dim, ds = 2, 4
    rng = np.random.default_rng(seed)
    nq = 4 * n_links + 1
    w = rng.uniform(1.0, 2.0, nq)
    c = rng.uniform(1.0, 2.0, nq)
    hA = np.zeros((nq, ds, ds))
    hA[:, :dim, dim:] = np.eye(dim)
    hA[:, dim:, :dim] = -(w ** 2)[:, None, None] * np.eye(dim)
    hA[:, dim:, dim:] = -c[:, None, None] * np.eye(dim)
    hB = np.zeros((nq, ds, dim))
    hB[:, dim:, :] = np.eye(dim)
    return hA, hB


def ltv_links(hA_q: np.ndarray, hB_q: np.ndarray, delta_t: float, ctx=None):
    """Phi, Q, K^-1 = Q^-1 of links with quarter-interval dynamics hA_q [n, 4, ds, ds], hB_q [n, 4, ds, nb].
    ctx = None: NumPy (Van Loan per quarter, ltv_transition_batch).  ctx = a capi.Context: the same on the device
    (gvib200_ltv_transition) -- per-link scaling, so the result does not depend on the batching."""
    if ctx is not None:
        Phi, Q, Kinv = ctx.ltv_transition(hA_q, hB_q, delta_t)
        return Phi, Q, 0.5 * (Kinv + np.transpose(Kinv, (0, 2, 1)))
    Phi, Q = ltv_transition_batch(hA_q, hB_q, delta_t)
    Kinv = np.linalg.inv(Q)
    return Phi, Q, 0.5 * (Kinv + np.transpose(Kinv, (0, 2, 1)))


def ltv_group(S: int, delta_t: float, nominal: np.ndarray, seed=3, T=1.0, T_high=10.0, ctx=None):
    """LTV_GP prior for i = 0..S-2 with the damped-oscillator dynamics of SURVEY 8(d) cfg3 (ltv_system).
    target_mean = -nominal (sign quirk of gp/LTV_prior.h:87-94: the prior is centred on -target_mean)."""
    n = S - 1
    ds = 4
    hA, hB = ltv_system(n, seed)
    idx = 4 * np.arange(n)[:, None] + np.arange(4)[None, :]
    Phi, Q, Kinv = ltv_links(hA[idx], hB[idx], delta_t, ctx)
    Lam = np.concatenate([-Phi, np.broadcast_to(np.eye(ds), (n, ds, ds))], axis=2)
    Psi = -Lam
    target = -nominal
    mu_t = np.concatenate([target[:-1], target[1:]], axis=1)
    g = LinGroupSpec(start=np.arange(n, dtype=np.int32), Lambda=Lam, Psi=Psi, mu_t=mu_t, Kinv=Kinv,
                     C=np.full(n, 0.5), T=T, T_high=T_high)
    return g, Phi, Q, (hA, hB)


# ----------------------------------------------------------------------------------------------
# the configurations of BASELINE.json
# ----------------------------------------------------------------------------------------------
def make_cfg1() -> ProblemSpec:
    """src/1d_example.cpp: one 1-D nonlinear factor, GH degree 10, mu0 = 20, precision0 = 1/9."""
    prm = capi.Stereo1DParams(20.0, 400.0, 0.1, 0.09, 9.0, -0.8)
    spec = ProblemSpec(S=1, d=1)
    spec.groups.append(GhGroupSpec(capi.COST_STEREO_1D, 1, 10, np.zeros(1, np.int32), prm, 1.0, 10.0))
    spec.mu0 = np.array([20.0])
    spec.prec0_D = np.array([[[1.0 / 9.0]]])
    spec.prec0_O = np.zeros((0, 1, 1))
    spec.meta = dict(name="cfg1", step_size_base=0.75, niters=10, niters_lowtemp=10)
    return spec


def make_cfg2(S: int = 1000, delta_t: float = 0.1, anchors_every: int = 10, prec0: float = 10.0) -> ProblemSpec:
    """All-linear 2-D point robot (d = 4): fixed priors at both ends, minimum-acceleration GP prior, weak anchor
    priors every `anchors_every` states (keeps kappa(Vddmu) ~ 3e5)."""
    d = 4
    start = np.array([-15.0, -5.0, 0.0, 0.0])
    goal = np.array([15.0, 14.0, 0.0, 0.0])
    tt = np.linspace(0.0, 1.0, S)[:, None]
    vel = (goal[:2] - start[:2]) / (max(S - 1, 1) * delta_t)
    mu0 = start[None, :] * (1 - tt) + goal[None, :] * tt
    mu0[:, 2:] = vel
    spec = ProblemSpec(S=S, d=d)
    spec.groups.append(fixed_prior_group([0, S - 1], np.stack([start, goal]), 1e-4 * np.eye(d), d))
    spec.groups.append(minacc_group(S, 0.8 * np.eye(2), delta_t))
    if anchors_every > 0:
        st = np.arange(anchors_every, S - 1, anchors_every)
        if len(st):
            spec.groups.append(fixed_prior_group(st, mu0[st], np.eye(d), d))
    spec.mu0 = mu0.reshape(-1)
    spec.prec0_D = np.tile(prec0 * np.eye(d), (S, 1, 1))
    spec.prec0_O = np.zeros((S - 1, d, d))
    spec.meta = dict(name="cfg2", step_size_base=0.55, niters=10, niters_lowtemp=10)
    return spec


def make_cfg3(N: int = 100_000, delta_t: float = 0.2, deg: int = 6, sigma: float = 0.1, prec0: float = 100.0,
              seed: int = 3, clearance: Optional[float] = 0.6, ctx=None) -> ProblemSpec:
    """Headline shape: S = N + 2 states, N single-state planar hinge-SDF factors (d = 4, sparse GH degree `deg`)
    at states 1..S-2, N + 1 LTV GP factors, two fixed priors."""
    d = 4
    S = N + 2
    nominal = lissajous_nominal(S, delta_t, clearance, disc_layout())
    spec = ProblemSpec(S=S, d=d)
    spec.sdf = disc_sdf()
    spec.groups.append(fixed_prior_group([0, S - 1], np.stack([nominal[0], nominal[-1]]), 1e-4 * np.eye(d), d))
    g, _, _, _ = ltv_group(S, delta_t, nominal, seed=seed, ctx=ctx)
    spec.groups.append(g)
    spec.groups.append(GhGroupSpec(capi.COST_PLANAR_HINGE, d, deg, np.arange(1, S - 1, dtype=np.int32),
                                   capi.HingeParams(sigma, 0.5, 1.0), 1.0, 10.0))
    spec.mu0 = nominal.reshape(-1).copy()
    spec.prec0_D = np.tile(prec0 * np.eye(d), (S, 1, 1))
    spec.prec0_O = np.zeros((S - 1, d, d))
    spec.meta = dict(name="cfg3", step_size_base=0.55, niters=10, niters_lowtemp=10, n_nodes=None)
    return spec


def make_factor_batch(N: int = 100_000, deg: int = 6, sigma: float = 0.1, seed: int = 11) -> ProblemSpec:
    """Factor-batch micro-input for the 1e-10 moment parity: N independent single-state hinge
    factors, mu_k ~ U(field), Sigma_k = R diag(lam) R^T, lam log-uniform in [1e-3, 1], R Haar."""
    d = 4
    rng = np.random.default_rng(seed)
    sdf = disc_sdf()
    data, origin, cell = sdf
    rows, cols = data.shape
    mu = np.zeros((N, d))
    mu[:, 0] = rng.uniform(origin[0], origin[0] + (cols - 1) * cell, N)
    mu[:, 1] = rng.uniform(origin[1], origin[1] + (rows - 1) * cell, N)
    mu[:, 2:] = rng.standard_normal((N, 2))
    lam = np.exp(rng.uniform(np.log(1e-3), np.log(1.0), (N, d)))
    G = rng.standard_normal((N, d, d))
    Qm, Rm = np.linalg.qr(G)
    Qm = Qm * np.sign(np.diagonal(Rm, axis1=1, axis2=2))[:, None, :]
    Sigma = (Qm * lam[:, None, :]) @ np.transpose(Qm, (0, 2, 1))
    Sigma = 0.5 * (Sigma + np.transpose(Sigma, (0, 2, 1)))
    prec = np.linalg.inv(Sigma)
    prec = 0.5 * (prec + np.transpose(prec, (0, 2, 1)))
    spec = ProblemSpec(S=N, d=d)
    spec.sdf = sdf
    spec.groups.append(GhGroupSpec(capi.COST_PLANAR_HINGE, d, deg, np.arange(N, dtype=np.int32),
                                   capi.HingeParams(sigma, 0.5, 1.0), 1.0, 10.0))
    spec.mu0 = mu.reshape(-1)
    spec.prec0_D = prec
    spec.prec0_O = np.zeros((N - 1, d, d))
    spec.meta = dict(name="factor_batch", Sigma=Sigma)
    return spec


def ball_sdf3d(nz: int = 40, rows: int = 60, cols: int = 80, origin=(-4.0, -3.0, -2.0), cell: float = 0.1, n_balls: int = 6,
               seed: int = 7):
    """Analytic 3-D signed distance to a few balls on a [nz, rows, cols] grid (x along columns, y along rows)."""
    rng = np.random.default_rng(seed)
    ext = np.array([(cols - 1) * cell, (rows - 1) * cell, (nz - 1) * cell])
    c = np.asarray(origin) + rng.uniform(0.1, 0.9, (n_balls, 3)) * ext
    r = rng.uniform(0.4, 1.0, n_balls)
    x = origin[0] + cell * np.arange(cols)
    y = origin[1] + cell * np.arange(rows)
    z = origin[2] + cell * np.arange(nz)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    sd = np.full(X.shape, np.inf)
    for k in range(n_balls):
        sd = np.minimum(sd, np.sqrt((X - c[k, 0]) ** 2 + (Y - c[k, 1]) ** 2 + (Z - c[k, 2]) ** 2) - r[k])
    return sd, tuple(origin), cell


def _random_spd(rng, N, d, lo=1e-3, hi=1.0):
    lam = np.exp(rng.uniform(np.log(lo), np.log(hi), (N, d)))
    G = rng.standard_normal((N, d, d))
    Qm, Rm = np.linalg.qr(G)
    Qm = Qm * np.sign(np.diagonal(Rm, axis1=1, axis2=2))[:, None, :]
    Sigma = (Qm * lam[:, None, :]) @ np.transpose(Qm, (0, 2, 1))
    return 0.5 * (Sigma + np.transpose(Sigma, (0, 2, 1)))


def example_arm(n_dof: int = 3, sigma: float = 0.5, epsilon: float = 0.5) -> "capi.ArmParams":
    """A small serial arm in Denavit-Hartenberg form for the CudaOperation_3dArm functor (helpers/CudaOperation.h:325-410,
    680-779): link lengths ~1.2, alternating twists, body spheres spread over the links (more spheres than 2 n_dof, so the
    reference's n_balls = theta.size() cap is exercised)."""
    a = [1.2, 1.0, 0.8][:n_dof]
    alpha = [np.pi / 2, -np.pi / 3, 0.4][:n_dof]
    dd = [0.3, 0.0, 0.2][:n_dof]
    bias = [0.1, -0.2, 0.3][:n_dof]
    frames, centers, radii = [], [], []
    for j in range(n_dof):
        for k, t in enumerate((-0.8, -0.4, 0.0)):
            frames.append(j)
            centers.append([t * a[j], 0.05 * (k - 1), 0.02 * j])
            radii.append(0.25 + 0.05 * ((j + k) % 3))
    return capi.ArmParams.make(a, alpha, dd, bias, frames, centers, radii, sigma, epsilon)


def make_factor_batch_functor(kind: int, N: int = 64, d: int = 6, deg: int = 3, sigma: float = 0.5, seed: int = 5) -> ProblemSpec:
    """Factor-batch input for the moment parity of the robot cost functors beyond the planar point robot:
    kind = COST_HINGE_3D (3-D point robot, x[0:3] position in a 3-D field) or COST_QUAD_HINGE (planar quadrotor
    (x, z, phi, ...) in the planar field).  Means are spread over the field (some sigma points leave it: the clamp is
    exercised), covariances random SPD."""
    rng = np.random.default_rng(seed)
    spec = ProblemSpec(S=N, d=d)
    mu = rng.standard_normal((N, d))
    if kind == capi.COST_HINGE_3D:
        spec.sdf3d = ball_sdf3d()
        data, origin, cell = spec.sdf3d
        nz, rows, cols = data.shape
        ext = np.array([(cols - 1) * cell, (rows - 1) * cell, (nz - 1) * cell])
        mu[:, :3] = np.asarray(origin) + rng.uniform(-0.05, 1.05, (N, 3)) * ext
    elif kind == capi.COST_QUAD_HINGE:
        spec.sdf = disc_sdf()
        data, origin, cell = spec.sdf
        rows, cols = data.shape
        mu[:, 0] = rng.uniform(origin[0] - 1.0, origin[0] + (cols - 1) * cell + 1.0, N)
        mu[:, 1] = rng.uniform(origin[1] - 1.0, origin[1] + (rows - 1) * cell + 1.0, N)
        mu[:, 2] = rng.uniform(-np.pi, np.pi, N)
    elif kind == capi.COST_ARM_3D:
        spec.sdf3d = ball_sdf3d()
        mu[:, :d // 2] = rng.uniform(-np.pi, np.pi, (N, d // 2))       # joint angles; velocities stay N(0, 1)
    else:
        raise ValueError(kind)
    Sigma = _random_spd(rng, N, d, 1e-3, 0.3)
    prec = np.linalg.inv(Sigma)
    prec = 0.5 * (prec + np.transpose(prec, (0, 2, 1)))
    params = example_arm(d // 2, sigma) if kind == capi.COST_ARM_3D else capi.HingeParams(sigma, 0.5, 1.0)
    spec.groups.append(GhGroupSpec(kind, d, deg, np.arange(N, dtype=np.int32), params, 1.0, 10.0))
    spec.mu0 = mu.reshape(-1)
    spec.prec0_D = prec
    spec.prec0_O = np.zeros((N - 1, d, d))
    spec.meta = dict(name="factor_batch_functor", Sigma=Sigma)
    return spec


# ----------------------------------------------------------------------------------------------
# multi-GPU: the cfg3 chain cut along the time axis; rank r owns the links [r m, (r+1) m), m = N + 1
# ----------------------------------------------------------------------------------------------
def make_cfg3_segment(rank: int, world: int, N: int = 100_000, delta_t: float = 0.2, deg: int = 6, sigma: float = 0.1,
                      prec0: float = 100.0, seed: int = 3, clearance: Optional[float] = 0.6) -> ProblemSpec:
    """Rank `rank`'s time segment of ONE cfg3 chain of world * (N + 1) + 1 states: N + 2 local states (the first / last
    are shared with the neighbouring ranks), the LTV links it owns, the hinge factors of the states it owns (a shared
    state's factor belongs to the right-hand rank) and its share of the initial precision (a shared diagonal block goes
    entirely to the right-hand rank: only the sums matter).  merge_segments() of all ranks' segments is the single-GPU
    problem."""
    d = 4
    m = N + 1
    S = world * m + 1
    g0 = rank * m
    nominal = lissajous_nominal(S, delta_t, clearance, disc_layout())
    loc = nominal[g0:g0 + m + 1]
    spec = ProblemSpec(S=m + 1, d=d)
    spec.sdf = disc_sdf()
    ends, ends_mu = [], []
    if rank == 0:
        ends.append(0)
        ends_mu.append(nominal[0])
    if rank == world - 1:
        ends.append(m)
        ends_mu.append(nominal[-1])
    if ends:
        spec.groups.append(fixed_prior_group(ends, np.stack(ends_mu), 1e-4 * np.eye(d), d))
    # LTV links: one global random stream (so that the merged problem does not depend on the partition), own slice
    n_links = S - 1
    rng = np.random.default_rng(seed)
    nq = 4 * n_links + 1
    w = rng.uniform(1.0, 2.0, nq)
    c = rng.uniform(1.0, 2.0, nq)
    q0 = 4 * g0
    wq, cq = w[q0:q0 + 4 * m + 1], c[q0:q0 + 4 * m + 1]
    dim, ds = 2, 4
    hA = np.zeros((4 * m + 1, ds, ds))
    hA[:, :dim, dim:] = np.eye(dim)
    hA[:, dim:, :dim] = -(wq ** 2)[:, None, None] * np.eye(dim)
    hA[:, dim:, dim:] = -cq[:, None, None] * np.eye(dim)
    hB = np.zeros((4 * m + 1, ds, dim))
    hB[:, dim:, :] = np.eye(dim)
    idx = 4 * np.arange(m)[:, None] + np.arange(4)[None, :]
    Phi, Q = ltv_transition_batch(hA[idx], hB[idx], delta_t)
    Kinv = np.linalg.inv(Q)
    Kinv = 0.5 * (Kinv + np.transpose(Kinv, (0, 2, 1)))
    Lam = np.concatenate([-Phi, np.broadcast_to(np.eye(ds), (m, ds, ds))], axis=2)
    target = -loc
    spec.groups.append(LinGroupSpec(start=np.arange(m, dtype=np.int32), Lambda=Lam, Psi=-Lam,
                                    mu_t=np.concatenate([target[:-1], target[1:]], axis=1), Kinv=Kinv, C=np.full(m, 0.5)))
    first = 1 if rank == 0 else 0
    spec.groups.append(GhGroupSpec(capi.COST_PLANAR_HINGE, d, deg, np.arange(first, m, dtype=np.int32),
                                   capi.HingeParams(sigma, 0.5, 1.0), 1.0, 10.0))
    spec.mu0 = loc.reshape(-1).copy()
    spec.prec0_D = np.tile(prec0 * np.eye(d), (m + 1, 1, 1))
    if rank < world - 1:
        spec.prec0_D[m] = 0.0  # the shared block belongs to the right-hand neighbour
    spec.prec0_O = np.zeros((m, d, d))
    spec.meta = dict(name="cfg3_segment", rank=rank, world=world, links=m, step_size_base=0.55, niters_lowtemp=10)
    return spec


def merge_segments(segs: List[ProblemSpec]) -> ProblemSpec:
    """The single-GPU problem equivalent to a list of time segments (tests)."""
    d = segs[0].d
    m = segs[0].S - 1
    world = len(segs)
    S = world * m + 1
    out = ProblemSpec(S=S, d=d)
    out.sdf = segs[0].sdf
    out.mu0 = np.concatenate([s.mu0.reshape(-1, d)[:m] for s in segs] + [segs[-1].mu0.reshape(-1, d)[m:]]).reshape(-1)
    out.prec0_D = np.zeros((S, d, d))
    out.prec0_O = np.zeros((S - 1, d, d))
    for r, s in enumerate(segs):
        out.prec0_D[r * m:r * m + m + 1] += s.prec0_D
        out.prec0_O[r * m:(r + 1) * m] = s.prec0_O
    # concatenate groups position-wise by kind: fixed priors, LTV links, hinge factors
    def cat(groups, r_of):
        g0 = groups[0]
        if isinstance(g0, GhGroupSpec):
            return GhGroupSpec(g0.kind, g0.dim, g0.deg, np.concatenate([g.start + r * m for g, r in zip(groups, r_of)]).astype(np.int32),
                               g0.params, g0.T, g0.T_high)
        return LinGroupSpec(start=np.concatenate([g.start + r * m for g, r in zip(groups, r_of)]).astype(np.int32),
                            Lambda=np.concatenate([g.Lambda for g in groups]), Psi=np.concatenate([g.Psi for g in groups]),
                            mu_t=np.concatenate([g.mu_t for g in groups]), Kinv=np.concatenate([g.Kinv for g in groups]),
                            C=np.concatenate([np.broadcast_to(np.asarray(g.C, float), (len(g.start),)) for g in groups]),
                            T=g0.T, T_high=g0.T_high)
    fixed, ltv, hinge = [], [], []
    for r, s in enumerate(segs):
        for g in s.groups:
            if isinstance(g, GhGroupSpec):
                hinge.append((g, r))
            elif g.Lambda.shape[2] == d:
                fixed.append((g, r))
            else:
                ltv.append((g, r))
    for lst in (fixed, ltv, hinge):
        if lst:
            out.groups.append(cat([g for g, _ in lst], [r for _, r in lst]))
    out.meta = dict(segs[0].meta, name="cfg3_merged")
    return out


def make_cfg4(S: int = 10_001, delta_t: float = 1.0, deg: int = 4, closed_form: bool = False) -> ProblemSpec:
    """Prox-GVI shape of BASELINE.json configs[3]: 3-D point robot, state 6 (position + velocity), S states, S - 1
    two-state factors of dim 12 evaluated by sparse GH degree `deg` (2649 nodes at deg 4) with psi = cost_linear_gp
    (exactly integrable: closed_form=True builds the same problem from ProxFactorizedLinear factors -- a built-in KAT),
    fixed priors at both ends.  Delta t = 1, Lambda_0 = 50 I and eta base 0.1 put the reference's factor-wise JKO
    iteration in a regime where every iteration accepts a trial and the cost decreases (with the demo's base 0.75 it
    exhausts its back-tracking every iteration; probed with the oracle)."""
    d, dim = 6, 3
    start = np.array([-2.0, -1.0, 0.5, 0.0, 0.0, 0.0])
    goal = np.array([2.0, 1.5, 1.0, 0.0, 0.0, 0.0])
    tt = np.linspace(0.0, 1.0, S)[:, None]
    mu0 = start[None, :] * (1 - tt) + goal[None, :] * tt
    mu0[:, 3:] = (goal[:3] - start[:3]) / (max(S - 1, 1) * delta_t)
    spec = ProblemSpec(S=S, d=d)
    spec.groups.append(fixed_prior_group([0, S - 1], np.stack([start, goal]), 0.5 * np.eye(d), d))
    lin = minacc_group(S, 0.8 * np.eye(dim), delta_t)
    if closed_form:
        spec.groups.append(lin)
    else:
        Phi = -lin.Lambda[0][:, :d]
        rec = np.concatenate([np.tile(Phi.T.reshape(-1), (S - 1, 1)), np.tile(lin.Kinv[0].T.reshape(-1), (S - 1, 1))], axis=1)
        spec.groups.append(GhGroupSpec(capi.COST_LINEAR_GP, 2 * d, deg, lin.start, rec, 1.0, 10.0))
    spec.mu0 = mu0.reshape(-1)
    spec.prec0_D = np.tile(50.0 * np.eye(d), (S, 1, 1))
    spec.prec0_O = np.zeros((S - 1, d, d))
    spec.meta = dict(name="cfg4", step_size_base=0.1, niters=5, niters_lowtemp=1 << 30)
    return spec


def make_cfg5(n_problems: int = 64, N: int = 1000, first_seed: int = 1000, ctx=None, **kw) -> ProblemSpec:
    """BASELINE.json configs[4]: independent copies of the cfg3 generator (seeds first_seed, first_seed + 1, ...) batched
    as ONE block-diagonal chain: problem b occupies the states [b (N+2), (b+1)(N+2)) and nothing couples consecutive
    problems (the off-diagonal block between them stays zero), so the chain engine, the sweeps and the assembly run over
    the whole batch in single launches.  The line search is shared by the batch (one step size, the summed cost); as long
    as every problem would accept the same trial this is exactly the independent iteration.

    ctx = a capi.Context: the LTV links of all problems are set up in one device launch (gvib200_ltv_transition) and the
    batch is assembled without a per-problem Python loop -- what bench.py --config cfg5 uses for 4096 problems; every
    problem is bit-identical to make_cfg3(N, seed=first_seed + b, ctx=ctx)."""
    if ctx is not None and not kw:
        return _make_cfg5_device(n_problems, N, first_seed, ctx)
    subs = [make_cfg3(N=N, seed=first_seed + b, ctx=ctx, **kw) for b in range(n_problems)]
    Sb = subs[0].S
    d = subs[0].d
    out = ProblemSpec(S=Sb * n_problems, d=d)
    out.sdf = subs[0].sdf
    out.mu0 = np.concatenate([s.mu0 for s in subs])
    out.prec0_D = np.concatenate([s.prec0_D for s in subs])
    out.prec0_O = np.zeros((out.S - 1, d, d))
    for b, s in enumerate(subs):
        out.prec0_O[b * Sb:b * Sb + Sb - 1] = s.prec0_O
    for gi in range(len(subs[0].groups)):
        g0 = subs[0].groups[gi]
        starts = np.concatenate([s.groups[gi].start + b * Sb for b, s in enumerate(subs)]).astype(np.int32)
        if isinstance(g0, GhGroupSpec):
            out.groups.append(GhGroupSpec(g0.kind, g0.dim, g0.deg, starts, g0.params, g0.T, g0.T_high))
        else:
            cat = lambda name: np.concatenate([getattr(s.groups[gi], name) for s in subs])
            out.groups.append(LinGroupSpec(start=starts, Lambda=cat("Lambda"), Psi=cat("Psi"), mu_t=cat("mu_t"), Kinv=cat("Kinv"),
                                           C=np.concatenate([np.broadcast_to(np.asarray(s.groups[gi].C, float), (len(s.groups[gi].start),)) for s in subs]),
                                           T=g0.T, T_high=g0.T_high))
    out.meta = dict(subs[0].meta, name="cfg5", n_problems=n_problems, states_per_problem=Sb)
    return out


def _make_cfg5_device(n_problems: int, N: int, first_seed: int, ctx) -> ProblemSpec:
    """make_cfg5 without the per-problem loop: the problems share everything but the LTV dynamics (the seed only draws
    w(t), c(t)), whose links are integrated in one device launch."""
    base = make_cfg3(N=N, seed=first_seed, ctx=ctx)
    Sb, d, n = base.S, base.d, base.S - 1
    fixed0, ltv0, hinge0 = base.groups
    idx = 4 * np.arange(n)[:, None] + np.arange(4)[None, :]
    delta_t = 0.2  # make_cfg3's default
    Phi = np.empty((n_problems * n, d, d))
    Kinv = np.empty((n_problems * n, d, d))
    CH = 128  # problems per device launch (bounds the host staging arrays)
    for b0 in range(0, n_problems, CH):
        nb_ = min(CH, n_problems - b0)
        hA = np.empty((nb_ * n, 4, d, d))
        hB = np.empty((nb_ * n, 4, d, 2))
        for b in range(nb_):
            a_, b_ = ltv_system(n, first_seed + b0 + b)
            hA[b * n:(b + 1) * n] = a_[idx]
            hB[b * n:(b + 1) * n] = b_[idx]
        Phi[b0 * n:(b0 + nb_) * n], _, Kinv[b0 * n:(b0 + nb_) * n] = ltv_links(hA, hB, delta_t, ctx)
    Lam = np.concatenate([-Phi, np.broadcast_to(np.eye(d), (n_problems * n, d, d))], axis=2)
    off = (np.arange(n_problems, dtype=np.int64) * Sb)[:, None]
    tile = lambda a: np.tile(a, (n_problems,) + (1,) * (a.ndim - 1))
    out = ProblemSpec(S=Sb * n_problems, d=d)
    out.sdf = base.sdf
    out.mu0 = np.tile(base.mu0, n_problems)
    out.prec0_D = tile(base.prec0_D)
    out.prec0_O = np.zeros((out.S - 1, d, d))
    out.groups.append(LinGroupSpec(start=(fixed0.start[None, :] + off).reshape(-1).astype(np.int32), Lambda=tile(fixed0.Lambda),
                                   Psi=tile(fixed0.Psi), mu_t=tile(fixed0.mu_t), Kinv=tile(fixed0.Kinv),
                                   C=np.tile(np.broadcast_to(np.asarray(fixed0.C, float), (len(fixed0.start),)), n_problems),
                                   T=fixed0.T, T_high=fixed0.T_high))
    out.groups.append(LinGroupSpec(start=(ltv0.start[None, :] + off).reshape(-1).astype(np.int32), Lambda=Lam, Psi=-Lam,
                                   mu_t=tile(ltv0.mu_t), Kinv=Kinv, C=np.full(n_problems * n, 0.5), T=ltv0.T, T_high=ltv0.T_high))
    out.groups.append(GhGroupSpec(hinge0.kind, hinge0.dim, hinge0.deg, (hinge0.start[None, :] + off).reshape(-1).astype(np.int32),
                                  hinge0.params, hinge0.T, hinge0.T_high))
    out.meta = dict(base.meta, name="cfg5", n_problems=n_problems, states_per_problem=Sb)
    return out
