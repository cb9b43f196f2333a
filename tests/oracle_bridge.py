"""Build the CPU-oracle twin (oracle/gvi_oracle.py) of a gaussianvi_b200.problems.ProblemSpec.
Test infrastructure: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT / "oracle") not in sys.path:
    sys.path.insert(0, str(ROOT / "oracle"))

import gvi_oracle as o  # noqa: E402
from gaussianvi_b200 import capi, problems  # noqa: E402


SHARED_KINDS = (capi.COST_STEREO_1D, capi.COST_PLANAR_HINGE, capi.COST_QUADRATIC, capi.COST_HINGE_3D, capi.COST_QUAD_HINGE,
                capi.COST_ARM_3D)


def psi_for_group(spec, g, i):
    """Vectorised psi(X) of GH factor i of group g."""
    if g.kind == capi.COST_STEREO_1D:
        p = g.params
        def psi(X, p=p):
            x = X[:, 0]
            y = p.f * p.b / p.mu_p + p.y_offset
            return (x - p.mu_p) * (x - p.mu_p) / p.sig_p_sq / 2 + (y - p.f * p.b / x) * (y - p.f * p.b / x) / p.sig_r_sq / 2
        return psi
    if g.kind == capi.COST_PLANAR_HINGE:
        data, origin, cell = spec.sdf
        sdf = o.PlanarSDF(np.asarray(origin), cell, data)
        return o.make_hinge_cost(sdf, g.params.sigma, g.params.epsilon, g.params.radius)
    if g.kind == capi.COST_QUAD_HINGE:
        data, origin, cell = spec.sdf
        sdf = o.PlanarSDF(np.asarray(origin), cell, data)
        return o.make_quad_hinge_cost(sdf, g.params.sigma, g.params.epsilon, g.params.radius)
    if g.kind == capi.COST_HINGE_3D:
        data, origin, cell = spec.sdf3d
        sdf = o.SignedDistanceField3D(np.asarray(origin), cell, data)
        return o.make_hinge3d_cost(sdf, g.params.sigma, g.params.epsilon, g.params.radius)
    if g.kind == capi.COST_ARM_3D:
        data, origin, cell = spec.sdf3d
        sdf = o.SignedDistanceField3D(np.asarray(origin), cell, data)
        q = g.params
        nd, ns = q.n_dof, q.n_spheres
        return o.make_arm_cost(sdf, list(q.a)[:nd], list(q.alpha)[:nd], list(q.d)[:nd], list(q.theta_bias)[:nd], list(q.frames)[:ns],
                               [[q.centers[i][k] for k in range(3)] for i in range(ns)], list(q.radii)[:ns], q.sigma, q.epsilon)
    if g.kind == capi.COST_LINEAR_GP:
        ds = g.dim // 2
        rec = np.asarray(g.params).reshape(len(g.start), 2, ds, ds)
        Phi = rec[i, 0].T  # records are column-major
        Qi = rec[i, 1].T
        return o.make_cost_linear_gp(Phi, Qi)
    if g.kind == capi.COST_FIXED_GP:
        rec = np.asarray(g.params).reshape(len(g.start), g.dim * g.dim + g.dim)
        Ki = rec[i, :g.dim * g.dim].reshape(g.dim, g.dim).T
        mu = rec[i, g.dim * g.dim:]
        return o.make_cost_fixed_gp(Ki, mu)
    if g.kind == capi.COST_QUADRATIC:
        c = float(np.asarray(g.params).reshape(-1)[0])
        return lambda X: c * (X ** 2).sum(1)
    raise ValueError(g.kind)


def build_factors(spec, fast=True, faithful_linear=False):
    factors = []
    for g in spec.groups:
        if isinstance(g, problems.GhGroupSpec):
            shared = psi_for_group(spec, g, 0) if g.kind in SHARED_KINDS else None
            for i, s in enumerate(g.start):
                psi = shared if shared is not None else psi_for_group(spec, g, i)
                factors.append(o.GHFactor(g.dim, spec.d, g.deg, psi, int(s), g.T, g.T_high, fast=fast))
        else:
            n = len(g.start)
            Cv = np.broadcast_to(np.asarray(g.C, float), (n,))
            for i, s in enumerate(g.start):
                m = o.LinearModel(g.Lambda[i], g.Psi[i], g.mu_t[i], g.Kinv[i], float(Cv[i]))
                factors.append(o.LinearFactorOpt(g.Lambda.shape[2], spec.d, m, int(s), g.T, g.T_high, faithful=faithful_linear))
    return factors


def build_oracle(spec, fast=True, faithful_linear=False, solver="direct", niters=None):
    """faithful_linear=True evaluates the linear factors' Vddmu through the reference's O(dim^4) fourth-moment
    expression (ngd/NGDFactorizedLinear.h:107-119).  It is algebraically 2 C A / T, but it feeds the rounding
    asymmetry of Sigma_k back into the precision and amplifies it ~1e3x per iteration (tests/test_oracle.py
    documents it), so multi-iteration parity runs use the closed form."""
    factors = build_factors(spec, fast=fast, faithful_linear=faithful_linear)
    opt = o.NGDGH(factors, spec.d, spec.S, niters or spec.meta.get("niters", 5), solver=solver)
    opt.set_step_size_base(spec.meta.get("step_size_base", 0.55))
    opt.set_niter_low_temperature(spec.meta.get("niters_lowtemp", 10))
    if spec.mu0 is not None:
        opt.set_initial_values(spec.mu0, o.BlockTri(np.array(spec.prec0_D, float), np.array(spec.prec0_O, float).reshape(max(spec.S - 1, 0), spec.d, spec.d)))
    return opt


def build_oracle_prox(spec, fast=True, niters=None):
    """Prox-GVI twin (oracle ProxGVIGH over ProxGHFactor / ProxLinearFactor) of a ProblemSpec."""
    factors = []
    for g in spec.groups:
        if isinstance(g, problems.GhGroupSpec):
            shared = psi_for_group(spec, g, 0) if g.kind in SHARED_KINDS else None
            for i, s in enumerate(g.start):
                psi = shared if shared is not None else psi_for_group(spec, g, i)
                factors.append(o.ProxGHFactor(g.dim, spec.d, g.deg, psi, int(s), g.T, g.T_high, fast=fast))
        else:
            n = len(g.start)
            Cv = np.broadcast_to(np.asarray(g.C, float), (n,))
            for i, s in enumerate(g.start):
                m = o.LinearModel(g.Lambda[i], g.Psi[i], g.mu_t[i], g.Kinv[i], float(Cv[i]))
                factors.append(o.ProxLinearFactor(g.Lambda.shape[2], spec.d, m, int(s), g.T, g.T_high, faithful=False))
    opt = o.ProxGVIGH(factors, spec.d, spec.S, niters or spec.meta.get("niters", 5))
    opt.set_step_size_base(spec.meta.get("step_size_base", 0.55))
    opt.set_niter_low_temperature(spec.meta.get("niters_lowtemp", 10))
    opt.set_initial_values(spec.mu0, o.BlockTri(np.array(spec.prec0_D, float), np.array(spec.prec0_O, float).reshape(max(spec.S - 1, 0), spec.d, spec.d)))
    return opt
