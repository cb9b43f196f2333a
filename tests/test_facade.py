"""C++ facade (gaussianvi_b200/cpp/gvi): the reference's class names on top of the C-ABI.
CPU: the examples compile (with and without the Eigen stand-in there is only the stand-in in this image).
GPU: examples/1d_example.cpp reproduces the reference's golden trace data/1d/*.csv; examples/planar_chain.cpp (factors
interleaved the way a planner creates them) matches the same problem driven through the ctypes mirror and the oracle."""
import pathlib
import subprocess

import numpy as np
import pytest

import oracle_bridge as ob
from gaussianvi_b200 import capi, problems

ROOT = ob.ROOT
GOLDEN = ROOT / "tests" / "golden"
BUILD = ROOT / "tests" / "cpp" / "_build"


def compile_example(name):
    BUILD.mkdir(exist_ok=True)
    out = BUILD / name
    src = ROOT / "examples" / f"{name}.cpp"
    lib = ROOT / "gaussianvi_b200"
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, (lib / "cpp" / "gvi" / "gvi.h").stat().st_mtime):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(lib / "cpp"), str(src), "-L", str(lib), "-lgvib200",
                        f"-Wl,-rpath,{lib}", "-o", str(out)], check=True)
    return out


@pytest.mark.parametrize("name", ["1d_example", "planar_chain", "1d_example_proxGVI", "point_robot_3d", "ltv_chain"])
def test_facade_examples_compile(name):
    assert compile_example(name).exists()


def test_unspecialised_cost_class_does_not_compile(tmp_path):
    """NGDFactorizedSimpleGH (an arbitrary host function, NoneType) has no device functor: a compile-time error, not a
    CPU fallback."""
    src = tmp_path / "bad.cpp"
    src.write_text('#include "ngd/NGDFactorizedSimpleGH.h"\n'
                   'double f(const gvi::VectorXd&, const gvi::NoneType&) { return 0; }\n'
                   'int main() { gvi::NGDFactorizedSimpleGH x(1, 1, 10, f, gvi::NoneType(), 1, 0, 1.0, 10.0); return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", str(ROOT / "gaussianvi_b200" / "cpp"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "DeviceCostTraits" in r.stderr


@pytest.mark.gpu
def test_facade_1d_example_golden_trace():
    exe = compile_example("1d_example")
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    rows = np.array([[float(x) for x in line.split()] for line in out.strip().splitlines()])
    assert rows.shape == (10, 5)
    g = lambda n: np.loadtxt(GOLDEN / "ref_1d" / f"{n}.csv", delimiter=",").reshape(-1)
    for col, name in ((1, "mean"), (2, "cov"), (3, "precision"), (4, "cost")):
        ref = g(name)
        assert np.abs(rows[:, col] - ref).max() / np.abs(ref).max() < 1e-10, name


@pytest.mark.gpu
def test_facade_1d_prox_example_golden_trace():
    exe = compile_example("1d_example_proxGVI")
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    rows = np.array([[float(x) for x in line.split()] for line in out.strip().splitlines()])
    assert rows.shape == (10, 5)
    g = lambda n: np.loadtxt(GOLDEN / "ref_1d_proxgvi" / f"{n}.csv", delimiter=",").reshape(-1)
    for col, name in ((1, "mean"), (2, "cov"), (3, "precision"), (4, "cost")):
        ref = g(name)
        assert np.abs(rows[:, col] - ref).max() / np.abs(ref).max() < 1e-10, name


@pytest.mark.gpu
def test_facade_planar_chain_matches_ctypes_mirror_and_oracle(gpu_ctx):
    S, niters = 40, 6
    exe = compile_example("planar_chain")
    out = subprocess.run([str(exe), str(S), str(niters)], capture_output=True, text=True, check=True).stdout
    costs = np.array([float(l.split()[2]) for l in out.splitlines() if l.startswith("cost")])
    mean = np.array([float(l.split()[2]) for l in out.splitlines() if l.startswith("mean")])
    sumfc = float([l for l in out.splitlines() if l.startswith("sumfc")][0].split()[1])
    # the same problem through the neutral spec
    spec = problems.make_cfg2(S=S)
    rows, cols, cell, origin = 120, 160, 0.25, (-20.0, -10.0)
    xs = origin[0] + cell * np.arange(cols)
    ys = origin[1] + cell * np.arange(rows)
    X, Y = np.meshgrid(xs, ys)
    spec.sdf = (np.hypot(X - 0.0, Y - 4.0) - 2.5, origin, cell)
    spec.groups.append(problems.GhGroupSpec(capi.COST_PLANAR_HINGE, 4, 6, np.arange(1, S - 1, dtype=np.int32),
                                            capi.HingeParams(0.1, 0.5, 1.0), 1.0, 10.0))
    p = problems.build_device_problem(gpu_ctx, spec)
    opts = capi.Problem.default_opts()
    stats = p.optimize(niters, opts)
    assert len(stats) == len(costs)
    assert np.abs(costs - np.array([s.cost for s in stats])).max() < 1e-9 * np.abs(costs).max()
    assert np.abs(mean - p.mean()).max() < 1e-9 * np.abs(mean).max()
    c, fc = p.cost()
    assert abs(sumfc - fc.sum()) < 1e-9 * abs(sumfc)
    ref = ob.build_oracle(spec, niters=niters)
    ref.optimize()
    assert np.abs(mean - ref.mean()).max() < 1e-7 * np.abs(mean).max()


def test_robot_cost_classes_compile(tmp_path):
    """Hinge3DCost / QuadHingeCost are device cost classes of the facade."""
    src = tmp_path / "robots.cpp"
    src.write_text('#include "ngd/NGDFactorizedBaseGH.h"\n#include "ngd/NGD-GH.h"\n'
                   'using namespace gvi;\n'
                   'double f3(const VectorXd&, const Hinge3DCost&) { return 0; }\n'
                   'double fq(const VectorXd&, const QuadHingeCost&) { return 0; }\n'
                   'int main() {\n'
                   '  auto sdf = std::make_shared<SignedDistanceField>();\n'
                   '  Hinge3DCost c3; c3.sdf = sdf;\n'
                   '  QuadHingeCost cq; cq.sdf = std::make_shared<PlanarSDF>();\n'
                   '  NGDFactorizedBaseGH<Hinge3DCost> a(6, 6, 3, f3, c3, 4, 1, 1.0, 10.0);\n'
                   '  NGDFactorizedBaseGH<QuadHingeCost> b(6, 6, 3, fq, cq, 4, 1, 1.0, 10.0);\n'
                   '  return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", str(ROOT / "gaussianvi_b200" / "cpp"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cuda_alias_api_compiles(tmp_path):
    """callers written against the reference's `_Cuda` names (NGDFactorizedBaseGH_Cuda<CudaOperation_*>,
    NGDFactorizedLinear_Cuda, classify_factors / set_alpha / cuda_init hooks) compile against the facade."""
    src = tmp_path / "cuda_alias.cpp"
    src.write_text('#include "ngd/NGDFactorizedBaseGH_Cuda.h"\n#include "ngd/NGD-GH-Cuda.h"\n'
                   '#include "gp/factorized_opts_linear_Cuda.h"\n#include "helpers/CudaOperation.h"\n'
                   'using namespace gvi;\n'
                   'int main() {\n'
                   '  auto cuda = std::make_shared<CudaOperation_PlanarPR>(15.5, 0.5, 1.0);\n'
                   '  cuda->set_sdf(std::make_shared<PlanarSDF>());\n'
                   '  using Col = NGDFactorizedBaseGH_Cuda<CudaOperation_PlanarPR>;\n'
                   '  std::vector<std::shared_ptr<GVIFactorizedBase>> fs;\n'
                   '  auto map = std::make_shared<QuadratureWeightsMap>();\n'
                   '  auto f = std::make_shared<Col>(4, 4, 6, 10, 1, 15.5, 0.5, 1.0, 1.0, 10.0, map, cuda);\n'
                   '  f->cuda_init(); f->cuda_free();\n'
                   '  auto q = std::make_shared<CudaOperation_Quad>();\n'
                   '  NGDFactorizedBaseGH_Cuda<CudaOperation_Quad> fq(6, 6, 3, 10, 1, 15.5, 0.5, 1.0, 1.0, 10.0, map, q);\n'
                   '  VectorXd v3 = VectorXd::Zero(3); std::vector<int> fr{0, 1, 2}; MatrixXd ctr = MatrixXd::Zero(3, 3);\n'
                   '  auto arm = std::make_shared<CudaOperation_3dArm>(v3, v3, v3, v3, v3, fr, ctr, 15.5, 0.5);\n'
                   '  arm->set_sdf(std::make_shared<SignedDistanceField>());\n'
                   '  NGDFactorizedBaseGH_Cuda<CudaOperation_3dArm> fa(6, 6, 3, 10, 1, 15.5, 0.5, 0.0, 1.0, 10.0, map, arm);\n'
                   '  auto r3 = std::make_shared<CudaOperation_3dpR>();\n'
                   '  NGDFactorizedBaseGH_Cuda<CudaOperation_3dpR> f3(6, 6, 3, 10, 1, 15.5, 0.5, 1.0, 1.0, 10.0, map, r3);\n'
                   '  std::vector<std::shared_ptr<Col>> v{f};\n'
                   '  NGDGH<Col> opt(v, 4, 10);\n'
                   '  opt.classify_factors(); opt.set_alpha(1.0);\n'
                   '  return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", str(ROOT / "gaussianvi_b200" / "cpp"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_facade_point_robot_3d_matches_ctypes_mirror_and_oracle(gpu_ctx):
    """examples/point_robot_3d.cpp (the reference's _Cuda class names, 3-D SignedDistanceField through the facade) against the
    same problem driven through the ctypes mirror and the oracle."""
    S, niters, d, dt = 12, 5, 6, 0.3
    exe = compile_example("point_robot_3d")
    out = subprocess.run([str(exe), str(S), str(niters)], capture_output=True, text=True, check=True).stdout
    costs = np.array([float(l.split()[2]) for l in out.splitlines() if l.startswith("cost")])
    mean = np.array([float(l.split()[2]) for l in out.splitlines() if l.startswith("mean")])
    nz, rows, cols, cell, origin = 40, 60, 80, 0.1, (-4.0, -3.0, -2.0)
    z, y, x = np.meshgrid(origin[2] + cell * np.arange(nz), origin[1] + cell * np.arange(rows), origin[0] + cell * np.arange(cols),
                          indexing="ij")
    spec = problems.ProblemSpec(S=S, d=d)
    spec.sdf3d = (np.sqrt(x * x + (y - 0.3) ** 2 + z * z) - 0.8, origin, cell)
    start = np.array([-3.5, -2.5, -1.5, 0, 0, 0.0])
    goal = np.array([3.0, 2.0, 1.0, 0, 0, 0.0])
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
    spec.meta = dict(niters=niters, step_size_base=0.55, niters_lowtemp=10)
    p = problems.build_device_problem(gpu_ctx, spec)
    stats = p.optimize(niters, capi.Problem.default_opts())
    assert len(costs) == niters
    assert np.abs(costs - np.array([s.cost for s in stats])).max() / np.abs(costs).max() < 1e-12
    assert np.abs(mean - p.mean()).max() / np.abs(mean).max() < 1e-12
    ref = ob.build_oracle(spec, niters=niters)
    ref.optimize()
    assert np.abs(mean - ref.mean()).max() / np.abs(mean).max() < 1e-7


def _ltv_chain_spec(S):
    """The problem of examples/ltv_chain.cpp as a neutral spec."""
    d, dt = 4, 0.2
    i = np.arange(S, dtype=np.float64)
    nominal = np.stack([-6.0 + 12.0 * i / (S - 1), 5.0 + 3.5 * np.sin(3.0 * i / (S - 1)),
                        np.full(S, 12.0 / ((S - 1) * dt)), 3.5 * 3.0 / ((S - 1) * dt) * np.cos(3.0 * i / (S - 1))], axis=1)
    nq = 4 * (S - 1) + 1
    q = np.arange(nq, dtype=np.float64)
    w, c = 1.5 + 0.4 * np.sin(0.37 * q), 1.4 + 0.3 * np.cos(0.21 * q)
    hA = np.zeros((nq, 4, 4))
    hA[:, :2, 2:] = np.eye(2)
    hA[:, 2:, :2] = -(w ** 2)[:, None, None] * np.eye(2)
    hA[:, 2:, 2:] = -c[:, None, None] * np.eye(2)
    hB = np.zeros((nq, 4, 2))
    hB[:, 2:, :] = np.eye(2)
    idx = 4 * np.arange(S - 1)[:, None] + np.arange(4)[None, :]
    Phi, Q, Kinv = problems.ltv_links(hA[idx], hB[idx], dt)
    Lam = np.concatenate([-Phi, np.broadcast_to(np.eye(d), (S - 1, d, d))], axis=2)
    target = -nominal
    spec = problems.ProblemSpec(S=S, d=d)
    rows, cols, cell, origin = 120, 160, 0.25, (-20.0, -10.0)
    X, Y = np.meshgrid(origin[0] + cell * np.arange(cols), origin[1] + cell * np.arange(rows))
    spec.sdf = (np.hypot(X - 1.0, Y - 5.0) - 2.0, origin, cell)
    spec.groups.append(problems.fixed_prior_group([0, S - 1], np.stack([nominal[0], nominal[-1]]), 1e-4 * np.eye(d), d))
    spec.groups.append(problems.LinGroupSpec(start=np.arange(S - 1, dtype=np.int32), Lambda=Lam, Psi=-Lam,
                                             mu_t=np.concatenate([target[:-1], target[1:]], axis=1), Kinv=Kinv, C=np.full(S - 1, 0.5)))
    spec.groups.append(problems.GhGroupSpec(capi.COST_PLANAR_HINGE, d, 6, np.arange(1, S - 1, dtype=np.int32),
                                            capi.HingeParams(0.1, 0.5, 1.0), 1.0, 10.0))
    spec.mu0 = nominal.reshape(-1).copy()
    spec.prec0_D = np.tile(100.0 * np.eye(d), (S, 1, 1))
    spec.prec0_O = np.zeros((S - 1, d, d))
    spec.meta = dict(step_size_base=0.55, niters_lowtemp=1 << 30)
    return spec


@pytest.mark.gpu
@pytest.mark.parametrize("alpha", [0.8, 1.0])
def test_facade_ltv_chain_ema_and_accessors_match_mirror_and_oracle(gpu_ctx, alpha):
    """examples/ltv_chain.cpp: facade LTV_GP (its own host-side Van Loan set-up) + _Cuda collision factors + the EMA update
    set_alpha + E_Phis / switch_to_high_temperature / SparseGaussHermite::update_parameters / sigmapts, against the same
    problem through the ctypes mirror (1e-9: the LTV blocks come from two implementations of the same exponential) and
    the oracle with the EMA of gvibase/GVI-GH-Cuda-impl.h:112-114 (1e-7)."""
    import gvi_oracle as o
    S, niters = 30, 6
    exe = compile_example("ltv_chain")
    out = subprocess.run([str(exe), str(S), str(niters), str(alpha)], capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    costs = np.array([float(l.split()[2]) for l in lines if l.startswith("cost ")])
    mean = np.array([float(l.split()[2]) for l in lines if l.startswith("mean ")])
    ephi = np.array([[float(x) for x in l.split()[2:]] for l in lines if l.startswith("ephi ")])
    c_low, c_high, temp = (float(x) for x in [l for l in lines if l.startswith("costs ")][0].split()[1:5:1] if x != "temperature")
    gh = [float(x) for x in [l for l in lines if l.startswith("gh ")][0].split()[1:]]
    spec = _ltv_chain_spec(S)
    p = problems.build_device_problem(gpu_ctx, spec)
    opts = capi.Problem.default_opts()
    opts.niters_lowtemp = 1 << 30
    opts.ema_alpha = alpha
    stats = p.optimize(niters, opts)
    assert len(costs) == niters and all(s.accepted for s in stats)
    assert np.abs(costs - np.array([s.cost for s in stats])).max() < 1e-9 * np.abs(costs).max()
    assert np.abs(mean - p.mean()).max() < 1e-9 * np.abs(mean).max()
    ref = ob.build_oracle(spec, niters=niters)
    ref.alpha = alpha
    recs = ref.optimize()
    assert all(r.accepted for r in recs)
    assert np.abs(costs - np.array([r.cost for r in recs])).max() < 1e-8 * np.abs(costs).max()
    assert np.abs(mean - ref.mean()).max() < 1e-7 * np.abs(mean).max()
    # per-factor expectations in the caller's (interleaved) order
    (E0, E1, E2), = p.moments()
    c, fc = p.cost()
    assert abs(c - c_low) < 1e-9 * abs(c)
    k = 0
    for i in range(S):
        if i == 0:
            assert abs(ephi[k, 0] - fc[0]) < 1e-9 * max(1.0, abs(fc[0])); k += 1
        if i == S - 1:
            assert abs(ephi[k, 0] - fc[1]) < 1e-9 * max(1.0, abs(fc[1])); k += 1
        if i < S - 1:
            assert abs(ephi[k, 0] - fc[2 + i]) < 1e-9 * max(1.0, abs(fc[2 + i])); k += 1
        if 0 < i < S - 1:
            j = i - 1
            assert abs(ephi[k, 0] - E0[j]) <= 1e-10 * max(abs(E0[j]), 1e-300) + 1e-300
            assert abs(ephi[k, 1] - E1[j, 0]) <= 1e-9 * np.abs(E1[j]).max() + 1e-300
            assert abs(ephi[k, 2] - E2[j, 1, 1]) <= 1e-9 * np.abs(E2[j]).max() + 1e-300
            k += 1
    assert k == len(ephi)
    p.switch_to_high_temperature()
    c2, _ = p.cost()
    assert abs(c2 - c_high) < 1e-9 * abs(c2) and temp == 10.0 and c2 != c
    # SparseGaussHermite after update_parameters(deg 6, dim 4): 953 sigma points, moments vs the oracle
    m4 = np.array([1.0, 3.5, 0.4, 0.6])
    P4 = np.full((4, 4), 0.05) + np.diag([0.45, 0.55, 0.65, 0.75])
    Z, w = o.table(4, 6)
    psi = ob.psi_for_group(spec, spec.groups[2], 0)
    r0, r1, r2 = o.moments(psi, m4, P4, Z, w)
    assert int(gh[0]) == 953
    assert abs(gh[1] - r0) < 1e-10 * abs(r0) and abs(gh[2] - r1[0]) < 1e-9 * np.abs(r1).max()
    assert abs(gh[3] - m4[1]) < 1e-12 and abs(gh[4] - P4[1, 2]) < 1e-12
