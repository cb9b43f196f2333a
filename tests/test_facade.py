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


@pytest.mark.parametrize("name", ["1d_example", "planar_chain", "1d_example_proxGVI", "point_robot_3d"])
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
