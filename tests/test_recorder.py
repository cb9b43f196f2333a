"""result recorder / CSV emitters compatible with helpers/DataRecorder.h:177-224 and
MatrixIO::saveData (helpers/MatrixHelper.h:52-61, CSVFormat helpers/CommonDefinitions.h:32).

CPU: the writer reproduces the reference's own CSV files BYTE FOR BYTE from their parsed values (15 significant digits,
", " separators, no trailing newline); trace_save lays the files out like VIMPResults::save_data.
GPU: examples/1d_example.cpp <prefix> (the reference's driver, one optimize() call + save) writes files that match
data/1d/*.csv in shape and to 1e-10 in value."""
import ctypes as C
import pathlib
import subprocess

import numpy as np
import pytest

import oracle_bridge as ob
from gaussianvi_b200 import capi

GOLD = ob.ROOT / "tests" / "golden" / "ref_1d"


@pytest.mark.parametrize("name", ["mean", "cov", "precision", "cost", "factor_costs", "costmap"])
def test_csv_writer_reproduces_reference_files(name, tmp_path):
    text = (GOLD / f"{name}.csv").read_text()
    a = np.loadtxt(GOLD / f"{name}.csv", delimiter=",", ndmin=2)
    if name == "cost":
        a = a.reshape(-1, 1)
    out = tmp_path / "o.csv"
    capi.csv_write(out, a)
    assert out.read_text() == text


def test_trace_save_layout(tmp_path):
    """VIMPResults::save_data: one column per iteration; zk_sdf is d x S, Sk_sdf is d*d x S of the last iteration;
    joint files are dense (S*d)^2 x niters built from the block-tridiagonal blocks."""
    S, d, nf, n = 3, 2, 4, 5
    rng = np.random.default_rng(0)
    rec = capi.ResultRecorder(n, d, S, nf, joint=True)
    for a in (rec.mean, rec.cov, rec.precision, rec.cov_off, rec.prec_off, rec.factor_costs):
        a[...] = rng.standard_normal(a.shape)
    rec.cost[...] = rng.standard_normal(n)
    rec.trace.n_recorded = 4   # early end: only four iterations are written
    rec.save_data(str(tmp_path) + "/", "run1")
    ld = lambda f: np.loadtxt(tmp_path / f"{f}_run1.csv", delimiter=",", ndmin=2)
    np.testing.assert_allclose(ld("mean"), rec.mean[:4].T, rtol=1e-14)
    np.testing.assert_allclose(ld("cov"), rec.cov[:4].T, rtol=1e-14)
    np.testing.assert_allclose(ld("precision"), rec.precision[:4].T, rtol=1e-14)
    np.testing.assert_allclose(ld("cost").reshape(-1), rec.cost[:4], rtol=1e-14)
    np.testing.assert_allclose(ld("factor_costs"), rec.factor_costs[:4].T, rtol=1e-14)
    np.testing.assert_allclose(ld("zk_sdf"), rec.mean[3].reshape(S, d).T, rtol=1e-14)
    np.testing.assert_allclose(ld("Sk_sdf"), rec.cov[3].reshape(S, d * d).T, rtol=1e-14)
    J = ld("joint_precision")
    assert J.shape == ((S * d) ** 2, 4)
    J3 = J[:, 3].reshape(S * d, S * d, order="F")
    for s in range(S):
        blk = rec.precision[3].reshape(S, d * d)[s].reshape(d, d, order="F")
        np.testing.assert_allclose(J3[s * d:(s + 1) * d, s * d:(s + 1) * d], blk, rtol=1e-14)
    off = rec.prec_off[3].reshape(S - 1, d * d)[0].reshape(d, d, order="F")
    np.testing.assert_allclose(J3[0:d, d:2 * d], off, rtol=1e-14)
    np.testing.assert_allclose(J3[d:2 * d, 0:d], off.T, rtol=1e-14)   # mirrored below the diagonal
    assert np.all(J3[0:d, 2 * d:] == 0)


def test_trace_save_rejects_empty(tmp_path):
    rec = capi.ResultRecorder(2, 1, 1, 1)
    with pytest.raises(capi.GviError):
        rec.save_data(str(tmp_path) + "/")


@pytest.mark.gpu
def test_1d_example_writes_the_reference_files(tmp_path):
    from test_facade import compile_example
    exe = compile_example("1d_example")
    subprocess.run([str(exe), str(tmp_path) + "/"], check=True, capture_output=True)
    for name in ("mean", "cov", "precision", "cost", "factor_costs"):
        ref = np.loadtxt(GOLD / f"{name}.csv", delimiter=",", ndmin=2)
        got = np.loadtxt(tmp_path / f"{name}.csv", delimiter=",", ndmin=2)
        assert got.shape == ref.shape, name
        assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-10, name
    # joint files of a 1-D problem equal the marginal ones; zk / Sk are the last iteration
    np.testing.assert_array_equal(np.loadtxt(tmp_path / "joint_cov.csv", delimiter=",", ndmin=2),
                                  np.loadtxt(tmp_path / "cov.csv", delimiter=",", ndmin=2))
    assert abs(float(np.loadtxt(tmp_path / "zk_sdf.csv")) - np.loadtxt(GOLD / "mean.csv", delimiter=",")[-1]) < 1e-9


@pytest.mark.gpu
def test_recorded_optimize_matches_plain_optimize():
    import gaussianvi_b200 as gv
    from gaussianvi_b200 import problems
    ctx = gv.Context(0)
    spec = problems.make_cfg3(N=40)
    p1 = problems.build_device_problem(ctx, spec)
    p2 = problems.build_device_problem(ctx, spec)
    opts = gv.Problem.default_opts()
    st1, fc, mt = p1.optimize(4, opts, want_traces=True)
    st2, rec = p2.optimize_recorded(4, opts, joint=False)
    assert rec.n_recorded == 4
    np.testing.assert_array_equal(rec.mean, mt)
    np.testing.assert_array_equal(rec.factor_costs[:, :fc.shape[1]], fc)
    np.testing.assert_array_equal(rec.cost, [s.cost for s in st1])
    np.testing.assert_array_equal(p1.mean(), p2.mean())
