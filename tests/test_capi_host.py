"""CPU tests of the product's host side: libgvib200.so loads and exports every symbol include/gvib200.h declares,
the native sparse-GH table generator is bit-identical to the oracle's restatement of nwspgr.m, the chain engine's
__host__ __device__ arithmetic (run by tests/cpp/libbt_host_emu.so) matches the oracle, and the library fails loudly
without a CUDA device (no CPU fallback)."""
import ctypes as C
import re

import numpy as np
import pytest

import gvi_oracle as o
import oracle_bridge as ob
from gaussianvi_b200 import capi

ROOT = ob.ROOT


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def test_header_symbols_are_exported():
    hdr = (ROOT / "include" / "gvib200.h").read_text()
    declared = sorted(set(re.findall(r"\b(gvib200_[A-Za-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gvib200.h but not exported by libgvib200.so"
    assert set(declared) == set(capi.EXPORTS), set(declared) ^ set(capi.EXPORTS)
    assert b"sm_100a" in lib.gvib200_version()


def test_no_cpu_fallback():
    """Without a usable CUDA device the product refuses to run (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the refusal path cannot be exercised")
    with pytest.raises(capi.GviError) as e:
        capi.Context(0)
    assert e.value.code == -2


@pytest.mark.parametrize("dim,deg", [(1, 10), (1, 6), (2, 10), (3, 8), (4, 3), (4, 6), (5, 2), (8, 4), (12, 4)])
def test_native_table_generator_bit_exact(dim, deg):
    """gvib200_table_generate (gaussianvi_b200/csrc/spgh_table.cpp) vs the oracle's nwspgr restatement: same node
    count, same ROW ORDER, bit-identical nodes and weights."""
    Z, w = o.table(dim, deg)
    Zc, wc = capi.table_generate(dim, deg)
    assert Zc.shape == Z.shape
    assert np.array_equal(Zc, Z)
    assert np.array_equal(wc, w)


def test_table_unavailable_is_an_error():
    lib = capi.load_library()
    assert lib.gvib200_table_size(4, 99) < 0


# ------------------------------------------------------------------ chain engine on the host
def _emu():
    lib = C.CDLL(str(ROOT / "tests" / "cpp" / "libbt_host_emu.so"))
    return lib


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


@pytest.mark.parametrize("n", [1, 2, 4, 8, 12])
def test_jacobi_sqrt_matches_eigh(n):
    """The device prologue's symmetric PSD root (quadrature/SparseGaussHermite.h:231-233) vs numpy eigh."""
    lib = _emu()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(n)
    for trial in range(20):
        lam = np.exp(rng.uniform(np.log(1e-6), np.log(10.0), n))
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        Sig = (Q * lam) @ Q.T
        Sig = 0.5 * (Sig + Sig.T)
        S, R = np.zeros((n, n)), np.zeros((n, n))
        assert lib.emu_sqrt_invsqrt(n, Sig.ctypes.data_as(dp), S.ctypes.data_as(dp), R.ctypes.data_as(dp)) == 0
        ref = o.sqrtm_psd(Sig)
        assert rel(S, ref) < 1e-12
        assert rel(R @ R, np.linalg.inv(Sig)) < 1e-9


# ------------------------------------------------------------------ chain engine (bt_cr.h) on the host
def emu_cr_run(S, d, D, O, rhs, force_T, smem=220 * 1024):
    lib = _emu()
    dp = C.POINTER(C.c_double)
    Dc = np.ascontiguousarray(np.transpose(D, (0, 2, 1)))
    Oc = np.ascontiguousarray(np.transpose(O, (0, 2, 1))) if S > 1 else np.zeros((1, d, d))
    x = np.zeros(S * d)
    cD = np.zeros((S, d, d))
    cO = np.zeros((max(S - 1, 1), d, d))
    ld = C.c_double()
    p = lambda a: a.ctypes.data_as(dp)
    rc = lib.emu_cr_blocktri(S, d, p(Dc), p(Oc), p(rhs), p(x), p(cD), p(cO), C.byref(ld), force_T, C.c_size_t(smem))
    return rc, x, np.transpose(cD, (0, 2, 1)), np.transpose(cO[:S - 1], (0, 2, 1)), ld.value


@pytest.mark.parametrize("S,d,force_T", [
    (1, 4, -1), (2, 4, -1), (3, 4, -1), (4, 1, -1), (9, 4, -1), (100, 4, -1), (257, 3, -1), (40, 6, -1),   # top only
    (3, 4, 2), (10, 4, 2), (10, 4, 3), (33, 2, 4), (100, 4, 7), (101, 4, 10), (1000, 4, 64), (1002, 4, 251),  # tiled
    (1000, 4, 0), (5000, 4, 0), (700, 6, 0), (3000, 1, 0), (2500, 2, 37), (777, 3, 0),                       # automatic
])
def test_cr_engine_host_matches_oracle(S, d, force_T):
    rng = np.random.default_rng(S * 10 + d)
    D, O = rand_spd_chain(rng, S, d)
    rhs = rng.standard_normal(S * d)
    rc, x, cD, cO, ld = emu_cr_run(S, d, D, O, rhs, force_T)
    assert rc == 0
    bt = o.BlockTri(D, O)
    ref = o.inverse_gbp(bt)
    assert rel(x, o.block_solve(bt, rhs)) < 1e-11
    assert rel(cD, ref.D) < 1e-11
    if S > 1:
        assert rel(cO, ref.O) < 1e-11
    assert abs(ld - o.logdet(bt)) < 1e-10 * max(1.0, abs(ld))


def test_cr_engine_reports_indefinite_and_too_long():
    D = np.tile(np.eye(2), (50, 1, 1))
    D[31] = -np.eye(2)
    O = np.zeros((49, 2, 2))
    rc, *_ = emu_cr_run(50, 2, D, O, np.zeros(100), 8)
    assert rc == -4
    rc, *_ = emu_cr_run(50, 2, D, O, np.zeros(100), -1)
    assert rc == -4
    # a forced tile size whose separator chain does not fit the top is refused, not mangled
    rng = np.random.default_rng(0)
    Dl, Ol = rand_spd_chain(rng, 4000, 4)
    rc, *_ = emu_cr_run(4000, 4, Dl, Ol, np.zeros(16000), 2, smem=16 * 1024)
    assert rc == -1


@pytest.mark.parametrize("S,d,smem", [(4000, 4, 16 * 1024), (20000, 2, 8 * 1024), (3000, 6, 48 * 1024)])
def test_cr_engine_three_levels(S, d, smem):
    """Chains too long for two levels (here: a tiny shared-memory budget) go through a tiled separator chain."""
    rng = np.random.default_rng(S + d)
    D, O = rand_spd_chain(rng, S, d)
    rhs = rng.standard_normal(S * d)
    rc, x, cD, cO, ld = emu_cr_run(S, d, D, O, rhs, 0, smem=smem)
    assert rc == 0
    bt = o.BlockTri(D, O)
    ref = o.inverse_gbp(bt)
    assert rel(x, o.block_solve(bt, rhs)) < 1e-10
    assert rel(cD, ref.D) < 1e-10 and rel(cO, ref.O) < 1e-10
    assert abs(ld - o.logdet(bt)) < 1e-9 * max(1.0, abs(ld))


# ------------------------------------------------------------------ multi-GPU chain pass replayed on the host
@pytest.mark.parametrize("P,m,d,T", [(2, 5, 4, 2), (2, 64, 4, 16), (3, 100, 2, 7), (4, 333, 4, 50), (8, 40, 4, 40),
                                     (8, 257, 1, 32), (2, 90, 6, 30), (5, 1, 4, 2)])
def test_distributed_chain_pass_matches_oracle(P, m, d, T):
    """P ranks, each owning m links and its share of the end blocks; one boundary all-gather per pass.  The result must
    equal the single-process solve / selected inverse of the global chain of P m + 1 nodes."""
    rng = np.random.default_rng(P * 1000 + m + d)
    S = P * m + 1
    D, O = rand_spd_chain(rng, S, d)
    rhs = rng.standard_normal(S * d).reshape(S, d)
    n = m + 1
    Dloc = np.zeros((P, n, d, d))
    gloc = np.zeros((P, n, d))
    Oloc = np.zeros((P, m, d, d))
    for r in range(P):
        Dloc[r] = D[r * m:r * m + n]
        gloc[r] = rhs[r * m:r * m + n]
        Oloc[r] = O[r * m:(r + 1) * m]
    # shared end blocks: split the diagonal block and the rhs between the two neighbours (any split must work)
    for r in range(1, P):
        share = rng.uniform(0.2, 0.8)
        Dloc[r - 1, m] = share * D[r * m]
        Dloc[r, 0] = D[r * m] - Dloc[r - 1, m]
        gloc[r - 1, m] = share * rhs[r * m]
        gloc[r, 0] = rhs[r * m] - gloc[r - 1, m]
    lib = _emu()
    dp = C.POINTER(C.c_double)
    p = lambda a: a.ctypes.data_as(dp)
    Dc = np.ascontiguousarray(np.transpose(Dloc, (0, 1, 3, 2)))
    Oc = np.ascontiguousarray(np.transpose(Oloc, (0, 1, 3, 2)))
    gc = np.ascontiguousarray(gloc)
    x = np.zeros((P, n, d))
    cD = np.zeros((P, n, d, d))
    cO = np.zeros((P, m, d, d))
    ld = C.c_double()
    rc = lib.emu_cr_distributed(P, m, d, p(Dc), p(Oc), p(gc), p(x), p(cD), p(cO), C.byref(ld), T)
    assert rc == 0
    bt = o.BlockTri(D, O)
    ref = o.inverse_gbp(bt)
    xe = o.block_solve(bt, rhs.reshape(-1)).reshape(S, d)
    cD = np.transpose(cD, (0, 1, 3, 2))
    cO = np.transpose(cO, (0, 1, 3, 2))
    for r in range(P):
        assert rel(x[r], xe[r * m:r * m + n]) < 1e-10
        assert rel(cD[r], ref.D[r * m:r * m + n]) < 1e-10
        assert rel(cO[r], ref.O[r * m:(r + 1) * m]) < 1e-10
    assert abs(ld.value - o.logdet(bt)) < 1e-9 * max(1.0, abs(ld.value))
