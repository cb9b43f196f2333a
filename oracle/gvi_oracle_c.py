"""ctypes wrapper of oracle/libgvi_oracle_c.so (the C restatement of the reference's NGD-GVI path).

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by gaussianvi_b200/.  Takes the same neutral ProblemSpec
(gaussianvi_b200.problems) the GPU path is built from, so both sides see identical numbers."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None
MAX_LIN_GROUPS = 4
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


class LinGroup(C.Structure):
    _fields_ = [("n", C.c_int), ("dim", C.c_int), ("m", C.c_int), ("kdim", C.c_int), ("start", _IP),
                ("Lambda", _DP), ("Psi", _DP), ("mu_t", _DP), ("Kinv", _DP), ("C", _DP), ("T", _DP)]


class OrcProblem(C.Structure):
    _fields_ = [("S", C.c_int), ("d", C.c_int), ("n_gh", C.c_int), ("gh_dim", C.c_int), ("n_nodes", C.c_int),
                ("cost_kind", C.c_int), ("gh_start", _IP), ("Z", _DP), ("w", _DP), ("gh_T", _DP), ("cp", C.c_double * 8),
                ("rows", C.c_int), ("cols", C.c_int), ("ox", C.c_double), ("oy", C.c_double), ("cell", C.c_double),
                ("sdf", _DP), ("n_lin_groups", C.c_int), ("lin", LinGroup * MAX_LIN_GROUPS)]


class OrcStats(C.Structure):
    _fields_ = [("cost", C.c_double), ("new_cost", C.c_double), ("step", C.c_double), ("n_backtrack", C.c_int),
                ("accepted", C.c_int), ("status", C.c_int), ("n_psi_sweeps", C.c_int), ("n_inversions", C.c_int)]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "libgvi_oracle_c.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
        _LIB = C.CDLL(str(so))
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _dp(a):
    return a.ctypes.data_as(_DP)


def _cm(blocks):
    """[n, r, c] row-major NumPy blocks -> packed column-major."""
    return np.ascontiguousarray(np.transpose(np.asarray(blocks, dtype=np.float64), (0, 2, 1)))


class COracle:
    """C-oracle twin of a ProblemSpec with at most one GH group (stereo / hinge / quadratic cost) and up to four
    linear groups.  Factor-cost order: the GH group first, then the linear groups in spec order."""

    def __init__(self, spec, table_fn):
        from gaussianvi_b200 import capi, problems  # neutral spec / constants only
        self.spec = spec
        self.S, self.d = spec.S, spec.d
        self._keep = []
        p = OrcProblem()
        p.S, p.d = spec.S, spec.d
        p.n_gh = 0
        p.n_lin_groups = 0
        self.order = []  # (is_gh, group index in spec, n) in C-oracle cost order
        lin_specs = []
        for gi, g in enumerate(spec.groups):
            if isinstance(g, problems.GhGroupSpec):
                assert p.n_gh == 0, "C oracle: one GH group"
                n = len(g.start)
                p.n_gh, p.gh_dim, p.cost_kind = n, g.dim, g.kind
                Z, w = table_fn(g.dim, g.deg)
                Z = np.ascontiguousarray(Z, dtype=np.float64)
                w = np.ascontiguousarray(w, dtype=np.float64)
                st = np.ascontiguousarray(g.start, dtype=np.int32)
                T = np.full(n, float(g.T))
                self._keep += [Z, w, st, T]
                p.n_nodes = len(w)
                p.Z, p.w, p.gh_start, p.gh_T = _dp(Z), _dp(w), st.ctypes.data_as(_IP), _dp(T)
                if g.kind == capi.COST_STEREO_1D:
                    q = g.params
                    vals = [q.mu_p, q.f, q.b, q.sig_r_sq, q.sig_p_sq, q.y_offset]
                elif g.kind == capi.COST_PLANAR_HINGE:
                    q = g.params
                    vals = [q.sigma, q.epsilon, q.radius]
                elif g.kind == capi.COST_QUADRATIC:
                    vals = [float(np.asarray(g.params).reshape(-1)[0])]
                else:
                    raise ValueError("C oracle: unsupported GH cost kind %d" % g.kind)
                for i, v in enumerate(vals):
                    p.cp[i] = v
                self.gh_spec_index = gi
            else:
                lin_specs.append((gi, g))
        if spec.sdf is not None:
            data, origin, cell = spec.sdf
            cm = np.ascontiguousarray(np.asarray(data, dtype=np.float64).T)  # column-major rows x cols
            self._keep.append(cm)
            p.rows, p.cols = data.shape
            p.ox, p.oy, p.cell = origin[0], origin[1], cell
            p.sdf = _dp(cm)
        assert len(lin_specs) <= MAX_LIN_GROUPS
        if p.n_gh:
            self.order.append((True, self.gh_spec_index, p.n_gh))
        for k, (gi, g) in enumerate(lin_specs):
            n = len(g.start)
            lg = p.lin[k]
            lg.n, lg.m, lg.dim, lg.kdim = n, g.Lambda.shape[1], g.Lambda.shape[2], g.Psi.shape[2]
            arrs = dict(start=np.ascontiguousarray(g.start, dtype=np.int32), Lambda=_cm(g.Lambda), Psi=_cm(g.Psi),
                        mu_t=np.ascontiguousarray(g.mu_t, dtype=np.float64), Kinv=_cm(g.Kinv),
                        C=np.ascontiguousarray(np.broadcast_to(np.asarray(g.C, float), (n,))), T=np.full(n, float(g.T)))
            self._keep.append(arrs)
            lg.start = arrs["start"].ctypes.data_as(_IP)
            for name in ("Lambda", "Psi", "mu_t", "Kinv", "C", "T"):
                setattr(lg, name, _dp(arrs[name]))
            self.order.append((False, gi, n))
        p.n_lin_groups = len(lin_specs)
        self.p = p
        self.n_factors = sum(n for _, _, n in self.order)
        # state
        self.mu = np.array(spec.mu0, dtype=np.float64).reshape(-1).copy()
        self.LD = _cm(spec.prec0_D)
        self.LO = np.zeros((max(self.S, 1), self.d, self.d))
        if self.S > 1:
            self.LO[:self.S - 1] = _cm(spec.prec0_O)
        self.cD = np.zeros((self.S, self.d, self.d))
        self.cO = np.zeros((max(self.S, 1), self.d, self.d))
        self.carried = C.c_double(0.0)
        self.have_carry = C.c_int(0)

    # ---- pieces ----
    def covariance_blocks(self):
        rc = lib().orc_inverse_gbp(self.S, self.d, _dp(self.LD), _dp(self.LO), _dp(self.cD), _dp(self.cO))
        assert rc == 0
        return np.transpose(self.cD, (0, 2, 1)).copy(), np.transpose(self.cO[:self.S - 1], (0, 2, 1)).copy()

    def moments(self, faithful=False):
        self.covariance_blocks()
        n, dim = self.p.n_gh, self.p.gh_dim
        E0, E1, E2 = np.zeros(n), np.zeros((n, dim)), np.zeros((n, dim, dim))
        lib().orc_gh_moments(C.byref(self.p), _dp(self.mu), _dp(self.cD), _dp(self.cO), int(faithful), _dp(E0), _dp(E1),
                             _dp(E2))
        return E0, E1, np.transpose(E2, (0, 2, 1)).copy()

    def cost_value(self):
        c = C.c_double()
        fc = np.zeros(max(self.n_factors, 1))
        rc = lib().orc_cost_value(C.byref(self.p), _dp(self.mu), _dp(self.LD), _dp(self.LO), C.byref(c), _dp(fc))
        assert rc == 0
        return c.value, fc[:self.n_factors]

    def factor_costs_in_spec_order(self, fc):
        """Reorder the C oracle's cost vector (GH group first) into the spec's id order."""
        out, off = {}, 0
        for is_gh, gi, n in self.order:
            out[gi] = fc[off:off + n]
            off += n
        return np.concatenate([out[gi] for gi in sorted(out)])

    def iterate(self, step_size_base=0.55, backtrack_ratio=0.75, max_backtrack=10, schedule=1) -> OrcStats:
        st = OrcStats()
        lib().orc_ngd_iterate(C.byref(self.p), _dp(self.mu), _dp(self.LD), _dp(self.LO), _dp(self.cD), _dp(self.cO),
                              C.c_double(step_size_base), C.c_double(backtrack_ratio), max_backtrack, schedule,
                              C.byref(self.carried), C.byref(self.have_carry), C.byref(st))
        return st

    def snapshot(self):
        """Copy of the optimizer state (bench.py's CPU arm rewinds to it like the GPU arm does on the device)."""
        return (self.mu.copy(), self.LD.copy(), self.LO.copy(), self.cD.copy(), self.cO.copy(), self.carried.value,
                self.have_carry.value)

    def restore(self, snap):
        mu, LD, LO, cD, cO, carried, have = snap
        self.mu[...] = mu
        self.LD[...] = LD
        self.LO[...] = LO
        self.cD[...] = cD
        self.cO[...] = cO
        self.carried.value = carried
        self.have_carry.value = have

    def mean(self):
        return self.mu.copy()

    def cov_blocks(self):
        return np.transpose(self.cD, (0, 2, 1)).copy(), np.transpose(self.cO[:self.S - 1], (0, 2, 1)).copy()

    def prec_blocks(self):
        return np.transpose(self.LD, (0, 2, 1)).copy(), np.transpose(self.LO[:self.S - 1], (0, 2, 1)).copy()


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))
