"""ctypes binding of libgvib200.so (include/gvib200.h) -- the host-side mirror used by the Python
tests and bench.py.  Everything numerical happens inside the shared library on the GPU; this
module only marshals NumPy arrays.  There is no fallback: if the library is missing or no CUDA
device is usable, construction raises."""
from __future__ import annotations

import ctypes as C
import os
import pathlib
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_PKG = pathlib.Path(__file__).resolve().parent
LIB_PATH = _PKG / "libgvib200.so"

COST_STEREO_1D = 1
COST_PLANAR_HINGE = 2
COST_LINEAR_GP = 3
COST_FIXED_GP = 4
COST_QUADRATIC = 5
COST_HINGE_3D = 6
COST_QUAD_HINGE = 7
COST_ARM_3D = 8
ARM_MAX_DOF, ARM_MAX_SPHERES = 3, 12

E_NOTSPD = -4

# every symbol include/gvib200.h declares (checked by tests/test_capi_symbols.py)
EXPORTS = [
    "gvib200_ctx_create", "gvib200_ctx_destroy", "gvib200_last_error", "gvib200_version", "gvib200_ctx_set_comm",
    "gvib200_table_size", "gvib200_table_generate", "gvib200_table_set", "gvib200_problem_create",
    "gvib200_problem_destroy", "gvib200_set_planar_sdf", "gvib200_add_gh_factors", "gvib200_add_linear_factors",
    "gvib200_problem_finalize", "gvib200_set_state", "gvib200_get_mean", "gvib200_get_prec_blocks",
    "gvib200_get_cov_blocks", "gvib200_moments", "gvib200_cost", "gvib200_gradients", "gvib200_get_V",
    "gvib200_default_opts", "gvib200_optimize", "gvib200_ngd_iterate", "gvib200_reset_schedule",
    "gvib200_selected_inverse", "gvib200_blocktri_solve", "gvib200_time_stage", "gvib200_fp64_peak",
    "gvib200_launch_count", "gvib200_timer_start", "gvib200_timer_stop", "gvib200_profile_begin", "gvib200_profile_end",
    "gvib200_kernel_class_name", "gvib200_problem_info", "gvib200_snapshot_save", "gvib200_snapshot_restore",
    "gvib200_problem_set_option", "gvib200_prox_iterate", "gvib200_prox_optimize",
    "gvib200_table_file_write", "gvib200_table_file_load", "gvib200_table_file_query", "gvib200_table_get",
    "gvib200_optimize_traced", "gvib200_csv_write", "gvib200_trace_save", "gvib200_set_sdf3d",
    "gvib200_evaluated_factors", "gvib200_ltv_transition", "gvib200_switch_to_high_temperature", "gvib200_ctx_mailbox_create", "gvib200_ctx_mailbox_connect",
    "gvib200_set_state_async", "gvib200_get_mean_async", "gvib200_get_prec_blocks_async", "gvib200_get_cov_blocks_async",
    "gvib200_sync", "gvib200_set_batch", "gvib200_batch_iterate", "gvib200_batch_costs", "gvib200_batch_optimize",
]


class GviError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gvib200 error {code}: {msg}")
        self.code = code


class Opts(C.Structure):
    _fields_ = [("step_size_base", C.c_double), ("backtrack_ratio", C.c_double), ("max_backtrack", C.c_int),
                ("niters_lowtemp", C.c_int), ("reuse_accepted_sweep", C.c_int), ("ema_alpha", C.c_double)]


class IterStats(C.Structure):
    _fields_ = [("cost", C.c_double), ("new_cost", C.c_double), ("step", C.c_double), ("n_backtrack", C.c_int),
                ("accepted", C.c_int), ("switched_high_T", C.c_int), ("converged", C.c_int), ("status", C.c_int),
                ("n_moment_sweeps", C.c_int), ("n_cost_sweeps", C.c_int)]


class Profile(C.Structure):
    _fields_ = [("n_classes", C.c_int), ("count", C.c_longlong * 16), ("ms", C.c_double * 16)]


class Trace(C.Structure):
    _fields_ = [("capacity", C.c_int), ("n_recorded", C.c_int), ("mean", C.POINTER(C.c_double)),
                ("cov_diag", C.POINTER(C.c_double)), ("prec_diag", C.POINTER(C.c_double)), ("cov_off", C.POINTER(C.c_double)),
                ("prec_off", C.POINTER(C.c_double)), ("cost", C.POINTER(C.c_double)), ("fac_costs", C.POINTER(C.c_double))]


class Info(C.Structure):
    _fields_ = [("num_states", C.c_int), ("dim_state", C.c_int), ("n_factors", C.c_int), ("n_gh_factors", C.c_int),
                ("n_linear_factors", C.c_int), ("chain_levels", C.c_int), ("chain_tiles", C.c_int), ("chain_tile_links", C.c_int),
                ("sigma_points_per_sweep", C.c_longlong)]


class Stereo1DParams(C.Structure):
    _fields_ = [("mu_p", C.c_double), ("f", C.c_double), ("b", C.c_double), ("sig_r_sq", C.c_double),
                ("sig_p_sq", C.c_double), ("y_offset", C.c_double)]


class HingeParams(C.Structure):
    _fields_ = [("sigma", C.c_double), ("epsilon", C.c_double), ("radius", C.c_double)]


class ArmParams(C.Structure):
    """gvib200_arm_params: Denavit-Hartenberg arm with body spheres (helpers/CudaOperation.h:325-410, 680-779)."""
    _fields_ = [("sigma", C.c_double), ("epsilon", C.c_double), ("n_dof", C.c_int), ("n_spheres", C.c_int),
                ("a", C.c_double * ARM_MAX_DOF), ("alpha", C.c_double * ARM_MAX_DOF), ("d", C.c_double * ARM_MAX_DOF),
                ("theta_bias", C.c_double * ARM_MAX_DOF), ("frames", C.c_int * ARM_MAX_SPHERES),
                ("centers", (C.c_double * 3) * ARM_MAX_SPHERES), ("radii", C.c_double * ARM_MAX_SPHERES)]

    @classmethod
    def make(cls, a, alpha, d, theta_bias, frames, centers, radii, sigma=15.5, epsilon=0.5):
        p = cls()
        p.sigma, p.epsilon, p.n_dof, p.n_spheres = sigma, epsilon, len(a), len(frames)
        for j in range(len(a)):
            p.a[j], p.alpha[j], p.d[j], p.theta_bias[j] = a[j], alpha[j], d[j], theta_bias[j]
        for i in range(len(frames)):
            p.frames[i], p.radii[i] = int(frames[i]), radii[i]
            for k in range(3):
                p.centers[i][k] = centers[i][k]
        return p


_lib = None
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)


def load_library() -> C.CDLL:
    """Load libgvib200.so (built in-tree by __graft_entry__.build()).  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    lib.gvib200_last_error.restype = C.c_char_p
    lib.gvib200_version.restype = C.c_char_p
    lib.gvib200_launch_count.restype = C.c_longlong
    lib.gvib200_launch_count.argtypes = [C.c_void_p]
    lib.gvib200_kernel_class_name.restype = C.c_char_p
    _lib = lib
    return lib


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_DP)


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_IP)


def _check(rc: int):
    if rc < 0:
        raise GviError(rc, load_library().gvib200_last_error().decode())
    return rc


def _blocks_to_c(blocks: np.ndarray) -> np.ndarray:
    """[n, d, d] NumPy (row-major per block) -> packed column-major blocks."""
    return np.ascontiguousarray(np.transpose(np.asarray(blocks, dtype=np.float64), (0, 2, 1)))


def _blocks_from_c(buf: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.transpose(buf, (0, 2, 1)))


def table_generate(dim: int, deg: int):
    lib = load_library()
    n = _check(lib.gvib200_table_size(dim, deg))
    nodes = np.zeros((n, dim))
    w = np.zeros(n)
    _check(lib.gvib200_table_generate(dim, deg, _dp(nodes), _dp(w), n))
    return nodes, w


def csv_write(path, a: np.ndarray):
    """MatrixIO::saveData (helpers/MatrixHelper.h:52-61): Eigen CSVFormat, 15 significant digits."""
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    f = np.asfortranarray(a)
    _check(load_library().gvib200_csv_write(str(path).encode(), a.shape[0], a.shape[1], f.ctypes.data_as(_DP)))


class ResultRecorder:
    """Host buffers of gvib200_trace + save_data() (VIMPResults::save_data, helpers/DataRecorder.h:177-224)."""

    def __init__(self, niters: int, dim_state: int, nstates: int, n_factors: int, joint: bool = False):
        self.niters, self.d, self.S, self.n_factors = niters, dim_state, nstates, n_factors
        d, S = dim_state, nstates
        self.mean = np.zeros((niters, S * d))
        self.cov = np.zeros((niters, S * d * d))
        self.precision = np.zeros((niters, S * d * d))
        self.cov_off = np.zeros((niters, max(S - 1, 1) * d * d)) if joint else None
        self.prec_off = np.zeros((niters, max(S - 1, 1) * d * d)) if joint else None
        self.cost = np.zeros(niters)
        self.factor_costs = np.zeros((niters, max(n_factors, 1)))
        self.joint = joint
        p = lambda a: a.ctypes.data_as(_DP) if a is not None else None
        self.trace = Trace(niters, 0, p(self.mean), p(self.cov), p(self.precision), p(self.cov_off), p(self.prec_off),
                           p(self.cost), p(self.factor_costs))

    @property
    def n_recorded(self) -> int:
        return self.trace.n_recorded

    def save_data(self, prefix: str = "", afterfix: str = "", dense_limit: int = 64):
        _check(load_library().gvib200_trace_save(C.byref(self.trace), self.S, self.d, self.n_factors, str(prefix).encode(),
                                                 str(afterfix).encode(), dense_limit if self.joint else 0))


def table_file_write(path: str, keys):
    """Write the rules [(dim, deg), ...] in the reference's cereal table format (quadrature/saveSparseGHWeightMap.h)."""
    lib = load_library()
    dims = np.ascontiguousarray([k[0] for k in keys], dtype=np.int32)
    degs = np.ascontiguousarray([k[1] for k in keys], dtype=np.int32)
    _check(lib.gvib200_table_file_write(str(path).encode(), len(keys), dims.ctypes.data_as(C.POINTER(C.c_int32)),
                                        degs.ctypes.data_as(C.POINTER(C.c_int32))))


def table_file_query(path: str):
    """[(dim, deg, n_nodes), ...] of a table file, in file order."""
    lib = load_library()
    n = _check(lib.gvib200_table_file_query(str(path).encode(), 0, None, None, None))
    dims, degs, sizes = (np.zeros(max(n, 1), dtype=np.int32) for _ in range(3))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    _check(lib.gvib200_table_file_query(str(path).encode(), n, ip(dims), ip(degs), ip(sizes)))
    return [(int(dims[i]), int(degs[i]), int(sizes[i])) for i in range(n)]


class Context:
    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        _check(self.lib.gvib200_ctx_create(device, C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.gvib200_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def table_set(self, dim, deg, nodes, w):
        nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        _check(self.lib.gvib200_table_set(self.h, dim, deg, len(w), _dp(nodes), _dp(w)))

    def table_file_load(self, path: str) -> int:
        n = C.c_int()
        _check(self.lib.gvib200_table_file_load(self.h, str(path).encode(), C.byref(n)))
        return n.value

    def table_get(self, dim: int, deg: int):
        n = _check(self.lib.gvib200_table_get(self.h, dim, deg, None, None, 0))
        nodes = np.zeros((n, dim))
        w = np.zeros(n)
        _check(self.lib.gvib200_table_get(self.h, dim, deg, _dp(nodes), _dp(w), n))
        return nodes, w

    def fp64_peak_tflops(self) -> float:
        v = C.c_double()
        _check(self.lib.gvib200_fp64_peak(self.h, C.byref(v)))
        return v.value

    def launch_count(self) -> int:
        return int(self.lib.gvib200_launch_count(self.h))

    def ltv_transition(self, A: np.ndarray, B: np.ndarray, delta_t: float, want_inverse: bool = True):
        """Device-side LTV prior set-up (gvib200_ltv_transition; gp/LTV_prior.h:123-197): A [n, 4, ds, ds], B [n, 4, ds, nb]
        piece-wise constant on the four quarter intervals -> Phi, Q (and Q^-1) as [n, ds, ds]."""
        A = np.asarray(A, dtype=np.float64)
        B = np.asarray(B, dtype=np.float64)
        n, _, ds, _ = A.shape
        nb = B.shape[3]
        Ac = np.ascontiguousarray(np.transpose(A, (0, 1, 3, 2)))  # column-major blocks
        Bc = np.ascontiguousarray(np.transpose(B, (0, 1, 3, 2)))
        Phi, Q = np.zeros((n, ds, ds)), np.zeros((n, ds, ds))
        Qi = np.zeros((n, ds, ds)) if want_inverse else None
        _check(self.lib.gvib200_ltv_transition(self.h, n, ds, nb, _dp(Ac), _dp(Bc), C.c_double(delta_t), _dp(Phi), _dp(Q),
                                               _dp(Qi) if want_inverse else None))
        Phi, Q = np.transpose(Phi, (0, 2, 1)), np.transpose(Q, (0, 2, 1))
        return (Phi, Q, np.transpose(Qi, (0, 2, 1))) if want_inverse else (Phi, Q)

    def selected_inverse(self, D: np.ndarray, O: np.ndarray):
        S, d = D.shape[0], D.shape[1]
        Dc, Oc = _blocks_to_c(D), _blocks_to_c(O) if S > 1 else np.zeros((1, d, d))
        cD = np.zeros((S, d, d))
        cO = np.zeros((max(S - 1, 1), d, d))
        ld = C.c_double()
        _check(self.lib.gvib200_selected_inverse(self.h, S, d, _dp(Dc), _dp(Oc), _dp(cD), _dp(cO), C.byref(ld)))
        return _blocks_from_c(cD), _blocks_from_c(cO[:S - 1]), ld.value

    def blocktri_solve(self, D: np.ndarray, O: np.ndarray, rhs: np.ndarray):
        S, d = D.shape[0], D.shape[1]
        Dc, Oc = _blocks_to_c(D), _blocks_to_c(O) if S > 1 else np.zeros((1, d, d))
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros(S * d)
        ld = C.c_double()
        _check(self.lib.gvib200_blocktri_solve(self.h, S, d, _dp(Dc), _dp(Oc), _dp(rhs), _dp(x), C.byref(ld)))
        return x, ld.value


class Problem:
    """Device-resident NGD-GVI problem (mirror of gvi::NGDGH construction + optimize)."""

    def __init__(self, ctx: Context, num_states: int, dim_state: int):
        self.ctx = ctx
        self.lib = ctx.lib
        self.S, self.d = num_states, dim_state
        self.h = C.c_void_p()
        self.n_factors = 0
        self.gh_dims = []  # (n, dim) per GH group in id order
        _check(self.lib.gvib200_problem_create(ctx.h, num_states, dim_state, C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.gvib200_problem_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- definition ----
    def set_planar_sdf(self, data: np.ndarray, origin, cell_size: float):
        rows, cols = data.shape
        col_major = np.ascontiguousarray(np.asarray(data, dtype=np.float64).T)  # [cols, rows] == column-major rows x cols
        _check(self.lib.gvib200_set_planar_sdf(self.h, rows, cols, C.c_double(origin[0]), C.c_double(origin[1]),
                                               C.c_double(cell_size), _dp(col_major)))

    def set_sdf3d(self, data: np.ndarray, origin, cell_size: float):
        """data [nz, rows, cols] (SignedDistanceField, helpers/CudaOperation.h:133-160)."""
        nz, rows, cols = data.shape
        flat = np.ascontiguousarray(np.transpose(np.asarray(data, dtype=np.float64), (0, 2, 1)))  # [z][c][r] -> r + c*rows + z*rows*cols
        _check(self.lib.gvib200_set_sdf3d(self.h, rows, cols, nz, C.c_double(origin[0]), C.c_double(origin[1]),
                                          C.c_double(origin[2]), C.c_double(cell_size), _dp(flat)))

    def add_gh_factors(self, kind: int, dim: int, deg: int, start_index, params, T=None, T_high=None) -> int:
        start = np.ascontiguousarray(start_index, dtype=np.int32)
        n = len(start)
        if isinstance(params, C.Structure):
            buf = (C.c_char * C.sizeof(params)).from_buffer_copy(params)
            nbytes = C.sizeof(params)
        else:
            arr = np.ascontiguousarray(params, dtype=np.float64)
            buf = arr.ctypes.data_as(C.c_void_p)
            nbytes = arr.nbytes
            self._keep = arr
        Tn = None if T is None else np.ascontiguousarray(np.broadcast_to(np.asarray(T, float), (n,)))
        Th = None if T_high is None else np.ascontiguousarray(np.broadcast_to(np.asarray(T_high, float), (n,)))
        first = C.c_int()
        _check(self.lib.gvib200_add_gh_factors(self.h, kind, dim, deg, n, _ip(start), _dp(Tn), _dp(Th), buf,
                                               C.c_size_t(nbytes), C.byref(first)))
        self.n_factors += n
        self.gh_dims.append((n, dim))
        return first.value

    def add_linear_factors(self, start_index, Lambda, Psi, mu_t, Kinv, Cc, T=None, T_high=None) -> int:
        """Lambda [n, m, dim], Psi [n, m, kdim], mu_t [n, kdim], Kinv [n, m, m], Cc [n] (NumPy row-major blocks)."""
        start = np.ascontiguousarray(start_index, dtype=np.int32)
        n = len(start)
        Lambda = np.asarray(Lambda, float)
        Psi = np.asarray(Psi, float)
        m, dim = Lambda.shape[1], Lambda.shape[2]
        kdim = Psi.shape[2]
        Lc = np.ascontiguousarray(np.transpose(Lambda, (0, 2, 1)))
        Pc = np.ascontiguousarray(np.transpose(Psi, (0, 2, 1)))
        Kc = np.ascontiguousarray(np.transpose(np.asarray(Kinv, float), (0, 2, 1)))
        mt = np.ascontiguousarray(mu_t, dtype=np.float64)
        Cv = np.ascontiguousarray(np.broadcast_to(np.asarray(Cc, float), (n,)))
        Tn = None if T is None else np.ascontiguousarray(np.broadcast_to(np.asarray(T, float), (n,)))
        Th = None if T_high is None else np.ascontiguousarray(np.broadcast_to(np.asarray(T_high, float), (n,)))
        first = C.c_int()
        _check(self.lib.gvib200_add_linear_factors(self.h, dim, m, kdim, n, _ip(start), _dp(Lc), _dp(Pc), _dp(mt),
                                                   _dp(Kc), _dp(Cv), _dp(Tn), _dp(Th), C.byref(first)))
        self.n_factors += n
        return first.value

    def finalize(self):
        _check(self.lib.gvib200_problem_finalize(self.h))

    # ---- state ----
    def set_state(self, mu=None, prec_D=None, prec_O=None):
        mu_c = None if mu is None else np.ascontiguousarray(mu, dtype=np.float64)
        Dc = None if prec_D is None else _blocks_to_c(prec_D)
        Oc = None if (prec_O is None or self.S == 1) else _blocks_to_c(prec_O)
        _check(self.lib.gvib200_set_state(self.h, _dp(mu_c), _dp(Dc), _dp(Oc)))

    # ---- raw variants: caller-owned buffers already in the C-ABI layout (column-major blocks), no marshalling; with
    # pinned buffers the library's cudaMemcpyAsync is a single DMA ----
    def set_state_raw(self, mu: np.ndarray, prec_diag: np.ndarray, prec_off: Optional[np.ndarray]):
        _check(self.lib.gvib200_set_state(self.h, _dp(mu), _dp(prec_diag), _dp(prec_off) if self.S > 1 else None))

    # ---- asynchronous variants (pinned buffers): enqueued on the handle's stream, valid after sync() ----
    def set_state_raw_async(self, mu: np.ndarray, prec_diag: np.ndarray, prec_off: Optional[np.ndarray]):
        _check(self.lib.gvib200_set_state_async(self.h, _dp(mu), _dp(prec_diag), _dp(prec_off) if self.S > 1 else None))

    def get_mean_into_async(self, mu: np.ndarray):
        _check(self.lib.gvib200_get_mean_async(self.h, _dp(mu)))

    def get_cov_blocks_into_async(self, diag: np.ndarray, off: np.ndarray):
        _check(self.lib.gvib200_get_cov_blocks_async(self.h, _dp(diag), _dp(off)))

    def sync(self):
        _check(self.lib.gvib200_sync(self.h))

    def get_mean_into(self, mu: np.ndarray):
        _check(self.lib.gvib200_get_mean(self.h, _dp(mu)))

    def get_cov_blocks_into(self, diag: np.ndarray, off: np.ndarray):
        _check(self.lib.gvib200_get_cov_blocks(self.h, _dp(diag), _dp(off)))

    def mean(self) -> np.ndarray:
        mu = np.zeros(self.S * self.d)
        _check(self.lib.gvib200_get_mean(self.h, _dp(mu)))
        return mu

    def _get_blocks(self, fn):
        D = np.zeros((self.S, self.d, self.d))
        O = np.zeros((max(self.S - 1, 1), self.d, self.d))
        _check(fn(self.h, _dp(D), _dp(O)))
        return _blocks_from_c(D), _blocks_from_c(O[:self.S - 1])

    def covariance(self):
        return self._get_blocks(self.lib.gvib200_get_cov_blocks)

    def precision(self):
        return self._get_blocks(self.lib.gvib200_get_prec_blocks)

    # ---- hot path pieces ----
    def moments(self):
        """Per GH factor (id order): list of (E0 [n], E1 [n, dim], E2 [n, dim, dim]) per group."""
        n0 = sum(n for n, _ in self.gh_dims)
        n1 = sum(n * dim for n, dim in self.gh_dims)
        n2 = sum(n * dim * dim for n, dim in self.gh_dims)
        E0, E1, E2 = np.zeros(max(n0, 1)), np.zeros(max(n1, 1)), np.zeros(max(n2, 1))
        _check(self.lib.gvib200_moments(self.h, _dp(E0), _dp(E1), _dp(E2)))
        out = []
        o0 = o1 = o2 = 0
        for n, dim in self.gh_dims:
            e2 = E2[o2:o2 + n * dim * dim].reshape(n, dim, dim).transpose(0, 2, 1)
            out.append((E0[o0:o0 + n].copy(), E1[o1:o1 + n * dim].reshape(n, dim).copy(), np.ascontiguousarray(e2)))
            o0 += n
            o1 += n * dim
            o2 += n * dim * dim
        return out

    def cost(self, mu=None, prec_D=None, prec_O=None, want_factor_costs=True):
        mu_c = None if mu is None else np.ascontiguousarray(mu, dtype=np.float64)
        Dc = None if prec_D is None else _blocks_to_c(prec_D)
        Oc = None if (prec_O is None or self.S == 1) else _blocks_to_c(prec_O)
        c = C.c_double()
        fc = np.zeros(max(self.n_factors, 1)) if want_factor_costs else None
        _check(self.lib.gvib200_cost(self.h, _dp(mu_c), _dp(Dc), _dp(Oc), C.byref(c), _dp(fc)))
        return c.value, (fc[:self.n_factors] if fc is not None else None)

    def gradients(self):
        dmu = np.zeros(self.S * self.d)
        dD = np.zeros((self.S, self.d, self.d))
        dO = np.zeros((max(self.S - 1, 1), self.d, self.d))
        _check(self.lib.gvib200_gradients(self.h, _dp(dmu), _dp(dD), _dp(dO)))
        return dmu, _blocks_from_c(dD), _blocks_from_c(dO[:self.S - 1])

    def get_V(self):
        v = np.zeros(self.S * self.d)
        D = np.zeros((self.S, self.d, self.d))
        O = np.zeros((max(self.S - 1, 1), self.d, self.d))
        _check(self.lib.gvib200_get_V(self.h, _dp(v), _dp(D), _dp(O)))
        return v, _blocks_from_c(D), _blocks_from_c(O[:self.S - 1])

    @staticmethod
    def default_opts() -> Opts:
        o = Opts()
        load_library().gvib200_default_opts(C.byref(o))
        return o

    def iterate(self, opts: Optional[Opts] = None) -> IterStats:
        st = IterStats()
        _check(self.lib.gvib200_ngd_iterate(self.h, C.byref(opts) if opts is not None else None, C.byref(st)))
        return st

    def optimize(self, n_iters: int, opts: Optional[Opts] = None, want_traces: bool = False):
        stats = (IterStats * n_iters)()
        done = C.c_int()
        fc = np.zeros((n_iters, max(self.n_factors, 1))) if want_traces else None
        mt = np.zeros((n_iters, self.S * self.d)) if want_traces else None
        _check(self.lib.gvib200_optimize(self.h, C.byref(opts) if opts is not None else None, n_iters, stats,
                                         C.byref(done), _dp(fc), _dp(mt)))
        out = [stats[i] for i in range(done.value)]
        if want_traces:
            return out, fc[:done.value, :self.n_factors], mt[:done.value]
        return out

    def optimize_recorded(self, n_iters: int, opts: Optional[Opts] = None, prox: bool = False, joint: bool = False):
        """Run n_iters iterations with the result recorder attached (VIMPResults, helpers/DataRecorder.h, in banded form).
        Returns (stats, ResultRecorder); ResultRecorder.save_data(prefix, afterfix) writes the reference's CSV files."""
        rec = ResultRecorder(n_iters, self.d, self.S, self.n_factors, joint=joint)
        stats = (IterStats * n_iters)()
        done = C.c_int()
        _check(self.lib.gvib200_optimize_traced(self.h, C.byref(opts) if opts is not None else None, n_iters, 1 if prox else 0,
                                                stats, C.byref(done), C.byref(rec.trace)))
        return [stats[i] for i in range(done.value)], rec

    # ---- batches of independent problems, line search per problem ----
    def set_batch(self, state_offsets):
        off = np.ascontiguousarray(state_offsets, dtype=np.int32)
        self._n_batch = len(off) - 1
        _check(self.lib.gvib200_set_batch(self.h, self._n_batch, off.ctypes.data_as(_IP)))

    def batch_iterate(self, opts: Optional[Opts] = None):
        """One iteration of every problem of the batch; returns (per-problem IterStats list, trial sweeps of the batch)."""
        stats = (IterStats * self._n_batch)()
        nt = C.c_int()
        _check(self.lib.gvib200_batch_iterate(self.h, C.byref(opts) if opts is not None else None, stats, C.byref(nt)))
        return [stats[i] for i in range(self._n_batch)], nt.value

    def batch_optimize(self, n_iters: int, opts: Optional[Opts] = None):
        """Up to n_iters lock-step iterations of every problem; returns a list (per iteration) of per-problem IterStats lists."""
        stats = (IterStats * (n_iters * self._n_batch))()
        done = C.c_int()
        _check(self.lib.gvib200_batch_optimize(self.h, C.byref(opts) if opts is not None else None, n_iters, stats, C.byref(done)))
        return [[stats[it * self._n_batch + q] for q in range(self._n_batch)] for it in range(done.value)]

    def batch_costs(self) -> np.ndarray:
        out = np.zeros(self._n_batch)
        _check(self.lib.gvib200_batch_costs(self.h, _dp(out)))
        return out

    def evaluated_factors(self, reset: bool = True) -> int:
        v = C.c_longlong()
        _check(self.lib.gvib200_evaluated_factors(self.h, C.byref(v), 1 if reset else 0))
        return v.value

    def prox_iterate(self, opts: Optional[Opts] = None) -> IterStats:
        st = IterStats()
        _check(self.lib.gvib200_prox_iterate(self.h, C.byref(opts) if opts is not None else None, C.byref(st)))
        return st

    def switch_to_high_temperature(self):
        _check(self.lib.gvib200_switch_to_high_temperature(self.h))

    def reset_schedule(self):
        _check(self.lib.gvib200_reset_schedule(self.h))

    def timer_start(self):
        _check(self.lib.gvib200_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _check(self.lib.gvib200_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile_begin(self):
        _check(self.lib.gvib200_profile_begin(self.h))

    def profile_end(self) -> dict:
        """{kernel class name: (launches, summed device ms)} for the launches since profile_begin()."""
        pr = Profile()
        _check(self.lib.gvib200_profile_end(self.h, C.byref(pr)))
        return {self.lib.gvib200_kernel_class_name(k).decode(): (int(pr.count[k]), float(pr.ms[k]))
                for k in range(pr.n_classes) if pr.count[k]}

    def info(self) -> Info:
        out = Info()
        _check(self.lib.gvib200_problem_info(self.h, C.byref(out)))
        return out

    def set_option(self, name: str, value: int):
        _check(self.lib.gvib200_problem_set_option(self.h, name.encode(), int(value)))

    def snapshot_save(self):
        _check(self.lib.gvib200_snapshot_save(self.h))

    def snapshot_restore(self):
        _check(self.lib.gvib200_snapshot_restore(self.h))

    def time_stage(self, stage: int, reps: int, opts: Optional[Opts] = None):
        ms = C.c_float()
        nl = C.c_longlong()
        _check(self.lib.gvib200_time_stage(self.h, stage, reps, C.byref(opts) if opts is not None else None,
                                           C.byref(ms), C.byref(nl)))
        return ms.value, nl.value
