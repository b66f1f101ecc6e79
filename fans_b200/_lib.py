"""ctypes binding of libfans_gpu.so (include/fans_gpu.h).  Pure plumbing: no numerics live here.

The library is built in-tree by `__graft_entry__.build()` (fans_b200/csrc/Makefile) into fans_b200/lib/.
There is no CPU fallback: every compute entry point fails loudly when the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FANS_GPU_LIB") or os.path.join(_HERE, "lib", "libfans_gpu.so")  # FANS_GPU_LIB: another build of the same library (kernel A/B runs)

FANS_MAX_PARAMS = 84

# enums (include/fans_gpu.h)
FE = {"HEX8": 0, "HEX8R": 1, "BBAR": 2}
FIELD = {"u": 0, "r": 1, "s": 2, "d": 3, "rnew": 4, "u_prev": 5}
MEASURE = {"L1": 0, "L2": 1, "Linfinity": 2}
ERRTYPE = {"absolute": 0, "relative": 1}
METHOD = {"cg": 0, "fp": 1}
MAT_LINEAR, MAT_PP_LIN, MAT_PP_NONLIN, MAT_J2_LIN, MAT_J2_NONLIN, MAT_J2NEW, MAT_SVK, MAT_NEOHOOKE = range(8)


class PhaseDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("local_mat", C.c_int32), ("group_n_mat", C.c_int32), ("reserved", C.c_int32),
                ("params", C.c_double * FANS_MAX_PARAMS)]


class Config(C.Structure):
    _fields_ = [("dims", C.c_int32 * 3), ("L", C.c_double * 3), ("howmany", C.c_int32), ("n_str", C.c_int32),
                ("fe_type", C.c_int32), ("world_size", C.c_int32), ("world_rank", C.c_int32), ("local_n0", C.c_int32),
                ("local_0_start", C.c_int32), ("local_n1", C.c_int32), ("local_1_start", C.c_int32), ("device", C.c_int32),
                ("nccl_comm", C.c_void_p), ("stream", C.c_void_p)]


class MixedBCDesc(C.Structure):
    _fields_ = [("n_F", C.c_int32), ("idx_F", C.c_int32 * 9), ("M", C.c_double * 81), ("P_target", C.c_double * 9)]


class SolveParams(C.Structure):
    _fields_ = [("method", C.c_int32), ("n_it", C.c_int32), ("tol", C.c_double), ("measure", C.c_int32),
                ("err_type", C.c_int32), ("ls_max_iter", C.c_int32), ("ls_tol", C.c_double), ("verbose", C.c_int32),
                ("force_nonlinear", C.c_int32)]


class SolveResult(C.Structure):
    _fields_ = [("iters", C.c_int32), ("n_residual_evals", C.c_int32), ("err_last", C.c_double),
                ("elapsed_ms", C.c_double), ("fft_ms", C.c_double), ("loop_ms", C.c_double)]


# every symbol include/fans_gpu.h declares (tests check that the library exports all of them)
EXPORTS = [
    "fans_create", "fans_destroy", "fans_last_error", "fans_version", "fans_set_microstructure", "fans_set_materials",
    "fans_set_reference_stiffness", "fans_set_gradient", "fans_get_gradient", "fans_set_mixed_bc", "fans_update_mixed_bc",
    "fans_field_upload", "fans_field_download", "fans_field_zero", "fans_field_copy", "fans_residual", "fans_apply_linear",
    "fans_convolution", "fans_dot", "fans_axpy", "fans_norm", "fans_solve", "fans_homogenized_stress", "fans_commit_history",
    "fans_extrapolate_displacement", "fans_get_field", "fans_strain_stress", "fans_strain_stress_gp", "fans_launch_count", "fans_set_profiling", "fans_get_profile", "fans_comm_unique_id", "fans_comm_create",
    "fans_comm_destroy", "fans_allreduce_sum", "fans_solve_batch", "fans_batch_load_displacement", "fans_batch_release",
]

_lib = None


class FansError(RuntimeError):
    pass


def load():
    """Load libfans_gpu.so (once) and declare prototypes. Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FansError("libfans_gpu.so not found at %s: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    dp = C.POINTER(C.c_double)
    lib.fans_create.argtypes = [C.POINTER(P), C.POINTER(Config)]
    lib.fans_destroy.argtypes = [P]
    lib.fans_destroy.restype = None
    lib.fans_last_error.argtypes = [P]
    lib.fans_last_error.restype = C.c_char_p
    lib.fans_set_microstructure.argtypes = [P, C.POINTER(C.c_uint16)]
    lib.fans_set_materials.argtypes = [P, C.c_int32, C.POINTER(PhaseDesc)]
    lib.fans_set_reference_stiffness.argtypes = [P, dp]
    lib.fans_set_gradient.argtypes = [P, dp]
    lib.fans_get_gradient.argtypes = [P, dp]
    lib.fans_set_mixed_bc.argtypes = [P, C.POINTER(MixedBCDesc)]
    lib.fans_update_mixed_bc.argtypes = [P]
    lib.fans_field_upload.argtypes = [P, C.c_int32, dp]
    lib.fans_field_download.argtypes = [P, C.c_int32, dp]
    lib.fans_field_zero.argtypes = [P, C.c_int32]
    lib.fans_field_copy.argtypes = [P, C.c_int32, C.c_int32]
    lib.fans_residual.argtypes = [P, C.c_int32, C.c_int32]
    lib.fans_apply_linear.argtypes = [P, C.c_int32, C.c_int32]
    lib.fans_convolution.argtypes = [P, C.c_int32, C.c_int32]
    lib.fans_dot.argtypes = [P, C.c_int32, C.c_int32, dp]
    lib.fans_axpy.argtypes = [P, C.c_int32, C.c_double, C.c_int32]
    lib.fans_norm.argtypes = [P, C.c_int32, C.c_int32, dp]
    lib.fans_solve.argtypes = [P, C.POINTER(SolveParams), C.POINTER(SolveResult), dp]
    lib.fans_homogenized_stress.argtypes = [P, dp]
    lib.fans_solve_batch.argtypes = [P, C.c_int32, dp, C.POINTER(SolveParams), C.POINTER(SolveResult), dp, dp]
    lib.fans_batch_load_displacement.argtypes = [P, C.c_int32, C.c_int32]
    lib.fans_batch_release.argtypes = [P]
    lib.fans_commit_history.argtypes = [P]
    lib.fans_extrapolate_displacement.argtypes = [P]
    lib.fans_get_field.argtypes = [P, C.c_char_p, C.c_void_p, C.c_size_t]
    lib.fans_strain_stress.argtypes = [P, dp, dp]
    lib.fans_strain_stress_gp.argtypes = [P, dp, dp, dp, dp]
    lib.fans_set_profiling.argtypes = [P, C.c_int32]
    lib.fans_get_profile.argtypes = [P, C.c_int32, C.POINTER(C.c_char_p), dp, C.POINTER(C.c_int64)]
    lib.fans_comm_unique_id.argtypes = [C.c_void_p]
    lib.fans_comm_create.argtypes = [C.POINTER(P), C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
    lib.fans_comm_destroy.argtypes = [P]
    lib.fans_allreduce_sum.argtypes = [P, dp, C.c_int32]
    lib.fans_launch_count.argtypes = [P]
    lib.fans_launch_count.restype = C.c_int64
    _lib = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """One fans_ctx (= one Solver instance on one GPU). Thin, argument-checking wrapper over the C ABI."""

    def __init__(self, dims, L, howmany, n_str, fe_type="HEX8", device=-1, comm=None):
        """dims: GLOBAL grid. comm: fans_b200.dist.SlabComm for world_size > 1 (this rank then owns dims[0]/P x-planes;
        every field / microstructure buffer passed to this object is the rank's slab [n_x/P][n_y][n_z])."""
        self.lib = load()
        self.gdims = tuple(int(d) for d in dims)
        P_, r_ = (comm.world_size, comm.rank) if comm is not None else (1, 0)
        self.dims = (self.gdims[0] // P_,) + self.gdims[1:]
        self.h, self.n_str = int(howmany), int(n_str)
        cfg = Config()
        cfg.dims[:] = self.gdims
        cfg.L[:] = [float(x) for x in L]
        cfg.howmany, cfg.n_str = self.h, self.n_str
        if fe_type not in FE:
            raise FansError("Unknown FE_type: '%s'. Supported types: HEX8, HEX8R, BBAR" % fe_type)
        cfg.fe_type = FE[fe_type]
        cfg.world_size, cfg.world_rank = P_, r_
        cfg.local_n0, cfg.local_0_start = self.gdims[0] // P_, r_ * (self.gdims[0] // P_)
        cfg.local_n1, cfg.local_1_start = self.gdims[1] // P_, r_ * (self.gdims[1] // P_)
        cfg.device = device
        cfg.nccl_comm = comm.handle if comm is not None else None
        self.comm = comm
        cfg.stream = None
        self.n_gp = 1 if fe_type == "HEX8R" else 8
        self.ptr = C.c_void_p()
        rc = self.lib.fans_create(C.byref(self.ptr), C.byref(cfg))
        if rc != 0:
            msg = self.lib.fans_last_error(None).decode()
            self.ptr = None
            raise FansError("fans_create failed (%d): %s" % (rc, msg))

    def close(self):
        if getattr(self, "ptr", None):
            self.lib.fans_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FansError("libfans_gpu error %d: %s" % (rc, self.lib.fans_last_error(self.ptr).decode()))

    @property
    def field_shape(self):
        return self.dims + (self.h,)

    # ---- problem data
    def set_microstructure(self, ms):
        ms = np.ascontiguousarray(ms, dtype=np.uint16)
        assert ms.shape == self.dims, (ms.shape, self.dims)
        self._ck(self.lib.fans_set_microstructure(self.ptr, ms.ctypes.data_as(C.POINTER(C.c_uint16))))

    def set_materials(self, descs):
        arr = (PhaseDesc * len(descs))(*descs)
        self._ck(self.lib.fans_set_materials(self.ptr, len(descs), arr))

    def set_reference_stiffness(self, kapparef):
        k = np.ascontiguousarray(kapparef, dtype=np.float64)
        assert k.shape == (self.n_str, self.n_str)
        self._ck(self.lib.fans_set_reference_stiffness(self.ptr, _dptr(k)))

    def set_gradient(self, g0):
        g = np.ascontiguousarray(g0, dtype=np.float64)
        assert g.shape == (self.n_str,)
        self._ck(self.lib.fans_set_gradient(self.ptr, _dptr(g)))

    def get_gradient(self):
        g = np.zeros(self.n_str)
        self._ck(self.lib.fans_get_gradient(self.ptr, _dptr(g)))
        return g

    def set_mixed_bc(self, idx_F=None, M=None, P_target=None):
        if idx_F is None:
            self._ck(self.lib.fans_set_mixed_bc(self.ptr, None))
            return
        d = MixedBCDesc()
        d.n_F = len(idx_F)
        for i, k in enumerate(idx_F):
            d.idx_F[i] = int(k)
            d.P_target[i] = float(P_target[i])
        Mf = np.asarray(M, dtype=np.float64).reshape(-1)
        for i, v in enumerate(Mf):
            d.M[i] = v
        self._ck(self.lib.fans_set_mixed_bc(self.ptr, C.byref(d)))

    def update_mixed_bc(self):
        self._ck(self.lib.fans_update_mixed_bc(self.ptr))

    # ---- fields
    def upload(self, field, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.field_shape, (a.shape, self.field_shape)
        self._ck(self.lib.fans_field_upload(self.ptr, FIELD[field], _dptr(a)))

    def download(self, field):
        a = np.empty(self.field_shape, dtype=np.float64)
        self._ck(self.lib.fans_field_download(self.ptr, FIELD[field], _dptr(a)))
        return a

    def download_into(self, field, a):
        """download into a caller-owned (e.g. pinned) C-contiguous float64 array of field_shape"""
        assert a.shape == self.field_shape and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        self._ck(self.lib.fans_field_download(self.ptr, FIELD[field], _dptr(a)))
        return a

    def zero(self, field):
        self._ck(self.lib.fans_field_zero(self.ptr, FIELD[field]))

    def copy(self, dst, src):
        self._ck(self.lib.fans_field_copy(self.ptr, FIELD[dst], FIELD[src]))

    # ---- operators
    def residual(self, out, u):
        self._ck(self.lib.fans_residual(self.ptr, FIELD[out], FIELD[u]))

    def apply_linear(self, out, d):
        self._ck(self.lib.fans_apply_linear(self.ptr, FIELD[out], FIELD[d]))

    def convolution(self, fin, fout):
        self._ck(self.lib.fans_convolution(self.ptr, FIELD[fin], FIELD[fout]))

    def dot(self, a, b):
        v = C.c_double()
        self._ck(self.lib.fans_dot(self.ptr, FIELD[a], FIELD[b], C.byref(v)))
        return v.value

    def axpy(self, y, alpha, x):
        self._ck(self.lib.fans_axpy(self.ptr, FIELD[y], float(alpha), FIELD[x]))

    def norm(self, field, measure="Linfinity"):
        if measure not in MEASURE:
            raise FansError("Unknown measure type: " + str(measure))
        v = C.c_double()
        self._ck(self.lib.fans_norm(self.ptr, FIELD[field], MEASURE[measure], C.byref(v)))
        return v.value

    # ---- drivers
    def solve(self, method="cg", n_it=100, tol=1e-10, measure="Linfinity", err_type="absolute", ls_max_iter=5, ls_tol=1e-2,
              verbose=False, force_nonlinear=False):
        if method not in METHOD:
            raise FansError(str(method) + " is not a valid method")
        if measure not in MEASURE:
            raise FansError("Unknown measure type: " + str(measure))
        if err_type not in ERRTYPE:
            raise FansError("Unknown error type: " + str(err_type))
        p = SolveParams(METHOD[method], int(n_it), float(tol), MEASURE[measure], ERRTYPE[err_type], int(ls_max_iter),
                        float(ls_tol), int(bool(verbose)), int(bool(force_nonlinear)))
        r = SolveResult()
        hist = np.zeros(int(n_it) + 1)
        self._ck(self.lib.fans_solve(self.ptr, C.byref(p), C.byref(r), _dptr(hist)))
        return {"iters": r.iters, "n_residual_evals": r.n_residual_evals, "err_last": r.err_last, "elapsed_ms": r.elapsed_ms,
                "loop_ms": r.loop_ms, "fft_ms": r.fft_ms, "err_all": hist[: r.iters + 1].copy()}

    def solve_batch(self, macro, n_it=100, tol=1e-10, measure="Linfinity", err_type="absolute", verbose=False):
        """n_b linear load cases as lanes of ONE CG loop (fans_solve_batch; the loop of get_homogenized_tangent, solver.h:762-775).
        macro: [n_b][n_str].  Returns (list of per-lane result dicts, homogenized stresses [n_b][n_str])."""
        macro = np.ascontiguousarray(macro, dtype=np.float64).reshape(-1, self.n_str)
        nb = macro.shape[0]
        if measure not in MEASURE:
            raise FansError("Unknown measure type: " + str(measure))
        if err_type not in ERRTYPE:
            raise FansError("Unknown error type: " + str(err_type))
        p = SolveParams(METHOD["cg"], int(n_it), float(tol), MEASURE[measure], ERRTYPE[err_type], 5, 1e-2, int(bool(verbose)), 0)
        res = (SolveResult * nb)()
        hist = np.zeros((nb, int(n_it) + 1))
        stress = np.zeros((nb, self.n_str))
        self._ck(self.lib.fans_solve_batch(self.ptr, nb, _dptr(macro), C.byref(p), res, _dptr(stress), _dptr(hist)))
        out = [{"iters": r.iters, "n_residual_evals": r.n_residual_evals, "err_last": r.err_last, "elapsed_ms": r.elapsed_ms,
                "loop_ms": r.loop_ms, "fft_ms": r.fft_ms, "err_all": hist[i, : r.iters + 1].copy()} for i, r in enumerate(res)]
        return out, stress

    def batch_displacement(self, lane, field="u"):
        """copies the displacement of one lane of the last solve_batch into a field of the context (default: u)"""
        self._ck(self.lib.fans_batch_load_displacement(self.ptr, int(lane), FIELD[field] if isinstance(field, str) else int(field)))

    def batch_release(self):
        self._ck(self.lib.fans_batch_release(self.ptr))

    def homogenized_stress(self):
        out = np.zeros(self.n_str)
        self._ck(self.lib.fans_homogenized_stress(self.ptr, _dptr(out)))
        return out

    def commit_history(self):
        self._ck(self.lib.fans_commit_history(self.ptr))

    def extrapolate_displacement(self):
        self._ck(self.lib.fans_extrapolate_displacement(self.ptr))

    def strain_stress(self):
        """(strain, stress) element averages from ONE sweep, each [x][y][z][n_str]."""
        e = np.empty(self.dims + (self.n_str,))
        s = np.empty(self.dims + (self.n_str,))
        self._ck(self.lib.fans_strain_stress(self.ptr, _dptr(e), _dptr(s)))
        return e, s

    def strain_stress_gp(self):
        """(strain, stress, strain_gp, stress_gp) from ONE sweep: element averages and all Gauss-point values [x][y][z][n_gp][n_str]"""
        e = np.empty(self.dims + (self.n_str,))
        s = np.empty(self.dims + (self.n_str,))
        eg = np.empty(self.dims + (self.n_gp, self.n_str))
        sg = np.empty(self.dims + (self.n_gp, self.n_str))
        self._ck(self.lib.fans_strain_stress_gp(self.ptr, _dptr(e), _dptr(s), _dptr(eg), _dptr(sg)))
        return e, s, eg, sg

    def get_field(self, name):
        nx, ny, nz = self.dims
        if name in ("strain", "stress"):
            out = np.empty((nx, ny, nz, self.n_str))
        elif name in ("strain_gp", "stress_gp"):
            out = np.empty((nx, ny, nz, self.n_gp, self.n_str))
        elif name == "plastic_flag":
            out = np.empty((nx, ny, nz), dtype=np.float32)
        elif name in ("plastic_strain", "kinematic_hardening_variable"):
            out = np.empty((nx, ny, nz, 6))
        elif name == "isotropic_hardening_variable":
            out = np.empty((nx, ny, nz))
        elif name in ("plastic_strain_gp", "kinematic_hardening_variable_gp"):   # every Gauss point, J2Plasticity.h:298-307
            out = np.empty((nx, ny, nz, self.n_gp, 6))
        elif name == "isotropic_hardening_variable_gp":
            out = np.empty((nx, ny, nz, self.n_gp))
        elif name == "fundamental_solution":  # global frequency grid; only this rank's y rows are filled when world_size > 1
            out = np.zeros((ny, self.gdims[0], nz // 2 + 1, self.h * (self.h + 1) // 2))
        else:
            raise FansError("unknown field " + name)
        self._ck(self.lib.fans_get_field(self.ptr, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def set_profiling(self, on=True):
        self._ck(self.lib.fans_set_profiling(self.ptr, int(bool(on))))

    def profile(self):
        """{kernel class: (total device ms, launches)} since set_profiling(True)"""
        out = {}
        for cls in range(16):
            name, ms, n = C.c_char_p(), C.c_double(), C.c_int64()
            self._ck(self.lib.fans_get_profile(self.ptr, cls, C.byref(name), C.byref(ms), C.byref(n)))
            if name.value and n.value:
                out[name.value.decode()] = (ms.value, n.value)
        return out

    def launch_count(self):
        return int(self.lib.fans_launch_count(self.ptr))
