"""ctypes front end + build recipe of oracle/cpu/fans_cpu.cpp, the multithreaded C++ restatement of the reference's linear solve
path (std::thread over x-slabs, own FFT).  TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by tests/, by bench.py's cpu_baseline
leg and by `bench.py --impl reference` — never by the product (fans_b200/).  Pinned against fans_oracle.py in tests/test_cpu_port.py."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu", "fans_cpu.cpp")
OUT = os.path.join(HERE, "build")
VARIANTS = {"avx2": ["-mavx2", "-mfma"], "generic": []}


def build(variant=None):
    """g++ -O3 -> oracle/build/libfans_cpu_<variant>.so (git-ignored; travels to the GPU box with the snapshot)"""
    os.makedirs(OUT, exist_ok=True)
    outs = {}
    for v in ([variant] if variant else list(VARIANTS)):
        out = os.path.join(OUT, "libfans_cpu_%s.so" % v)
        if not (os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(SRC)):
            subprocess.run(["g++", "-std=c++17", "-O3", "-fPIC", "-shared", "-pthread", "-fno-math-errno"] + VARIANTS[v] + [SRC, "-o", out], check=True)
        outs[v] = out
    return outs


def _cpu_has_avx2():
    try:
        flags = open("/proc/cpuinfo").read()
        return " avx2" in flags and " fma" in flags
    except OSError:
        return False


_lib = None


def load():
    global _lib
    if _lib is None:
        v = "avx2" if _cpu_has_avx2() else "generic"
        lib = C.CDLL(build(v)[v])
        dp = C.POINTER(C.c_double)
        lib.fcpu_create.restype = C.c_void_p
        lib.fcpu_create.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int]
        lib.fcpu_destroy.argtypes = [C.c_void_p]
        lib.fcpu_threads.argtypes = [C.c_void_p]
        lib.fcpu_set_microstructure.argtypes = [C.c_void_p, C.POINTER(C.c_uint16)]
        lib.fcpu_set_phases.argtypes = [C.c_void_p, C.c_int, dp]
        lib.fcpu_set_reference.argtypes = [C.c_void_p, dp]
        lib.fcpu_get_u.argtypes = [C.c_void_p, dp]
        lib.fcpu_zero_u.argtypes = [C.c_void_p]
        lib.fcpu_convolution.argtypes = [C.c_void_p, dp, dp]
        lib.fcpu_apply_linear.argtypes = [C.c_void_p, dp, dp]
        lib.fcpu_solve_cg.argtypes = [C.c_void_p, dp, C.c_int, C.c_double, C.c_int, dp, dp, dp]
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class CpuSolver:
    """Linear two-phase (or n-phase) CG solve on the host cores.  ms: uint16 [x][y][z]; tangents: per phase n_str x n_str."""

    def __init__(self, ms, L, tangents, reference, howmany=3, threads=0):
        self.lib = load()
        ms = np.ascontiguousarray(ms, dtype=np.uint16)
        self.shape, self.h = ms.shape, howmany
        self.n_str = 3 if howmany == 1 else 6
        Lv = np.asarray(L, dtype=np.float64)
        self.p = self.lib.fcpu_create(ms.shape[0], ms.shape[1], ms.shape[2], _dp(Lv), howmany, threads)
        if not self.p:
            raise ValueError("fcpu_create: grid dimensions must be powers of two >= 4")
        self.lib.fcpu_set_microstructure(self.p, ms.ctypes.data_as(C.POINTER(C.c_uint16)))
        t = np.ascontiguousarray(tangents, dtype=np.float64)
        assert t.shape[1:] == (self.n_str, self.n_str)
        self.lib.fcpu_set_phases(self.p, t.shape[0], _dp(t))
        ref = np.ascontiguousarray(reference, dtype=np.float64)
        self.lib.fcpu_set_reference(self.p, _dp(ref))
        self.threads = self.lib.fcpu_threads(self.p)

    def solve(self, g0, n_it, tol, measure="Linfinity"):
        g = np.ascontiguousarray(g0, dtype=np.float64)
        hist = np.zeros(n_it + 1)
        sig = np.zeros(self.n_str)
        times = np.zeros(2)
        it = self.lib.fcpu_solve_cg(self.p, _dp(g), int(n_it), float(tol), {"L1": 0, "L2": 1, "Linfinity": 2}[measure], _dp(hist), _dp(sig), _dp(times))
        return {"iters": it, "err_all": hist[: it + 1].copy(), "sigma": sig, "loop_s": times[0], "fft_s": times[1]}

    def u(self):
        out = np.empty(self.shape + (self.h,))
        self.lib.fcpu_get_u(self.p, _dp(out))
        return out

    def zero_u(self):
        self.lib.fcpu_zero_u(self.p)

    def convolution(self, field):
        a = np.ascontiguousarray(field, dtype=np.float64)
        out = np.empty_like(a)
        self.lib.fcpu_convolution(self.p, _dp(a), _dp(out))
        return out

    def apply_linear(self, field):
        a = np.ascontiguousarray(field, dtype=np.float64)
        out = np.empty_like(a)
        self.lib.fcpu_apply_linear(self.p, _dp(a), _dp(out))
        return out

    def close(self):
        if self.p:
            self.lib.fcpu_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def elastic_tangent(lam, mu):
    k = np.zeros((6, 6))
    k[:3, :3] = lam
    k += 2.0 * mu * np.eye(6)
    return k


def two_phase_elastic(ms, L, bulk, shear, threads=0):
    """the bench workload: LinearElasticIsotropic phases, reference = (max + min) / 2 of lambda and mu (LinearElastic.h:55-68)"""
    bulk, mu = np.asarray(bulk, dtype=np.float64), np.asarray(shear, dtype=np.float64)
    lam = bulk - 2.0 / 3.0 * mu
    tang = np.stack([elastic_tangent(lam[i], mu[i]) for i in range(len(bulk))])
    return CpuSolver(ms, L, tang, elastic_tangent((lam.max() + lam.min()) / 2, (mu.max() + mu.min()) / 2), 3, threads)
