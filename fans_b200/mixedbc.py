"""Host-side mixed stress/strain boundary conditions for the ctypes front end: the index sets, M = (Q_F^T C0 Q_F)^+ and the per-step
bookkeeping of MixedBCController::activate (include/mixedBCs.h:30-46, 180-226).  The iteration itself (g0 += Q_F M (P_F - Q_F^T Pbar))
runs inside libfans_gpu (fans_update_mixed_bc); this file only prepares the small n_str-sized host data, like
fans_b200/host/mixedbc.hpp does for the C++ front end."""
import numpy as np


class MixedBC:
    def __init__(self, strain_indices, stress_indices, strain_path, stress_path, n_str):
        self.idx_E = [int(i) for i in strain_indices]
        self.idx_F = [int(i) for i in stress_indices]
        if sorted(self.idx_E + self.idx_F) != list(range(n_str)):
            raise ValueError("mixed BC: strain_indices and stress_indices must partition 0..%d" % (n_str - 1))
        self.n_str = n_str
        self.F_E_path = np.asarray(strain_path, dtype=np.float64).reshape(-1, len(self.idx_E)) if self.idx_E else np.zeros((0, 0))
        self.P_F_path = np.asarray(stress_path, dtype=np.float64).reshape(-1, len(self.idx_F)) if self.idx_F else np.zeros((0, 0))
        self.n_steps = max(len(self.F_E_path), len(self.P_F_path))
        self.M = np.zeros((0, 0))

    def finalize(self, C0):
        """M = pinv(Q_F^T C0 Q_F): the rows/columns idx_F of the reference stiffness (mixedBCs.h:42-45)"""
        C0 = np.asarray(C0, dtype=np.float64)
        self.M = np.linalg.pinv(C0[np.ix_(self.idx_F, self.idx_F)]) if self.idx_F else np.zeros((0, 0))


class MixedBCController:
    """activate(ctx, t) before every solve of time step t; the library updates g0 during the iterations."""

    def __init__(self, mbc, kapparef):
        self.mbc = mbc
        mbc.finalize(kapparef)
        self.g0 = self.g0_prev = None

    def activate(self, ctx, t):
        m = self.mbc
        if t == 0:
            self.g0 = np.zeros(m.n_str)
            if m.n_str == 9:
                self.g0[[0, 4, 8]] = 1.0          # F = I (mixedBCs.h:193-199)
            self.g0_prev = self.g0.copy()
        else:
            self.g0 = ctx.get_gradient()          # what the last solve left (stress-controlled components moved)
            delta = self.g0 - self.g0_prev
            self.g0_prev = self.g0.copy()
            for k in m.idx_F:                     # linear extrapolation of the stress-controlled components (mixedBCs.h:205-212)
                self.g0[k] += delta[k]
        for i, k in enumerate(m.idx_E):
            self.g0[k] = m.F_E_path[t, i]
        ctx.set_gradient(self.g0)
        ctx.set_mixed_bc(m.idx_F, m.M, m.P_F_path[t] if m.idx_F else [])
        ctx.update_mixed_bc()                     # one update to adjust the stress-controlled components (mixedBCs.h:224)
