"""CPU checks of the Python host helpers above the C ABI against the oracle's restatement of the reference: mixed-BC bookkeeping
(include/mixedBCs.h:30-46) and the closed-form finite-strain reference medium (LargeStrainMechModel.h:105-180 at F = I)."""
import numpy as np

import fans_oracle as fo
from fans_b200 import mixedbc, simple


def test_spatial_tangent_closed_form():
    lam, mu = 40.0, 28.0
    C = np.zeros((6, 6))
    C[:3, :3] = lam
    C += 2 * mu * np.eye(6)
    A = fo.compute_spatial_tangent(np.eye(3), np.zeros((3, 3)), C)
    assert np.abs(A - simple.spatial_tangent_at_identity(lam, mu)).max() < 1e-13


def test_mixed_bc_matches_oracle():
    rng = np.random.default_rng(0)
    C0 = rng.standard_normal((9, 9))
    C0 = C0 @ C0.T + 9 * np.eye(9)
    idx_E, idx_F = [1, 2, 3, 5, 6, 7, 8], [0, 4]
    strain = [[0, 0, 0, 0, 0, 0, 1.1], [0, 0, 0, 0, 0, 0, 1.2]]
    stress = [[0.0, 0.0], [0.1, 0.0]]
    a = mixedbc.MixedBC(idx_E, idx_F, strain, stress, 9)
    a.finalize(C0)
    b = fo.MixedBC(idx_E, idx_F, strain, stress, 9)
    b.finalize(C0)
    assert np.abs(a.M - b.M).max() < 1e-12 * np.abs(b.M).max()
    assert np.array_equal(a.F_E_path, b.F_E_path) and np.array_equal(a.P_F_path, b.P_F_path) and a.n_steps == b.n_steps == 2
