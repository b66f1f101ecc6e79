"""Slab-decomposed (one rank per GPU) parity: the worker runs under torchrun on P GPUs, this process gathers the slabs and
compares them with the oracle (with the oracle's slab emulation of the MAX-norm quirk where it matters).
Needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import subprocess
import sys

import numpy as np
import pytest

import fans_oracle as fo
import golden_util as gu
import util
from util import rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


# (ranks, fused transposes, variant): fused 1 = the y passes push / pull their rows through peer-mapped memory over NVLink (the product
# path, with the component pipeline of solve.cu conv_run_pipelined), 0 = plain ncclSend/ncclRecv block all-to-all (FANS_P2P=0);
# variant "persist": the pipelined y passes loop over their tiles on 5 CTAs, "seq": fused transposes without the pipeline (FANS_PIPE=0)
VARIANT_ENV = {"": {}, "persist": {"FANS_Y_GRID": "5"}, "seq": {"FANS_PIPE": "0"}}


@pytest.fixture(scope="module", params=[(2, 1, ""), (2, 1, "persist"), (2, 1, "seq"), (2, 0, ""), (4, 1, ""), (8, 1, "")],
                ids=lambda p: "P%d-%s%s" % (p[0], "fused" if p[1] else "nccl", "-" + p[2] if p[2] else ""))
def run(request, tmp_path_factory):
    P, fused, variant = request.param
    if _ngpu() < P:
        pytest.skip("needs %d GPUs" % P)
    out = tmp_path_factory.mktemp("mgpu%d_%d%s" % (P, fused, variant))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(P), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + 2 * P + fused + 20 * list(VARIANT_ENV).index(variant)), os.path.join(HERE, "mgpu_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, FANS_P2P=str(fused), **VARIANT_ENV[variant]))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return P, str(out)


def gather(out, case, P, key):
    parts = [np.load(os.path.join(out, "%s_rank%d.npz" % (case, r))) for r in range(P)]
    assert [int(p["x0"]) for p in parts] == sorted(int(p["x0"]) for p in parts)
    return np.concatenate([p[key] for p in parts], axis=0), parts


def test_linear_elastic_cg(run):
    P, out = run
    sol = fo.OracleSolver(gu.sphere32(), [1.0, 1.0, 1.0], "mechanical", util.ELASTIC, "HEX8", "cg", "small", util.EP, 100, n_ranks=P)
    sol.set_gradient([0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001])
    sol.solve()
    u, parts = gather(out, "elastic", P, "u")
    for p in parts:  # every rank reports the same global scalars
        assert abs(int(p["iters"]) - sol.iter) <= 1
        assert rel_err(p["sig"], sol.get_homogenized_stress()) < 1e-9
        n = min(int(p["iters"]), sol.iter)
        assert rel_err(p["err_all"][: n + 1], sol.err_all[: n + 1]) < 1e-6
    assert rel_err(u, sol.u) < 1e-8
    strain, stress, _, _ = sol.strain_stress()
    assert rel_err(gather(out, "elastic", P, "strain")[0].reshape(-1, 6), strain) < 1e-8
    assert rel_err(gather(out, "elastic", P, "stress")[0].reshape(-1, 6), stress) < 1e-8


def test_operators(run):
    P, out = run
    sol = fo.OracleSolver(gu.sphere32(), [1.0, 1.0, 1.0], "mechanical", util.ELASTIC, "HEX8", "cg", "small", util.EP, 0)
    sol.set_gradient([0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001])
    rng = np.random.default_rng(5)
    rfull = rng.standard_normal(sol.u.shape)
    ufull = rng.standard_normal(sol.u.shape) * 1e-3
    assert rel_err(gather(out, "ops", P, "conv")[0], sol.convolution(rfull)) < 1e-8
    assert rel_err(gather(out, "ops", P, "resid")[0], sol.compute_residual(ufull)) < 1e-8
    kd = sol.apply_linear(ufull)
    assert rel_err(gather(out, "ops", P, "kd")[0], kd) < 1e-8
    assert rel_err(gather(out, "ops", P, "kd_sweep")[0], kd) < 1e-8
    _, parts = gather(out, "ops", P, "conv")
    slabs = np.array_split(ufull, P, axis=0)
    for p in parts:
        assert abs(float(p["dot"]) - (ufull * rfull).sum()) <= 1e-12 * np.abs(ufull * rfull).sum()   # SUM over the slabs
        assert abs(float(p["l1"]) - max(np.abs(s).sum() for s in slabs)) <= 1e-12 * np.abs(ufull).sum()  # MAX (solver.h:430)
        assert abs(float(p["l2"]) - max(np.sqrt((s * s).sum()) for s in slabs)) <= 1e-12
        assert float(p["linf"]) == np.abs(ufull).max()


def test_thermal_l2_max_norm_quirk(run):
    P, out = run
    ep = {"measure": "L2", "type": "relative", "tolerance": 1e-8}
    sol = fo.OracleSolver(gu.sphere32(), [1.0, 1.0, 1.0], "thermal", util.THERMAL, "HEX8R", "cg", "small", ep, 100, n_ranks=P)
    sol.set_gradient([0.01, 0.02, -0.01])
    sol.solve()
    u, parts = gather(out, "thermal", P, "u")
    for p in parts:
        assert abs(int(p["iters"]) - sol.iter) <= 1
        n = min(int(p["iters"]), sol.iter)
        assert rel_err(p["err_all"][: n + 1], sol.err_all[: n + 1]) < 1e-6   # per-slab maxima, not the global L2 norm
        assert rel_err(p["sig"], sol.get_homogenized_stress()) < 1e-9


def test_j2_plasticity(run):
    P, out = run
    mats = [{"phases": [0], "matmodel": "J2ViscoPlastic_LinearIsotropicHardening",
             "material_properties": {"bulk_modulus": [62.5], "shear_modulus": [28.8462], "yield_stress": [0.1],
                                     "isotropic_hardening_parameter": [3.0], "kinematic_hardening_parameter": [2.0], "viscosity": [1.0],
                                     "time_step": 0.01}},
            {"phases": [1], "matmodel": "LinearElasticIsotropic", "material_properties": {"bulk_modulus": [222.222], "shear_modulus": [166.6667]}}]
    ms = util.two_phase_ms(0, 11, (16, 8, 32))
    if ms.shape[0] // 4 < P:
        pytest.skip("n_x/4 < world_size (reader.cpp:306)")
    sol = fo.OracleSolver(ms, [1.0, 1.5, 2.0], "mechanical", mats, "HEX8", "cg", "small", util.EP, 200, n_ranks=P)
    for t, g in enumerate([[0.002, -0.001, -0.001, 0.0005, 0, 0], [0.004, -0.002, -0.002, 0.001, 0, 0]]):
        sol.set_gradient(g)
        sol.solve()
        sig = sol.get_homogenized_stress()
        u, parts = gather(out, "j2", P, "u%d" % t)
        for p in parts:
            assert abs(int(p["iters%d" % t]) - sol.iter) <= 1
            assert rel_err(p["sig%d" % t], sig) < 1e-9
        assert rel_err(u, sol.u) < 1e-8
        ep = gather(out, "j2", P, "ep%d" % t)[0]
        assert rel_err(ep, sol.models[0].ep_t.mean(1).reshape(ep.shape)) < 1e-8  # J2Plasticity.h:245-322: GP mean of the committed values
        sol.extrapolate_displacement()


def test_neohooke_mixed_bc(run):
    """BASELINE config 5's path on slabs: finite-strain Neo-Hooke, stress-controlled components through the mixed-BC update (its
    homogenized-stress sweep is all-reduced over the ranks), CG with line search; two load steps with displacement extrapolation"""
    P, out = run
    ms = util.two_phase_ms(0, 12, (16, 16, 16))
    if ms.shape[0] // 4 < P:
        pytest.skip("n_x/4 < world_size (reader.cpp:306)")
    ro, _ = fo.run_load_cases(ms, util.NH_MIXED_CFG, on_step=lambda sol, lc, t, res: res.update(u=sol.u.copy(), sig=sol.get_homogenized_stress()))
    for t, b in enumerate(ro[0]):
        u, parts = gather(out, "nhmixed", P, "u%d" % t)
        for p in parts:
            assert abs(int(p["iters%d" % t]) - b["iters"]) <= 1
            assert rel_err(p["sig%d" % t], b["sig"]) < 1e-9
            assert rel_err(p["g0_%d" % t], b["g0"]) < 1e-9
        assert rel_err(u, b["u"]) < 1e-8
