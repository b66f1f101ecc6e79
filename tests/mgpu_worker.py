"""Worker of tests/test_multi_gpu.py: run under `python -m torch.distributed.run --nproc-per-node P`, one rank per GPU.
Every rank drives its slab through the C ABI and writes its results to <outdir>/<case>_rank<r>.npz; the parent test
compares the gathered slabs with the oracle."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import fans_oracle as fo  # noqa: E402  (parameter parsing only: descriptors, kapparef)
import golden_util as gu  # noqa: E402
import util  # noqa: E402
from fans_b200 import _lib as L, dist as fdist  # noqa: E402


def slab_ctx(par, comm):
    x0, n0 = fdist.slab(par.nx, comm.world_size, comm.rank)
    ctx = L.Context((par.nx, par.ny, par.nz), par.L, par.h, par.n_str, par.FE_type, comm=comm if comm.world_size > 1 else None)
    ctx.set_materials(util.phase_descs_from_oracle(par))
    ctx.set_microstructure(par.ms[x0:x0 + n0])
    ctx.set_reference_stiffness(par.kapparef)
    return ctx, x0, n0


def main():
    outdir = sys.argv[1]
    comm = fdist.init()
    r = comm.rank
    ep = {"measure": "Linfinity", "type": "absolute", "tolerance": 1e-10}

    # ---- 1. linear elastic CG on sphere32 (stencil fast path, halo both ways, transposes, scalar all-reduces)
    par = fo.OracleSolver(gu.sphere32(), [1.0, 1.0, 1.0], "mechanical", util.ELASTIC, "HEX8", "cg", "small", ep, 0)
    ctx, x0, n0 = slab_ctx(par, comm)
    g0 = [0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001]
    ctx.set_gradient(g0)
    res = ctx.solve("cg", 100, 1e-10, "Linfinity", "absolute")
    sig = ctx.homogenized_stress()
    strain, stress = ctx.strain_stress()
    np.savez(os.path.join(outdir, "elastic_rank%d.npz" % r), u=ctx.download("u"), iters=res["iters"], err_all=res["err_all"], sig=sig,
             strain=strain, stress=stress, x0=x0)
    # ---- 2. operators on random fields: convolution, residual, K.d (both the stencil and the sweep form), norms, dot
    rng = np.random.default_rng(5)
    rfull = rng.standard_normal((par.nx, par.ny, par.nz, par.h))
    ufull = rng.standard_normal((par.nx, par.ny, par.nz, par.h)) * 1e-3
    ctx.upload("r", rfull[x0:x0 + n0])
    ctx.convolution("r", "s")
    conv = ctx.download("s")
    ctx.upload("u", ufull[x0:x0 + n0])
    ctx.residual("r", "u")
    resid = ctx.download("r")
    ctx.apply_linear("rnew", "u")
    kd = ctx.download("rnew")
    os.environ["FANS_LINEAR_SWEEP"] = "1"
    ctx.apply_linear("rnew", "u")
    kd_sweep = ctx.download("rnew")
    del os.environ["FANS_LINEAR_SWEEP"]
    ctx.upload("s", rfull[x0:x0 + n0])
    np.savez(os.path.join(outdir, "ops_rank%d.npz" % r), conv=conv, resid=resid, kd=kd, kd_sweep=kd_sweep, dot=ctx.dot("u", "s"),
             l1=ctx.norm("u", "L1"), l2=ctx.norm("u", "L2"), linf=ctx.norm("u", "Linfinity"), x0=x0)
    ctx.close()

    # ---- 3. thermal, L2 error measure: the MAX-over-ranks quirk (solver.h:430) decides the iteration count
    ep2 = {"measure": "L2", "type": "relative", "tolerance": 1e-8}
    par = fo.OracleSolver(gu.sphere32(), [1.0, 1.0, 1.0], "thermal", util.THERMAL, "HEX8R", "cg", "small", ep2, 0)
    ctx, x0, n0 = slab_ctx(par, comm)
    ctx.set_gradient([0.01, 0.02, -0.01])
    res = ctx.solve("cg", 100, 1e-8, "L2", "relative")
    np.savez(os.path.join(outdir, "thermal_rank%d.npz" % r), u=ctx.download("u"), iters=res["iters"], err_all=res["err_all"],
             sig=ctx.homogenized_stress(), x0=x0)
    ctx.close()

    # ---- 4. J2 plasticity, 3 load steps into the plastic regime (nonlinear sweep with history, line search, halo add)
    mats = [{"phases": [0], "matmodel": "J2ViscoPlastic_LinearIsotropicHardening",
             "material_properties": {"bulk_modulus": [62.5], "shear_modulus": [28.8462], "yield_stress": [0.1],
                                     "isotropic_hardening_parameter": [3.0], "kinematic_hardening_parameter": [2.0], "viscosity": [1.0],
                                     "time_step": 0.01}},
            {"phases": [1], "matmodel": "LinearElasticIsotropic", "material_properties": {"bulk_modulus": [222.222], "shear_modulus": [166.6667]}}]
    ms = util.two_phase_ms(0, 11, (16, 8, 32))
    par = fo.OracleSolver(ms, [1.0, 1.5, 2.0], "mechanical", mats, "HEX8", "cg", "small", ep, 0)
    if par.nx // 4 < comm.world_size:  # reader.cpp:306
        comm.close()
        return
    ctx, x0, n0 = slab_ctx(par, comm)
    out = {"x0": x0}
    for t, g in enumerate([[0.002, -0.001, -0.001, 0.0005, 0, 0], [0.004, -0.002, -0.002, 0.001, 0, 0]]):
        ctx.set_gradient(g)
        res = ctx.solve("cg", 200, 1e-10, "Linfinity", "absolute")
        out["iters%d" % t] = res["iters"]
        out["sig%d" % t] = ctx.homogenized_stress()
        out["u%d" % t] = ctx.download("u")
        out["ep%d" % t] = ctx.get_field("plastic_strain")
        ctx.extrapolate_displacement()
    np.savez(os.path.join(outdir, "j2_rank%d.npz" % r), **out)
    ctx.close()

    # ---- 5. compressible Neo-Hooke with mixed boundary conditions (BASELINE config 5's path): finite-strain sweep, sigma-bar sweep +
    #         all-reduce inside every mixed-BC update, line search
    ms = util.two_phase_ms(0, 12, (16, 16, 16))
    out = {}

    def make_ctx(par):
        c, x0, n0 = slab_ctx(par, comm)
        out["x0"] = x0
        return c

    def on_step(c, lc, t, res):
        out["iters%d" % t], out["sig%d" % t], out["u%d" % t], out["g0_%d" % t] = res["iters"], c.homogenized_stress(), c.download("u"), res["g0"]

    _, ctx = util.run_gpu_load_cases(ms, util.NH_MIXED_CFG, on_step=on_step, make_ctx=make_ctx)
    np.savez(os.path.join(outdir, "nhmixed_rank%d.npz" % r), **out)
    ctx.close()
    comm.close()


if __name__ == "__main__":
    main()
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
