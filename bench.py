#!/usr/bin/env python
"""bench.py — headline benchmark of the FANS per-iteration solve loop on B200.

Metric (BASELINE.json): voxel-DOF updates/s (and CG iterations/s) of a two-phase linear-elastic CG solve (HEX8, spherical
inclusion) on 512^3 voxels PER GPU, with the fraction of the measured HBM roofline.  A "step" is ONE CG iteration
(convolution: 5 FFT passes with the fused Green operator; fused direction update + K.d stencil; fused r/u update + norms).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size n] [--impl ours|reference] [--workload NAME]

--workload (default elastic512 = the headline, BASELINE.json metric) selects the other BASELINE.json configs, each printed as one JSON
line of the same shape: config2 (256^3 sphere, 6 unit load cases solved to 1e-10), config3 (J2 plasticity fibre image, load
increments), config4 (Voronoi "polycrystal-like" image, the headline protocol on it), config5 (Neo-Hooke, mixed BCs, --method cg|fp).

N > 1 (launched by torchrun, one rank per GPU): the grid grows with N (weak scaling, 512^3 voxels per GPU):
N=2: 1024x512x512, N=4: 1024x1024x512, N=8: 1024^3 (BASELINE config 4's grid), decomposed into x-slabs like the reference.
One JSON line on stdout (rank 0).  `--impl reference` times the CPU restatement of the reference algorithm (oracle/) on a
bounded sample on the host cores (the reference itself needs MPI/FFTW/HDF5/Eigen and cannot be built in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_BULK, G_SHEAR = [62.5, 222.222], [28.8462, 166.6667]   # test_LinearElastic.json / SURVEY 8d config 2
G0 = [0.001, -0.002, 0.003, 0.0015, -0.0025, 0.001]
BYTES_PER_VOXEL_ITER_H3 = 554.0  # 23 F + 2 N with F = 24 B/voxel
# algorithmic bytes per voxel and launch of the seven passes (h = 3, F = 24 B/voxel)
ALG_BYTES_PER_VOXEL = {"fft_z_fwd": 48.0, "fft_y_fwd": 48.0, "fft_x_gamma": 72.0, "fft_y_inv": 48.0, "fft_z_inv": 72.0,
                       "sweep_linear": 98.0, "cg_update": 168.0}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu=0):
        self.rows, self.stop, self.gpu = [], False, gpu
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        import statistics
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def host_memory_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return float(ln.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def cpu_port_run(dims, iters, warm=0, threads=0):
    """The CPU arm: oracle/cpu/fans_cpu.cpp, a multithreaded C++ restatement of the reference's linear solve loop (std::thread over
    x-slabs, own FFT; the reference itself needs MPI/FFTW/HDF5/Eigen and cannot be built in this image) on the SAME workload: two-phase
    ellipsoid, LinearElasticIsotropic, HEX8, CG, exactly `iters` iterations (tol = 0).  Returns rate, iterations, loop seconds, threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fans_cpu
    from fans_b200 import simple   # microstructure generator only (NumPy)
    ms = simple.ellipsoid_microstructure(dims)
    cpu = fans_cpu.two_phase_elastic(ms, [1.0, 1.0, 1.0], K_BULK, G_SHEAR, threads)
    if warm:
        cpu.solve(G0, warm, 0.0)
        cpu.zero_u()
    r = cpu.solve(G0, iters, 0.0)
    nthreads = cpu.threads
    cpu.close()
    dof = 3.0 * dims[0] * dims[1] * dims[2]
    return dof * r["iters"] / r["loop_s"], r["iters"], r["loop_s"], nthreads, r["fft_s"]


def cpu_grid(full_dims, budget_s, iters):
    """largest power-of-two cube <= the workload's grid that the host can hold (about 185 B/voxel) and finish within budget_s, from a
    measured 64^3 probe of this machine"""
    probe = [64, 64, 64]
    rate, _, _, _, _ = cpu_port_run(probe, 2)
    n = min(full_dims)
    mem = host_memory_gb()
    while n > 64 and (185.0 * n ** 3 / 1e9 > 0.5 * mem or 3.0 * n ** 3 * iters / rate > budget_s):
        n //= 2
    return [n, n, n] if n < min(full_dims) else list(full_dims)


def cpu_baseline_object(args):
    dims = cpu_grid([args.size] * 3, 25.0, 3) if not args.cpu_size else [args.cpu_size] * 3
    probe, _, _, _, _ = cpu_port_run(dims, 1)
    iters = int(min(30, max(3, 12.0 * probe / (3.0 * dims[0] * dims[1] * dims[2]))))   # about 10-15 s of CPU work
    rate, it, dt, nthreads, fft = cpu_port_run(dims, iters)
    return {"value": rate, "unit": "voxel-DOF/s", "cores": nthreads, "kind": "port",
            "sample": "%d CG iterations of the same workload on a %dx%dx%d grid, %.1f s (%.0f %% FFT): multithreaded C++ restatement of the "
                      "reference loop (oracle/cpu/fans_cpu.cpp, std::thread over x-slabs, own FFT), not the FANS binary"
                      % (it, dims[0], dims[1], dims[2], dt, 100.0 * fft / dt)}


# BASELINE.json `configs` (SURVEY.md 8d gives the concrete inputs); aliases map onto the canonical names
WORKLOADS = {"elastic512": "elastic512", "config2": "config2", "elastic256x6": "config2", "config3": "config3", "j2_fibre": "config3",
             "config4": "config4", "voronoi": "config4", "config5": "config5", "neohooke_mixed": "config5"}


def workload_name(dims):
    return "linear-elastic two-phase ellipsoidal inclusion (semi-axes 0.4 n), CG, HEX8, %dx%dx%d, 1 load case" % tuple(dims)


def run_reference(args, dims):
    """`--impl reference`: the reference's CPU implementation of the path on every host thread.  The FANS binary cannot exist here
    (MPI, FFTW, HDF5, Eigen absent), so this is the C++ restatement; a step is one CG iteration of the SAME workload, on the same
    grid when the host can hold it and finish within a few minutes, else on the largest cube that does (stated in `config`)."""
    K, W = max(args.steps, 1), max(args.warmup, 0)
    g = [args.ref_size] * 3 if args.ref_size else cpu_grid(dims, 150.0, K + min(W, 3))
    rate, it, dt, nthreads, fft = cpu_port_run(g, K, warm=min(W, 3))
    sample = ("%d CG iterations on %dx%dx%d after %d warm-up iterations, %.1f s, %.0f %% of it inside convolution(); multithreaded C++ "
              "restatement of the reference loop (oracle/cpu/fans_cpu.cpp), not the FANS binary (MPI/FFTW/HDF5/Eigen: unbuildable here)"
              % (it, g[0], g[1], g[2], min(W, 3), dt, 100.0 * fft / dt))
    line = {"impl": "reference", "metric": "voxel_dof_updates_per_s", "value": rate, "unit": "voxel-DOF/s", "n_gpus": args.gpus,
            "steps": it, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(it, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(dims), "grid": dims, "cpu_grid": g, "same_grid": list(g) == list(dims)},
            "cg_iterations_per_s": it / dt,
            "cpu_baseline": {"value": rate, "unit": "voxel-DOF/s", "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "voxel-DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def grid_for(n, gpus):
    """weak scaling: n^3 voxels per GPU; the grid doubles along x, then y, then z"""
    dims = [n, n, n]
    g, ax = gpus, 0
    while g > 1:
        dims[ax] *= 2
        ax = (ax + 1) % 3
        g //= 2
    return dims


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=0, help="voxels per axis and GPU (default: the BASELINE size of the workload)")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-size", type=int, default=0)
    ap.add_argument("--cpu-size", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--grid", default="", help="nx,ny,nz: override the weak-scaling grid (experiments only)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (experiments only)")
    ap.add_argument("--workload", default="elastic512", choices=sorted(WORKLOADS))
    ap.add_argument("--method", default="cg", help="config5: cg | fp")
    ap.add_argument("--load-steps", type=int, default=3, help="config3 / config5: load increments to run")
    ap.add_argument("--no-selfcheck", action="store_true")
    ap.add_argument("--sequential", action="store_true", help="config2: one solve per load case instead of fans_solve_batch")
    ap.add_argument("--no-ncu", action="store_true", help="do not measure roofline.traffic live with ncu")
    args = ap.parse_args()
    args.workload = WORKLOADS[args.workload]
    if not args.size:
        args.size = 256 if args.workload == "config2" else 512
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dims = grid_for(args.size, max(world, args.gpus) if args.impl == "reference" else world)
    if args.grid:
        dims = [int(x) for x in args.grid.split(",")]
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, dims)
        return

    import torch
    from fans_b200 import dist as fdist

    comm = fdist.init()
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    env = Env(args, rank, world, dev, comm, torch, dims)
    if args.workload in ("elastic512", "config4"):
        line = run_linear_headline(env)
    elif args.workload == "config2":
        line = run_config2(env)
    else:
        line = run_nonlinear(env)
    if world > 1 and not args.no_selfcheck:
        sc = selfcheck_slabs(env)
        if rank == 0:
            line["selfcheck"] = sc
    if rank == 0:
        print(json.dumps(line))
    comm.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


class Env:
    """per-process state shared by the workloads"""

    def __init__(self, args, rank, world, dev, comm, torch, dims):
        self.args, self.rank, self.world, self.dev, self.comm, self.torch, self.dims = args, rank, world, dev, comm, torch, dims
        from fans_b200 import dist as fdist
        self.x0, self.n0 = fdist.slab(dims[0], world, rank)
        self.nloc = float(self.n0) * dims[1] * dims[2]
        self.nvox = float(dims[0]) * dims[1] * dims[2]

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.torch.distributed.barrier()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.torch.distributed.all_reduce(t, op=self.torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def slab_comm(self):
        return self.comm if self.world > 1 else None


def base_line(env, value, t_loop, K, W, workload, extra_cfg=None):
    cfg = {"workload": workload, "grid": env.dims, "voxels_per_gpu": env.nloc,
           "decomposition": "x-slabs, %d plane(s) of %dx%d per GPU" % (env.n0, env.dims[1], env.dims[2]),
           "l2": "fields (%.2f GB each per GPU) are far larger than the 126 MB L2; no flush needed" % (24.0 * env.nloc / 1e9),
           "timing": "CUDA events on the library stream around the iteration loop (host polls of the error / line-search scalars "
                     "included), max over ranks, barrier + synchronize on both sides"}
    cfg.update(extra_cfg or {})
    return {"metric": "voxel_dof_updates_per_s", "value": value, "unit": "voxel-DOF/s", "n_gpus": env.world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_loop / max(K, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "cg_iterations_per_s": K / t_loop}


# ------------------------------------------------------------------------------------------------------------------------------
# headline (BASELINE.json metric) and config 4: fixed number of linear-elastic CG iterations on the two-phase image
# ------------------------------------------------------------------------------------------------------------------------------
def run_linear_headline(env):
    from fans_b200 import simple
    args, dims, world, rank, dev = env.args, env.dims, env.world, env.rank, env.dev
    K, W = args.steps, max(args.warmup, 3)
    if args.workload == "config4":
        ms = simple.voronoi_microstructure(dims, x0=env.x0, n0=env.n0)
        wl = "linear-elastic polycrystal-like Voronoi image (512 seeds per 1024^3, phase = grain mod 2), CG, HEX8, %dx%dx%d" % tuple(dims)
    else:
        ms = simple.ellipsoid_microstructure(dims, env.x0, env.n0)
        wl = workload_name(dims)
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], K_BULK, G_SHEAR, "HEX8", dev, gdims=dims, comm=env.slab_comm())
    ctx.set_gradient(G0)
    # warm-up: W iterations of a fresh solve (tol = 0 forces exactly n_it iterations)
    ctx.zero("u")
    ctx.solve("cg", W, 0.0, "Linfinity", "absolute")
    ctx.zero("u")
    env.barrier()
    l0 = ctx.launch_count()
    with ClockSampler(dev) as cs:
        res = ctx.solve("cg", K, 0.0, "Linfinity", "absolute")
        env.barrier()
    l1 = ctx.launch_count()
    assert res["iters"] == K, res
    t_loop = env.max_over_ranks(res["loop_ms"]) * 1e-3   # CUDA events on the library stream around exactly K iterations
    dof = 3.0 * env.nvox
    value = dof * K / t_loop

    # per-kernel device times (separate, untimed run so the event pairs do not perturb the number above)
    ctx.zero("u")
    ctx.set_profiling(True)
    ctx.solve("cg", max(3, min(K, 5)), 0.0, "Linfinity", "absolute")
    prof = ctx.profile()
    ctx.set_profiling(False)
    nloc = env.nloc
    alg = {k: v * nloc for k, v in ALG_BYTES_PER_VOXEL.items()}
    # the y passes carry the x<->y transposes over NVLink when world > 1: they are NVLink-bound there (nvlink_roofline below), the
    # HBM roofline object is then quoted for the dominant HBM-bound kernel
    iter_classes = {k: v for k, v in prof.items() if k in alg and not (world > 1 and k.startswith("fft_y"))}
    dom = max(iter_classes, key=lambda k: iter_classes[k][0] / iter_classes[k][1])
    dom_ms = iter_classes[dom][0] / iter_classes[dom][1]
    peaks, which = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    achieved = alg[dom] / (dom_ms * 1e-3) / 1e9
    iter_gbs = BYTES_PER_VOXEL_ITER_H3 * nloc * K / t_loop / 1e9   # per GPU

    e2e_obj = None
    if args.no_e2e:
        sig = ctx.homogenized_stress()
    else:
        e2e_obj, sig = e2e_leg(env, ctx, ms, K, dof)
    sc1 = None
    if world == 1 and not args.no_selfcheck:
        sc1 = selfcheck_operator(env, ctx)
    ctx.close()
    if rank != 0:
        return None
    line = base_line(env, value, t_loop, K, W, wl)
    traffic, tsrc = dram_traffic(dom, nloc, args)
    line.update({"hbm_roofline_iteration": {"bytes_per_voxel_iter": BYTES_PER_VOXEL_ITER_H3, "achieved_gbs_per_gpu": iter_gbs, "peak_gbs": peak,
                                            "frac": iter_gbs / peak, "peak_source": which},
                 "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                              "traffic": traffic, "traffic_source": tsrc, "peak_source": which, "ms_per_launch": dom_ms,
                              "algorithmic_bytes": ALG_BYTES_PER_VOXEL[dom] * nloc},
                 "kernel_ms": {k: v[0] / v[1] for k, v in prof.items()},
                 "clocks": cs.summary(), "e2e": e2e_obj, "gpu_launches": l1 - l0, "homogenized_stress": [float(x) for x in sig]})
    if sc1 is not None:
        line["selfcheck_1gpu"] = sc1
    if world > 1:
        # bytes each GPU sends (forward push) / receives (inverse pull) per transpose: its spectrum minus the block it keeps
        kzp = (dims[2] // 2 + 1 + 7) // 8 * 8
        tb = 16.0 * 3 * env.n0 * dims[1] * kzp * (world - 1) / world
        km = line["kernel_ms"]
        t_y = (km["fft_y_fwd"] + km["fft_y_inv"]) * 1e-3
        line["nvlink_roofline"] = {"bound": "nvlink", "kernels": "fft_y_fwd + fft_y_inv (transposes fused into the y passes, peer stores / loads)",
                                   "bytes_per_gpu_per_transpose": tb, "achieved": 2.0 * tb / t_y / 1e9, "peak": 770.0, "unit": "GB/s per direction",
                                   "frac": 2.0 * tb / t_y / 1e9 / 770.0,
                                   "peak_source": "measured peer copy per direction per GPU (B200_PROFILING.md); nominal NVLink 5: 900",
                                   "ms_both_transposes": 1e3 * t_y,
                                   "note": "kernel times from the profiling run (passes one after the other); in the timed loop the z passes "
                                           "of the neighbouring component overlap them (component pipeline)"}
    if not args.no_cpu and world == 1:
        line["cpu_baseline"] = cpu_baseline_object(args)
    return line


def e2e_leg(env, ctx, ms, K, dof):
    """the reference-facing call sequence with HOST buffers (pinned): microstructure + start field in, K iterations, homogenized
    stress + displacement field out; all copies inside the timed region, wall clock, max over ranks"""
    torch, dims, world = env.torch, env.dims, env.world
    u_host = torch.zeros((env.n0, dims[1], dims[2], 3), dtype=torch.float64, pin_memory=True).numpy()
    ms = torch.from_numpy(ms.view("int16")).pin_memory().numpy().view("uint16")   # every host buffer of the timed region is pinned
    ctx.upload("u", u_host)
    ctx.solve("cg", 1, 0.0, "Linfinity", "absolute")   # untimed: first-touch of the staging buffer
    env.barrier()
    t0 = time.perf_counter()
    ctx.set_microstructure(ms)               # H2D: phase image
    ctx.set_gradient(G0)
    ctx.zero("u")                            # a new Solver starts from v_u = 0 on its own (solver.h:128-131): the image is the input
    r2 = ctx.solve("cg", K, 0.0, "Linfinity", "absolute")
    sig = ctx.homogenized_stress()           # D2H: n_str doubles
    ctx.download_into("u", u_host)           # D2H: fluctuation field
    env.barrier()
    t_e2e = env.max_over_ranks(time.perf_counter() - t0)
    e2e = dof * r2["iters"] / t_e2e
    return {"value": e2e, "unit": "voxel-DOF/s", "h2d_bytes_per_step": world * ms.nbytes / K,
            "d2h_bytes_per_step": world * (u_host.nbytes + sig.nbytes) / K,
            "what": "set_microstructure (phase image H2D) + zero start field + K CG iterations + homogenized stress + download of the "
                    "displacement field, pinned host buffers, wall clock"}, sig


def selfcheck_operator(env, ctx):
    """Full-size self-consistency of the linear operator on this GPU: K.d of a pseudo-random field through the 27-point stencil
    kernel against the element-sweep form (FANS_LINEAR_SWEEP=1), and the dense-product sweep against the sum-factorised one for the
    residual — two independent kernels per operator, compared at the bench size (the oracle parity tests stop at 128^3)."""
    import numpy as np
    torch = env.torch
    n0, ny, nz = env.n0, env.dims[1], env.dims[2]
    g = torch.Generator(device="cpu").manual_seed(7)
    host = torch.empty((n0, ny, nz, 3), dtype=torch.float64, pin_memory=True)
    host.normal_(generator=g)
    host.mul_(1e-3)
    ctx.upload("u", host.numpy())
    out = {}
    ctx.apply_linear("rnew", "u")
    a = torch.from_numpy(ctx.download("rnew"))
    os.environ["FANS_LINEAR_SWEEP"] = "1"
    ctx.apply_linear("rnew", "u")
    del os.environ["FANS_LINEAR_SWEEP"]
    b = torch.from_numpy(ctx.download("rnew"))
    out["Kd_stencil_vs_sweep_rel"] = float((a - b).abs().max() / b.abs().max())
    del a, b
    ctx.residual("r", "u")
    a = torch.from_numpy(ctx.download("r"))
    os.environ["FANS_SWEEP_DENSE"] = "1"
    ctx.residual("r", "u")
    del os.environ["FANS_SWEEP_DENSE"]
    b = torch.from_numpy(ctx.download("r"))
    out["residual_sumfact_vs_dense_rel"] = float((a - b).abs().max() / b.abs().max())
    out["gate"] = 1e-12
    out["ok"] = bool(max(out["Kd_stencil_vs_sweep_rel"], out["residual_sumfact_vs_dense_rel"]) < 1e-12)
    out["grid"] = [n0, ny, nz]
    ctx.zero("u")
    return out


def selfcheck_slabs(env):
    """Multi-GPU parity inside the bench run (the driver's GPU test box has one GPU): a 64^3 two-phase sphere solved to 1e-10 on the
    N-rank slab path (halo exchange, fused NVLink transposes, scalar all-reduces) and, on rank 0 alone, on a single GPU.
    Gates: homogenized stress 1e-9, displacement field 1e-8 (relative, max norm), CG iterations +-1."""
    import numpy as np
    from fans_b200 import simple, dist as fdist
    torch, world, rank, dev = env.torch, env.world, env.rank, env.dev
    n = 64
    dims = [n, n, n]
    x0, n0 = fdist.slab(n, world, rank)
    ms_full = simple.sphere_microstructure(n)
    ctx = simple.linear_elastic_context(ms_full[x0:x0 + n0], [1.0, 1.0, 1.0], K_BULK, G_SHEAR, "HEX8", dev, gdims=dims, comm=env.comm)
    ctx.set_gradient(G0)
    res = ctx.solve("cg", 200, 1e-10, "Linfinity", "absolute")
    sig = ctx.homogenized_stress()
    u = torch.from_numpy(ctx.download("u")).cuda()
    ctx.close()
    parts = [torch.empty_like(u) for _ in range(world)]
    torch.distributed.all_gather(parts, u)
    out = None
    if rank == 0:
        u_slabs = torch.cat(parts, 0).cpu().numpy()
        c1 = simple.linear_elastic_context(ms_full, [1.0, 1.0, 1.0], K_BULK, G_SHEAR, "HEX8", dev)
        c1.set_gradient(G0)
        r1 = c1.solve("cg", 200, 1e-10, "Linfinity", "absolute")
        s1 = c1.homogenized_stress()
        u1 = c1.download("u")
        c1.close()
        out = {"case": "64^3 sphere, linear elastic CG to 1e-10: %d-rank slab path vs one GPU" % world,
               "sigma_rel_err": float(np.abs(sig - s1).max() / np.abs(s1).max()),
               "u_rel_err": float(np.abs(u_slabs - u1).max() / np.abs(u1).max()),
               "iters_diff": int(res["iters"] - r1["iters"]), "iters": int(res["iters"]), "gates": [1e-9, 1e-8, 1]}
        out["ok"] = bool(out["sigma_rel_err"] < 1e-9 and out["u_rel_err"] < 1e-8 and abs(out["iters_diff"]) <= 1)
    torch.distributed.barrier()
    return out


# ------------------------------------------------------------------------------------------------------------------------------
# config 2: 256^3 sphere, six unit load cases solved to 1e-10 (the homogenized tangent of solver.h:739-778)
# ------------------------------------------------------------------------------------------------------------------------------
def run_config2(env):
    import numpy as np
    from fans_b200 import simple
    args, dims, dev = env.args, env.dims, env.dev
    ms = simple.ellipsoid_microstructure(dims, env.x0, env.n0)   # the sphere of SURVEY 8d config 2 on a cube
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], K_BULK, G_SHEAR, "HEX8", dev, gdims=dims, comm=env.slab_comm())
    ctx.set_gradient(G0)
    ctx.solve("cg", 3, 0.0, "Linfinity", "absolute")   # warm-up
    torch = env.torch
    u_host = torch.zeros((env.n0, dims[1], dims[2], 3), dtype=torch.float64, pin_memory=True).numpy()
    env.barrier()

    def sequential():   # a new Solver (u = 0) per load case, main.cpp:15-16
        iters, t_loop, t_solve, C = 0, 0.0, 0.0, np.zeros((6, 6))
        for i in range(6):
            g = np.zeros(6)
            g[i] = 1e-3
            ctx.upload("u", u_host)          # H2D: zero start field
            ctx.set_gradient(g)
            r = ctx.solve("cg", 500, 1e-10, "Linfinity", "absolute")
            C[:, i] = ctx.homogenized_stress() / 1e-3
            iters += r["iters"]
            t_loop += r["loop_ms"] * 1e-3
            t_solve += r["elapsed_ms"] * 1e-3
        return iters, t_loop, t_solve, C

    def batched():      # the six load cases as lanes of one CG loop (fans_solve_batch); lanes start from u = 0 on the device
        res, sig = ctx.solve_batch(1e-3 * np.eye(6), 500, 1e-10, "Linfinity", "absolute")
        return sum(r["iters"] for r in res), res[0]["loop_ms"] * 1e-3, res[0]["elapsed_ms"] * 1e-3, sig.T / 1e-3

    use_batch = env.world == 1 and not args.sequential
    seq = None
    if use_batch:       # the one-after-the-other form next to it, for the record
        batched()       # warm-up: lane buffers
        ts = time.perf_counter()
        it_s, tl_s, _, C_s = sequential()
        seq = {"iters": it_s, "ms_per_step": 1e3 * tl_s / it_s, "wall_s_total": time.perf_counter() - ts}
    l0 = ctx.launch_count()
    with ClockSampler(dev) as cs:
        time.sleep(0.2)   # the sampler's first nvidia-smi process is up: its start-up does not sit inside the wall-clock region
        t0 = time.perf_counter()
        iters, t_loop, t_solve, C = batched() if use_batch else sequential()
        env.barrier()
        t_wall = time.perf_counter() - t0
    t_wall = env.max_over_ranks(t_wall)
    t_loop = env.max_over_ranks(t_loop)
    l1 = ctx.launch_count()
    if seq is not None:
        seq["tangent_max_rel_diff"] = float(np.abs(C - C_s).max() / np.abs(C_s).max())
    ctx.close()
    if env.rank != 0:
        return None
    dof = 3.0 * env.nvox
    line = base_line(env, dof * iters / t_loop, t_loop, iters, 3,
                     "config 2: linear-elastic two-phase sphere %dx%dx%d, CG to Linf 1e-10, 6 unit load cases 1e-3 e_i" % tuple(dims),
                     {"load_cases": 6, "batched": bool(use_batch)})
    peaks, which = measured_peaks()
    gbs = BYTES_PER_VOXEL_ITER_H3 * env.nloc * iters / t_loop / 1e9
    line.update({"hbm_roofline_iteration": {"bytes_per_voxel_iter": BYTES_PER_VOXEL_ITER_H3, "achieved_gbs_per_gpu": gbs,
                                            "peak_gbs": float(peaks["hbm_gbs"]), "frac": gbs / float(peaks["hbm_gbs"]), "peak_source": which},
                 "e2e": {"value": dof * iters / t_wall, "unit": "voxel-DOF/s",
                         "h2d_bytes_per_step": (288.0 if use_batch else env.world * 6.0 * u_host.nbytes) / iters,
                         "d2h_bytes_per_step": 6.0 * 48 / iters,
                         "what": ("fans_solve_batch: 6 macroscopic strains in, 6 homogenized stresses out, wall clock" if use_batch
                                  else "6 x (upload u, solve, homogenized stress), wall clock")},
                 "sequential": seq,
                 "solve_s_total": t_solve, "wall_s_total": t_wall, "clocks": cs.summary(), "gpu_launches": l1 - l0,
                 "homogenized_tangent": [[float(x) for x in row] for row in 0.5 * (C + C.T)]})
    return line


# ------------------------------------------------------------------------------------------------------------------------------
# configs 3 and 5: nonlinear CG (secant line search) / fixed point with history variables or mixed BCs, load increments to 1e-10
# ------------------------------------------------------------------------------------------------------------------------------
def run_nonlinear(env):
    import numpy as np
    from fans_b200 import simple, mixedbc
    args, dims, dev = env.args, env.dims, env.dev
    n = dims[1]
    steps = args.load_steps
    sl = slice(env.x0, env.x0 + env.n0)
    ctrl = None
    if args.workload == "config3":
        if dims[0] != dims[1]:
            raise SystemExit("config3 uses a cubic grid: run it with --grid n,n,n for more than one GPU")
        ms = simple.fiber_microstructure(n)[sl]
        ctx = simple.j2_fiber_context(ms, [1.0, 1.0, 1.0], "HEX8", dev, gdims=dims, comm=env.slab_comm())
        method, n_it = "cg", 500
        loads = [[1e-3 * (t + 1), 0, 0, 0, 0, 0] for t in range(steps)]   # every 10th increment of the config's 1e-4 ramp
        wl = ("config 3: J2ViscoPlastic_NonLinearIsotropicHardening matrix (test_J2Plasticity.json parameters) + elastic fibres along z "
              "(64 discs, vf 0.4), %dx%dx%d, HEX8, CG + secant line search to Linf 1e-10, eps11 = 1e-3 .. %g" % (tuple(dims) + (1e-3 * steps,)))
        hist_bytes = 1664.0   # 13 doubles read + 13 written per Gauss point, 8 Gauss points
    else:
        ms = simple.ellipsoid_microstructure(dims, env.x0, env.n0)
        ctx = simple.neohooke_context(ms, [1.0, 1.0, 1.0], K_BULK, G_SHEAR, "HEX8", dev, gdims=dims, comm=env.slab_comm())
        method, n_it = args.method, 2000
        lam = np.asarray(K_BULK) - 2.0 / 3.0 * np.asarray(G_SHEAR)
        kref = np.mean([simple.spatial_tangent_at_identity(lam[i], G_SHEAR[i]) for i in range(2)], axis=0)
        dF = 0.1
        if method == "fp":
            # the basic scheme diverges with the reference's own finite-strain reference medium (halved shear entries,
            # LargeStrainMechModel.h:142-171) — in the reference as well; like tests/test_nonlinear_gpu.py::test_large_strain it gets a
            # stiff user "reference_material" (MaterialManager.h:179-196: lambda, mu of the stiff phase) and 2 % stretch increments
            kref = np.zeros((9, 9))
            for i in range(3):
                for j in range(3):
                    kref[3 * i + i, 3 * j + j] += lam.max()
                    kref[3 * i + j, 3 * i + j] += max(G_SHEAR)
            ctx.set_reference_stiffness(kref)
            dF = 0.02
        # test_MixedBCs_LargeStrain.json load case 1: F33 ramp with P11 = P22 = 0, all other F components held at the identity
        mbc = mixedbc.MixedBC([1, 2, 3, 5, 6, 7, 8], [0, 4], [[0, 0, 0, 0, 0, 0, 1.0 + dF * (t + 1)] for t in range(steps)],
                              [[0.0, 0.0]] * steps, 9)
        ctrl = mixedbc.MixedBCController(mbc, kref)
        loads = [None] * steps
        wl = ("config 5: CompressibleNeoHookean two-phase sphere %dx%dx%d, HEX8, mixed BCs (F33 = %.2f .. %.2f, P11 = P22 = 0), %s to "
              "Linf 1e-10" % (tuple(dims) + (1.0 + dF, 1.0 + dF * steps, method)))
        hist_bytes = 0.0
    env.barrier()   # no separate warm-up: every load step is a full solve of seconds; the first one carries the one-off costs
    per_step, iters, evals, t_loop = [], 0, 0, 0.0
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    with ClockSampler(dev) as cs:
        for t in range(steps):
            if ctrl is not None:
                ctrl.activate(ctx, t)
            else:
                ctx.set_gradient(loads[t])
            if t == steps - 1:
                ctx.set_profiling(True)
            r = ctx.solve(method, n_it, 1e-10, "Linfinity", "absolute")
            if t == steps - 1:
                prof = ctx.profile()
                ctx.set_profiling(False)
            sig = ctx.homogenized_stress()
            ms_loop = env.max_over_ranks(r["loop_ms"])
            per_step.append({"iters": r["iters"], "residual_evals": r["n_residual_evals"], "loop_ms": ms_loop,
                             "ms_per_iter": ms_loop / max(r["iters"], 1), "fft_ms": r["fft_ms"], "err_last": r["err_last"],
                             "stress": [float(x) for x in sig], "gradient": [float(x) for x in ctx.get_gradient()]})
            iters += r["iters"]
            evals += r["n_residual_evals"]
            t_loop += ms_loop * 1e-3
            ctx.extrapolate_displacement()
        env.barrier()
    t_wall = env.max_over_ranks(time.perf_counter() - t0)
    l1 = ctx.launch_count()
    ctx.close()
    if env.rank != 0:
        return None
    dof = 3.0 * env.nvox
    line = base_line(env, dof * iters / t_loop, t_loop, iters, 0, wl, {"load_steps": steps, "method": method})
    peaks, which = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    km = {k: v[0] / v[1] for k, v in prof.items()}
    # dominant kernel: the residual sweep.  Algorithmic bytes per evaluation: u in + r out (2F) + phase image (2N) + history of
    # the history-bearing elements (config 3: the J2 matrix, 60 % of the voxels)
    frac_hist = 0.6 if hist_bytes else 0.0
    alg = (48.0 + 2.0 + hist_bytes * frac_hist) * env.nloc
    ach = alg / (km["sweep_residual"] * 1e-3) / 1e9
    line.update({"roofline": {"bound": "hbm", "kernel": "sweep_residual", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                              "traffic": None, "peak_source": which, "ms_per_launch": km["sweep_residual"], "algorithmic_bytes": alg,
                              "note": "the sweep is FP64-pipe / latency bound (profiles/r2_ncu_sweep_*.txt), HBM is the stated roofline"},
                 "residual_evals": evals, "kernel_ms": km, "kernel_calls_last_step": {k: v[1] for k, v in prof.items()},
                 "load_step_results": per_step, "clocks": cs.summary(), "gpu_launches": l1 - l0,
                 "e2e": {"value": dof * iters / t_wall, "unit": "voxel-DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8.0 * 9 * steps / max(iters, 1),
                         "what": "wall clock over all load steps incl. mixed-BC activation, history commit, homogenized stress readback; "
                                 "fields stay on the device between load steps like Solver::v_u in the reference"}})
    return line


def dram_traffic(kernel, nloc, args):
    """per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel: measured live with one short ncu
    pass of tools/kbench.py on the same grid when ncu is on PATH, else the table committed under profiles/ (bytes per voxel of an
    `ncu --set full` capture of this workload at 512^3), scaled to this run's voxels per GPU"""
    import shutil
    kname = {"sweep_linear": "k_stencil_linear", "fft_x_gamma": "k_fft_xg", "cg_update": "k_cg_update", "fft_z_inv": "k_fft_zi",
             "fft_z_fwd": "k_fft_zf", "fft_y_fwd": "k_fft_y", "fft_y_inv": "k_fft_y"}.get(kernel)
    if kname and shutil.which("ncu") and not args.no_ncu and args.workload == "elastic512" and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        try:
            cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:" + kname,
                   "--launch-skip", "3", "--launch-count", "1", "--csv", sys.executable, os.path.join(ROOT, "tools", "kbench.py"),
                   "--size", str(args.size), "--steps", "1", "--no-profile"]
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=240).stdout
            tot = 0.0
            for ln in out.splitlines():
                if "dram__bytes_" in ln:
                    cells = [c.strip('"') for c in ln.split('","')]
                    val, unit = float(cells[-1].replace(",", "")), cells[-2].lower()
                    tot += val * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            if tot > 0:
                return tot, "ncu live (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
        except Exception:
            pass
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic_per_voxel.json")))
        return t["kernels"][kernel] * nloc, "profiles/dram_traffic_per_voxel.json (ncu --set full capture) x voxels"
    except Exception:
        return None, None


if __name__ == "__main__":
    main()
