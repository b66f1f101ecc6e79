#!/usr/bin/env python
"""bench.py — headline benchmark of the FANS per-iteration solve loop on B200.

Metric (BASELINE.json): voxel-DOF updates/s (and CG iterations/s) of a two-phase linear-elastic CG solve (HEX8, spherical
inclusion) on 512^3 voxels PER GPU, with the fraction of the measured HBM roofline.  A "step" is ONE CG iteration
(convolution: 5 FFT passes with the fused Green operator; fused direction update + K.d stencil; fused r/u update + norms).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size n] [--impl ours|reference]

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


def cpu_port_rate(n, iters):
    """CPU restatement (oracle: NumPy/BLAS + scipy.fft on every host thread) timed on an n^3 sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fans_oracle as fo
    ms = fo.sphere_microstructure(n)
    mats = [{"phases": [0, 1], "matmodel": "LinearElasticIsotropic", "material_properties": {"bulk_modulus": K_BULK, "shear_modulus": G_SHEAR}}]
    sol = fo.OracleSolver(ms, [1.0, 1.0, 1.0], "mechanical", mats, "HEX8", "cg", "small",
                          {"measure": "Linfinity", "type": "absolute", "tolerance": 0.0}, iters)
    sol.set_gradient(G0)
    t0 = time.perf_counter()
    sol.solve()
    dt = time.perf_counter() - t0
    return 3.0 * n ** 3 * sol.iter / dt, sol.iter, dt


def workload_name(dims):
    return "linear-elastic two-phase ellipsoidal inclusion (semi-axes 0.4 n), CG, HEX8, %dx%dx%d, 1 load case" % tuple(dims)


def run_reference(args, dims):
    n = args.ref_size
    for _ in range(max(1, args.warmup // 3)):
        cpu_port_rate(n, 1)
    rate, it, dt = cpu_port_rate(n, max(args.steps, 1))
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": "voxel_dof_updates_per_s", "value": rate, "unit": "voxel-DOF/s", "n_gpus": args.gpus,
            "steps": it, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(it, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(dims), "sample": "%d^3 sample of it" % n},
            "cg_iterations_per_s": it / dt,
            "cpu_baseline": {"value": rate, "unit": "voxel-DOF/s", "cores": cores, "kind": "port",
                             "sample": "%d CG iterations on a %d^3 sample, NumPy/scipy.fft restatement of the reference (oracle/), "
                                       "not the FANS binary (it needs MPI/FFTW/HDF5/Eigen: unbuildable here)" % (it, n)},
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
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-size", type=int, default=96)
    ap.add_argument("--cpu-size", type=int, default=96)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--grid", default="", help="nx,ny,nz: override the weak-scaling grid (experiments only)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (experiments only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dims = grid_for(args.size, max(world, args.gpus) if args.impl == "reference" else world)
    if args.grid:
        dims = [int(x) for x in args.grid.split(",")]
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, dims)
        return

    import numpy as np
    import torch
    from fans_b200 import simple, dist as fdist

    comm = fdist.init()
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    n, K, W = args.size, args.steps, max(args.warmup, 3)
    x0, n0 = fdist.slab(dims[0], world, rank)
    ms = simple.ellipsoid_microstructure(dims, x0, n0)
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], K_BULK, G_SHEAR, "HEX8", dev, gdims=dims, comm=comm if world > 1 else None)
    ctx.set_gradient(G0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # warm-up: W iterations of a fresh solve (tol = 0 forces exactly n_it iterations)
    ctx.zero("u")
    ctx.solve("cg", W, 0.0, "Linfinity", "absolute")
    ctx.zero("u")
    barrier()
    l0 = ctx.launch_count()
    with ClockSampler(dev) as cs:
        res = ctx.solve("cg", K, 0.0, "Linfinity", "absolute")
        barrier()
    l1 = ctx.launch_count()
    assert res["iters"] == K, res
    t_loop = max_over_ranks(res["loop_ms"]) * 1e-3   # CUDA events on the library stream around exactly K iterations
    nvox = float(dims[0]) * dims[1] * dims[2]
    dof = 3.0 * nvox
    value = dof * K / t_loop

    # per-kernel device times (separate, untimed run so the event pairs do not perturb the number above)
    ctx.zero("u")
    ctx.set_profiling(True)
    ctx.solve("cg", max(3, min(K, 5)), 0.0, "Linfinity", "absolute")
    prof = ctx.profile()
    ctx.set_profiling(False)
    nloc = float(n0) * dims[1] * dims[2]
    F = 8.0 * 3 * nloc
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

    # e2e: the reference-facing call sequence with HOST buffers (pinned): microstructure + start field in, K iterations,
    # homogenized stress + displacement field out; all copies inside the timed region, wall clock, max over ranks
    e2e_obj = None
    if args.no_e2e:
        sig = ctx.homogenized_stress()
    else:
        e2e_obj, sig = e2e_leg(torch, ctx, ms, n0, dims, K, dof, world, barrier, max_over_ranks)
    finish(args, rank, world, ctx, comm, torch, dims, n0, nloc, K, W, value, t_loop, iter_gbs, peak, which, dom, achieved, dom_ms, prof, cs,
           e2e_obj, l1 - l0, sig)


def e2e_leg(torch, ctx, ms, n0, dims, K, dof, world, barrier, max_over_ranks):
    u_host = torch.zeros((n0, dims[1], dims[2], 3), dtype=torch.float64, pin_memory=True).numpy()
    ms = torch.from_numpy(ms.view("int16")).pin_memory().numpy().view("uint16")   # every host buffer of the timed region is pinned
    ctx.upload("u", u_host)
    ctx.solve("cg", 1, 0.0, "Linfinity", "absolute")   # untimed: first-touch of the staging buffer
    u_host[...] = 0.0
    barrier()
    t0 = time.perf_counter()
    ctx.set_microstructure(ms)               # H2D: phase image
    ctx.set_gradient(G0)
    ctx.upload("u", u_host)                  # H2D: start field (Solver::v_u)
    r2 = ctx.solve("cg", K, 0.0, "Linfinity", "absolute")
    sig = ctx.homogenized_stress()           # D2H: n_str doubles
    ctx.download_into("u", u_host)           # D2H: fluctuation field
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e = dof * r2["iters"] / t_e2e
    return {"value": e2e, "unit": "voxel-DOF/s", "h2d_bytes_per_step": world * (ms.nbytes + u_host.nbytes) / K,
            "d2h_bytes_per_step": world * (u_host.nbytes + sig.nbytes) / K,
            "what": "set_microstructure + upload u + K CG iterations + homogenized stress + download u, pinned host buffers, wall clock"}, sig


def dram_traffic(kernel, nloc):
    """per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one `ncu --set full` capture of bench.py's workload,
    committed under profiles/ as bytes per voxel of the capture (512^3); scaled to this run's voxels per GPU"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic_per_voxel.json")))
        return t["kernels"][kernel] * nloc
    except Exception:
        return None


def finish(args, rank, world, ctx, comm, torch, dims, n0, nloc, K, W, value, t_loop, iter_gbs, peak, which, dom, achieved, dom_ms, prof, cs,
           e2e_obj, launches, sig):
    if rank == 0:
        line = {"metric": "voxel_dof_updates_per_s", "value": value, "unit": "voxel-DOF/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * t_loop / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(dims), "grid": dims, "voxels_per_gpu": nloc,
                           "decomposition": "x-slabs, %d plane(s) of %dx%d per GPU" % (n0, dims[1], dims[2]),
                           "l2": "fields (3.2 GB each per GPU) are far larger than the 126 MB L2; no flush needed",
                           "timing": "CUDA events on the library stream around exactly K iterations (one host poll of the error per "
                                     "iteration included), max over ranks, barrier + synchronize on both sides"},
                "cg_iterations_per_s": K / t_loop,
                "hbm_roofline_iteration": {"bytes_per_voxel_iter": BYTES_PER_VOXEL_ITER_H3, "achieved_gbs_per_gpu": iter_gbs, "peak_gbs": peak,
                                            "frac": iter_gbs / peak, "peak_source": which},
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": dram_traffic(dom, nloc), "peak_source": which, "ms_per_launch": dom_ms,
                             "algorithmic_bytes": ALG_BYTES_PER_VOXEL[dom] * nloc},
                "kernel_ms": {k: v[0] / v[1] for k, v in prof.items()},
                "clocks": cs.summary(),
                "e2e": e2e_obj,
                "gpu_launches": launches,
                "homogenized_stress": [float(x) for x in sig]}
        if world > 1:
            # bytes each GPU sends (forward push) / receives (inverse pull) per transpose: its spectrum minus the block it keeps
            kzp = (dims[2] // 2 + 1 + 7) // 8 * 8
            tb = 16.0 * 3 * n0 * dims[1] * kzp * (world - 1) / world
            km = line["kernel_ms"]
            t_y = (km["fft_y_fwd"] + km["fft_y_inv"]) * 1e-3
            line["nvlink_roofline"] = {"bound": "nvlink", "kernels": "fft_y_fwd + fft_y_inv (transposes fused into the y passes, peer stores / loads)",
                                       "bytes_per_gpu_per_transpose": tb, "achieved": 2.0 * tb / t_y / 1e9, "peak": 900.0, "unit": "GB/s per direction",
                                       "frac": 2.0 * tb / t_y / 1e9 / 900.0, "peak_source": "NVLink 5 nominal, per direction per GPU",
                                       "ms_both_transposes": 1e3 * t_y,
                                       "note": "kernel times from the profiling run (passes one after the other); in the timed loop the z passes "
                                               "of the neighbouring component overlap them (component pipeline)"}
        if not args.no_cpu and world == 1:
            rate, it, dt = cpu_port_rate(args.cpu_size, 6)
            line["cpu_baseline"] = {"value": rate, "unit": "voxel-DOF/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "%d CG iterations on a %d^3 sample of the workload, NumPy/scipy.fft restatement (oracle/), %.1f s"
                                              % (it, args.cpu_size, dt)}
        print(json.dumps(line))
    ctx.close()
    comm.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
