#!/usr/bin/env python
"""bench.py — headline benchmark of the FANS per-iteration solve loop on B200.

Metric (BASELINE.json): CG iterations/s and voxel-DOF updates/s of a 512^3 two-phase linear-elastic CG solve
(HEX8, sphere inclusion), with the fraction of the measured HBM roofline.  A "step" is ONE CG iteration
(convolution: 5 FFT passes with the fused Green operator; fused direction update + K.d sweep; fused r/u update + norms).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size n] [--impl ours|reference]

One JSON line on stdout (rank 0).  `--impl reference` times the CPU restatement of the reference algorithm
(oracle/) on a bounded sample on the host cores.
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
            time.sleep(0.1)

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
    """CPU restatement (oracle, NumPy, all host threads NumPy/BLAS uses) timed on an n^3 sample of the same workload."""
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


def run_reference(args):
    n = args.ref_size
    vals = []
    for _ in range(max(1, args.warmup // 3)):
        cpu_port_rate(n, 1)
    rate, it, dt = cpu_port_rate(n, args.steps)
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": "voxel_dof_updates_per_s", "value": rate, "unit": "voxel-DOF/s", "n_gpus": args.gpus,
            "steps": it, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(it, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "linear-elastic two-phase sphere, CG, HEX8, 512^3 (timed on a %d^3 sample of it)" % n},
            "cg_iterations_per_s": it / dt,
            "cpu_baseline": {"value": rate, "unit": "voxel-DOF/s", "cores": cores, "kind": "port",
                             "sample": "%d CG iterations on a %d^3 sample, NumPy restatement of the reference (oracle/), not the FANS binary" % (it, n)},
            "e2e": {"value": rate, "unit": "voxel-DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-size", type=int, default=64)
    ap.add_argument("--cpu-size", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    if args.gpus != 1 or int(os.environ.get("WORLD_SIZE", "1")) != 1:
        if rank == 0:
            print(json.dumps({"metric": "voxel_dof_updates_per_s", "n_gpus": args.gpus, "error": "multi-GPU slab path not built in this round"}))
        return

    import numpy as np
    from fans_b200 import simple

    n, K, W = args.size, args.steps, max(args.warmup, 3)
    ms = simple.sphere_microstructure(n)
    ctx = simple.linear_elastic_context(ms, [1.0, 1.0, 1.0], K_BULK, G_SHEAR, "HEX8", 0)
    ctx.set_gradient(G0)
    launches0 = ctx.launch_count()
    # warm-up: W iterations of a fresh solve (tol = 0 forces exactly n_it iterations)
    ctx.zero("u")
    ctx.solve("cg", W, 0.0, "Linfinity", "absolute")
    ctx.zero("u")
    l0 = ctx.launch_count()
    with ClockSampler(0) as cs:
        res = ctx.solve("cg", K, 0.0, "Linfinity", "absolute")
    l1 = ctx.launch_count()
    assert res["iters"] == K, res
    t_loop = res["loop_ms"] * 1e-3
    dof = 3.0 * n ** 3
    value = dof * K / t_loop
    # per-kernel device times (separate, untimed run so the event pairs do not perturb the number above)
    ctx.zero("u")
    ctx.set_profiling(True)
    ctx.solve("cg", max(3, min(K, 5)), 0.0, "Linfinity", "absolute")
    prof = ctx.profile()
    ctx.set_profiling(False)
    F = 8.0 * 3 * n ** 3
    alg = {"fft_z_fwd": 2 * F, "fft_y_fwd": 2 * F, "fft_x_gamma": 3 * F, "fft_y_inv": 2 * F, "fft_z_inv": 3 * F,
           "sweep_linear": 4 * F + 2.0 * n ** 3, "cg_update": 7 * F}
    iter_classes = {k: v for k, v in prof.items() if k in alg}
    dom = max(iter_classes, key=lambda k: iter_classes[k][0] / iter_classes[k][1])
    dom_ms = iter_classes[dom][0] / iter_classes[dom][1]
    peaks, which = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    achieved = alg[dom] / (dom_ms * 1e-3) / 1e9
    iter_gbs = BYTES_PER_VOXEL_ITER_H3 * n ** 3 * K / t_loop / 1e9

    # e2e: host buffers in, host buffers out, copies inside the timed region
    u_host = None
    ctx.zero("u")
    t0 = time.perf_counter()
    ctx.set_microstructure(ms)               # H2D: phase image
    ctx.set_gradient(G0)
    r2 = ctx.solve("cg", K, 0.0, "Linfinity", "absolute")
    sig = ctx.homogenized_stress()
    u_host = ctx.download("u")               # D2H: fluctuation field
    t_e2e = time.perf_counter() - t0
    e2e = dof * r2["iters"] / t_e2e

    line = {"metric": "voxel_dof_updates_per_s", "value": value, "unit": "voxel-DOF/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_loop / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "linear-elastic two-phase sphere (r=0.4n), CG, HEX8, %d^3, 1 load case" % n, "grid": [n, n, n],
                       "l2": "fields (3.2 GB each at 512^3) are far larger than the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the library stream around exactly K iterations (one host poll of the error per iteration included)"},
            "cg_iterations_per_s": K / t_loop,
            "hbm_roofline_iteration": {"bytes_per_voxel_iter": BYTES_PER_VOXEL_ITER_H3, "achieved_gbs": iter_gbs, "peak_gbs": peak,
                                        "frac": iter_gbs / peak, "peak_source": which},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": which, "ms_per_launch": dom_ms},
            "kernel_ms": {k: v[0] / v[1] for k, v in prof.items()},
            "clocks": cs.summary(),
            "e2e": {"value": e2e, "unit": "voxel-DOF/s", "h2d_bytes_per_step": ms.nbytes / K, "d2h_bytes_per_step": (u_host.nbytes + sig.nbytes) / K,
                    "what": "set_microstructure + K CG iterations + homogenized stress + download of u, wall clock"},
            "gpu_launches": l1 - l0,
            "homogenized_stress": [float(x) for x in sig]}
    if not args.no_cpu:
        rate, it, dt = cpu_port_rate(args.cpu_size, 8)
        line["cpu_baseline"] = {"value": rate, "unit": "voxel-DOF/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "%d CG iterations on a %d^3 sample of the workload, NumPy restatement (oracle/), %.1f s" % (it, args.cpu_size, dt)}
    print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
