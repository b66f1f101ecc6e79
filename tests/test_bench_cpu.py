"""CPU checks of bench.py's contract pieces that need no GPU: the per-pass algorithmic bytes add up to SURVEY.md 8(d)'s 554 B/voxel,
the weak-scaling grids, and the reference arm (`--impl reference`) prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_add_up():
    # 23 F + 2 N with F = 24 B/voxel (h = 3): P1..P7 of DESIGN.md section 4
    assert sum(bench.ALG_BYTES_PER_VOXEL.values()) == bench.BYTES_PER_VOXEL_ITER_H3 == 23 * 24 + 2


def test_weak_scaling_grids():
    assert bench.grid_for(512, 1) == [512, 512, 512]
    assert bench.grid_for(512, 2) == [1024, 512, 512]
    assert bench.grid_for(512, 4) == [1024, 1024, 512]
    assert bench.grid_for(512, 8) == [1024, 1024, 1024]     # the 1024^3 run of BASELINE.json at 8 GPUs


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-size", "16"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "voxel-DOF/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "voxel-DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_profile_tools_on_committed_evidence():
    """tools/launch_shares.py digests the committed ncu launch list; the kernels of one CG iteration are all there and their shares of
    the serialised launch list agree with the CUDA-event times bench.py reported for the same command (profiles/r1z_*)."""
    csv_path = os.path.join(ROOT, "profiles", "r1z_launches.csv")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_shares.py"), csv_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ms = {}
    for line in r.stdout.splitlines():
        parts = line.split()
        if len(parts) >= 5 and parts[0].startswith("k_"):
            ms[line[:50].strip()] = float(parts[-2])       # ms per launch
    for k in ("k_stencil_linear<3, 2, 1>", "k_cg_update", "k_fft_xg_seq<512, 8>", "k_fft_zi<256>", "k_fft_zf<256>", "k_fft_y<512, 8, 0>",
              "k_fft_y<512, 8, 1>"):
        assert k in ms, (k, sorted(ms))
    line = json.loads([l for l in open(os.path.join(ROOT, "profiles", "r1z_bench.json")) if l.startswith("{")][0])
    km = line["kernel_ms"]
    pairs = [("k_stencil_linear<3, 2, 1>", "sweep_linear"), ("k_cg_update", "cg_update"), ("k_fft_xg_seq<512, 8>", "fft_x_gamma"),
             ("k_fft_zi<256>", "fft_z_inv"), ("k_fft_zf<256>", "fft_z_fwd")]
    for a, b in pairs:
        assert abs(ms[a] - km[b]) / km[b] < 0.05, (a, ms[a], km[b])
    traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic_per_voxel.json")))["kernels"]
    for k, alg in bench.ALG_BYTES_PER_VOXEL.items():
        assert 0.97 * alg <= traffic[k] <= 1.15 * alg, (k, traffic[k], alg)   # DRAM traffic ~ algorithmic bytes: nothing is re-read
