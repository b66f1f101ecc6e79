#!/usr/bin/env python
"""DRAM bytes per voxel and launch of the seven CG-iteration kernels from one `ncu --set full` capture of bench.py's workload
(512^3 unless --voxels says otherwise) -> profiles/dram_traffic_per_voxel.json, read by bench.py for `roofline.traffic`.
Usage: python tools/traffic_from_ncu.py gpurun_out/xyz.ncu-rep [--voxels N] > profiles/dram_traffic_per_voxel.json"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
vox = float(sys.argv[sys.argv.index("--voxels") + 1]) if "--voxels" in sys.argv else 512.0 ** 3
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, k):
    return float(r[col[k]].replace(",", "")) * SCALE[units[col[k]]]


CLASS = [("k_fft_zf", "fft_z_fwd"), ("k_fft_zi", "fft_z_inv"), ("k_fft_xg", "fft_x_gamma"), ("k_stencil_linear", "sweep_linear"),
         ("k_cg_update", "cg_update")]
res, ny = {}, 0
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    cls = next((c for k, c in CLASS if k in name), None)
    if cls is None and "k_fft_y" in name:
        cls = "fft_y_inv" if ny else "fft_y_fwd"   # launch order inside one iteration
        ny += 1
    if cls and cls not in res:
        res[cls] = b / vox
print(json.dumps({"source": rep.split("/")[-1], "voxels_of_capture": vox, "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch / voxels",
                  "kernels": res}, indent=1))
