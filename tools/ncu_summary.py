#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) per kernel: duration, DRAM bytes, pipe utilisation, occupancy and the
top warp-stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return float("nan")


KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__block_size", "block"), ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "warp_insts"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "bank_conf_ld"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "bank_conf_st"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct")]
for r in rows[2:]:
    print("==", r[col["Kernel Name"]][:110])
    for k, nm in KEYS:
        if k in col:
            print("   %-16s %s %s" % (nm, r[col[k]], units[col[k]]))
    st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), num(r, h)) for h in hdr
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    st.sort(key=lambda x: -x[1])
    print("   stalls (warps per issue-active cycle): " + ", ".join("%s=%.2f" % s for s in st[:6]))
