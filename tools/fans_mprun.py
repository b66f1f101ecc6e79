#!/usr/bin/env python
"""mpiexec stand-in for the C++ front end: start P ranks of a command, one per GPU, with the rank environment FANS_gpu reads
(RANK, WORLD_SIZE, LOCAL_RANK) and a fresh FANS_COMM_FILE for the NCCL id rendezvous.
Usage: python tools/fans_mprun.py -n P -- tests/_build/FANS_gpu input.json results_dir [ms.u16 nx ny nz]"""
import argparse
import os
import subprocess
import sys
import tempfile

ap = argparse.ArgumentParser()
ap.add_argument("-n", type=int, required=True)
ap.add_argument("cmd", nargs=argparse.REMAINDER)
a = ap.parse_args()
cmd = a.cmd[1:] if a.cmd and a.cmd[0] == "--" else a.cmd
with tempfile.TemporaryDirectory() as tmp:
    procs = []
    for r in range(a.n):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(a.n), LOCAL_RANK=str(r), FANS_COMM_FILE=os.path.join(tmp, "nccl_id"))
        procs.append(subprocess.Popen(cmd, env=env, stdout=None if r == 0 else subprocess.DEVNULL))
    rc = 0
    for p in procs:
        rc = rc or p.wait()
sys.exit(rc)
