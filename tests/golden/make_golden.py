#!/usr/bin/env python
"""Regenerates the committed fixtures under tests/golden/ from the reference checkout (run in the BUILD container,
where /root/reference exists; the GPU box never reads the reference):

  sphere32_ms.npz          test/microstructures/sphere32.h5, dataset /sphere/32x32x32/ms, decoded (single deflate chunk),
                           stored in the on-disk z-y-x order exactly as the reference's Reader sees it before its
                           zyx->xyz transpose (src/reader.cpp:385-394)
  reference_inputs.json    the seven test/input_files/test_*.json scenarios, keyed by name (the two embedded
                           known answers live in MixedBCs / MixedBCs_LargeStrain)
  oracle_scenarios.json    outputs of oracle/fans_oracle.py on those scenarios (bounded number of steps): iteration
                           counts, macroscopic gradient, homogenized stress, error history — the numbers the GPU
                           parity tests are compared with on the GPU box without re-running the slow NumPy oracle.

Usage: python tests/golden/make_golden.py [--reference /root/reference] [--skip-oracle]
"""
import argparse
import glob
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# scenario -> number of time steps per load case the oracle output is recorded for
STEPS = {"LinearThermal": 1, "LinearElastic": 1, "PseudoPlastic": 4, "J2Plasticity": 4, "CompressibleNeoHookean": 2,
         "MixedBCs": 2, "MixedBCs_LargeStrain": 1}


def decode_sphere32(path):
    """The file is a v0-superblock HDF5 with ONE chunked uint8 dataset stored as a single deflate chunk."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    for off in range(len(b) - 2):  # locate the zlib stream (0x78 header) that inflates to 32^3 bytes
        if b[off] == 0x78:
            try:
                raw = zlib.decompressobj().decompress(b[off:])
            except zlib.error:
                continue
            if len(raw) == 32 ** 3:
                return np.frombuffer(raw, dtype=np.uint8).reshape(32, 32, 32).copy(), off
    raise RuntimeError("no 32^3 deflate chunk found")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--skip-oracle", action="store_true")
    a = ap.parse_args()
    zyx, off = decode_sphere32(os.path.join(a.reference, "test/microstructures/sphere32.h5"))
    print("sphere32: chunk at file offset", off, "phase-1 voxels", int(zyx.sum()))
    np.savez_compressed(os.path.join(HERE, "sphere32_ms.npz"), ms_zyx=zyx)
    inputs = {}
    for f in sorted(glob.glob(os.path.join(a.reference, "test/input_files/test_*.json"))):
        inputs[os.path.basename(f)[5:-5]] = json.load(open(f))
    json.dump(inputs, open(os.path.join(HERE, "reference_inputs.json"), "w"), indent=1)
    if a.skip_oracle:
        return
    import fans_oracle as fo
    ms = np.ascontiguousarray(zyx.transpose(2, 1, 0)).astype(np.uint16)
    out = {}
    for name, cfg in inputs.items():
        rec = []

        def on_step(sol, lc, t, res):
            res["stress_average"] = sol.get_homogenized_stress()
            rec.append({"load_case": lc, "step": t, "iters": int(res["iters"]), "n_residual_evals": int(res["n_residual_evals"]),
                        "g0": [float(x) for x in res["g0"]], "stress_average": [float(x) for x in res["stress_average"]],
                        "err_all": [float(x) for x in res["err_all"]],
                        "u_absmax": float(np.abs(sol.u).max()), "u_l2": float(np.sqrt((sol.u ** 2).sum()))})
            print(name, lc, t, res["iters"], flush=True)
        fo.run_load_cases(ms, cfg, max_steps=STEPS[name], on_step=on_step)
        out[name] = rec
    json.dump(out, open(os.path.join(HERE, "oracle_scenarios.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
