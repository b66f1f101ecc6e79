#!/bin/bash
# round 2, first GPU pass: parity tests, sweep A/B (dense vs sum-factorised), ncu of the new sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_pytest.log
python tools/sweepbench.py --size 256 --tag sf > gpurun_out/r2a_sweep_sf_256.json 2> gpurun_out/r2a_sweep_sf_256.err
FANS_SWEEP_DENSE=1 python tools/sweepbench.py --size 256 --tag dense > gpurun_out/r2a_sweep_dense_256.json 2> gpurun_out/r2a_sweep_dense_256.err
python tools/sweepbench.py --size 512 --laws linear,neohooke --tag sf > gpurun_out/r2a_sweep_sf_512.json 2> gpurun_out/r2a_sweep_sf_512.err
cat gpurun_out/r2a_sweep_*.json
ncu --set full --clock-control none --import-source on -k regex:k_sweep_sf --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2a_sweep \
    python tools/sweepbench.py --size 256 --laws linear --reps 1 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_ncu.log
