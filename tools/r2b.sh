#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_pytest.log
python tools/sweepbench.py --size 256 --tag sf2 > gpurun_out/r2b_sweep_256.json 2> gpurun_out/r2b_sweep_256.err
python tools/sweepbench.py --size 512 --laws linear,neohooke --tag sf2 > gpurun_out/r2b_sweep_512.json 2> gpurun_out/r2b_sweep_512.err
cat gpurun_out/r2b_sweep_*.json
ncu --set full --clock-control none --import-source on -k regex:k_sweep_sf --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2b_sweep \
    python tools/sweepbench.py --size 256 --laws linear --reps 1 > gpurun_out/r2b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_sf --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2b_sweep_j2 \
    python tools/sweepbench.py --size 256 --laws j2 --reps 1 > gpurun_out/r2b_ncu_j2.log 2>&1
tail -3 gpurun_out/r2b_ncu_j2.log
