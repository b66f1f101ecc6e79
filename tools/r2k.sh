#!/bin/bash
mkdir -p gpurun_out
T=r2k
timeout 1500 python -m pytest tests -m gpu -x -q -rf 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
python tools/kbench.py --steps 5 --tag head >> gpurun_out/${T}_kbench.txt 2>&1
FANS_GPU_LIB=$PWD/fans_b200/lib/libfans_gpu_mt.so python tools/kbench.py --steps 5 --tag mixedtable >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag head_vor256 --ms voronoi --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
FANS_GPU_LIB=$PWD/fans_b200/lib/libfans_gpu_mt.so python tools/kbench.py --steps 5 --tag mixedtable_vor256 --ms voronoi --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
cut -c1-330 gpurun_out/${T}_kbench.txt
python tools/sweepbench.py --size 256 --laws j2 --tag head > gpurun_out/${T}_sweep_j2.json 2>&1
cat gpurun_out/${T}_sweep_j2.json
ncu --set full --clock-control none --import-source on -k regex:k_sweep_sf --launch-skip 2 --launch-count 1 -f -o gpurun_out/${T}_sweep_j2 \
    python tools/sweepbench.py --size 256 --laws j2 --reps 1 > gpurun_out/${T}_ncu_j2.log 2>&1
tail -2 gpurun_out/${T}_ncu_j2.log
