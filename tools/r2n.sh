#!/bin/bash
mkdir -p gpurun_out
T=r2n
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_scenarios_gpu.py tests/test_midsize_gpu.py -m gpu -x -q -rf 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
python tools/kbench.py --steps 5 --tag blend >> gpurun_out/${T}_kbench.txt 2>&1
FANS_GPU_LIB=$PWD/fans_b200/lib/libfans_gpu_noblend.so python tools/kbench.py --steps 5 --tag noblend >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag blend_vor256 --ms voronoi --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
FANS_GPU_LIB=$PWD/fans_b200/lib/libfans_gpu_noblend.so python tools/kbench.py --steps 5 --tag noblend_vor256 --ms voronoi --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag blend_vor512 --ms voronoi >> gpurun_out/${T}_kbench.txt 2>&1
cut -c1-330 gpurun_out/${T}_kbench.txt
