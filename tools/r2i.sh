#!/bin/bash
mkdir -p gpurun_out
T=r2i
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_scenarios_gpu.py tests/test_midsize_gpu.py -m gpu -x -q -rf 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
python tools/kbench.py --steps 5 --tag defer >> gpurun_out/${T}_kbench.txt 2>&1
FANS_GPU_LIB=$PWD/fans_b200/lib/libfans_gpu_nodefer.so python tools/kbench.py --steps 5 --tag nodefer >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag defer_vor256 --ms voronoi --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
FANS_GPU_LIB=$PWD/fans_b200/lib/libfans_gpu_nodefer.so python tools/kbench.py --steps 5 --tag nodefer_vor256 --ms voronoi --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag defer_hom --ms homogeneous >> gpurun_out/${T}_kbench.txt 2>&1
cut -c1-300 gpurun_out/${T}_kbench.txt
