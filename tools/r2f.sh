#!/bin/bash
mkdir -p gpurun_out
T=r2f
timeout 900 python -m pytest tests/test_invariants_gpu.py tests/test_host_cpp_gpu.py tests/test_midsize_gpu.py -m gpu -q -rf 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
for v in base reg r104; do
  FANS_GPU_LIB=$PWD/fans_b200/lib/libfans_gpu_$v.so python tools/kbench.py --steps 5 --tag $v >> gpurun_out/${T}_kbench.txt 2>&1
  FANS_GPU_LIB=$PWD/fans_b200/lib/libfans_gpu_$v.so python tools/kbench.py --steps 5 --tag ${v}_voronoi --ms voronoi --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
done
cat gpurun_out/${T}_kbench.txt
