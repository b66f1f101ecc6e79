#!/bin/bash
mkdir -p gpurun_out
T=r2t
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_scenarios_gpu.py tests/test_nonlinear_gpu.py tests/test_anysize_gpu.py -m gpu -x -q -rf 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
for n in 32 64 128 256 512; do python tools/kbench.py --steps 20 --tag xchunk$n --size $n >> gpurun_out/${T}_kbench.txt 2>&1; done
cut -c1-330 gpurun_out/${T}_kbench.txt
python tools/nlbench.py --size 32 --steps 2 > gpurun_out/${T}_nl32.json 2>&1; cut -c1-400 gpurun_out/${T}_nl32.json
