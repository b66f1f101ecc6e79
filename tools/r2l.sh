#!/bin/bash
mkdir -p gpurun_out
T=r2l
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_scenarios_gpu.py tests/test_nonlinear_gpu.py -m gpu -x -q -rf 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
python tools/kbench.py --steps 5 --tag zi_prefetch >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag zi_prefetch_256 --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
cut -c1-330 gpurun_out/${T}_kbench.txt
