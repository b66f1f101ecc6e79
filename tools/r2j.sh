#!/bin/bash
mkdir -p gpurun_out
T=r2j
timeout 1200 python -m pytest tests/test_anysize_gpu.py tests/test_kernels_gpu.py -m gpu -q -rf 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
