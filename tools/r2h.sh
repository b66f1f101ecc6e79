#!/bin/bash
mkdir -p gpurun_out
T=r2h
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_scenarios_gpu.py tests/test_midsize_gpu.py -m gpu -x -q -rf 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
python tools/kbench.py --steps 5 --tag rows2 >> gpurun_out/${T}_kbench.txt 2>&1
FANS_STENCIL_ROWS=1 python tools/kbench.py --steps 5 --tag rows1 >> gpurun_out/${T}_kbench.txt 2>&1
FANS_XG_T=4 python tools/kbench.py --steps 5 --tag rows2_xgT4 >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag rows2_vor256 --ms voronoi --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
FANS_STENCIL_ROWS=1 python tools/kbench.py --steps 5 --tag rows1_vor256 --ms voronoi --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag rows2_hom --ms homogeneous >> gpurun_out/${T}_kbench.txt 2>&1
FANS_STENCIL_CTAS=1776 python tools/kbench.py --steps 5 --tag rows2_ctas1776 >> gpurun_out/${T}_kbench.txt 2>&1
FANS_STENCIL_CTAS=7104 python tools/kbench.py --steps 5 --tag rows2_ctas7104 >> gpurun_out/${T}_kbench.txt 2>&1
cut -c1-420 gpurun_out/${T}_kbench.txt
