#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_graph_gpu.py tests/test_batch_gpu.py tests/test_pyfans_gpu.py -m gpu -q -x -rf 2>&1 | tail -8
timeout 300 python tools/batchbench.py 16 32 64 128
echo "== FANS_GRAPH=0"
FANS_GRAPH=0 timeout 300 python tools/batchbench.py 16 32 64 128
} > gpurun_out/r2gs_batch_graph.txt 2>&1
cat gpurun_out/r2gs_batch_graph.txt
