#!/bin/bash
# the GPU suite on the non-default paths: plain launches instead of graph replay, and the opt-in TMA forward z pass
mkdir -p gpurun_out
{
echo "== FANS_GRAPH=0"; FANS_GRAPH=0 timeout 900 python -m pytest tests -m gpu -q -x -rf --deselect tests/test_graph_gpu.py 2>&1 | tail -5
echo "== FANS_Z_TMA=1"; FANS_Z_TMA=1 timeout 900 python -m pytest tests -m gpu -q -x -rf 2>&1 | tail -5
} > gpurun_out/r2va_variants_pytest.txt 2>&1
cat gpurun_out/r2va_variants_pytest.txt
