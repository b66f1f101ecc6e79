#!/bin/bash
# CUDA-graph replay of the linear iteration: parity tests, small-grid timing with and without, then the whole GPU suite
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rf 2>&1 | tail -8
for n in 16 32 64 128; do
  for gsw in 1 0; do
    FANS_GRAPH=$gsw timeout 120 python tools/kbench.py --size $n --steps 50 --no-profile --tag graph${gsw}_$n 2>&1 | tail -1 | cut -c1-120
  done
done
timeout 900 python -m pytest tests -m gpu -q -x -rf 2>&1 | tail -6
} > gpurun_out/r2gr_graph.txt 2>&1
cat gpurun_out/r2gr_graph.txt
