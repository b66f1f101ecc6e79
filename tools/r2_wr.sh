#!/bin/bash
# stencil warp footprint A/B: 1x64 (default build) against 2x32, 4x16, 8x8 (alt libraries), ellipsoid / Voronoi / homogeneous images at 512^3
mkdir -p gpurun_out
for k in 1 2 4 8; do
  if [ $k = 1 ]; then unset FANS_GPU_LIB; else export FANS_GPU_LIB=$PWD/fans_b200/lib/alt_wr$k.so; fi
  echo "== warp rows $k" 
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -2
  for ms in ellipsoid voronoi homogeneous; do
    timeout 300 python tools/kbench.py --steps 10 --ms $ms --tag wr${k}_$ms 2>&1 | tail -1 | cut -c1-330
  done
done > gpurun_out/r2wr_footprint.txt 2>&1
cat gpurun_out/r2wr_footprint.txt
