#!/bin/bash
# generic A/B of the default library against alternative builds: tools/r2_ab.sh OUT alt1.so alt2.so ... ; kbench on three images each
OUT=$1; shift
mkdir -p gpurun_out
for lib in default "$@"; do
  if [ $lib = default ]; then unset FANS_GPU_LIB; else export FANS_GPU_LIB=$PWD/fans_b200/lib/$lib; fi
  echo "== $lib"
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -1
  for ms in ellipsoid voronoi homogeneous; do
    timeout 300 python tools/kbench.py --steps 10 --ms $ms --tag ${lib}_$ms 2>&1 | tail -1 | cut -c1-330
  done
done > gpurun_out/$OUT 2>&1
cat gpurun_out/$OUT
