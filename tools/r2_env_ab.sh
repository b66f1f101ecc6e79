#!/bin/bash
# A/B of an environment switch of the library: tools/r2_env_ab.sh OUT VAR ; kbench with VAR unset and VAR=0, three images, then the GPU suite
OUT=$1; VAR=$2
mkdir -p gpurun_out
{
for v in on off; do
  if [ $v = on ]; then unset $VAR; else export $VAR=0; fi
  echo "== $VAR $v"
  for ms in ellipsoid homogeneous; do
    timeout 300 python tools/kbench.py --steps 10 --ms $ms --tag ${VAR}_${v}_$ms 2>&1 | tail -1 | cut -c1-330
  done
  timeout 120 python tools/kbench.py --steps 10 --size 256 --tag ${VAR}_${v}_256 2>&1 | tail -1 | cut -c1-330
done
unset $VAR
timeout 900 python -m pytest tests -m gpu -q -x -rf 2>&1 | tail -6
} > gpurun_out/$OUT 2>&1
cat gpurun_out/$OUT
