#!/bin/bash
# usage: tools/r1_profile_1gpu.sh TAG  -- the single-GPU evidence set of a round: GPU tests, bench line, reference arm, ncu launch list,
# ncu --set full of one CG iteration (7 kernels).  Everything lands in gpurun_out/TAG_*.
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_fft_z|k_fft_y|k_fft_xg|k_stencil_linear|k_cg_update' \
    --launch-skip 14 --launch-count 7 -f -o gpurun_out/${TAG}_full python tools/kbench.py --steps 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
cat gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_bench.json
