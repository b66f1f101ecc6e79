#!/bin/bash
# usage: tools/r2_profile_1gpu.sh TAG -- the single-GPU evidence set of a round: GPU tests, bench line (+ reference arm), the BASELINE
# configs at their stated sizes, ncu launch list of the bench command, ncu --set full of one CG iteration and of the element sweep.
# Everything lands in gpurun_out/TAG_*; copy what is to be judged into profiles/.
T=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
tail -c 700 gpurun_out/${T}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-selfcheck --no-ncu > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fft_z|k_fft_y|k_fft_xg|k_stencil_linear|k_cg_update' \
    --launch-skip 14 --launch-count 7 -f -o gpurun_out/${T}_full python tools/kbench.py --steps 2 --no-profile > gpurun_out/${T}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_sf --launch-skip 2 --launch-count 2 -f -o gpurun_out/${T}_sweep \
    python tools/sweepbench.py --size 256 --laws linear --reps 1 > gpurun_out/${T}_ncu_sweep.log 2>&1
timeout 300 python bench.py --workload config2 > gpurun_out/${T}_config2_256.json 2> gpurun_out/${T}_config2.err
timeout 900 python bench.py --workload config3 --load-steps 3 > gpurun_out/${T}_config3_512.json 2> gpurun_out/${T}_config3.err
timeout 600 python bench.py --workload config4 --no-cpu > gpurun_out/${T}_config4_512.json 2> gpurun_out/${T}_config4.err
timeout 900 python bench.py --workload config5 --load-steps 2 > gpurun_out/${T}_config5_512.json 2> gpurun_out/${T}_config5.err
for f in config2_256 config3_512 config4_512 config5_512; do python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${T}_$f.json').read().strip().splitlines()[-1]); print('$f', round(j['ms_per_step'],3), 'ms/step', '%.3e'%j['value'])
except Exception as e: print('$f failed', e)
PY
done
