#!/bin/bash
# final single-GPU evidence at HEAD: smoke, GPU tests, bench line, ncu launch list of the bench command, ncu --set full of one CG iteration,
# configs 2 (batched) and 4
T=$1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 500 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-selfcheck --no-ncu > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fft_z|k_fft_y|k_fft_xg|k_stencil_linear|k_cg_update' \
    --launch-skip 14 --launch-count 7 -f -o gpurun_out/${T}_full python tools/kbench.py --steps 2 --no-profile > gpurun_out/${T}_ncu_full.log 2>&1
timeout 300 python bench.py --workload config2 > gpurun_out/${T}_config2_256.json 2> gpurun_out/${T}_config2.err
timeout 600 python bench.py --workload config4 --no-cpu > gpurun_out/${T}_config4_512.json 2> gpurun_out/${T}_config4.err
for f in bench config2_256 config4_512; do python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${T}_$f.json').read().strip().splitlines()[-1]); print('$f', round(j['ms_per_step'],3), 'ms/step', '%.3e'%j['value'], 'e2e %.3e'%j['e2e']['value'])
except Exception as e: print('$f failed', e)
PY
done
