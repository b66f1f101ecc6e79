#!/bin/bash
mkdir -p gpurun_out
T=r2u
timeout 1500 python -m pytest tests -m gpu -x -q -rf 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
python tools/nlbench.py --size 32 --steps 3 > gpurun_out/${T}_nl32.json 2>&1
python - <<PY
import json
j=json.loads(open('gpurun_out/${T}_nl32.json').read().strip().splitlines()[-1])
for s in j['steps']: print(s['iters'], s['residual_evals'], round(s['ms_per_iter'],3), s['kernel_calls'])
PY
python tools/kbench.py --steps 10 --tag head >> gpurun_out/${T}_kbench.txt 2>&1; cut -c1-330 gpurun_out/${T}_kbench.txt
