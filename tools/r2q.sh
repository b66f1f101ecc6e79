#!/bin/bash
mkdir -p gpurun_out
T=r2q
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rf -x 2>&1 | tail -6 > gpurun_out/${T}_pytest_2gpu.log
cat gpurun_out/${T}_pytest_2gpu.log
for C in 4 0 2; do
FANS_CHUNKS=$C timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$C bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/${T}_bench_2gpu_chunks$C.json 2> gpurun_out/${T}_bench_2gpu_chunks$C.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${T}_bench_2gpu_chunks$C.json').read().strip().splitlines()[-1])
    print('chunks=$C ms/it', round(j['ms_per_step'],3), 'selfcheck', j['selfcheck']['ok'], j['selfcheck']['sigma_rel_err'], j['selfcheck']['iters_diff'])
except Exception as e:
    print('chunks=$C failed', e); print(open('gpurun_out/${T}_bench_2gpu_chunks$C.err').read()[-1500:])
PY
done
for Y in 64 120; do
FANS_CHUNKS=4 FANS_Y_GRID=$Y timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$((Y/60)) bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-selfcheck > gpurun_out/${T}_bench_2gpu_y$Y.json 2> gpurun_out/${T}_bench_2gpu_y$Y.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${T}_bench_2gpu_y$Y.json').read().strip().splitlines()[-1])
    print('ygrid=$Y ms/it', round(j['ms_per_step'],3))
except Exception as e:
    print('ygrid=$Y failed', e)
PY
done
