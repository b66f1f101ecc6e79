#!/bin/bash
mkdir -p gpurun_out
T=r2r
run() {  # N chunks ygrid tag
  FANS_CHUNKS=$2 FANS_Y_GRID=$3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $1 --steps 20 --warmup 5 --no-e2e --no-selfcheck > gpurun_out/${T}_n$1_$4.json 2> gpurun_out/${T}_n$1_$4.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${T}_n$1_$4.json').read().strip().splitlines()[-1])
    print('N=$1 $4 ms/it', round(j['ms_per_step'],3))
except Exception as e:
    print('N=$1 $4 failed', e); print(open('gpurun_out/${T}_n$1_$4.err').read()[-800:])
PY
}
run 8 0 96 old
run 8 4 96 c4y96
run 8 4 64 c4y64
run 8 4 48 c4y48
run 8 3 64 c3y64
run 4 0 96 old
run 4 4 64 c4y64
run 4 4 48 c4y48
