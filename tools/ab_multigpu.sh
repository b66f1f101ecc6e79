#!/bin/bash
# usage: tools/ab_multigpu.sh NGPU "grid" "ENV1" "ENV2" ...   -- one short bench per environment variant, prints ms/iteration + kernel times
N=$1; GRID=$2; shift 2
port=29700
for v in "$@"; do
  port=$((port+1))
  env $v timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 6 --warmup 3 --grid $GRID --no-e2e --no-cpu 2>/dev/null > /tmp/ab.json
  python - "$v" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open('/tmp/ab.json') if l.startswith('{')][0])
    print(sys.argv[1], "ms/it=%.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["kernel_ms"].items() if not k.startswith("sweep_res")})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
