#!/bin/bash
# BASELINE configs at their stated GPU counts: config 4 (1024^3 Voronoi image) and config 5 (512^3 Neo-Hooke, mixed BCs, CG and FP) on 8 GPUs
T=$1
mkdir -p gpurun_out
run() { # name, port, args...
  n=$1; port=$2; shift 2
  timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
  echo "$n rc=$?"; tail -c 400 gpurun_out/${T}_$n.json; grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/${T}_$n.err | tail -3
}
run config5_512_cg_8gpu 29621 --workload config5 --grid 512,512,512 --load-steps 2 --no-cpu
run config5_512_fp_8gpu 29622 --workload config5 --grid 512,512,512 --load-steps 1 --method fp --no-cpu
run config4_1024_8gpu 29623 --workload config4 --steps 10 --warmup 3 --no-cpu
