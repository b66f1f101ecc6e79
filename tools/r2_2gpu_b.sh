#!/bin/bash
# 2-GPU: slab parity incl. the Neo-Hooke mixed-BC case (first fixture variant only), config 5's bench path on slabs at a small size
T=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rf -k "P2-fused and not persist and not seq" 2>&1 | tail -8 > gpurun_out/${T}_pytest_2gpu_nh.log
cat gpurun_out/${T}_pytest_2gpu_nh.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --workload config5 --grid 128,128,128 --load-steps 1 --no-cpu > gpurun_out/${T}_config5_128_2gpu.json 2> gpurun_out/${T}_config5_128_2gpu.err
tail -c 600 gpurun_out/${T}_config5_128_2gpu.json; tail -3 gpurun_out/${T}_config5_128_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --workload config4 --grid 256,256,256 --no-cpu --steps 5 > gpurun_out/${T}_config4_256_2gpu.json 2> gpurun_out/${T}_config4_256_2gpu.err
tail -c 600 gpurun_out/${T}_config4_256_2gpu.json; tail -3 gpurun_out/${T}_config4_256_2gpu.err
