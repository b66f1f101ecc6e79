#!/bin/bash
# 2-GPU validation at HEAD: slab parity tests, the 2-rank C++ CLI, bench.py at N=2 with its self-check
T=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp_gpu.py -m gpu -q -rf 2>&1 | tail -8 > gpurun_out/${T}_pytest_2gpu.log
cat gpurun_out/${T}_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
tail -c 1500 gpurun_out/${T}_bench_2gpu.json; tail -3 gpurun_out/${T}_bench_2gpu.err
