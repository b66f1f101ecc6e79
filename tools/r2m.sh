#!/bin/bash
# 8-GPU pass: multi-GPU parity tests (2/4/8 ranks) + the weak-scaling bench lines with the N-rank selfcheck
mkdir -p gpurun_out
T=r2m
nvidia-smi -L | wc -l
timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp_gpu.py::test_cli_two_ranks -m gpu -q -rs -rf 2>&1 | tail -15 > gpurun_out/${T}_pytest_8gpu.log
cat gpurun_out/${T}_pytest_8gpu.log
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
tail -c 2600 gpurun_out/${T}_bench_${N}gpu.json | cut -c1-2600; tail -2 gpurun_out/${T}_bench_${N}gpu.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
tail -c 1500 gpurun_out/${T}_bench_1gpu.json
