#!/bin/bash
# 2-GPU pass: multi-GPU parity tests (log kept under profiles/), bench with the N-rank selfcheck, config 3 on 2 GPUs
mkdir -p gpurun_out
T=r2e
nvidia-smi -L
timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp_gpu.py tests/test_invariants_gpu.py -m gpu -q -rs 2>&1 | tail -25 > gpurun_out/${T}_pytest_2gpu.log
cat gpurun_out/${T}_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
tail -c 1800 gpurun_out/${T}_bench_2gpu.json; tail -3 gpurun_out/${T}_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --workload config3 --grid 512,512,512 --load-steps 2 > gpurun_out/${T}_config3_512_2gpu.json 2> gpurun_out/${T}_config3_512_2gpu.err
tail -c 1500 gpurun_out/${T}_config3_512_2gpu.json; tail -3 gpurun_out/${T}_config3_512_2gpu.err
