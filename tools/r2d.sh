#!/bin/bash
mkdir -p gpurun_out
T=r2d
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 1500 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 900 python bench.py --workload config3 --load-steps 3 > gpurun_out/${T}_config3_512.json 2> gpurun_out/${T}_config3_512.err
tail -c 2500 gpurun_out/${T}_config3_512.json; tail -3 gpurun_out/${T}_config3_512.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
timeout 900 python bench.py --workload config5 --load-steps 2 > gpurun_out/${T}_config5_512.json 2> gpurun_out/${T}_config5_512.err
tail -c 2000 gpurun_out/${T}_config5_512.json; tail -3 gpurun_out/${T}_config5_512.err
timeout 900 python bench.py --workload config5 --method fp --load-steps 1 --size 256 > gpurun_out/${T}_config5_fp_256.json 2> gpurun_out/${T}_config5_fp_256.err
tail -c 1200 gpurun_out/${T}_config5_fp_256.json; tail -3 gpurun_out/${T}_config5_fp_256.err
timeout 600 python bench.py --workload config4 --no-cpu > gpurun_out/${T}_config4_512.json 2> gpurun_out/${T}_config4_512.err
tail -c 1500 gpurun_out/${T}_config4_512.json; tail -3 gpurun_out/${T}_config4_512.err
