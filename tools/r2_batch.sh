#!/bin/bash
# batched-solve evidence: new tests first, then the whole GPU suite, the lane benchmark, config 2 (batched + sequential) and the headline
T=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch_gpu.py -m gpu -q -x -rf 2>&1 | tail -15 > gpurun_out/${T}_pytest_batch.log
cat gpurun_out/${T}_pytest_batch.log
timeout 600 python tools/batchbench.py 16 32 64 128 256 > gpurun_out/${T}_batchbench.txt 2>&1
cat gpurun_out/${T}_batchbench.txt
timeout 300 python bench.py --workload config2 > gpurun_out/${T}_config2_256.json 2> gpurun_out/${T}_config2.err
tail -c 1500 gpurun_out/${T}_config2_256.json; tail -3 gpurun_out/${T}_config2.err
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 900 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
