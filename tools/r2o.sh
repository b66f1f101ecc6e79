#!/bin/bash
mkdir -p gpurun_out
T=r2o
timeout 1500 python -m pytest tests -m gpu -x -q -rf 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 1800 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
