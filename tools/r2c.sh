#!/bin/bash
mkdir -p gpurun_out
T=r2c
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 300 python tools/sweepbench.py --size 256 --tag sf3 > gpurun_out/${T}_sweep_256.json 2> gpurun_out/${T}_sweep_256.err
timeout 300 python tools/sweepbench.py --size 512 --laws linear,neohooke --tag sf3 > gpurun_out/${T}_sweep_512.json 2> gpurun_out/${T}_sweep_512.err
cat gpurun_out/${T}_sweep_*.json
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 3000 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 300 python bench.py --workload config2 > gpurun_out/${T}_config2.json 2> gpurun_out/${T}_config2.err
tail -c 1500 gpurun_out/${T}_config2.json; tail -3 gpurun_out/${T}_config2.err
timeout 600 python bench.py --workload config3 --size 256 --load-steps 2 > gpurun_out/${T}_config3_256.json 2> gpurun_out/${T}_config3_256.err
tail -c 2500 gpurun_out/${T}_config3_256.json; tail -3 gpurun_out/${T}_config3_256.err
timeout 600 python bench.py --workload config5 --size 256 --load-steps 1 > gpurun_out/${T}_config5_256.json 2> gpurun_out/${T}_config5_256.err
tail -c 2500 gpurun_out/${T}_config5_256.json; tail -3 gpurun_out/${T}_config5_256.err
nproc; free -g | head -2
