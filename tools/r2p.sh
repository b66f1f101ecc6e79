#!/bin/bash
mkdir -p gpurun_out
T=r2p
python tools/kbench.py --steps 5 --tag poly16_256 --ms poly16 --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag poly16_512 --ms poly16 >> gpurun_out/${T}_kbench.txt 2>&1
FANS_LINEAR_SWEEP=1 python tools/kbench.py --steps 3 --tag poly16_512_sweepform --ms poly16 >> gpurun_out/${T}_kbench.txt 2>&1
cut -c1-330 gpurun_out/${T}_kbench.txt
ncu --set full --clock-control none --import-source on -k 'regex:k_stencil_linear' --launch-skip 3 --launch-count 1 -f -o gpurun_out/${T}_stencil_vor python tools/kbench.py --steps 2 --no-profile --ms voronoi --size 256 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
