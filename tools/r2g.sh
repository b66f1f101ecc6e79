#!/bin/bash
mkdir -p gpurun_out
T=r2g
timeout 1500 python -m pytest tests -m gpu -x -q -rf 2>&1 | tail -12 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
python tools/kbench.py --steps 5 --tag pf >> gpurun_out/${T}_kbench.txt 2>&1
FANS_XG_PF=0 python tools/kbench.py --steps 5 --tag nopf >> gpurun_out/${T}_kbench.txt 2>&1
python tools/kbench.py --steps 5 --tag pf256 --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
FANS_XG_PF=0 python tools/kbench.py --steps 5 --tag nopf256 --size 256 >> gpurun_out/${T}_kbench.txt 2>&1
cat gpurun_out/${T}_kbench.txt
ncu --set full --clock-control none --import-source on -k 'regex:k_stencil_linear|k_fft_xg' --launch-skip 6 --launch-count 2 -f -o gpurun_out/${T}_full python tools/kbench.py --steps 2 --no-profile > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
