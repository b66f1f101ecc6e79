#!/bin/bash
mkdir -p gpurun_out
T=r2s
for n in 32 64 128; do python tools/kbench.py --steps 20 --tag small$n --size $n >> gpurun_out/${T}_kbench.txt 2>&1; done
cut -c1-330 gpurun_out/${T}_kbench.txt
python tools/nlbench.py --size 32 --steps 2 > gpurun_out/${T}_nl32.json 2>&1; cut -c1-900 gpurun_out/${T}_nl32.json
