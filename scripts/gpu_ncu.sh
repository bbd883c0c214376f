#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_check_se2 -s 3 -c 2 -f -o gpurun_out/prof python bench.py --checks 20000 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof.ncu-rep; tail -2 gpurun_out/ncu_full.log | cut -c1-200
