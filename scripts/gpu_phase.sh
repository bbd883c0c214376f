#!/bin/bash
mkdir -p gpurun_out
IPC_B200_LIB=$PWD/ipc_b200/libipc_b200_prof.so timeout 300 python scripts/phase_clocks.py 200000 > gpurun_out/phase_clocks.json 2> gpurun_out/phase_clocks.err; echo rc=$?
cat gpurun_out/phase_clocks.json; tail -3 gpurun_out/phase_clocks.err
