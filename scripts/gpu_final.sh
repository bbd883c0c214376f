#!/bin/bash
# lean end-of-round record: full-workload bench line (2 timed steps), then the phase-clock split of the final kernel
mkdir -p gpurun_out
timeout 330 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
cut -c1-1800 gpurun_out/bench_full.json; tail -2 gpurun_out/bench_full.err
IPC_B200_LIB=$PWD/ipc_b200/libipc_b200_prof.so timeout 60 python scripts/phase_clocks.py 200000 > gpurun_out/phase_clocks.json 2> gpurun_out/phase_clocks.err; echo "phase rc=$?"
cat gpurun_out/phase_clocks.json | tr -d '\n' | cut -c1-1500
