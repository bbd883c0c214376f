#!/bin/bash
# end-of-round record: tests, smoke, full bench, launch list of the same command, ncu full of the dominant kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cut -c1-2500 gpurun_out/bench_full.json; tail -2 gpurun_out/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_check_se2 -s 3 -c 2 -f -o gpurun_out/prof python bench.py --checks 20000 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof.ncu-rep; tail -2 gpurun_out/ncu_full.log | cut -c1-200
