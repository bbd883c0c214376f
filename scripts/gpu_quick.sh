#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --checks 400000 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_400k.json 2> gpurun_out/bench_400k.err; cut -c1-300 gpurun_out/bench_400k.json; tail -3 gpurun_out/bench_400k.err
