#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cut -c1-1500 gpurun_out/bench_full.json; tail -2 gpurun_out/bench_full.err
timeout 900 python bench.py --config sphere --checks 300000 --steps 2 --warmup 3 > gpurun_out/bench_sphere.json 2> gpurun_out/bench_sphere.err; cut -c1-1200 gpurun_out/bench_sphere.json; tail -2 gpurun_out/bench_sphere.err
timeout 900 python bench.py --config intel --steps 3 --warmup 3 > gpurun_out/bench_intel.json 2> gpurun_out/bench_intel.err; cut -c1-1200 gpurun_out/bench_intel.json; tail -2 gpurun_out/bench_intel.err
