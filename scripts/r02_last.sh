#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k world1 > gpurun_out/r02_last_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_last_pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 95 python bench.py --steps 1 --warmup 1 --no-cpu --no-full-retries --stream-seconds 0 > gpurun_out/r02_bench_overlap_1step.json 2> gpurun_out/r02_bench_overlap_1step.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_overlap_1step.json; tail -2 gpurun_out/r02_bench_overlap_1step.err
