#!/bin/bash
# round 2, call 1: sharded C-ABI tests on one GPU + short N=1 bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_c1_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c1_pytest.log
tail -15 gpurun_out/r02_c1_pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 --stream-seconds 5 > gpurun_out/r02_c1_bench1.json 2> gpurun_out/r02_c1_bench1.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_c1_bench1.err; cat gpurun_out/r02_c1_bench1.json
