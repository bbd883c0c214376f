#!/bin/bash
# round 2, call 2 (2 GPUs): threaded 2-handle test of the sharded C ABI + the driver's multi-GPU bench command
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_c2_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/r02_c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c2_pytest.log
tail -5 gpurun_out/r02_c2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_c2_bench2.json 2> gpurun_out/r02_c2_bench2.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_c2_bench2.err; cat gpurun_out/r02_c2_bench2.json
