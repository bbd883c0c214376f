#!/bin/bash
# the driver's launch line at N = 4 (shorter step count, bounded CPU legs)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 5 --warmup 3 --cpu-seconds 5 --stream-seconds 0 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r02_bench_4gpu.json; tail -5 gpurun_out/r02_bench_4gpu.err
