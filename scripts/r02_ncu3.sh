#!/bin/bash
# ncu launch list of the bench's OWN first step (every chain_check launch of the full M3500 matrix): duration, DRAM bytes, fp64 instruction counts.
# Feeds roofline.traffic / roofline_fp64 of bench.py (profiles/r02_bench_launch_metrics.csv -> scripts/make_traffic_json.py).
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,launch__registers_per_thread
timeout 1200 ncu --metrics $M --clock-control none -k regex:chain_check -c 6 --csv --log-file gpurun_out/r02_bench_launch_metrics.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-full-retries > gpurun_out/r02_ncu3_bench.json 2> gpurun_out/r02_ncu3_bench.err
echo "ncu metrics rc=$?"; tail -2 gpurun_out/r02_ncu3_bench.err
