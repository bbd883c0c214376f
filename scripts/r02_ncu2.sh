#!/bin/bash
# round 2 ncu evidence for the shipped one-warp-per-check kernels (one GPU):
#  (1) per-launch time / fp64 instruction counts / DRAM bytes of the bench's OWN first batch (all bucket launches of the full matrix)
#  (2) --set full capture (source-level stalls) of the bucket launches of a 30 000-check sample (kernels short enough for every counter)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 1500 ncu --metrics $M --clock-control none -k regex:chain_check -c 6 --csv --log-file gpurun_out/r02_ncu2_bench_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-full-retries > gpurun_out/r02_ncu2_bench.json 2> gpurun_out/r02_ncu2_bench.err
echo "ncu metrics rc=$?"; tail -2 gpurun_out/r02_ncu2_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_check_se2 -c 6 -f -o gpurun_out/r02_prof2_se2 python scripts/run_sample.py 30000 m3500 > /dev/null 2> gpurun_out/r02_ncu2_full.err
echo "ncu full rc=$?"; tail -2 gpurun_out/r02_ncu2_full.err; ls -la gpurun_out/r02_prof2_se2.ncu-rep
