#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ab.py 200000 > gpurun_out/ab.log 2>&1; echo "ab rc=$?" >> gpurun_out/ab.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/ab.log; tail -15 gpurun_out/pytest_gpu.log
