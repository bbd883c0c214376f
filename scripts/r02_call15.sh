#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/stream_bench.py m3500 1.0 --depth=8 > gpurun_out/r02_stream_m3500_d8.json 2> gpurun_out/r02_c15.err; tail -2 gpurun_out/r02_c15.err; cut -c1-600 gpurun_out/r02_stream_m3500_d8.json
timeout 300 python scripts/stream_bench.py intel 1.0 --depth=8 --oracle > gpurun_out/r02_stream_intel_d8.json 2>> gpurun_out/r02_c15.err; cut -c1-900 gpurun_out/r02_stream_intel_d8.json
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02_c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c15_pytest.log
tail -15 gpurun_out/r02_c15_pytest.log
