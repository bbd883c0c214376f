#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q -m gpu > gpurun_out/r02_c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c9_pytest.log
tail -5 gpurun_out/r02_c9_pytest.log
timeout 600 python scripts/stream_bench.py m3500 1.0 --limit=1200 --depth=8 2>&1 | tail -2
timeout 600 python scripts/stream_bench.py m3500 1.0 --limit=400 2>&1 | tail -2
