#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py -x -q -m gpu -k "stream or cli or add_edge" > gpurun_out/r02_c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c7_pytest.log
tail -12 gpurun_out/r02_c7_pytest.log
timeout 600 python scripts/stream_bench.py m3500 1.0 --limit=800 --depth=8 2>&1 | tail -2
