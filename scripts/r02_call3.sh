#!/bin/bash
# round 2, call 3: persistent stream solver (hand-written Cholesky) — stream parity tests + stream throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream or cli or matrix" > gpurun_out/r02_c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c3_pytest.log
tail -15 gpurun_out/r02_c3_pytest.log
timeout 300 python scripts/stream_bench.py intel 1.0 --oracle > gpurun_out/r02_c3_stream_intel.json 2> gpurun_out/r02_c3_stream_intel.err; echo "intel rc=$?"; cat gpurun_out/r02_c3_stream_intel.json; tail -3 gpurun_out/r02_c3_stream_intel.err
timeout 600 python scripts/stream_bench.py m3500 1.0 > gpurun_out/r02_c3_stream_m3500.json 2> gpurun_out/r02_c3_stream_m3500.err; echo "m3500 rc=$?"; cat gpurun_out/r02_c3_stream_m3500.json; tail -3 gpurun_out/r02_c3_stream_m3500.err
