#!/bin/bash
mkdir -p gpurun_out
timeout 60 python scripts/ab_overlap.py > gpurun_out/r02_ab_overlap.log 2> gpurun_out/r02_ab_overlap.err; echo "ab rc=$?"; cut -c1-260 gpurun_out/r02_ab_overlap.log; tail -3 gpurun_out/r02_ab_overlap.err
timeout 100 python -m pytest tests -x -q -m gpu -k "test_long_windows_match_oracle or test_pair_batch_matches_golden or test_full_size_properties or test_launch_variants or test_se3_pair_batch or test_consistency_matrix" > gpurun_out/r02_overlap_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_overlap_pytest.log
