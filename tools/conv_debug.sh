#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detector.py -x -q 2>&1 | tail -25 | tee gpurun_out/conv_test.log
timeout 600 python tools/conv_perf.py 8 2>&1 | tee gpurun_out/conv_perf.log
