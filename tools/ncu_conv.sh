#!/bin/bash
export PATH=/usr/local/cuda/bin:$PATH
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_detector.py -x -q 2>&1 | grep -E "Error|assert|passed|failed" | head -8
timeout 600 python tools/conv_sweep.py 2>&1 | tee gpurun_out/conv_sweep.log | cut -c1-150
bash tools/pipe.sh 2>&1 | cut -c1-1500
