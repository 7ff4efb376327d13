#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_detector.py -m gpu -x -q > gpurun_out/pytest_det.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_det.log
timeout 300 python tools/conv_perf.py 8 > gpurun_out/conv_perf.log 2>&1; echo "perf rc=$?"; cat gpurun_out/conv_perf.log | cut -c1-200
timeout 400 python bench.py --steps 60 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
