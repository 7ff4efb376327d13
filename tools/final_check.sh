#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 150 python bench.py --steps 60 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
