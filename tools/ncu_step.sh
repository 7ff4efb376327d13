#!/bin/bash
# ncu --set full capture of one steady-state track_step launch (+ source import); writes gpurun_out/prof_step.ncu-rep
export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_step -s ${SKIP:-75} -c 1 -f -o gpurun_out/prof_step \
    python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/ncu_step.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_step.log
