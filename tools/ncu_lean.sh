#!/bin/bash
# lean ncu pass: pipeline launch list + one full-set capture of 24 conv launches, exported to CSV on the box
set -u
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1 ; echo "ncu-list rc=$?"
timeout 240 ncu --profile-from-start off --set full --clock-control none -k regex:conv_bias_act -c 24 -f -o /tmp/prof_conv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_conv.log 2>&1 ; echo "ncu-conv rc=$?"
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2> gpurun_out/prof_conv_raw.err ; echo "export rc=$?"
ls -la gpurun_out | tail -5
