#!/bin/bash
# ncu passes only (timed region of bench.py via cudaProfilerStart/Stop)
set -u
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1 ; echo "ncu-list rc=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tracker.csv \
    python bench.py --workload tracker --steps 20 --warmup 3 > gpurun_out/ncu_launch_trk.log 2>&1 ; echo "ncu-list-trk rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_bias_act -c 96 -f -o gpurun_out/prof_conv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_conv.log 2>&1 ; echo "ncu-conv rc=$?"
