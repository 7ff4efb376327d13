#!/bin/bash
# round-2 profile pass on the GPU box.  Everything lands in gpurun_out/ as small CSV / text (reports stay on the box).
set -u
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
BENCH="python bench.py --steps 2 --warmup 3 --no-sub --no-cpu"
# 1. every launch of the timed region with its device time, tensor-pipe activity and DRAM bytes (light metric set: all 96 conv launches)
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 330 --csv --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_ncu_launch.log 2>&1 ; echo "launch-list rc=$?"
# 2. --set full: the first conv launches of a step (stem, stride-2, 1x1, resident-weight 64 ch, halo 128 ch) and the non-conv kernels of the pipeline
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_bias_act -c 14 -f -o /tmp/r02_conv $BENCH > gpurun_out/r02_ncu_conv.log 2>&1 ; echo "conv-full rc=$?"
ncu -i /tmp/r02_conv.ncu-rep --page raw --csv > gpurun_out/r02_conv_raw.csv 2>/dev/null
timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:'track_step|nms_greedy|filter_raw|letterbox_reorg|spp_pool|upsample' -c 8 -f -o /tmp/r02_misc $BENCH > gpurun_out/r02_ncu_misc.log 2>&1 ; echo "misc-full rc=$?"
ncu -i /tmp/r02_misc.ncu-rep --page raw --csv > gpurun_out/r02_misc_raw.csv 2>/dev/null
# 3. the association branch's op-level kernels (C5 sweep + Kalman over 1 M tracks): numbers, then one full capture each
timeout 300 python tools/micro_bench.py --dtype f64 --batch 64 --reps 5 > gpurun_out/r02_micro_f64.log 2>&1 ; echo "micro rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:'kalman_|iou_cost|lap_' -c 14 -f -o /tmp/r02_assoc python tools/micro_bench.py --dtype f64 --batch 64 --reps 1 > gpurun_out/r02_ncu_assoc.log 2>&1 ; echo "assoc-full rc=$?"
ncu -i /tmp/r02_assoc.ncu-rep --page raw --csv > gpurun_out/r02_assoc_raw.csv 2>/dev/null
timeout 120 python tools/phase_profile.py bytetrack > gpurun_out/r02_phase.log 2>&1 ; echo "phase rc=$?"
ls -la gpurun_out | grep r02_ | tail -20
