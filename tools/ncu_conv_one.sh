#!/bin/bash
# ncu --set full of single conv configurations; exports raw CSV pages (small) into gpurun_out/
set -u
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
i=0
while read -r line; do
  [ -z "$line" ] && continue
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_bias_act -s 1 -c 1 -f -o /tmp/one_$i python tools/conv_one.py $line > gpurun_out/ncu_one_$i.log 2>&1
  echo "cfg $i: $line rc=$?"
  ncu -i /tmp/one_$i.ncu-rep --page raw --csv > gpurun_out/ncu_one_${i}_raw.csv 2>/dev/null
  ncu -i /tmp/one_$i.ncu-rep --page details --csv > gpurun_out/ncu_one_${i}_details.csv 2>/dev/null
  cp /tmp/one_$i.ncu-rep gpurun_out/ncu_one_$i.ncu-rep
done <<< "$CONFIGS"
ls -la gpurun_out | grep ncu_one
