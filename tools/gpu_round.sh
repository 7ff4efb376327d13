#!/bin/bash
# Runs on the GPU box under gpurun: smoke, GPU parity tests, bench (both arms), ncu launch list and
# one full capture of the dominant kernel.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" | tee -a gpurun_out/rc.txt
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/rc.txt
tail -15 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" | tee -a gpurun_out/rc.txt
cat gpurun_out/bench.json
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; echo "benchref rc=$?" | tee -a gpurun_out/rc.txt
cat gpurun_out/bench_ref.json
echo "== bench tracker-only workload (C3)" ; timeout 600 python bench.py --workload tracker > gpurun_out/bench_tracker.json 2> gpurun_out/bench_tracker.err ; echo "benchtrk rc=$?" | tee -a gpurun_out/rc.txt
cat gpurun_out/bench_tracker.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
# reports stay on the box (only their CSV export comes back): gpurun copies back at most 64 MiB
bash tools/ncu_lean.sh
fi
cat gpurun_out/rc.txt
