#!/bin/bash
# round-2 GPU call: what runs is selected by the words in $STEPS (probe tests parity smoke bench ncu)
set -u
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
STEPS="${STEPS:-tests smoke bench}"
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
if has probe; then timeout 200 tools/probe/probe_bw ${PROBE_SEL:-8} > gpurun_out/r02_probe_bw2.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02_probe_bw2.log; fi
if has smoke; then timeout 600 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r02_smoke.log; fi
if has tests; then timeout 1500 python -m pytest ${TESTS:-tests} -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -${TAIL:-25} gpurun_out/r02_pytest.log; fi
if has parity; then timeout 900 python tools/parity_probe.py ${PARITY_ARGS:-} > gpurun_out/r02_parity_probe.log 2> gpurun_out/r02_parity_probe.err; echo "parity rc=$?"; cat gpurun_out/r02_parity_probe.log; tail -5 gpurun_out/r02_parity_probe.err; fi
if has bench; then timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; cat gpurun_out/r02_bench.json; tail -5 gpurun_out/r02_bench.err; fi
if has layers; then timeout 600 python tools/conv_layer_bench.py ${LAYER_ARGS:-} > gpurun_out/r02_layers.log 2> gpurun_out/r02_layers.err; echo "layers rc=$?"; cat gpurun_out/r02_layers.log; tail -5 gpurun_out/r02_layers.err; fi
