#!/bin/bash
# One GPU-box session.  Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [full]'
#   default: GPU test-suite, backward-variant bench, bench.py (both arms, with the e2e timeline on stderr)
#   full:    additionally the micro-benches of the "next" rows and the whole-iteration bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/tests_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -4 gpurun_out/tests_gpu.log
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; cat gpurun_out/$name.json; tail -8 gpurun_out/$name.err; }
run bench_variants python tools/bench_variants.py
BENCH_E2E_TRACE=1 run bench python bench.py
run bench_ref python bench.py --impl reference
run bench_semantic python tools/bench_semantic.py
if [ "$1" = "full" ]; then
  run bench_loss python tools/bench_loss.py
  run bench_epilogue python tools/bench_epilogue.py
  run bench_adam python tools/bench_adam.py
  ADAM_DENSE=1 run bench_adam_dense python tools/bench_adam.py
  run bench_iteration python tools/bench_iteration.py
  run pcie python tools/pcie_probe.py
fi
ls -la gpurun_out
