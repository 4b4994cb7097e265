#!/bin/bash
# One GPU-box session.  Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [full]'
#   default: GPU test-suite, backward-variant bench, bench.py (both arms, with the e2e timeline on stderr)
#   full:    additionally the micro-benches of the "next" rows and the whole-iteration bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/tests_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -4 gpurun_out/tests_gpu.log
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; cat gpurun_out/$name.json; tail -8 gpurun_out/$name.err; }
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
BENCH_E2E_TRACE=1 run bench python bench.py
run bench_ref python bench.py --impl reference
run bench_activation python tools/bench_activation.py
run bench_iteration python tools/bench_iteration.py
if [ "$1" = "full" ]; then
  run bench_loss python tools/bench_loss.py
  run bench_epilogue python tools/bench_epilogue.py
  run bench_adam python tools/bench_adam.py
  ADAM_DENSE=1 run bench_adam_dense python tools/bench_adam.py
  run bench_semantic python tools/bench_semantic.py
  run pcie python tools/pcie_probe.py
fi
if [ "$1" = "iteration-profile" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r01_iteration_launches.csv python tools/bench_iteration.py > gpurun_out/ncu_iteration.log 2>&1; echo "ncu iteration rc=$?"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:activate_ -c 2 -o gpurun_out/r01f_activate -f python tools/bench_activation.py > gpurun_out/ncu_activate.log 2>&1; echo "ncu activate rc=$?"
fi
if [ "$1" = "full" ] || [ "$1" = "profile" ]; then
  # ncu passes (never a source of bench numbers): launch list of the benchmark step, full captures of the blend kernels
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_ -s 6 -c 3 -o gpurun_out/r01f_render -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_render.log 2>&1; echo "ncu render rc=$?"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_ -c 2 -o gpurun_out/r01f_classes -f python tools/bench_semantic.py > gpurun_out/ncu_classes.log 2>&1; echo "ncu classes rc=$?"
fi
ls -la gpurun_out
