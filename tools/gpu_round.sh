#!/bin/bash
# One GPU-box session: new-row tests first, then the full GPU suite, micro-benches, bench.py and ncu captures.
# Usage (from the repo root): gpurun --timeout 900 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 400 python -m pytest tests/test_color_passes_gpu.py tests/test_loss_gpu.py tests/test_adam_gpu.py -q -m gpu > gpurun_out/tests_new.log 2>&1; echo "new tests rc=$?"
tail -5 gpurun_out/tests_new.log
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_loss_gpu.py --deselect tests/test_adam_gpu.py --deselect tests/test_color_passes_gpu.py > gpurun_out/tests_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -3 gpurun_out/tests_gpu.log
timeout 120 python tools/bench_loss.py > gpurun_out/bench_loss.json 2> gpurun_out/bench_loss.err; cat gpurun_out/bench_loss.json
timeout 120 python tools/bench_adam.py > gpurun_out/bench_adam.json 2> gpurun_out/bench_adam.err; cat gpurun_out/bench_adam.json
timeout 120 python tools/bench_epilogue.py > gpurun_out/bench_epilogue.json 2> gpurun_out/bench_epilogue.err; cat gpurun_out/bench_epilogue.json
timeout 200 python tools/bench_semantic.py > gpurun_out/bench_semantic.json 2> gpurun_out/bench_semantic.err; cat gpurun_out/bench_semantic.json; tail -3 gpurun_out/bench_semantic.err
timeout 300 python tools/bench_iteration.py > gpurun_out/bench_iteration.json 2> gpurun_out/bench_iteration.err; cat gpurun_out/bench_iteration.json; tail -3 gpurun_out/bench_iteration.err
timeout 60 python tools/pcie_probe.py > gpurun_out/pcie.json 2>&1; cat gpurun_out/pcie.json
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/iteration_launches.csv python tools/bench_iteration.py > gpurun_out/ncu_iteration.log 2>&1; echo "ncu iteration rc=$?"
ls -la gpurun_out
