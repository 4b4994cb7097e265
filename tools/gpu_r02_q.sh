#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "radix or golden or full_size or live_against" > gpurun_out/r02q_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02q_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strong > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; grep "^{" gpurun_out/r02q_bench.json | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['stage_ms'], l['gpu_launches_per_step']['hand_written'])"
