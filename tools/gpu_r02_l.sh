#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02l_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strong > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"; grep "^{" gpurun_out/r02l_bench.json | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['stage_ms'], l['e2e']['value'])"
