#!/bin/bash
# Round 2, session B: GPU tests (incl. new parity / caller tests) + the full default bench line + reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -s -k "live_against or sky_frame or reference_render or variants" > gpurun_out/r02b_newtests.log 2>&1; echo "new tests rc=$?"; grep -E "passed|failed|error|parity" gpurun_out/r02b_newtests.log | tail -12
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02b_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; cat gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02b_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02b_smoke.log
