#!/bin/bash
# Round 2, session A: reference per-kernel baseline (ncu launch list with DRAM bytes), our launch list, tests
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 260 --csv \
   --log-file gpurun_out/r02_ref_launches.csv python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_ref_ncu.log 2>&1
echo "ref ncu rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
   --log-file gpurun_out/r02_ours_launches_start.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ours_ncu.log 2>&1
echo "ours ncu rc=$?"
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02a_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; cat gpurun_out/r02a_bench.json
