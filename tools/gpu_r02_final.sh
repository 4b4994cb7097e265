#!/bin/bash
# Round 2 final single-GPU session: what the driver runs (tests, smoke, both bench arms) + the ncu evidence of the final build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02z_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02z_smoke.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02z_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err; echo "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 240 --csv \
   --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r02z_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_bwd" -s 4 -c 4 -o gpurun_out/r02_render -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r02z_ncu2.log 2>&1; echo "full capture rc=$?"
