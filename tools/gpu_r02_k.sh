#!/bin/bash
# Round 2 profile session (1 GPU): launch list of the final build + ncu --set full of the two blend kernels
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 330 --csv \
   --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r02k_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_ -s 6 -c 3 -o gpurun_out/r02_render -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r02k_ncu2.log 2>&1; echo "full capture rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"preprocess_|digit_scatter|emit_instances" -s 8 -c 4 -o gpurun_out/r02_bandwidth -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r02k_ncu3.log 2>&1; echo "bw capture rc=$?"
ls -la gpurun_out/*.ncu-rep
