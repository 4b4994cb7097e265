#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"preprocess_bwd|preprocess_fwd" -s 4 -c 2 -o gpurun_out/r02_k8 -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r02m_ncu.log 2>&1; echo "capture rc=$?"
