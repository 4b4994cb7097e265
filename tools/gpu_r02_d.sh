#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/calibrate_partition.py 16000000 8 0 16 64 256 > gpurun_out/r02d_calib16M.log 2>&1; echo rc=$?; cat gpurun_out/r02d_calib16M.log
timeout 600 python tools/calibrate_partition.py 8000000 8 16 64 > gpurun_out/r02d_calib8M.log 2>&1; echo rc=$?; cat gpurun_out/r02d_calib8M.log
