#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exchange_gpu.py -x -q -s > gpurun_out/r02c_exchange.log 2>&1; echo "exchange tests rc=$?"; tail -30 gpurun_out/r02c_exchange.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02c_tests.log
