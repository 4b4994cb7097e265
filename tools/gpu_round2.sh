#!/bin/bash
# Two-GPU session: gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round2.sh'
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu -k "sharded_two" > gpurun_out/tests_2gpu.log 2>&1; echo "2-GPU tests rc=$?"
tail -4 gpurun_out/tests_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
SURFEL_SHARD_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_2gpu_phases.json 2> gpurun_out/bench_2gpu_phases.err; tail -25 gpurun_out/bench_2gpu_phases.err
