#!/bin/bash
# Multi-GPU session: gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_r02_n.sh N'
N=${1:-2}
mkdir -p gpurun_out
for mode in all color_alpha; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py 600000 $mode > gpurun_out/r02_multigpu_check_n$N_$mode.log 2>&1; echo "multigpu_check $mode rc=$?"; tail -$((N+1)) gpurun_out/r02_multigpu_check_n$N_$mode.log
done
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"; cat gpurun_out/r02_bench_n$N.json; tail -5 gpurun_out/r02_bench_n$N.err
SURFEL_SHARD_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 6 --warmup 3 --no-strong > gpurun_out/r02_bench_n${N}_phases.json 2> gpurun_out/r02_bench_n${N}_phases.err; grep "shard phases" gpurun_out/r02_bench_n${N}_phases.err
SURFEL_SHARD_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 3 --total 8000000 > gpurun_out/r02_bench_n${N}_strong_phases.json 2> gpurun_out/r02_bench_n${N}_strong_phases.err; grep "shard phases" gpurun_out/r02_bench_n${N}_strong_phases.err
