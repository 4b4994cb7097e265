#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_exchange_gpu.py -x -q > gpurun_out/r02e_exchange.log 2>&1; echo "exchange tests rc=$?"; tail -5 gpurun_out/r02e_exchange.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py 600000 all > gpurun_out/r02e_multigpu_check_n$N.log 2>&1; echo "multigpu_check rc=$?"; tail -$((N+1)) gpurun_out/r02e_multigpu_check_n$N.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02e_bench_n$N.json 2> gpurun_out/r02e_bench_n$N.err; echo "bench rc=$?"; python - <<PY
import json
l=json.load(open("gpurun_out/r02e_bench_n$N.json"))
print({k:l[k] for k in ("value","ms_per_step","ms_per_step_median")}, l["e2e"]["value"], l["parity_selfcheck"]["weak_scene"]["ok"], l.get("config5_strong"))
PY
tail -3 gpurun_out/r02e_bench_n$N.err
SURFEL_SHARD_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 12 --warmup 6 --no-strong > gpurun_out/r02e_bench_n${N}_phases.json 2> gpurun_out/r02e_bench_n${N}_phases.err; grep "shard phases" gpurun_out/r02e_bench_n${N}_phases.err
SURFEL_SHARD_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 12 --warmup 6 --total 8000000 > gpurun_out/r02e_bench_n${N}_strong_phases.json 2> gpurun_out/r02e_bench_n${N}_strong_phases.err; grep "shard phases" gpurun_out/r02e_bench_n${N}_strong_phases.err
