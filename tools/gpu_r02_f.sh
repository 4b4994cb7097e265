#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py 2000000 all > gpurun_out/r02f_multigpu_check_n$N.log 2>&1; echo "multigpu_check rc=$?"; tail -$((N+1)) gpurun_out/r02f_multigpu_check_n$N.log | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02f_bench_n$N.json 2> gpurun_out/r02f_bench_n$N.err; echo "bench rc=$?"; grep "^{" gpurun_out/r02f_bench_n$N.json | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print({k:l[k] for k in ('value','ms_per_step','ms_per_step_median')}, 'e2e', l['e2e']['value'], 'selfcheck', {k:v['ok'] for k,v in l['parity_selfcheck'].items()})
print(l.get('config5_strong'))
"
tail -3 gpurun_out/r02f_bench_n$N.err
SURFEL_SHARD_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 12 --warmup 6 --no-strong > gpurun_out/r02f_bench_n${N}_phases.json 2> gpurun_out/r02f_bench_n${N}_phases.err; grep "shard phases" gpurun_out/r02f_bench_n${N}_phases.err
SURFEL_SHARD_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 12 --warmup 6 --total 8000000 > gpurun_out/r02f_bench_n${N}_strong_phases.json 2> gpurun_out/r02f_bench_n${N}_strong_phases.err; grep "shard phases" gpurun_out/r02f_bench_n${N}_strong_phases.err
