#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py 1000000 color_alpha > gpurun_out/r02p_multigpu_check_n$N.log 2>&1; echo "multigpu_check rc=$?"; tail -1 gpurun_out/r02p_multigpu_check_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02p_bench_n$N.json 2> gpurun_out/r02p_bench_n$N.err; echo "bench rc=$?"; grep "^{" gpurun_out/r02p_bench_n$N.json | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print({k:l[k] for k in ('value','ms_per_step','ms_per_step_median')}, 'e2e', l['e2e']['value'], 'selfcheck', {k:v['ok'] for k,v in l['parity_selfcheck'].items()})
c=l.get('config5_strong'); print({k:c[k] for k in ('ms_per_step','value','n1_ms_per_step','speedup_vs_n1')}); print(l.get('shard_balance'))
"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus $N --steps 5 --warmup 3 | cut -c1-300
