#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_exchange_gpu.py -x -q > gpurun_out/r02h_exchange.log 2>&1; echo "exchange tests rc=$?"; tail -5 gpurun_out/r02h_exchange.log
for mode in auto nccl; do
SURFEL_ROW_EXCHANGE=$mode timeout 300 python -W default -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py 600000 all > gpurun_out/r02h_check_${mode}_n$N.log 2>&1; echo "multigpu_check $mode rc=$?"; grep -E "MULTIGPU|exchange|Warning|Error|error" gpurun_out/r02h_check_${mode}_n$N.log | cut -c1-200 | head -8
done
for mode in auto nccl; do
SURFEL_ROW_EXCHANGE=$mode SURFEL_SHARD_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 12 --warmup 6 --no-strong > gpurun_out/r02h_bench_n${N}_${mode}.json 2> gpurun_out/r02h_bench_n${N}_${mode}.err; echo "$mode rc=$?"; grep "shard phases" gpurun_out/r02h_bench_n${N}_${mode}.err; grep "^{" gpurun_out/r02h_bench_n${N}_${mode}.json | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['parity_selfcheck']['weak_scene']['ok'], l['config'].get('image_exchange'), l['config'].get('row_exchange'))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err; echo "bench rc=$?"; grep "^{" gpurun_out/r02h_bench_n$N.json | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print({k:l[k] for k in ('value','ms_per_step','ms_per_step_median')}, 'e2e', l['e2e']['value'], 'selfcheck', {k:v['ok'] for k,v in l['parity_selfcheck'].items()})
c=l.get('config5_strong'); print({k:c[k] for k in ('ms_per_step','value','n1_ms_per_step','speedup_vs_n1')})
"
