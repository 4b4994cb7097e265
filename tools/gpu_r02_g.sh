#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for mode in auto peers; do
SURFEL_IMAGE_EXCHANGE=$mode timeout 300 python -W default -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py 600000 all > gpurun_out/r02g_check_${mode}_n$N.log 2>&1; echo "multigpu_check $mode rc=$?"; grep -E "MULTIGPU|image exchange|Warning|Error|error" gpurun_out/r02g_check_${mode}_n$N.log | head -8
done
for mode in auto peers allreduce; do
SURFEL_IMAGE_EXCHANGE=$mode SURFEL_SHARD_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 12 --warmup 6 --no-strong > gpurun_out/r02g_bench_n${N}_${mode}.json 2> gpurun_out/r02g_bench_n${N}_${mode}.err; echo "$mode rc=$?"; grep "shard phases" gpurun_out/r02g_bench_n${N}_${mode}.err; grep "^{" gpurun_out/r02g_bench_n${N}_${mode}.json | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['parity_selfcheck']['weak_scene']['ok'], l['config'].get('image_exchange'))"
done
