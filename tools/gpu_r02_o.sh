#!/bin/bash
# register-cap / ring-depth experiments for the forward blend: rebuild render_fwd.cu on the box with compile-time switches
mkdir -p gpurun_out
cd streetunveiler_b200/csrc
run() { echo "=== variant: $1"; rm -f build/render_fwd.o; make EXTRA="$1" -j8 > /dev/null 2>&1 || echo BUILD FAILED; cuobjdump -res-usage ../libsurfel_b200.so 2>/dev/null | grep -A1 "render_fwd_kernelILb1ELb0ELb0" | grep -o "REG:[0-9]* STACK:[0-9]*"; (cd ../.. && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strong 2>/dev/null | grep "^{" | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['stage_ms']['render_fwd'], l['stage_ms']['render_bwd'])"); }
run ""
run "-DSURFEL_FWD_MAXREG=64"
run "-DSURFEL_FWD_MAXREG=72 -DSURFEL_FWD_STAGES=8"
run "-DSURFEL_FWD_MAXREG=48"
