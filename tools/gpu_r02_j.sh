#!/bin/bash
# bisect the 1e-7 forward difference: rebuild the blend kernels with compile-time switches on the box
mkdir -p gpurun_out
cd streetunveiler_b200/csrc
run() { echo "=== variant: $1"; rm -f build/render_fwd.o build/render_bwd.o; make EXTRA="$1" -j8 > /dev/null 2>&1 || echo BUILD FAILED; (cd ../.. && timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -s -k "live_against and 500000" 2>&1 | grep -E "live parity|passed|failed" | cut -c1-260); }
run ""
run "-DSURFEL_GENERIC_LDS"
run "-DSURFEL_CULL_IEEE_RCP"
run "-DSURFEL_GENERIC_LDS -DSURFEL_CULL_IEEE_RCP"
