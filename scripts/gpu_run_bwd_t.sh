#!/bin/bash
# each parametrisation in its own process: a device-side trap in one shape must not hide the others
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/bwd_t_tests.log
: > $out
ids=$(python -m pytest tests/test_gpu_parity.py --collect-only -q -k "transposed" 2>/dev/null | grep "::")
for id in $ids; do
  echo "=== $id" >> $out
  timeout 180 python -m pytest "$id" -x -q 2>&1 | tail -n 15 >> $out
done
grep -c "passed" $out
grep -B2 -A12 "failed\|Error\|error" $out | head -150
timeout 600 python scripts/gpu_gcn_bwd_bench.py > gpurun_out/bwd_t_bench.log 2>&1
cat gpurun_out/bwd_t_bench.log | tail -20
