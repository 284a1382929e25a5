#!/bin/bash
cd "$(dirname "$0")/.."
for d in 0 16 8 11; do
  echo "== GWN_BT_DEBUG=$d"
  GWN_BT_DEBUG=$d timeout 300 python scripts/gpu_gcn_bwd_bench.py 2>&1 | grep "Lout=12 drop=0.3\|Lout=6 drop=0.3 sa=2"
done
