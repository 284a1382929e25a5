#!/bin/bash
# one GPU call: parity tests, default bench, launch list (ncu, after the plain run exited 0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/pytest_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.log 2>&1; rc=$?; echo "bench rc=$rc"
tail -n 3 gpurun_out/bench_$tag.log
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-roofline --no-extra > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
  python scripts/launch_summary.py gpurun_out/launches_$tag.csv 40 > gpurun_out/launch_summary_$tag.md 2>&1
  cat gpurun_out/launch_summary_$tag.md
fi
