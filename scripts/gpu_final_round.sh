#!/bin/bash
# Final evidence of a build (one GPU call): parity suite, smoke, the driver-style bench line, and ncu launch lists
# (gpu__time_duration per launch) of one config-2 and one config-3 step.   usage: scripts/gpu_final_round.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-final}
out=gpurun_out/final_$tag
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; tail -2 $out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_20.log 2>&1
tail -1 $out/bench_20.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'bwd frac', d['roofline']['frac'], 'gate', d['roofline_gate']['frac'], 'hop', d['roofline_diffusion_v3100']['frac'], 'c3', d['config3']['samples_per_s'], 'c5', d['config5']['samples_per_s'], 'cpu', d['cpu_baseline']['value'])"
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-roofline --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_c2.csv $B > $out/ncu_c2.log 2>&1
python scripts/launch_summary.py $out/launches_c2.csv 45 > $out/launch_summary_c2.md 2>&1; head -8 $out/launch_summary_c2.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file $out/launches_c3.csv $B --config c3 > $out/ncu_c3.log 2>&1
python scripts/launch_summary.py $out/launches_c3.csv 30 > $out/launch_summary_c3.md 2>&1; head -8 $out/launch_summary_c3.md
ls -la $out
