#!/bin/bash
# A/B of experiment libraries (GWN_VARIANT builds): step time and the live kernel rooflines of bench.py for each.   usage: gpu_lib_ab.sh "" park0 park2
cd "$(dirname "$0")/.."
for v in "$@"; do
  lib=multimodal_outage_b200/libgwn${v:+_$v}.so
  GWN_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu --no-extra > gpurun_out/bench_ab_${v:-base}.log 2>&1
  python - "$v" <<PY
import json, sys
v = sys.argv[1] or 'base'
try:
    d = json.loads(open(f'gpurun_out/bench_ab_{v}.log').read().strip().splitlines()[-1])
    print(v, 'ms/step %.4f' % d['ms_per_step'], ' '.join('%s=%.1fus' % (k[9:] or 'bwd', d[k]['ms_per_launch'] * 1e3) for k in d if k.startswith('roofline')))
except Exception as e:
    print(v, 'failed', e)
PY
done
