#!/bin/bash
# Evidence capture for profiles/ (one GPU call): launch list of the default bench + `ncu --set full` of the four main
# kernels at their config-2 layer-0 launches (third eager warm-up step of bench.py), raw pages exported as CSV.
cd "$(dirname "$0")/.."
tag=${1:-r2}
out=gpurun_out/prof_$tag
mkdir -p $out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-roofline --no-extra"
timeout 300 $B > $out/plain.log 2>&1 || { echo "plain run failed"; tail -n 5 $out/plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches.csv $B > $out/ncu_list.log 2>&1
python scripts/launch_summary.py $out/launches.csv 45 > $out/launch_summary.md 2>&1
# (kernel regex, launches to skip: layer 0 of the third eager step - forward kernels run layer 0 first, backward last)
cap() { timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c 1 -o $out/$3 $B > $out/ncu_$3.log 2>&1; \
        ncu -i $out/$3.ncu-rep --page raw --csv > $out/$3_raw.csv 2>/dev/null; \
        ncu -i $out/$3.ncu-rep --page details --csv > $out/$3_details.csv 2>/dev/null; rm -f $out/$3.ncu-rep; }
cap "gcn_bwd_t_kernel" 20 gcn_bwd_t_L0
cap "gcn_fwd_t_kernel" 16 gcn_fwd_t_L0
# (-k matches the base name: 16 pos_gemm_tc launches per step - 8 gate forwards, then 8 gate data gradients, layer 0 last)
cap "pos_gemm_tc_kernel" 32 gate_fwd_L0
cap "pos_gemm_tc_kernel" 47 gate_bwd_L0
ls -la $out
