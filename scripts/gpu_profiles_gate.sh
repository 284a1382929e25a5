cd "$(dirname "$0")/.."
out=gpurun_out/prof_${1:-r2g}
mkdir -p $out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-roofline --no-extra"
cap() { timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c 1 -o $out/$3 $B > $out/ncu_$3.log 2>&1; \
        ncu -i $out/$3.ncu-rep --page raw --csv > $out/$3_raw.csv 2>/dev/null; \
        ncu -i $out/$3.ncu-rep --page details --csv > $out/$3_details.csv 2>/dev/null; rm -f $out/$3.ncu-rep; }
timeout 300 $B > $out/plain.log 2>&1 && cap "pos_gemm_tc_kernel" 32 gate_fwd_L0 && cap "pos_gemm_tc_kernel" 47 gate_bwd_L0
ls -la $out | grep gate
