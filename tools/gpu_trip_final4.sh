#!/bin/bash
# --set full captures of the top kernels, summarised ON THE BOX into small csv files (the reports exceed the return limit)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
NCU="ncu --set full --clock-control none --profile-from-start off"
full() {
    local tag=$1; shift
    echo "=== ncu_$tag" | tee -a $O/summary.txt
    timeout 900 $NCU "$@" -o /tmp/prof_$tag python tools/profile_step.py --eager > $O/ncu_$tag.log 2>&1
    echo "rc=$? $(tail -n 1 $O/ncu_$tag.log)" | tee -a $O/summary.txt
    python tools/ncu_summary.py full /tmp/prof_$tag.ncu-rep $O/ncu_full_$tag.csv
    wc -c $O/ncu_full_$tag.csv | tee -a $O/summary.txt
}
full persist -k regex:"tc_fwd_persist" -s 60 -c 6
full wgrad   -k regex:"tc_wgrad_kernel" -s 20 -c 3
full msda    -k regex:"msda_" -c 4
full attn    -k regex:"^(fwd|dq|dkv)_kernel" -s 3 -c 3
full small   -k regex:"fdr_head|stem_|dwconv3x3s2|matcher|layernorm_bwd" -c 8
cat $O/summary.txt
