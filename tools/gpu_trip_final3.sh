#!/bin/bash
# ncu evidence of the final state, summarised ON THE BOX (the .ncu-rep files of five --set full captures exceed the
# 64 MiB return limit): launch list, DRAM traffic per family, full-metric rows of the top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 1200 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
run ncu_list    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
run ncu_traffic ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"tc_fwd_persist|tc_wgrad_kernel|msda_" --csv --log-file $O/traffic.csv python tools/profile_step.py --eager
NCU="ncu --set full --clock-control none --profile-from-start off"
full() { name=$1; shift; run ncu_$name $NCU "$@" -o /tmp/prof_$name python tools/profile_step.py --eager; python tools/ncu_summary.py full /tmp/prof_$name.ncu-rep $O/ncu_full_$name.csv; }
full persist -k regex:"tc_fwd_persist" -s 60 -c 8
full wgrad   -k regex:"tc_wgrad_kernel" -s 20 -c 4
full msda    -k regex:"msda_" -c 8
full attn    -k regex:"^(fwd|dq|dkv)_kernel" -s 3 -c 3
full small   -k regex:"fdr_head|stem_|dwconv3x3s2|bn_finalize_apply|matcher|layernorm_bwd" -c 10
python tools/ncu_summary.py launches $O/launches.csv $O/launches.md
python tools/ncu_summary.py traffic $O/traffic.csv $O/traffic.json
rm -f $O/launches.csv $O/traffic.csv
cat $O/summary.txt
