#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
run tests       $PT tests
run smoke       python __graft_entry__.py smoke
run bench_tc3   python bench.py --steps 10 --warmup 3 --no-cpu-baseline
DFINE_GEMM=tc run bench_tc python bench.py --steps 10 --warmup 3 --no-cpu-baseline
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
DFINE_GEMM=tc run ncu_p128 $NCU -k regex:"tc_fwd_persist" -s 40 -c 6 -o $O/prof_r1_persist python tools/profile_step.py --eager
DFINE_GEMM=tc run ncu_wgrad $NCU -k regex:"tc_wgrad_kernel" -s 20 -c 3 -o $O/prof_r1_wgrad2 python tools/profile_step.py --eager
run ncu_list    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
cat $O/summary.txt
