#!/bin/bash
# End-of-round evidence run: tests, smoke, both bench arms, launch list, DRAM traffic of the dominant kernel family,
# --set full captures of the top kernels.  Summaries are converted into profiles/ by tools/ncu_summary.py here.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 1200 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
run smoke       python __graft_entry__.py smoke
run bench_ref   python bench.py --impl reference --steps 2 --warmup 1
run bench_tc3   python bench.py --steps 20 --warmup 3
DFINE_GEMM=tc run bench_tc python bench.py --steps 20 --warmup 3 --no-cpu-baseline
run ncu_list    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
run ncu_traffic ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"tc_fwd_persist|tc_wgrad_kernel|msda_" --csv --log-file $O/traffic.csv python tools/profile_step.py --eager
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
run ncu_persist $NCU -k regex:"tc_fwd_persist" -s 60 -c 8 -o $O/prof_final_persist python tools/profile_step.py --eager
run ncu_wgrad   $NCU -k regex:"tc_wgrad_kernel" -s 20 -c 4 -o $O/prof_final_wgrad python tools/profile_step.py --eager
run ncu_msda    $NCU -k regex:"msda_" -c 4 -o $O/prof_final_msda python tools/profile_step.py --eager
DFINE_GEMM=tc run conv_tc   python tools/bench_conv.py
run conv_tc3   python tools/bench_conv.py
cat $O/summary.txt
