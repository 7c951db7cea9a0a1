#!/bin/bash
# Round-1 re-measure: tests, smoke, torch profile, ncu launch list, ncu --set full of the top kernels, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider"
run ops         $PT tests/test_ops_gpu.py tests/test_matcher_gpu.py
run model       $PT tests/test_model_gpu.py
run smoke       python __graft_entry__.py smoke
run torchprof   python tools/profile_step.py --torch --eager
run ncu_list    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
run ncu_msda    $NCU -k regex:"msda_(fwd|bwd)_kernel" -c 8 -o $O/prof_r1_msda python tools/profile_step.py --eager
run ncu_tcfwd   $NCU -k regex:"tc_fwd_kernel" -s 30 -c 8 -o $O/prof_r1_tcfwd python tools/profile_step.py --eager
run ncu_tcwgrad $NCU -k regex:"tc_wgrad_kernel" -s 10 -c 6 -o $O/prof_r1_tcwgrad python tools/profile_step.py --eager
run bench       python bench.py --steps 10 --warmup 3
run bench_ref   python bench.py --impl reference --steps 2 --warmup 1
cat $O/summary.txt
