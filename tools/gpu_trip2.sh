#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 600 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
run norm        $PT tests/test_ops_gpu.py -k "layernorm or maxpool"
DFINE_GEMM=simt run conv_simt   $PT tests/test_ops_gpu.py -k "conv"
run tc_vs_simt  $PT tests/test_ops_gpu.py -k "tc_matches_simt"
run conv_tc     $PT tests/test_ops_gpu.py -k "conv"
run linear_tc   $PT tests/test_ops_gpu.py -k "linear"
run model       $PT tests/test_model_gpu.py
run smoke       python __graft_entry__.py smoke
run torchprof   python tools/profile_step.py --torch
run ncu_list    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py
run bench_tc    python bench.py --steps 10 --warmup 3
cat $O/summary.txt
