#!/bin/bash
# Validate: generalized tcgen05 kernels (stride 2 / taps / 3xTF32), fused optimizer, 3-graph step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
run tc          $PT tests/test_ops_gpu.py -k "tc_matches_simt"
run optim       $PT tests/test_optim_gpu.py
run ops         $PT tests/test_ops_gpu.py tests/test_matcher_gpu.py -k "not tc_matches_simt"
run model       $PT tests/test_model_gpu.py
run smoke       python __graft_entry__.py smoke
run bench_tc3   python bench.py --steps 10 --warmup 3 --no-cpu-baseline
DFINE_GEMM=tc run bench_tc python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run ncu_list    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
cat $O/summary.txt
