#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
DFINE_TC_PERSIST=0 DFINE_TC_DBG=0 run diag_np0 python tools/diag_tf32.py
DFINE_TC_PERSIST=0 DFINE_TC_DBG=1 run diag_np1 python tools/diag_tf32.py
DFINE_TC_PERSIST=0 DFINE_TC_DBG=3 run diag_np3 python tools/diag_tf32.py
DFINE_TC_PERSIST=1 DFINE_TC_DBG=0 run diag_p0 python tools/diag_tf32.py
run tc          $PT tests/test_ops_gpu.py -k "tc_matches_simt"
run tests       $PT tests --deselect tests/test_ops_gpu.py::test_tc_matches_simt
run smoke       python __graft_entry__.py smoke
DFINE_GEMM=tc run bench_tc python bench.py --steps 10 --warmup 3 --no-cpu-baseline
DFINE_GEMM=tc DFINE_TC_PERSIST=0 run bench_tc_np python bench.py --steps 10 --warmup 3 --no-cpu-baseline
DFINE_GEMM=tc run ncu_list    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
cat $O/summary.txt
