#!/bin/bash
# First GPU bring-up: every op family in its own process (a faulting kernel poisons the CUDA context).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 420 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
PT="python -m pytest -x -q -m gpu -p no:cacheprovider"
run msda        $PT tests/test_ops_gpu.py -k "msda"
run matcher     $PT tests/test_matcher_gpu.py
run norm        $PT tests/test_ops_gpu.py -k "layernorm or maxpool"
run attention   $PT tests/test_ops_gpu.py -k "attention"
DFINE_GEMM=simt run conv_simt   $PT tests/test_ops_gpu.py -k "conv"
DFINE_GEMM=simt run linear_simt $PT tests/test_ops_gpu.py -k "linear"
run tc_vs_simt  $PT tests/test_ops_gpu.py -k "tc_matches_simt"
run conv_tc     $PT tests/test_ops_gpu.py -k "conv"
run linear_tc   $PT tests/test_ops_gpu.py -k "linear"
run model_simt  $PT tests/test_model_gpu.py -k "simt"
run model_tc    $PT tests/test_model_gpu.py -k "tc"
run smoke       python __graft_entry__.py smoke
DFINE_GEMM=simt run bench_simt  python bench.py --steps 2 --warmup 3 --no-cpu-baseline
run bench_tc    python bench.py --steps 5 --warmup 3
cat $O/summary.txt
