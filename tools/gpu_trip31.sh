#!/bin/bash
# final parity run (incl. D-FINE-m with the reference checkpoint) + bench line with the spin-bracketed kernel timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
run smoke       python __graft_entry__.py smoke
run bench_tc3   python bench.py --steps 20 --warmup 5
cat $O/summary.txt
