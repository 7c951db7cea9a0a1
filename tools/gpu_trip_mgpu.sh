#!/bin/bash
# 2-GPU check of the data-parallel path: torchrun bench (NCCL all-reduce of the flat gradient arenas between graphs).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary_mgpu.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary_mgpu.txt; }
rm -f $O/summary_mgpu.txt
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt 2>&1
run bench_n$N python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3
run bench_n1 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline
run bench_ref_n$N python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1
cat $O/summary_mgpu.txt
