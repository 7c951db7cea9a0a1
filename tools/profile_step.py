"""One train step of a bench.py workload (default D-FINE-m, batch 16, 640x640) bracketed by cudaProfilerStart/Stop for
`ncu --profile-from-start off`, or summarised with torch.profiler (--torch)."""
import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--torch", action="store_true")
ap.add_argument("--config", default="m640", choices=sorted(bench.WORKLOADS))
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--graph", action="store_true", help="CUDA-graph replay instead of eager launches")
ap.add_argument("--out", default="gpurun_out/torch_profile.txt")
args = ap.parse_args()

args.batch = bench.select_workload(args.config, args.batch)
dev = torch.device("cuda", 0)
step = bench.build_step(dev, 1, 0, eager=not args.graph)
x, l, b = bench.synthetic(args.batch, 1234)
m = bench.rect_masks(b, bench.HW).to(dev) if bench.SEG else None
x, l, b = x.to(dev), l.to(dev), b.to(dev)
targets = bench.to_targets(l, b, m)
for _ in range(6 if args.graph else 3):
    step(x, targets)
torch.cuda.synchronize()
if args.torch:
    from torch.profiler import ProfilerActivity, profile
    t0 = time.perf_counter()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step(x, targets)
        torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    with open(args.out, "w") as f:
        f.write(f"wall {wall * 1e3:.1f} ms (with profiler overhead)\n")
        f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=70, max_name_column_width=90))
else:
    torch.cuda.cudart().cudaProfilerStart()
    step(x, targets)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
