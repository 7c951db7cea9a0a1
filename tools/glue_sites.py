"""Where do the step's non-library launches come from?  Runs one eager train step of a bench.py workload under the torch
profiler with Python stacks and attributes every ATen kernel (add / copy_ / cat / fill_ / ...) to the innermost frame
inside custom_d_fine_b200/ (or autograd for gradient accumulation), summed by call site.

    python tools/glue_sites.py [--config m640] [--top 40]
"""
import argparse
import collections
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="m640")
ap.add_argument("--top", type=int, default=45)
args = ap.parse_args()
B = bench.select_workload(args.config)
dev = torch.device("cuda", 0)
step = bench.build_step(dev, 1, 0, eager=True)
x, l, b = bench.synthetic(B, 1234)
m = bench.rect_masks(b, bench.HW).to(dev) if bench.SEG else None
x, l, b = x.to(dev), l.to(dev), b.to(dev)
targets = bench.to_targets(l, b, m)
for _ in range(3):
    step(x, targets)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step(x, targets)
    torch.cuda.synchronize()

sites = collections.defaultdict(lambda: [0, 0.0])
ops = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
for ev in prof.events():
    if ev.device_type.name != "CPU" or not ev.name.startswith("aten::"):
        continue
    cuda_t = sum(k.duration for k in ev.kernels) if ev.kernels else 0.0
    if cuda_t == 0.0 or ev.cpu_children and any(c.kernels for c in ev.cpu_children):
        continue            # count leaf ops only (the op that actually launched the kernels)
    frame = None
    for f in ev.stack or []:              # innermost frame first
        if "custom_d_fine_b200/" in f:
            frame = f.split("custom_d_fine_b200/")[-1].strip()
            break
    if frame is None:                     # autograd engine thread (no Python frames): the operand shapes identify the tensor
        frame = "autograd " + str([tuple(s) for s in (ev.input_shapes or []) if s])[:100]
    sites[(ev.name, frame)][0] += 1
    sites[(ev.name, frame)][1] += cuda_t
    ops[ev.name][0] += 1
    ops[ev.name][1] += cuda_t
    total += cuda_t
print(f"# {args.config}: ATen leaf ops with device time, one eager step: {total / 1e3:.2f} ms in {sum(v[0] for v in ops.values())} ops")
print("\n## by op")
for k, (n, t) in sorted(ops.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"{k:40s} {n:5d} {t / 1e3:8.3f} ms")
print("\n## by call site")
for (name, frame), (n, t) in sorted(sites.items(), key=lambda kv: -kv[1][1])[:args.top]:
    print(f"{t / 1e3:8.3f} ms {n:5d}x {name:28s} {frame[:110]}")
