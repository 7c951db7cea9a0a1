"""The step when the per-image target counts do NOT repeat (a real loader): GraphedTrainStep keys its full-step CUDA graphs
on (input shape, per-image counts), so such steps take the HYBRID path — backbone + encoder forward / backward replayed as
graphs captured once per input shape, decoder / matcher / criterion eager (DFINE_HYBRID_GRAPH=0: fully eager).
Times D-FINE-m, batch 16, 640x640 with counts drawn uniformly from 1..20 anew every step, next to the fixed-count
(10 per image) graph-replay number that bench.py reports.

    python tools/bench_variable_targets.py [--steps 20]
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
B = bench.select_workload("m640")
dev = torch.device("cuda", 0)
step = bench.build_step(dev, 1, 0)
g = torch.Generator().manual_seed(7)
x = torch.rand(B, 3, bench.HW, bench.HW, generator=g).to(dev)


def targets():
    out = []
    for _ in range(B):
        n = int(torch.randint(1, 21, (1,), generator=g))
        cxcy = torch.rand(n, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(n, 2, generator=g) * 0.25 + 0.05
        out.append({"labels": torch.randint(0, bench.NUM_CLASSES, (n,), generator=g).to(dev),
                    "boxes": torch.cat([cxcy, wh], -1).to(dev)})
    return out


def run(n, make):
    batches = [make() for _ in range(n)]
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for t in batches:
        step(x, t)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


fixed = targets()
for _ in range(6):            # eager warm-up + capture of the fixed key
    step(x, fixed)
ms_fixed = run(args.steps, lambda: fixed)
run(3, targets)
ms_var = run(args.steps, targets)
print(json.dumps({"workload": "D-FINE-m train step, batch 16, 640x640", "fixed_counts_graph_replay_ms": round(ms_fixed, 2),
                  "fixed_counts_img_s": round(B / ms_fixed * 1e3, 1), "variable_counts_ms": round(ms_var, 2),
                  "variable_counts_img_s": round(B / ms_var * 1e3, 1), "steps": args.steps,
                  "hybrid": os.environ.get("DFINE_HYBRID_GRAPH", "1") != "0",
                  "note": "counts ~ U{1..20} per image, new every step: no full-step graph key repeats"}))
