"""One conv / linear geometry through the host launchers a few times — the target of `ncu --set full` source-level captures.

    python tools/one_conv.py --shape 16x80x80x768x256x1 --what fwd --mode hf3
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from custom_d_fine_b200 import cuda_ops as co  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="16x80x80x768x256x1", help="BxHxWxCinxCoutxk")
ap.add_argument("--what", default="fwd", choices=["fwd", "dgrad", "wgrad"])
ap.add_argument("--mode", default="hf3")
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
co.CudaOps()
co.set_gemm_mode(args.mode)
B, H, W, Cin, Cout, k = [int(v) for v in args.shape.split("x")]
pad = ((k - 1) // 2,) * 4
geom = (B, H, W, Cin, H, W, Cout, k, 1, pad)
x = torch.randn(B, H, W, Cin, device="cuda")
w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
dy = torch.randn(B, H, W, Cout, device="cuda")
y = torch.empty(B, H, W, Cout, device="cuda")
dx = torch.empty(B, H, W, Cin, device="cuda")
stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
cache = co._WCache()
fn = {"fwd": lambda: co._conv_fwd(x, Cin, w, cache.getter(w), None, y, Cout, geom, 0, stats),
      "dgrad": lambda: co._conv_dgrad(dy, Cout, w, cache.getter(w), dx, Cin, geom),
      "wgrad": lambda: co._conv_wgrad(dy, Cout, x, Cin, geom)}[args.what]
for _ in range(args.iters):
    fn()
torch.cuda.synchronize()
print("done")
