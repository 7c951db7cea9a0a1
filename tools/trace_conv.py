"""Phase timeline of the 3xFP16 forward kernel (tc_fwd_ts) on one geometry: dfine_tc_trace makes every CTA write
globaltimer stamps at its phase boundaries; two launches back to back show the launch gap between dependent kernels.

    python tools/trace_conv.py --shape 16x40x40x128x128x1 [--shape ...]
"""
import argparse
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from custom_d_fine_b200 import cuda_ops as co  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", action="append", default=[], help="BxHxWxCinxCoutxk")
args = ap.parse_args()
co.CudaOps()
SLOTS = ["entry", "prologue", "dep-wait", "tma0", "landed0", "planes0", "acc0", "tile0-st", "tileN-st", "roles-done",
         "stats", "finalize", "e:tmem-ld", "e:sts", "e:stores", "e:stats"]
lib = co.lib()
for shape in args.shape or ["16x40x40x128x128x1"]:
    B, H, W, Cin, Cout, k = [int(v) for v in shape.split("x")]
    pad = ((k - 1) // 2,) * 4
    geom = (B, H, W, Cin, H, W, Cout, k, 1, pad)
    x = torch.randn(B, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    y = torch.empty(B, H, W, Cout, device="cuda")
    cache = co._WCache()
    bn = [torch.ones(Cout, device="cuda"), torch.zeros(Cout, device="cuda"), torch.zeros(Cout, device="cuda"),
          torch.ones(Cout, device="cuda")] + [torch.empty(Cout, device="cuda") for _ in range(4)]

    def run(buf):
        stats = torch.zeros(2 * Cout + 1, dtype=torch.float64, device="cuda")
        lib.dfine_tc_trace(ctypes.c_void_p(buf.data_ptr() if buf is not None else 0))
        co._conv_fwd(x, Cin, w, cache.getter(w), None, y, Cout, geom, 0, stats,
                     bn_fin=(stats[2 * Cout:], bn[0], bn[1], bn[2], bn[3], bn[4], bn[5], bn[6], bn[7], 0.1, 1e-5))

    for _ in range(3):
        run(None)
    bufs = [torch.zeros(256 * 16, dtype=torch.int64, device="cuda") for _ in range(3)]
    torch.cuda.synchronize()
    for b in bufs:
        run(b)
    lib.dfine_tc_trace(ctypes.c_void_p(0))
    torch.cuda.synchronize()
    t = [b.cpu().view(256, 16) for b in bufs]
    base = int(t[0][:, 0][t[0][:, 0] > 0].min())
    print(f"== {shape}: times in us relative to the first launch's first CTA entry; columns min / median / max over CTAs")
    for li, tb in enumerate(t):
        live = tb[:, 0] > 0
        print(f" launch {li}: {int(live.sum())} CTAs")
        for si, name in enumerate(SLOTS):
            v = tb[live, si]
            v = v[v > 0].double()
            if v.numel() == 0:
                continue
            v = (v - base) / 1e3
            print(f"   {name:11s} {v.min():8.2f} {v.median():8.2f} {v.max():8.2f}   (n={v.numel()})")
