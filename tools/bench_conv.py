"""Time single conv / linear geometries of D-FINE-m through the host launchers (CUDA events, L2 flushed between
iterations) and report achieved GB/s (algorithmic in+out+weights) and TFLOP/s.  Env switches are read by the
library once per process: DFINE_GEMM, DFINE_TC_PERSIST, DFINE_TC_PREFETCH, DFINE_TC_DBG (10 = loads only,
11 = MMAs only, 12 = weight loads + MMAs, 13 = activation loads + MMAs)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from custom_d_fine_b200 import cuda_ops as co  # noqa: E402

co.CudaOps()
SHAPES = [  # name, B, H, W, Cin, Cout, k, stride, pad
    ("fpn cv4 768->256 1x1 @80", 16, 80, 80, 768, 256, 1, 1, (0, 0, 0, 0)),
    ("csp 128->128 3x3 @80", 16, 80, 80, 128, 128, 3, 1, (1, 1, 1, 1)),
    ("csp 128->128 1x1 @80", 16, 80, 80, 128, 128, 1, 1, (0, 0, 0, 0)),
    ("agg 1280->384 1x1 @40", 16, 40, 40, 1280, 384, 1, 1, (0, 0, 0, 0)),
    ("agg 1792->768 1x1 @20", 16, 20, 20, 1792, 768, 1, 1, (0, 0, 0, 0)),
    ("stage1 32->32 3x3 @160", 16, 160, 160, 32, 32, 3, 1, (1, 1, 1, 1)),
    ("stem2a 24->12 2x2 @320", 16, 320, 320, 24, 12, 2, 1, (0, 0, 1, 1)),
    ("enc_output 256->256 linear 134400 rows", 1, 1, 134400, 256, 256, 1, 1, (0, 0, 0, 0)),
]
if len(sys.argv) > 1:        # extra geometries: BxHxWxCinxCoutxk (stride 1, "same" padding), e.g. 1x1x2000x256x1024x1
    SHAPES = []
    for a in sys.argv[1:]:
        B_, H_, W_, ci_, co_, k_ = [int(v) for v in a.split("x")]
        SHAPES.append((a, B_, H_, W_, ci_, co_, k_, 1, ((k_ - 1) // 2,) * 4))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print({k: os.environ.get(k) for k in ("DFINE_GEMM", "DFINE_TC_PERSIST", "DFINE_TC_PREFETCH", "DFINE_TC_DBG", "DFINE_TMA_TF32")})
for name, B, H, W, Cin, Cout, k, stride, pad in SHAPES:
    OH = (H + pad[0] + pad[2] - k) // stride + 1
    OW = (W + pad[1] + pad[3] - k) // stride + 1
    geom = (B, H, W, Cin, OH, OW, Cout, k, stride, pad)
    x = torch.randn(B, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    dy = torch.randn(B, OH, OW, Cout, device="cuda")
    y = torch.empty(B, OH, OW, Cout, device="cuda")
    dx = torch.empty(B, H, W, Cin, device="cuda")
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    cache = co._WCache()
    fns = {"fwd": lambda: co._conv_fwd(x, Cin, w, cache.getter(w), None, y, Cout, geom, 0, stats),
           "dgrad": lambda: co._conv_dgrad(dy, Cout, w, cache.getter(w), dx, Cin, geom),
           "wgrad": lambda: co._conv_wgrad(dy, Cout, x, Cin, geom)}
    nbytes = 4 * (x.numel() + y.numel() + w.numel())
    flops = 2.0 * B * OH * OW * Cout * Cin * k * k
    out = []
    for fname, fn in fns.items():
        for _ in range(2):
            fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        t = sorted(ts)[len(ts) // 2]
        out.append(f"{fname} {t:7.1f} us {nbytes / t / 1e3:6.0f} GB/s {flops / t / 1e6:6.1f} TF/s")
    print(f"{name:42s} | " + " | ".join(out))
