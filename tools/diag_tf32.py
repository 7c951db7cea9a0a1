"""Pin down the accuracy of the tcgen05 tf32 paths: exact-representable operands (accumulation error only),
3xTF32 with one / both operands carrying low bits, against an fp64 matmul."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from custom_d_fine_b200 import cuda_ops as co  # noqa: E402

co.CudaOps()


def rn_tf32(t):
    i = t.view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF)
    return r.view(torch.float32)


def run(mode, x, w):
    co.set_gemm_mode(mode)
    M, K = x.shape
    N = w.shape[0]
    y = torch.zeros(M, N, device="cuda")
    geom = (1, 1, M, K, 1, M, N, 1, 1, (0, 0, 0, 0))
    cache = co._WCache()
    co._conv_fwd(x, K, w, cache.getter(w), None, y, N, geom, 0, None)
    torch.cuda.synchronize()
    return y


import os
print("DFINE_TC_PERSIST", os.environ.get("DFINE_TC_PERSIST"), "DFINE_TC_DBG", os.environ.get("DFINE_TC_DBG"))
g = torch.Generator().manual_seed(0)
for (M, K, N) in [(512, 4, 512), (512, 32, 128), (1024, 256, 128), (2048, 1152, 128)]:
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    xr, wr = rn_tf32(x.clone()), rn_tf32(w.clone())

    def err(y, a, b):
        ref = a.double() @ b.double().t()
        return float((y.double() - ref).abs().max() / ref.abs().max())

    print(f"M{M} K{K} N{N}: "
          f"tc exact-operands {err(run('tc', xr, wr), xr, wr):.2e} | "
          f"tc general {err(run('tc', x, w), x, w):.2e} | "
          f"tc3 exact-operands {err(run('tc3', xr, wr), xr, wr):.2e} | "
          f"tc3 x-low-bits only {err(run('tc3', x, wr), x, wr):.2e} | "
          f"tc3 w-low-bits only {err(run('tc3', xr, w), xr, w):.2e} | "
          f"tc3 general {err(run('tc3', x, w), x, w):.2e} | "
          f"bf3 general {err(run('bf3', x, w), x, w):.2e} | hf3 general {err(run('hf3', x, w), x, w):.2e} | "
          f"hf3 small-x (x*1e-3) {err(run('hf3', x * 1e-3, w), x * 1e-3, w):.2e} | tc3 small-x {err(run('tc3', x * 1e-3, w), x * 1e-3, w):.2e} | tch general {err(run('tch', x, w), x, w):.2e} | "
          f"simt {err(run('simt', x, w), x, w):.2e} | tc3 out absmax {float(run('tc3', x, w).abs().max()):.3e} "
          f"ref absmax {float((x.double() @ w.double().t()).abs().max()):.3e}")
