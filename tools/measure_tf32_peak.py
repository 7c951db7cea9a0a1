"""Measured tf32 tensor peak of this pool's B200, the way MEASURED_PEAKS.json measures bf16: cuBLAS 8192^3 through
torch.matmul with tf32 enabled — best of 10 (burst) and back to back for 4 s (sustained, under the power cap).
Writes gpurun_out/tf32_peak.json (copied to profiles/r2_tf32_peak.json; bench.py's by-bound split reads it)."""
import json
import time
from pathlib import Path

import torch

torch.backends.cuda.matmul.allow_tf32 = True
N = 8192
a = torch.randn(N, N, device="cuda")
b = torch.randn(N, N, device="cuda")
flops = 2.0 * N ** 3
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    a @ b
    e.record()
    torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n, t0 = 0, time.time()
s.record()
while time.time() - t0 < 4.0:
    for _ in range(20):
        a @ b
    n += 20
    torch.cuda.synchronize()
e.record()
torch.cuda.synchronize()
sus = s.elapsed_time(e) / n
# bf16 the same way, same box, for the ratio
a16, b16 = a.bfloat16(), b.bfloat16()
for _ in range(3):
    a16 @ b16
torch.cuda.synchronize()
best16 = 1e9
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    a16 @ b16
    e.record()
    torch.cuda.synchronize()
    best16 = min(best16, s.elapsed_time(e))
out = {"tf32_tflops": round(flops / best / 1e9, 1), "tf32_tflops_sustained": round(flops / sus / 1e9, 1),
       "bf16_tflops_same_box": round(flops / best16 / 1e9, 1), "gpu": torch.cuda.get_device_name(0),
       "how": "torch.matmul fp32 inputs with allow_tf32 (cuBLAS tf32), 8192^3: best of 10 (burst) and back to back for 4 s"}
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/tf32_peak.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out))
