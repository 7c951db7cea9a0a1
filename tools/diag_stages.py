"""Module-by-module forward comparison of the CUDA library against the CPU oracle on the same seeded weights and batch
(forward hooks on every leaf block of the host graph): prints the relative L2 error of each module's output in
execution order, so a parity loss at a given size can be attributed to a layer (kernel bug) or seen to grow smoothly
(rounding amplification).

    python tools/diag_stages.py --size x --hw 1280 --batch 1 --mode tc3 [--all]
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from custom_d_fine_b200 import cuda_ops as co  # noqa: E402
from custom_d_fine_b200 import kernels  # noqa: E402
from custom_d_fine_b200.blocks import ConvUnit  # noqa: E402
from custom_d_fine_b200.decoder import DecoderLayer  # noqa: E402
from custom_d_fine_b200.encoder import AIFILayer  # noqa: E402
from custom_d_fine_b200.model import build_model  # noqa: E402
from oracle.torch_ops import OracleOps  # noqa: E402
from tests.golden.common import seeded_fill, synthetic_batch  # noqa: E402
from tests.test_model_gpu import _host_rng  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="x")
ap.add_argument("--hw", type=int, default=1280)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--mode", default="tc3")
ap.add_argument("--seed", type=int, default=11)
ap.add_argument("--all", action="store_true", help="print every module (default: only error increases > 1.5x and the tail)")
args = ap.parse_args()

x, targets = synthetic_batch(args.batch, args.hw, args.hw, seed=1234 + args.seed, T=(10, 7, 3, 10))
co.set_gemm_mode(args.mode)
outs = {}
for dev in ("cpu", "cuda"):
    torch.manual_seed(0)
    model = build_model(args.size, 80, False, dev, img_size=(args.hw, args.hw))
    seeded_fill(model, args.seed)
    model.train()
    rec = []

    def hook(name, rec=rec):
        def f(mod, inp, out):
            o = out[0] if isinstance(out, tuple) else out
            if isinstance(o, (list, tuple)):
                o = o[-1]
            if torch.is_tensor(o):
                rec.append((name, o.detach().float().cpu()))
        return f

    for name, m in model.named_modules():
        if isinstance(m, (ConvUnit, AIFILayer, DecoderLayer)):
            m.register_forward_hook(hook(name))
    xs = x.to(dev)
    tg = [{k: v.to(dev) for k, v in t.items()} for t in targets]
    torch.manual_seed(7)
    with _host_rng(), torch.no_grad():
        if dev == "cpu":
            with kernels.use(OracleOps()):
                out = model(xs, targets=tg)
        else:
            out = model(xs, targets=tg)
            torch.cuda.synchronize()
    outs[dev] = (rec, {k: v.detach().float().cpu() for k, v in out.items() if torch.is_tensor(v)})

(rc, oc), (rg, og) = outs["cpu"], outs["cuda"]
assert [n for n, _ in rc] == [n for n, _ in rg]
print(f"# {args.size} {args.hw}x{args.hw} batch {args.batch} mode {args.mode}: {len(rc)} module outputs")
prev = 0.0
for i, ((n, a), (_, b)) in enumerate(zip(rc, rg)):
    e = float((b.double() - a.double()).norm() / a.double().norm().clamp_min(1e-30))
    mx = float((b - a).abs().max() / a.abs().max().clamp_min(1e-30))
    flag = "  <<<" if e > 1.5 * max(prev, 1e-7) else ""
    if args.all or flag or i >= len(rc) - 12:
        print(f"{i:4d} {n:60s} {tuple(a.shape)!s:24s} relL2 {e:.2e} max {mx:.2e}{flag}")
    prev = max(prev, e)
for k in ("pred_logits", "pred_boxes"):
    a, b = oc[k], og[k]
    print(k, "relL2 (rows unpaired)", float((b.double() - a.double()).norm() / a.double().norm()))
