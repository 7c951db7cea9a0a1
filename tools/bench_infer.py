"""Inference throughput / latency of the deployed model on one B200 — the reference's only published numbers are for
this path (README.md:105-172: D-FINE-m 640x640, Torch FP32 latency and a batch sweep on an RTX 5070 Ti), so this puts
like beside like: `ours` = build_model(...).deploy() (conv+BN folded, RepVGG merged, decoder truncated) + the device
post-processor, replayed from a CUDA graph, timed end to end from a pinned uint8 host batch (H2D + uint8->float kernel
+ forward + post-process + D2H of labels / boxes / scores); `reference` = the unmodified reference model from
baseline/_ref in eval mode on the same GPU (eager fp32, torch defaults) with the same post-processing in torch ops.

    python tools/bench_infer.py [--size m] [--hw 640] [--batches 1,2,4,8,16,32] [--impl ours|reference|both]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser()
ap.add_argument("--size", default="m")
ap.add_argument("--hw", type=int, default=640)
ap.add_argument("--batches", default="1,2,4,8,16,32")
ap.add_argument("--impl", default="both")
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
dev = torch.device("cuda", 0)
res = {"size": args.size, "hw": args.hw, "gpu": torch.cuda.get_device_name(0), "rows": []}


def timeit(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


def ours(B):
    from custom_d_fine_b200.model import build_model
    from custom_d_fine_b200.postprocess import DFINEPostProcessor, prepare_inputs
    ck = ROOT / "baseline" / "_ref" / f"dfine_{args.size}_coco.pth"
    torch.manual_seed(0)
    model = build_model(args.size, 80, False, "cuda", img_size=(args.hw, args.hw),
                        pretrained_model_path=str(ck) if ck.exists() else None).deploy()
    post = DFINEPostProcessor(80)
    host = torch.randint(0, 256, (B, args.hw, args.hw, 3), dtype=torch.uint8).pin_memory()
    sx = host.to(dev)
    out_host = [torch.empty((B, 300), dtype=torch.int64).pin_memory(), torch.empty((B, 300, 4)).pin_memory(),
                torch.empty((B, 300)).pin_memory()]

    def fwd():
        with torch.no_grad():
            x = prepare_inputs(sx, None, bgr=True)
            return post(model(x), args.hw, args.hw)

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            outs = fwd()
    torch.cuda.synchronize()

    def e2e():
        sx.copy_(host, non_blocking=True)
        g.replay()
        for h, o in zip(out_host, outs):
            h.copy_(o, non_blocking=True)
        torch.cuda.synchronize()

    def resident():
        g.replay()

    return timeit(e2e, args.iters), timeit(resident, args.iters)


def reference(B):
    ref = ROOT / "baseline" / "_ref"
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    keep = list(sys.path)
    sys.path[:] = [str(ref)] + [q for q in sys.path if q not in ("", str(ROOT))]
    try:
        try:
            from loguru import logger
            logger.remove()
        except Exception:  # noqa: BLE001
            pass
        from src.d_fine.dfine import build_model
    finally:
        sys.path[:] = keep
    ck = ref / f"dfine_{args.size}_coco.pth"
    model = build_model(args.size, 80, False, "cuda", img_size=(args.hw, args.hw),
                        pretrained_model_path=str(ck) if ck.exists() else None).eval()
    host = torch.randint(0, 256, (B, 3, args.hw, args.hw), dtype=torch.uint8).pin_memory()

    def e2e():
        with torch.no_grad():
            x = host.to(dev, non_blocking=True).float().div_(255.0)          # infer/torch_model.py:283-286
            out = model(x)
            flat = torch.sigmoid(out["pred_logits"]).flatten(1)
            sc, idx = torch.topk(flat, 300, dim=-1)
            lab, q = idx % 80, idx // 80
            bx = out["pred_boxes"].gather(1, q[..., None].expand(-1, -1, 4))
            r = (lab.cpu(), bx.cpu(), sc.cpu())
        torch.cuda.synchronize()
        return r

    return timeit(e2e, max(args.iters // 2, 10)), None


for B in [int(b) for b in args.batches.split(",")]:
    row = {"batch": B}
    if args.impl in ("ours", "both"):
        e, r = ours(B)
        row["ours"] = {"e2e_ms": round(e, 3), "e2e_img_s": round(B / e * 1e3, 1), "resident_ms": round(r, 3),
                       "resident_img_s": round(B / r * 1e3, 1)}
    if args.impl in ("reference", "both"):
        e, _ = reference(B)
        row["reference_eager_fp32"] = {"e2e_ms": round(e, 3), "e2e_img_s": round(B / e * 1e3, 1)}
    print(json.dumps(row), flush=True)
    res["rows"].append(row)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"bench_infer_{args.size}{args.hw}.json").write_text(json.dumps(res, indent=1))
