"""Informational (last line): the UNMODIFIED reference (baseline/_ref/src/d_fine, see
tools/install_reference.py) timed on ONE B200 through its own stock PyTorch path — eager fp32 (torch defaults: cuDNN
tf32 convolutions, fp32 matmuls) and AMP fp16 (autocast + GradScaler, the shipped default, train.py:570-576) — on the
headline workload (D-FINE-m, batch 16, 640x640, 10 boxes per image): forward, criterion, backward, clip_grad_norm_,
AdamW step, zero_grad (no EMA: ModelEMA lives in the un-importable src/dl/train.py).  Not a bench.py arm; the output
goes to profiles/ as the GPU-side comparison for this repo's `value`.

    python tools/ref_on_gpu.py [--config m640] [--steps 10]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="m640")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=5)
args = ap.parse_args()
B = bench.select_workload(args.config)
REF = ROOT / "baseline" / "_ref"
for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
    del sys.modules[k]
sys.path[:] = [str(REF)] + [q for q in sys.path if q not in ("", str(ROOT))]
try:
    from loguru import logger
    logger.remove()
except Exception:  # noqa: BLE001
    pass
from src.d_fine.dfine import build_loss, build_model, build_optimizer  # noqa: E402

sys.path.insert(1, str(ROOT))
dev = torch.device("cuda", 0)
res = {"config": args.config, "batch": B, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
for amp in (False, True):
    torch.manual_seed(0)
    ck = REF / f"dfine_{bench.MODEL}_coco.pth"
    use_ck = ck.exists() and not bench.SEG           # the reference's own COCO checkpoint where it ships one (m)
    model = build_model(bench.MODEL, 80, bench.SEG, "cuda", img_size=(bench.HW, bench.HW),
                        pretrained_model_path=str(ck) if use_ck else None)
    res["weights"] = "pretrained COCO checkpoint" if use_ck else "seeded default init, zero heads randomised"
    if not use_ck:
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for p in model.parameters():
                if p.dim() >= 2 and float(p.abs().max()) == 0.0:
                    p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))
    model.train()
    loss_fn = build_loss(bench.MODEL, 80, 0.0, bench.SEG)
    opt = build_optimizer(model, lr=1.5e-4, backbone_lr=2e-5, betas=(0.9, 0.999), weight_decay=1.25e-4, base_lr=1.5e-4)
    scaler = torch.amp.GradScaler("cuda", enabled=amp)
    x, l, b = bench.synthetic(B, 1234)
    masks = bench.rect_masks(b, bench.HW) if bench.SEG else None
    x = x.to(dev)
    targets = bench.to_targets(l.to(dev), b.to(dev), masks.to(dev) if masks is not None else None)

    def step():
        if amp:
            with torch.autocast("cuda", cache_enabled=True):
                out = model(x, targets=targets)
            with torch.autocast("cuda", enabled=False):
                loss = sum(loss_fn(out, targets).values())
            scaler.scale(loss).backward()
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
            scaler.step(opt)
            scaler.update()
        else:
            out = model(x, targets=targets)
            loss = sum(loss_fn(out, targets).values())
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
            opt.step()
        opt.zero_grad()
        return loss

    key = "amp_fp16" if amp else "eager_fp32_tf32conv"
    try:
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            step()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / args.steps
        res[key] = {"ms_per_step": round(ms, 2), "images_per_s": round(B / ms * 1e3, 1)}
    except Exception as ex:  # noqa: BLE001  (fp16 autocast overflows on an untrained network trip the matcher's box assert)
        res[key] = {"failed": f"{type(ex).__name__}: {str(ex)[:200]}"}
    print(key, res[key], flush=True)
    del model, opt, loss_fn
    torch.cuda.empty_cache()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"ref_on_gpu_{args.config}.json").write_text(json.dumps(res, indent=1))
print(json.dumps(res))
