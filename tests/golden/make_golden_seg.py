"""Segmentation fixture: the UNMODIFIED reference (/root/reference) with the mask
head enabled, run on CPU.  Build container only:  python tests/golden/make_golden_seg.py"""
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
# the reference's `src` is a namespace package: this repo's `src/` shim (a regular package) would shadow it, so the
# repo root must NOT be importable while the reference is imported; the shared helpers are loaded by file path
sys.path = [p for p in sys.path if Path(p or ".").resolve() != HERE.parents[1]]
sys.path.insert(0, "/root/reference")
from src.d_fine.dfine import build_loss, build_model  # noqa: E402
import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("golden_common", HERE / "common.py")
_common = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_common)
seeded_fill, synthetic_batch, rect_masks = _common.seeded_fill, _common.synthetic_batch, _common.rect_masks
assert "/root/reference" in build_model.__code__.co_filename, build_model.__code__.co_filename


def seg_case(size, B, hw, seed):
    torch.manual_seed(0)
    model = build_model(size, 80, True, "cpu", img_size=(hw, hw))
    seeded_fill(model, seed)
    model.train()
    x, targets = synthetic_batch(B, hw, hw, seed=1234 + seed)
    for t in targets:
        t["masks"] = rect_masks(t["boxes"], hw, hw)
    torch.manual_seed(7)
    out = model(x, targets=targets)
    crit = build_loss(size, 80, 0.0, True)
    losses = crit(out, targets)
    sum(losses.values()).backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    with torch.no_grad():
        idx = crit.matcher({k: v for k, v in out.items() if "aux" not in k}, targets)["indices"]
    return {
        "size": size, "B": B, "hw": hw, "seed": seed,
        "losses": {k: float(v) for k, v in losses.items()},
        "pred_logits": out["pred_logits"].detach(), "pred_boxes": out["pred_boxes"].detach(),
        "pred_masks_q0_20": out["pred_masks"].detach()[:, :20].clone(),
        "pred_masks_absmax": float(out["pred_masks"].abs().max()),
        "dn_pred_masks_q0_8": out["dn_pred_masks"].detach()[:, :8].clone(),
        "indices": [(i.clone(), j.clone()) for i, j in idx],
        "grad_norms": {k: float(v.double().norm()) for k, v in grads.items()},
        "grads": {k: grads[k] for k in ("decoder.mask_head.layers.2.weight", "decoder.mask_decoder.up_conv.weight",
                                          "decoder.mask_decoder.lateral.0.weight", "decoder.mask_decoder.bn.1.weight")},
        "mask_keys": sorted(k for k in model.state_dict() if "mask" in k),
    }


if __name__ == "__main__":
    torch.set_num_threads(8)
    if "--x" in sys.argv:       # BASELINE config 4 family (D-FINE-x detect): losses + indices only
        torch.manual_seed(0)
        model = build_model("x", 80, False, "cpu", img_size=(320, 320))
        seeded_fill(model, 11)
        model.train()
        x, targets = synthetic_batch(2, 320, 320, seed=1234 + 11)
        torch.manual_seed(7)
        out = model(x, targets=targets)
        crit = build_loss("x", 80, 0.0, False)
        losses = crit(out, targets)
        with torch.no_grad():
            idx = crit.matcher({k: v for k, v in out.items() if "aux" not in k}, targets)["indices"]
        torch.save({"size": "x", "B": 2, "hw": 320, "seed": 11, "losses": {k: float(v) for k, v in losses.items()},
                    "indices": [(i.clone(), j.clone()) for i, j in idx]}, HERE / "model_x_320.pt")
        print(len(losses), (HERE / "model_x_320.pt").stat().st_size)
        sys.exit(0)
    if "--l" in sys.argv:       # BASELINE config 3 family (D-FINE-l segment): losses + indices only (small fixture)
        fix = seg_case("l", 2, 320, 5)
        small = {k: fix[k] for k in ("size", "B", "hw", "seed", "losses", "indices", "pred_masks_absmax", "mask_keys")}
        small["grad_norms"] = {k: v for k, v in fix["grad_norms"].items() if "mask" in k}
        torch.save(small, HERE / "model_l_seg_320.pt")
        print(len(small["losses"]), (HERE / "model_l_seg_320.pt").stat().st_size)
        sys.exit(0)
    fix = seg_case("s", 2, 320, 2)
    fix["grads"].pop("decoder.mask_decoder.up_conv.weight", None)       # 2.4 MB; its norm is kept
    fix["pred_masks_q0_20"] = fix["pred_masks_q0_20"][:, :10].clone()
    fix["dn_pred_masks_q0_8"] = fix["dn_pred_masks_q0_8"][:, :4].clone()
    torch.save(fix, HERE / "model_s_seg_320.pt")
    print({k: v for k, v in fix["losses"].items() if "mask" in k})
    print(len(fix["losses"]), fix["mask_keys"][:6], (HERE / "model_s_seg_320.pt").stat().st_size)
