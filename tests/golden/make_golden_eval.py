"""Inference fixture: the UNMODIFIED reference (/root/reference) in eval mode after `deploy()` (dfine.py:43-49:
RepVGG / ConvNormLayer_fuse re-parameterisation, decoder truncated to eval_idx), detect and segment heads.
Build container only:  python tests/golden/make_golden_eval.py"""
import importlib.util
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path = [p for p in sys.path if Path(p or ".").resolve() != HERE.parents[1]]   # see make_golden_seg.py
sys.path.insert(0, "/root/reference")
from src.d_fine.dfine import build_model  # noqa: E402

_spec = importlib.util.spec_from_file_location("golden_common", HERE / "common.py")
_common = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_common)
assert "/root/reference" in build_model.__code__.co_filename


def eval_case(size, seg, B, hw, seed):
    torch.manual_seed(0)
    model = build_model(size, 80, seg, "cpu", img_size=(hw, hw))
    _common.seeded_fill(model, seed)
    model.eval()
    x, _ = _common.synthetic_batch(B, hw, hw, seed=1234 + seed)
    with torch.no_grad():
        plain = model(x)
        out = model.deploy()(x)
    fix = {"size": size, "seg": seg, "B": B, "hw": hw, "seed": seed, "keys": sorted(out.keys()),
           "pred_logits": out["pred_logits"], "pred_boxes": out["pred_boxes"],
           "deploy_vs_plain": float((out["pred_logits"] - plain["pred_logits"]).abs().max())}
    if seg:
        fix["pred_masks_q0_6"] = out["pred_masks"][:, :6].clone()
    return fix


if __name__ == "__main__":
    torch.set_num_threads(8)
    cases = [eval_case("s", False, 2, 320, 3), eval_case("s", True, 1, 320, 4)]
    torch.save(cases, HERE / "eval_s_320.pt")
    print([(c["keys"], c["deploy_vs_plain"]) for c in cases], (HERE / "eval_s_320.pt").stat().st_size)
