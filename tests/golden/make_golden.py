"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python tests/golden/make_golden.py
The fixtures pin oracle/ (and the host graph that drives it) to the real reference; they are small
(outputs only — weights and inputs are regenerated from seeds by tests/golden/common.py).
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, "/root/reference")
from src.d_fine.arch.utils import deformable_attention_core_func_v2  # noqa: E402
from src.d_fine.dfine import build_loss, build_model  # noqa: E402
from src.d_fine.matcher import HungarianMatcher  # noqa: E402
from tests.golden.common import seeded_fill, synthetic_batch  # noqa: E402


def model_case(size, B, hw, seed):
    torch.manual_seed(0)
    model = build_model(size, 80, False, "cpu", img_size=(hw, hw))
    seeded_fill(model, seed)
    model.train()
    x, targets = synthetic_batch(B, hw, hw, seed=1234 + seed)
    torch.manual_seed(7)
    out = model(x, targets=targets)
    crit = build_loss(size, 80, 0.0, False)
    losses = crit(out, targets)
    total = sum(losses.values())
    total.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    keep = ["backbone.stem.stem1.conv.weight", "backbone.stages.1.blocks.0.aggregation.1.conv.weight",
            "encoder.fpn_blocks.0.cv4.conv.weight", "encoder.encoder.0.layers.0.self_attn.in_proj_weight",
            "decoder.decoder.layers.0.cross_attn.sampling_offsets.weight", "decoder.decoder.layers.1.linear1.weight",
            "decoder.dec_bbox_head.1.layers.2.weight", "decoder.enc_score_head.weight",
            "decoder.denoising_class_embed.weight", "decoder.query_pos_head.layers.0.weight"]
    gnorm = {k: float(v.double().norm()) for k, v in grads.items()}
    with torch.no_grad():
        m = HungarianMatcher(weight_dict={"cost_class": 2, "cost_bbox": 5, "cost_giou": 2}, use_focal_loss=True,
                             alpha=0.25, gamma=2.0)
        idx = m({"pred_logits": out["pred_logits"], "pred_boxes": out["pred_boxes"]}, targets)["indices"]
    n_dn = out["dn_meta"]["dn_num_split"][0]
    fix = {
        "size": size, "B": B, "hw": hw, "seed": seed,
        "losses": {k: float(v) for k, v in losses.items()},
        "pred_logits": out["pred_logits"].detach(), "pred_boxes": out["pred_boxes"].detach(),
        "pred_corners_absmax": float(out["pred_corners"].abs().max()),
        "enc_logits_rowmax": out["enc_aux_outputs"][0]["pred_logits"].detach().max(-1).values,
        "dn_logits_last": out["dn_outputs"][-1]["pred_logits"].detach(),
        "dn_boxes_last": out["dn_outputs"][-1]["pred_boxes"].detach(),
        "n_dn": n_dn,
        "indices": [(i.clone(), j.clone()) for i, j in idx],
        "grads": {k: grads[k] for k in keep if k in grads and grads[k].numel() < 200000},
        "grad_norms": gnorm,
        "running_mean_stem1": model.state_dict()["backbone.stem.stem1.bn.running_mean"].clone(),
    }
    return fix


def msda_case(seed):
    """The reference MSDeformableAttention module (dfine_decoder.py:49-178) end to end, with gradients."""
    from src.d_fine.arch.dfine_decoder import MSDeformableAttention
    g = torch.Generator().manual_seed(seed)
    B, Q, heads, d = 2, 50, 8, 256
    shapes, points = [(20, 20), (10, 10), (5, 5)], [3, 6, 3]
    L = sum(h * w for h, w in shapes)
    torch.manual_seed(0)
    mod = MSDeformableAttention(d, heads, 3, points)
    mod.sampling_offsets.weight.data = torch.randn(mod.sampling_offsets.weight.shape, generator=g) * 0.05
    mod.attention_weights.weight.data = torch.randn(mod.attention_weights.weight.shape, generator=g) * 0.05
    mod.attention_weights.bias.data = torch.randn(mod.attention_weights.bias.shape, generator=g) * 0.1
    memory = torch.randn(B, L, d, generator=g).requires_grad_()
    query = torch.randn(B, Q, d, generator=g).requires_grad_()
    ref = torch.rand(B, Q, 1, 4, generator=g)
    ref[..., 2:] = ref[..., 2:] * 0.5 + 0.02
    value = memory.reshape(B, L, heads, d // heads).permute(0, 2, 3, 1).split([h * w for h, w in shapes], dim=-1)
    out = mod(query, ref, value, shapes)
    go = torch.randn(out.shape, generator=g)
    out.backward(go)
    return {"seed": seed, "shapes": shapes, "points": points, "out": out.detach(),
            "gmem": memory.grad.clone(), "gquery": query.grad.clone(),
            "g_off_w": mod.sampling_offsets.weight.grad.clone(), "g_attn_b": mod.attention_weights.bias.grad.clone(),
            "off_bias": mod.sampling_offsets.bias.detach().clone()}


def lsap_cases():
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(0)
    cases = []
    mats = [np.zeros((4, 2)), np.array([[1, 1], [1, 1], [0, 0], [1, 1]], dtype=float),
            rng.random((300, 7)).astype(np.float32), rng.integers(0, 3, (30, 9)).astype(float),
            rng.integers(0, 2, (12, 12)).astype(float), rng.random((5, 40)), rng.integers(0, 4, (300, 25)).astype(float)]
    for m in mats:
        r, c = linear_sum_assignment(m)
        cases.append({"cost": torch.as_tensor(np.asarray(m, dtype=np.float64)), "rows": torch.as_tensor(r),
                      "cols": torch.as_tensor(c)})
    return cases


if __name__ == "__main__":
    torch.set_num_threads(8)
    torch.save(model_case("n", 2, 320, 0), HERE / "model_n_320.pt")
    torch.save(model_case("s", 2, 320, 1), HERE / "model_s_320.pt")
    torch.save(msda_case(3), HERE / "msda_ref.pt")
    torch.save(lsap_cases(), HERE / "lsap_scipy.pt")
    for f in HERE.glob("*.pt"):
        print(f.name, f.stat().st_size)
