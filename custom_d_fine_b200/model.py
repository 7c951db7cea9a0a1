"""DFINE model facade + builders: the drop-in surface of ``src/d_fine/dfine.py``.

``build_model / build_loss / build_optimizer`` keep the reference's signatures, argument
meaning and side effects (/root/reference/src/d_fine/dfine.py:51-124), and the module tree
keeps the reference's state-dict keys, shapes and dtypes so ``pretrained/dfine_*_coco.pth``
loads with nothing missed or unmatched (src/d_fine/utils.py:140-181).
"""
from __future__ import annotations

from copy import deepcopy
from pathlib import Path

import torch
import torch.nn as nn
import torch.optim as optim

from .backbone import HGNetv2
from .decoder import DFINETransformer
from .encoder import HybridEncoder
from .specs import models


class DFINE(nn.Module):
    def __init__(self, backbone, encoder, decoder):
        super().__init__()
        self.backbone = backbone
        self.decoder = decoder
        self.encoder = encoder

    def forward(self, x, targets=None):
        """x: float32 [B,3,H,W] in [0,1] (NCHW at the boundary, like the reference)."""
        x = x.permute(0, 2, 3, 1)  # NHWC view; the first conv kernel reads it with these strides
        return self.decoder(self.encoder(self.backbone(x)), targets)

    def deploy(self):
        """Inference form (dfine.py:43-48): eval mode + every module's `convert_to_deploy` — conv + BatchNorm folded
        into one biased conv, RepVGG branches merged into one 3x3 conv, decoder layers / heads past `eval_idx` dropped.
        Parents convert before their children (a RepVGG block folds its two branches itself)."""
        self.eval()
        for m in list(self.modules()):
            if m is not self and hasattr(m, "convert_to_deploy"):
                m.convert_to_deploy()
        return self


def build_model(model_name, num_classes, enable_mask_head, device, img_size=None, pretrained_model_path=None):
    cfg = deepcopy(models[model_name])
    cfg["HybridEncoder"]["eval_spatial_size"] = img_size
    cfg["DFINETransformer"]["eval_spatial_size"] = img_size
    cfg["DFINETransformer"]["enable_mask_head"] = enable_mask_head
    model = DFINE(HGNetv2(**cfg["HGNetv2"]), HybridEncoder(**cfg["HybridEncoder"]),
                  DFINETransformer(num_classes=num_classes, **cfg["DFINETransformer"]))
    if pretrained_model_path:
        if not Path(pretrained_model_path).exists():
            raise FileNotFoundError(f"{pretrained_model_path} does not exist")
        model = load_tuning_state(model, str(pretrained_model_path))
    return model.to(device)


def build_loss(model_name, num_classes, label_smoothing, enable_mask_head):
    from .criterion import DFINECriterion
    from .matcher import HungarianMatcher

    cfg = models[model_name]
    if enable_mask_head:
        # reference side effect kept on purpose: mutates the shared table (dfine.py:75-76)
        cfg["DFINECriterion"]["losses"].append("masks")
    matcher = HungarianMatcher(**cfg["matcher"])
    return DFINECriterion(matcher, num_classes=num_classes, label_smoothing=label_smoothing,
                          **cfg["DFINECriterion"])


def param_groups(model, backbone_lr, base_lr):
    """Four AdamW groups (dfine.py:87-124): backbone / backbone norms (wd 0) /
    enc-dec norms+biases (wd 0) / rest."""
    g = [[], [], [], []]
    for name, p in model.named_parameters():
        is_norm = "norm" in name or "bn" in name
        if "backbone" in name:
            g[1 if is_norm else 0].append(p)
        elif ("encoder" in name or "decoder" in name) and (is_norm or "bias" in name):
            g[2].append(p)
        else:
            g[3].append(p)
    return [
        {"params": g[0], "lr": backbone_lr, "initial_lr": backbone_lr},
        {"params": g[1], "lr": backbone_lr, "weight_decay": 0.0, "initial_lr": backbone_lr},
        {"params": g[2], "weight_decay": 0.0, "lr": base_lr, "initial_lr": base_lr},
        {"params": g[3], "lr": base_lr, "initial_lr": base_lr},
    ]


def build_optimizer(model, lr, backbone_lr, betas, weight_decay, base_lr):
    """AdamW over the reference's four groups (dfine.py:87-124).  On the GPU this is the flat-arena optimizer
    (custom_d_fine_b200/optim.py: clip + AdamW + EMA + zero_grad fused, graph-replayable, scheduler-compatible);
    on a CPU-resident model (host-logic tests) it is torch.optim.AdamW."""
    groups = param_groups(model, backbone_lr, base_lr)
    if all(p.is_cuda for p in model.parameters()):
        from .optim import FusedAdamW
        return FusedAdamW(groups, lr=lr, betas=betas, weight_decay=weight_decay)
    return optim.AdamW(groups, lr=lr, betas=betas, weight_decay=weight_decay)


# ---- checkpoint loading (src/d_fine/utils.py:92-181) ------------------------------------------
def _obj365_ids():
    # COCO-80 class -> Objects365 id table shipped with the reference checkpoints' tooling.
    return [0, 46, 5, 58, 114, 55, 116, 65, 21, 40, 176, 127, 249, 24, 56, 139, 92, 78, 99, 96, 144, 295,
            178, 180, 38, 39, 13, 43, 120, 219, 148, 173, 165, 154, 137, 113, 145, 146, 204, 8, 35, 10, 88,
            84, 93, 26, 112, 82, 265, 104, 141, 152, 234, 143, 150, 97, 2, 50, 25, 75, 98, 153, 37, 73, 115,
            132, 106, 61, 163, 134, 277, 81, 133, 18, 94, 30, 169, 70, 328, 226]


def _remap_class_rows(cur, pre):
    if pre.size() == cur.size():
        return pre
    out = cur.clone()
    out.requires_grad = False
    ids = _obj365_ids()
    if pre.size() > cur.size():
        for coco, obj in enumerate(ids):
            out[coco] = pre[obj + 1]
    else:
        for coco, obj in enumerate(ids):
            out[obj + 1] = pre[coco]
    return out


def load_tuning_state(model, path):
    state = torch.load(path, map_location="cpu", weights_only=True)
    if "ema" in state:
        pre = state["ema"]["module"]
    elif "model" in state:
        pre = state["model"]
    else:
        pre = state
    cur = model.state_dict()
    try:
        k = "decoder.denoising_class_embed.weight"
        if k in pre and k in cur and pre[k].size() != cur[k].size():
            del pre[k]
        names = ["decoder.enc_score_head.weight", "decoder.enc_score_head.bias"]
        for i in range(8):
            names += [f"decoder.dec_score_head.{i}.weight", f"decoder.dec_score_head.{i}.bias"]
        for n in names:
            if n in cur and n in pre:
                pre[n] = _remap_class_rows(cur[n], pre[n])
    except Exception:  # noqa: BLE001 - mirror the reference's best-effort behaviour
        pass
    matched = {k: pre[k] for k, v in cur.items() if k in pre and v.shape == pre[k].shape}
    info = {"missed": [k for k in cur if k not in pre],
            "unmatched": [k for k, v in cur.items() if k in pre and v.shape != pre[k].shape]}
    model.load_state_dict(matched, strict=False)
    model._load_info = info
    return model
