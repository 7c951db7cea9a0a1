"""Model-size specifications (n/s/m/l/x).

Facts restated from the reference hyper-parameter tables
(/root/reference/src/d_fine/configs.py:1-213 and the HGNetv2 stage tables at
src/d_fine/arch/hgnetv2.py:344-422).  ``models`` keeps the reference's nested
dict shape (``models[size]["HGNetv2"|"HybridEncoder"|"DFINETransformer"|
"DFINECriterion"|"matcher"]``) because ``build_model``/``build_loss`` callers and
the reference's global-mutation quirk in ``build_loss`` depend on it.
"""
from __future__ import annotations

from copy import deepcopy

# name -> (stem [in, mid, out], stages: (in, mid, out, blocks, downsample, light, k, layers))
BACKBONES = {
    "B0": ([3, 16, 16], [(16, 16, 64, 1, False, False, 3, 3), (64, 32, 256, 1, True, False, 3, 3),
                         (256, 64, 512, 2, True, True, 5, 3), (512, 128, 1024, 1, True, True, 5, 3)]),
    "B1": ([3, 24, 32], [(32, 32, 64, 1, False, False, 3, 3), (64, 48, 256, 1, True, False, 3, 3),
                         (256, 96, 512, 2, True, True, 5, 3), (512, 192, 1024, 1, True, True, 5, 3)]),
    "B2": ([3, 24, 32], [(32, 32, 96, 1, False, False, 3, 4), (96, 64, 384, 1, True, False, 3, 4),
                         (384, 128, 768, 3, True, True, 5, 4), (768, 256, 1536, 1, True, True, 5, 4)]),
    "B3": ([3, 24, 32], [(32, 32, 128, 1, False, False, 3, 5), (128, 64, 512, 1, True, False, 3, 5),
                         (512, 128, 1024, 3, True, True, 5, 5), (1024, 256, 2048, 1, True, True, 5, 5)]),
    "B4": ([3, 32, 48], [(48, 48, 128, 1, False, False, 3, 6), (128, 96, 512, 1, True, False, 3, 6),
                         (512, 192, 1024, 3, True, True, 5, 6), (1024, 384, 2048, 1, True, True, 5, 6)]),
    "B5": ([3, 32, 64], [(64, 64, 128, 1, False, False, 3, 6), (128, 128, 512, 2, True, False, 3, 6),
                         (512, 256, 1024, 5, True, True, 5, 6), (1024, 512, 2048, 2, True, True, 5, 6)]),
    "B6": ([3, 48, 96], [(96, 96, 192, 2, False, False, 3, 6), (192, 192, 512, 3, True, False, 3, 6),
                         (512, 384, 1024, 6, True, True, 5, 6), (1024, 768, 2048, 3, True, True, 5, 6)]),
}

_LOSS_W = dict(loss_vfl=1, loss_bbox=5, loss_giou=2, loss_fgl=0.15, loss_ddf=1.5,
               loss_mask_bce=1, loss_mask_dice=1)
_COST_W = dict(cost_class=2, cost_bbox=5, cost_giou=2, cost_mask=1, cost_mask_dice=1)

base_cfg = {
    "HGNetv2": dict(pretrained=False, local_model_dir="weight/hgnetv2/", freeze_stem_only=True),
    "HybridEncoder": dict(num_encoder_layers=1, nhead=8, dropout=0.0, enc_act="gelu", act="silu"),
    "DFINETransformer": dict(eval_idx=-1, num_queries=300, num_denoising=100, label_noise_ratio=0.5,
                             box_noise_scale=1.0, reg_max=32, layer_scale=1,
                             cross_attn_method="default", query_select_method="default"),
    "DFINECriterion": dict(weight_dict=dict(_LOSS_W), losses=["vfl", "boxes", "local"],
                           alpha=0.75, gamma=2.0, reg_max=32),
    "matcher": dict(weight_dict=dict(_COST_W), alpha=0.25, gamma=2.0, use_focal_loss=True),
}


def _size(backbone, ret, freeze_at, freeze_norm, lab, in_ch, strides, hd_enc, enc_idx, ffn_enc,
          expansion, depth, feat_ch, levels, layers, reg_scale, points, ffn_dec=None, hd_dec=None):
    dec = dict(feat_channels=list(feat_ch), feat_strides=list(strides), hidden_dim=hd_dec or hd_enc,
               num_levels=levels, num_layers=layers, reg_scale=reg_scale, num_points=list(points),
               mask_dim=256)
    if ffn_dec is not None:
        dec["dim_feedforward"] = ffn_dec
    return {
        "HGNetv2": dict(name=backbone, return_idx=list(ret), freeze_at=freeze_at,
                        freeze_norm=freeze_norm, use_lab=lab),
        "HybridEncoder": dict(in_channels=list(in_ch), feat_strides=list(strides), hidden_dim=hd_enc,
                              use_encoder_idx=list(enc_idx), dim_feedforward=ffn_enc,
                              expansion=expansion, depth_mult=depth),
        "DFINETransformer": dec,
    }


sizes_cfg = {
    "n": _size("B0", [2, 3], -1, False, True, [512, 1024], [16, 32], 128, [1], 512, 0.34, 0.5,
               [128, 128], 2, 3, 4, [6, 6], ffn_dec=512),
    "s": _size("B0", [1, 2, 3], -1, False, True, [256, 512, 1024], [8, 16, 32], 256, [2], 1024, 0.5,
               0.34, [256, 256, 256], 3, 3, 4, [3, 6, 3]),
    "m": _size("B2", [1, 2, 3], -1, False, True, [384, 768, 1536], [8, 16, 32], 256, [2], 1024, 1.0,
               0.67, [256, 256, 256], 3, 4, 4, [3, 6, 3], ffn_dec=1024),
    "l": _size("B4", [1, 2, 3], 0, True, False, [512, 1024, 2048], [8, 16, 32], 256, [2], 1024, 1.0,
               1.0, [256, 256, 256], 3, 6, 4, [3, 6, 3], ffn_dec=1024),
    "x": _size("B5", [1, 2, 3], 0, True, False, [512, 1024, 2048], [8, 16, 32], 384, [2], 2048, 1.0,
               1.0, [384, 384, 384], 3, 6, 8, [3, 6, 3], ffn_dec=1024, hd_dec=256),
}
sizes_cfg["m"]["DFINETransformer"]["enable_mask_head"] = False


def merge_configs(base, size_specific):
    """Recursive dict merge, size-specific keys win (configs.py:203-210)."""
    out = dict(base)
    for key, val in size_specific.items():
        if isinstance(out.get(key), dict):
            out[key] = merge_configs(out[key], val)
        else:
            out[key] = val
    return out


models = {size: merge_configs(deepcopy(base_cfg), cfg) for size, cfg in sizes_cfg.items()}
# Reference quirk (configs.py:213 + dfine.py:75-76): the nested DFINECriterion dicts of all sizes
# alias the *same* base dict, so build_loss(enable_mask_head=True) leaks "masks" into every size.
for _s in models:
    models[_s]["DFINECriterion"] = base_cfg["DFINECriterion"]
    models[_s]["matcher"] = base_cfg["matcher"]
