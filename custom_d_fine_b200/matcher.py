"""HungarianMatcher — drop-in for /root/reference/src/d_fine/matcher.py:74-257.

The reference builds a [B*Q, sum(T)] cost matrix on the device with ~25 ATen kernels, copies
it to the host (a device sync per call, 6 calls per step for D-FINE-m) and runs SciPy's
single-threaded LSAP per image.  Here ``K.match`` computes only the per-image diagonal
blocks and solves every (layer, image) assignment problem in ONE launch (one CTA per
problem, fp32 costs, fp64 duals, SciPy's tie-breaking), so a training step needs a single
small D2H of the index table instead of six synchronising matrix copies.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .kernels import K


class HungarianMatcher(nn.Module):
    def __init__(self, weight_dict, use_focal_loss=False, alpha=0.25, gamma=2.0):
        super().__init__()
        self.cost_class = weight_dict["cost_class"]
        self.cost_bbox = weight_dict["cost_bbox"]
        self.cost_giou = weight_dict["cost_giou"]
        self.cost_mask = weight_dict.get("cost_mask", 0)
        self.cost_mask_dice = weight_dict.get("cost_mask_dice", 0)
        self.use_focal_loss, self.alpha, self.gamma = use_focal_loss, alpha, gamma
        assert self.cost_class != 0 or self.cost_bbox != 0 or self.cost_giou != 0, "all costs cant be 0"
        if not use_focal_loss:
            raise NotImplementedError("softmax class cost is not on any shipped config's path")

    @torch.no_grad()
    def match_layers(self, outputs_list, targets):
        """Match several prediction sets (decoder layers) against the same targets at once.
        Returns, per layer, a list over images of (query_idx ascending, target_idx) int64 CPU tensors."""
        for o in outputs_list:
            if o.get("pred_masks") is not None and (self.cost_mask > 0 or self.cost_mask_dice > 0) and any(
                    t.get("masks") is not None and t["masks"].numel() > 0 for t in targets):
                raise NotImplementedError("mask matching cost (matcher.py:175-237) is a SURVEY §8(f) 'next' row")
        return K.match([o["pred_logits"] for o in outputs_list], [o["pred_boxes"] for o in outputs_list],
                       targets, self.alpha, self.gamma, float(self.cost_class), float(self.cost_bbox),
                       float(self.cost_giou))

    @torch.no_grad()
    def match_layers_raw(self, outputs_list, targets):
        """Device-side half of ``match_layers``: launches the matcher and returns the provider's raw result
        (device int64 [n_layers, sumT] x2 on the CUDA path) without touching the host."""
        for o in outputs_list:
            if o.get("pred_masks") is not None and (self.cost_mask > 0 or self.cost_mask_dice > 0) and any(
                    t.get("masks") is not None and t["masks"].numel() > 0 for t in targets):
                raise NotImplementedError("mask matching cost (matcher.py:175-237) is a SURVEY §8(f) 'next' row")
        return K.match_raw([o["pred_logits"] for o in outputs_list], [o["pred_boxes"] for o in outputs_list],
                           targets, self.alpha, self.gamma, float(self.cost_class), float(self.cost_bbox),
                           float(self.cost_giou))

    @staticmethod
    def raw_to_host(raw, plan):
        """Host int64 arrays [n_sets, sumT] (q, t) from the provider's raw result: the step's one D2H."""
        return K.match_raw_to_host(raw, plan)

    @torch.no_grad()
    def forward(self, outputs, targets, return_topk=False):
        if return_topk:
            raise NotImplementedError("get_top_k_matches has no caller in the reference (matcher.py:259-285)")
        return {"indices": self.match_layers([outputs], targets)[0]}
