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
import torch.nn.functional as F

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
    def mask_cost(self, outputs_list, targets):
        """Segmentation term of the matching cost (matcher.py:19-71, 175-237): per layer that predicts masks and per
        image, cost_mask_dice * (1 - pairwise Dice(sigmoid(pred), gt)) + cost_mask * pixel-wise sigmoid focal cost, GT
        masks bilinearly resized to the prediction size.  Returns None (no layer / no masks) or [n_layers, Q*sumT] in
        the matcher's cost layout (image b's [Q, T_b] block row-major at Q*offset_b; zero rows for layers without
        masks).  Device tensor ops: two [Q, HW] x [HW, T_b] GEMMs per (layer, image)."""
        if not (self.cost_mask > 0 or self.cost_mask_dice > 0):
            return None
        if not any(o.get("pred_masks") is not None for o in outputs_list):
            return None
        if not any(t.get("masks") is not None and t["masks"].numel() > 0 for t in targets):
            return None
        sizes = [len(t["boxes"]) for t in targets]
        Q = outputs_list[0]["pred_logits"].shape[1]
        dev = outputs_list[0]["pred_logits"].device
        if dev.type == "cuda" and hasattr(K, "mask_cost_layer") and all(
                t.get("masks") is not None and t["masks"].dim() == 3 for t in targets):
            return self._mask_cost_kernel(outputs_list, targets, sizes, Q, dev)
        extra = torch.zeros((len(outputs_list), Q * sum(sizes)), device=dev, dtype=torch.float32)
        resized = {}
        for li, o in enumerate(outputs_list):
            pm = o.get("pred_masks")
            if pm is None:
                continue
            if pm.shape[1] != Q:                    # denoising queries in front of the matching queries (matcher.py:192-198)
                pm = pm[:, pm.shape[1] - Q:]
            Hm, Wm = pm.shape[-2:]
            off = 0
            for b, t in enumerate(targets):
                n = sizes[b]
                m = t.get("masks")
                if n == 0 or m is None or m.numel() == 0:
                    off += n
                    continue
                key = (b, Hm, Wm)
                if key not in resized:
                    g = m.float().to(dev)
                    if g.shape[-2:] != (Hm, Wm):
                        g = F.interpolate(g.unsqueeze(1), size=(Hm, Wm), mode="bilinear", align_corners=False).squeeze(1)
                    resized[key] = g.flatten(1)                                      # [T_b, HW]
                gt = resized[key]
                logit = pm[b].flatten(1).float()                                     # [Q, HW]
                prob = logit.sigmoid()
                cost = torch.zeros((Q, n), device=dev, dtype=torch.float32)
                if self.cost_mask_dice > 0:
                    num = 2 * (prob @ gt.t())
                    den = prob.sum(1, keepdim=True) + gt.sum(1)
                    cost = cost + self.cost_mask_dice * (1 - (num + 1e-6) / (den + 1e-6))
                if self.cost_mask > 0:
                    neg = (1 - self.alpha) * (prob ** self.gamma) * (-(1 - prob + 1e-8).log())
                    pos = self.alpha * ((1 - prob) ** self.gamma) * (-(prob + 1e-8).log())
                    cost = cost + self.cost_mask * ((pos @ gt.t() + neg @ (1 - gt).t()) / logit.shape[1])
                extra[li, Q * off:Q * (off + n)] = cost.reshape(-1)
                off += n
        return extra

    def _mask_cost_kernel(self, outputs_list, targets, sizes, Q, dev):
        """CUDA path: one fused launch per layer (csrc/seg.cu: sigmoid / focal terms and the products with the GT masks in
        one pass over the mask logits) instead of ~40 device ops and three fp32 GEMMs per (layer, image)."""
        extra = torch.zeros((len(outputs_list), Q * sum(sizes)), device=dev, dtype=torch.float32)
        toff = K.toff_device(sizes, dev)
        gts = {}
        for li, o in enumerate(outputs_list):
            pm = o.get("pred_masks")
            if pm is None:
                continue
            if pm.shape[1] != Q:
                pm = pm[:, pm.shape[1] - Q:]
            Hm, Wm = pm.shape[-2:]
            if (Hm, Wm) not in gts:         # every GT mask resized once per step and prediction size
                g = torch.cat([t["masks"] for t in targets]).float()
                if g.shape[-2:] != (Hm, Wm):
                    g = F.interpolate(g.unsqueeze(1), size=(Hm, Wm), mode="bilinear", align_corners=False).squeeze(1)
                g = g.flatten(1).contiguous()
                gts[(Hm, Wm)] = (g, g.sum(1).contiguous())
            g, gsum = gts[(Hm, Wm)]
            extra[li] = K.mask_cost_layer(pm, g, gsum, toff, sizes, self.alpha, self.gamma, float(self.cost_mask_dice),
                                          float(self.cost_mask))
        return extra

    @torch.no_grad()
    def match_layers(self, outputs_list, targets):
        """Match several prediction sets (decoder layers) against the same targets at once.
        Returns, per layer, a list over images of (query_idx ascending, target_idx) int64 CPU tensors."""
        extra = self.mask_cost(outputs_list, targets)
        kw = {} if extra is None else {"extra_cost": extra}
        return K.match([o["pred_logits"] for o in outputs_list], [o["pred_boxes"] for o in outputs_list],
                       targets, self.alpha, self.gamma, float(self.cost_class), float(self.cost_bbox),
                       float(self.cost_giou), **kw)

    @torch.no_grad()
    def match_layers_raw(self, outputs_list, targets):
        """Device-side half of ``match_layers``: launches the matcher and returns the provider's raw result
        (device int64 [n_layers, sumT] x2 on the CUDA path) without touching the host."""
        extra = self.mask_cost(outputs_list, targets)
        kw = {} if extra is None else {"extra_cost": extra}
        return K.match_raw([o["pred_logits"] for o in outputs_list], [o["pred_boxes"] for o in outputs_list],
                           targets, self.alpha, self.gamma, float(self.cost_class), float(self.cost_bbox),
                           float(self.cost_giou), **kw)

    @staticmethod
    def raw_to_host(raw, plan):
        """Host int64 arrays [n_sets, sumT] (q, t) from the provider's raw result: the step's one D2H."""
        return K.match_raw_to_host(raw, plan)

    @torch.no_grad()
    def forward(self, outputs, targets, return_topk=False):
        if return_topk:
            raise NotImplementedError("get_top_k_matches has no caller in the reference (matcher.py:259-285)")
        return {"indices": self.match_layers([outputs], targets)[0]}
