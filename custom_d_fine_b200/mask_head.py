"""MaskDecoder — fuses the PAN features of the HybridEncoder into 1/4-resolution mask features.

Behaviour follows /root/reference/src/d_fine/arch/dfine_decoder.py:315-370: 1x1 lateral conv + GroupNorm(32) per level,
coarser levels bilinearly resized to the 1/8 map and summed, 3x3 conv + GroupNorm + ReLU, bilinear x2, 3x3 conv +
GroupNorm + ReLU.  NHWC inside the graph; ``nn.Conv2d`` / ``nn.GroupNorm`` are parameter containers with the
reference's state-dict keys (``lateral.i.weight``, ``bn.i.*``, ``fusion_conv.weight``, ``fusion_norm.*``,
``up_conv.weight``, ``bn1.*``); all math goes through the kernel table ``K``.
"""
from __future__ import annotations

import torch.nn as nn
import torch.nn.init as init

from .kernels import K


class MaskDecoder(nn.Module):
    def __init__(self, in_chs, out_ch=256):
        super().__init__()
        n_groups = 32
        self.lateral = nn.ModuleList([nn.Conv2d(c, out_ch, 1, bias=False) for c in in_chs])
        self.bn = nn.ModuleList([nn.GroupNorm(n_groups, out_ch) for _ in in_chs])
        self.fusion_conv = nn.Conv2d(out_ch, out_ch, 3, padding=1, bias=False)
        self.fusion_norm = nn.GroupNorm(n_groups, out_ch)
        self.up_conv = nn.Conv2d(out_ch, out_ch, 3, padding=1, bias=False)
        self.bn1 = nn.GroupNorm(n_groups, out_ch)
        init.kaiming_normal_(self.up_conv.weight, mode="fan_out", nonlinearity="relu")

    @staticmethod
    def _gn(x, norm, act=None):
        return K.group_norm(x, norm.num_groups, norm.weight, norm.bias, norm.eps, act=act)

    def forward(self, feats):
        """feats: [F_s8, F_s16, F_s32] NHWC -> [B, H/4, W/4, out_ch]."""
        x = self._gn(K.conv2d(feats[0], self.lateral[0].weight, 1, (0, 0, 0, 0)), self.bn[0])
        size = (x.shape[1], x.shape[2])
        for i in range(1, len(feats)):
            t = self._gn(K.conv2d(feats[i], self.lateral[i].weight, 1, (0, 0, 0, 0)), self.bn[i])
            x = x + K.resize_bilinear(t, size)
        x = self._gn(K.conv2d(x, self.fusion_conv.weight, 1, (1, 1, 1, 1)), self.fusion_norm, act="relu")
        x = K.resize_bilinear(x, (2 * x.shape[1], 2 * x.shape[2]))
        return self._gn(K.conv2d(x, self.up_conv.weight, 1, (1, 1, 1, 1)), self.bn1, act="relu")
