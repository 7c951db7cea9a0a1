"""HGNetv2 backbone graph, NHWC, built on fused conv units.

Behaviour follows /root/reference/src/d_fine/arch/hgnetv2.py (StemBlock 115-166,
HG_Block 192-275 with the agg="se" squeeze/excite pair that HG_Stage's default selects
— 290,320, HG_Stage 278-337, HGNetv2 425-568).  Module/attribute names are chosen so
the state-dict keys equal the reference's.
"""
from __future__ import annotations

import os

import torch.nn as nn

from .blocks import ConvUnit
from .kernels import K
from .specs import BACKBONES


class Stem(nn.Module):
    def __init__(self, cin, mid, cout, lab, frozen):
        super().__init__()
        kw = dict(act="relu", lab=lab, frozen_norm=frozen)
        self.stem1 = ConvUnit(cin, mid, 3, 2, **kw)
        # the reference zero-pads right/bottom by one before each 2x2 conv (hgnetv2.py:158-161)
        self.stem2a = ConvUnit(mid, mid // 2, 2, 1, pad=(0, 0, 1, 1), **kw)
        self.stem2b = ConvUnit(mid // 2, mid, 2, 1, pad=(0, 0, 1, 1), **kw)
        self.stem3 = ConvUnit(mid * 2, mid, 3, 2, **kw)
        self.stem4 = ConvUnit(mid, cout, 1, 1, **kw)

    def forward(self, x):
        x = self.stem1(x)
        x2 = self.stem2b(self.stem2a(x))
        x1 = K.maxpool2x2_s1_padbr(x)  # MaxPool2d(k=2,s=1,ceil) over the right/bottom zero-padded map
        return self.stem4(self.stem3(K.cat([x1, x2])))


class LightConv(nn.Module):
    """1x1 (no act) followed by depthwise kxk + ReLU + LAB (hgnetv2.py:83-112)."""

    def __init__(self, cin, cout, k, lab, frozen):
        super().__init__()
        self.conv1 = ConvUnit(cin, cout, 1, act=None, lab=lab, frozen_norm=frozen)
        self.conv2 = ConvUnit(cout, cout, k, groups=cout, act="relu", lab=lab, frozen_norm=frozen)

    def forward(self, x, tap=False, out=None):
        if tap:
            y, x_alias = self.conv1(x, tap=True)
            return self.conv2(y, out=out), x_alias
        return self.conv2(self.conv1(x), out=out)


class HGBlock(nn.Module):
    def __init__(self, cin, mid, cout, n_layers, k, residual, light, lab, frozen):
        super().__init__()
        self.residual = residual
        self.layers = nn.ModuleList()
        for i in range(n_layers):
            c = cin if i == 0 else mid
            if light:
                self.layers.append(LightConv(c, mid, k, lab, frozen))
            else:
                self.layers.append(ConvUnit(c, mid, k, act="relu", lab=lab, frozen_norm=frozen))
        total = cin + n_layers * mid
        self.cin, self.mid, self.total, self.cout = cin, mid, total, cout
        self.aggregation = nn.Sequential(
            ConvUnit(total, cout // 2, 1, act="relu", lab=lab, frozen_norm=frozen),
            ConvUnit(cout // 2, cout, 1, act="relu", lab=lab, frozen_norm=frozen),
        )

    def forward(self, x, buf=None, out=None):
        """buf: concatenation buffer [B,H,W,total] whose first `cin` channels ARE x (x is an alias of that slice): every
        layer then writes its output into its own channel slice and the concat (hgnetv2.py:265-275) is the buffer
        itself — no copy kernel.  out: where the block's result goes (the first slice of the next block's buffer)."""
        if buf is not None:
            feats, y, off = [x], x, self.cin
            for layer in self.layers:
                y = layer(y, out=buf[..., off:off + self.mid])
                feats.append(y)
                off += self.mid
            y = self.aggregation[0](K.cat_alias(buf, feats))
            return self.aggregation[1](y, post_add=x if self.residual else None, out=out)
        # every layer input is also a member of the concat: the concat reads the layer's `tap` alias of its input so
        # that the concat's gradient slice is added inside that layer's data-gradient kernel
        feats = []
        y = x
        for layer in self.layers:
            y, y_in = layer(y, tap=True)
            feats.append(y_in)
        feats.append(y)
        y = self.aggregation[0](K.cat(feats))
        return self.aggregation[1](y, post_add=x if self.residual else None)


class HGStage(nn.Module):
    def __init__(self, cin, mid, cout, n_blocks, n_layers, downsample, light, k, lab, frozen):
        super().__init__()
        if downsample:
            self.downsample = ConvUnit(cin, cin, 3, 2, groups=cin, act=None, lab=lab, frozen_norm=frozen)
        else:
            self.downsample = None
        self.blocks = nn.Sequential(*[
            HGBlock(cin if i == 0 else cout, mid, cout, n_layers, k, i > 0, light, lab, frozen)
            for i in range(n_blocks)
        ])

    def _sliced(self, x):
        blk = self.blocks[0]
        return (hasattr(K, "concat_buffer") and os.environ.get("DFINE_CAT_ALIAS", "1") != "0" and x.is_cuda
                and not hasattr(blk.aggregation[0], "conv_bn_fused") and blk.cin % 4 == 0 and blk.mid % 4 == 0)

    def forward(self, x):
        if not self._sliced(x):
            if self.downsample is not None:
                x = self.downsample(x)
            for blk in self.blocks:
                x = blk(x)
            return x
        # concat by channel slice: every block input is produced straight into the first slice of that block's buffer
        H, W = x.shape[1], x.shape[2]
        if self.downsample is not None:
            H, W = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        buf = K.concat_buffer(x, H, W, self.blocks[0].total)
        first = buf[..., :self.blocks[0].cin]
        x = self.downsample(x, out=first) if self.downsample is not None else K.copy_into(x, first)
        for i, blk in enumerate(self.blocks):
            nxt = K.concat_buffer(x, H, W, self.blocks[i + 1].total) if i + 1 < len(self.blocks) else None
            x = blk(x, buf=buf, out=None if nxt is None else nxt[..., :blk.cout])
            buf = nxt
        return x


class HGNetv2(nn.Module):
    def __init__(self, name, use_lab=False, return_idx=(1, 2, 3), freeze_stem_only=True,
                 freeze_at=0, freeze_norm=True, pretrained=False, local_model_dir=None):
        super().__init__()
        if pretrained:
            raise NotImplementedError("stage-1 backbone download needs network; load a full checkpoint instead")
        stem_ch, stages = BACKBONES[name]
        self.return_idx = list(return_idx)
        self.stem = Stem(*stem_ch, use_lab, freeze_norm)
        self.stages = nn.ModuleList(
            HGStage(cin, mid, cout, nb, nl, ds, light, k, use_lab, freeze_norm)
            for (cin, mid, cout, nb, ds, light, k, nl) in stages
        )
        if freeze_at >= 0:
            frozen = [self.stem]
            if not freeze_stem_only:
                frozen += list(self.stages[: freeze_at + 1])
            for m in frozen:
                for p in m.parameters():
                    p.requires_grad = False

    def forward(self, x_nhwc):
        x = self.stem(x_nhwc)
        outs = []
        for i, stage in enumerate(self.stages):
            x = stage(x)
            if i in self.return_idx:
                outs.append(x)
        return outs
