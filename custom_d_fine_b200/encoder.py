"""HybridEncoder graph (AIFI + CCFF), NHWC.

Behaviour follows /root/reference/src/d_fine/arch/hybrid_encoder.py: input_proj 345-356,
AIFI TransformerEncoderLayer 243-290 (post-norm, GELU FFN), sincos position embedding
425-441, top-down FPN 463-476 and bottom-up PAN 478-484 built from RepNCSPELAN4
(181-206) / CSPLayer (209-239) / VGGBlock (106-121) / SCDown (96-103).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .blocks import ConvUnit, mha
from .kernels import K


def _cn(cin, cout, k, stride=1, groups=1, act=None):
    return ConvUnit(cin, cout, k, stride, groups, act=act, norm_name="norm")


class RepVGG(nn.Module):
    """act(BN(conv3x3 x) + BN(conv1x1 x)) — train-time two-branch form."""

    def __init__(self, cin, cout, act):
        super().__init__()
        self.conv1 = _cn(cin, cout, 3)
        self.conv2 = _cn(cin, cout, 1, act=act)  # act runs after the branch sum (pre_add)

    def convert_to_deploy(self):
        """RepVGG re-parameterisation (VGGBlock.convert_to_deploy, hybrid_encoder.py:123-137): both branches with their
        BatchNorms folded become one biased 3x3 conv (the 1x1 kernel zero-padded into the centre tap)."""
        if hasattr(self, "conv"):
            return
        w3, b3 = self.conv1.fused_kernel_bias()
        w1, b1 = self.conv2.fused_kernel_bias()
        cout, cin = w3.shape[:2]
        conv = nn.Conv2d(cin, cout, 3, 1, padding=1)
        conv.weight.data = (w3 + torch.nn.functional.pad(w1, (1, 1, 1, 1))).contiguous()
        conv.bias.data = (b3 + b1).contiguous()
        conv.requires_grad_(False)
        self._act = self.conv2.act
        self.conv = conv
        del self.conv1
        del self.conv2

    def forward(self, x):
        if hasattr(self, "conv"):
            return K.conv_bias_act(x, self.conv.weight, self.conv.bias, 1, (1, 1, 1, 1), 1, self._act)
        # x feeds both branches: the 1x1 branch reads the 3x3 conv's `tap` alias of x, so its data gradient is added
        # inside the 3x3 conv's data-gradient kernel instead of by a separate accumulation kernel
        a, xa = self.conv1(x, tap=True)
        return self.conv2(xa, pre_add=a)


class CSP(nn.Module):
    def __init__(self, cin, cout, n, act):
        super().__init__()
        self.conv1 = _cn(cin, cout, 1, act=act)
        self.conv2 = _cn(cin, cout, 1, act=act)
        self.bottlenecks = nn.Sequential(*[RepVGG(cout, cout, act) for _ in range(n)])
        # expansion == 1.0 everywhere => hidden == out => the reference's conv3 is Identity

    def forward(self, x, tap=False):
        a, xa = self.conv1(x, tap=True)
        for blk in self.bottlenecks:
            a = blk(a)
        if tap:      # hand the alias chain on: a further reader of x (the ELAN concat) routes its gradient through here
            return self.conv2(xa, post_add=a, tap=True)
        return self.conv2(xa, post_add=a)


class ELANBlock(nn.Module):
    """RepNCSPELAN4: split -> (CSP -> 3x3) x2 -> concat -> 1x1."""

    def __init__(self, c1, c2, c3, c4, n, act="silu"):
        super().__init__()
        self.c = c3 // 2
        self.cv1 = _cn(c1, c3, 1, act=act)
        self.cv2 = nn.Sequential(CSP(c3 // 2, c4, n, act), _cn(c4, c4, 3, act=act))
        self.cv3 = nn.Sequential(CSP(c4, c4, n, act), _cn(c4, c4, 3, act=act))
        self.cv4 = _cn(c3 + 2 * c4, c2, 1, act=act)

    def forward(self, x):
        if hasattr(K, "concat_buffer") and not hasattr(self.cv1, "conv_bn_fused") and os.environ.get("DFINE_CAT_ALIAS", "1") != "0":
            # concat by channel slice: the three members are written straight into one buffer by their producers
            c3, c4 = self.cv1.conv.out_channels, self.cv2[1].conv.out_channels
            if c3 % 4 == 0 and c4 % 4 == 0:
                buf = K.concat_buffer(x, x.shape[1], x.shape[2], c3 + 2 * c4)
                y = self.cv1(x, out=buf[..., :c3])
                y3 = self.cv2[1](self.cv2[0](y[..., self.c:]), out=buf[..., c3:c3 + c4])
                y4 = self.cv3[1](self.cv3[0](y3), out=buf[..., c3 + c4:])
                return self.cv4(K.cat_alias(buf, [y, y3, y4]))
        y = self.cv1(x)
        y2 = y[..., self.c:]
        y3 = self.cv2[1](self.cv2[0](y2))
        y4 = self.cv3[1](self.cv3[0](y3))
        return self.cv4(K.cat([y, y3, y4]))


class SCDown(nn.Module):
    def __init__(self, c1, c2, k, s):
        super().__init__()
        self.cv1 = _cn(c1, c2, 1)
        self.cv2 = _cn(c2, c2, k, s, groups=c2)

    def forward(self, x):
        return self.cv2(self.cv1(x))


class AIFILayer(nn.Module):
    def __init__(self, d, nhead, ffn):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, nhead, 0.0, batch_first=True)
        self.linear1 = nn.Linear(d, ffn)
        self.linear2 = nn.Linear(ffn, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)

    def forward(self, src, pos):
        a = mha(self.self_attn, src + pos, src)
        src = K.layernorm(a, self.norm1.weight, self.norm1.bias, self.norm1.eps, residual=src)
        f = K.linear(src, self.linear1.weight, self.linear1.bias, act="gelu")
        f = K.linear(f, self.linear2.weight, self.linear2.bias)
        return K.layernorm(f, self.norm2.weight, self.norm2.bias, self.norm2.eps, residual=src)


class AIFIStack(nn.Module):
    def __init__(self, d, nhead, ffn, n):
        super().__init__()
        self.layers = nn.ModuleList(AIFILayer(d, nhead, ffn) for _ in range(n))

    def forward(self, src, pos):
        for layer in self.layers:
            src = layer(src, pos)
        return src


def sincos_pos_embed(w, h, dim, temperature=10000.0):
    """[1, w*h, dim] 2-D sin/cos embedding (hybrid_encoder.py:425-441).  Note the reference
    flattens a (w, h) 'ij' meshgrid, i.e. token t <-> (t // h, t % h); kept verbatim."""
    gw, gh = torch.meshgrid(torch.arange(int(w), dtype=torch.float32),
                            torch.arange(int(h), dtype=torch.float32), indexing="ij")
    pd = dim // 4
    omega = 1.0 / (temperature ** (torch.arange(pd, dtype=torch.float32) / pd))
    ow = gw.flatten()[:, None] @ omega[None]
    oh = gh.flatten()[:, None] @ omega[None]
    return torch.cat([ow.sin(), ow.cos(), oh.sin(), oh.cos()], 1)[None]


class HybridEncoder(nn.Module):
    def __init__(self, in_channels=(512, 1024, 2048), feat_strides=(8, 16, 32), hidden_dim=256, nhead=8,
                 dim_feedforward=1024, dropout=0.0, enc_act="gelu", use_encoder_idx=(2,),
                 num_encoder_layers=1, pe_temperature=10000, expansion=1.0, depth_mult=1.0, act="silu",
                 eval_spatial_size=None):
        super().__init__()
        assert dropout == 0.0 and enc_act == "gelu" and act == "silu"
        self.in_channels, self.feat_strides = list(in_channels), list(feat_strides)
        self.hidden_dim, self.use_encoder_idx = hidden_dim, list(use_encoder_idx)
        self.num_encoder_layers, self.pe_temperature = num_encoder_layers, pe_temperature
        self.eval_spatial_size = eval_spatial_size
        self.out_channels = [hidden_dim] * len(in_channels)
        self.out_strides = list(feat_strides)
        self._pos_cache = {}

        self.input_proj = nn.ModuleList(_cn(c, hidden_dim, 1) for c in in_channels)
        self.encoder = nn.ModuleList(
            AIFIStack(hidden_dim, nhead, dim_feedforward, num_encoder_layers) for _ in use_encoder_idx)
        n_lvl = len(in_channels)
        c4 = round(expansion * hidden_dim // 2)
        nb = round(3 * depth_mult)
        self.lateral_convs = nn.ModuleList(_cn(hidden_dim, hidden_dim, 1) for _ in range(n_lvl - 1))
        self.fpn_blocks = nn.ModuleList(
            ELANBlock(hidden_dim * 2, hidden_dim, hidden_dim * 2, c4, nb) for _ in range(n_lvl - 1))
        self.downsample_convs = nn.ModuleList(
            nn.Sequential(SCDown(hidden_dim, hidden_dim, 3, 2)) for _ in range(n_lvl - 1))
        self.pan_blocks = nn.ModuleList(
            ELANBlock(hidden_dim * 2, hidden_dim, hidden_dim * 2, c4, nb) for _ in range(n_lvl - 1))

    def _pos(self, w, h, device):
        # The reference rebuilds this on the host and copies it H2D every training step
        # (hybrid_encoder.py:452-456); it only depends on (w, h) so it is cached on device.
        key = (w, h, str(device))
        if key not in self._pos_cache:
            self._pos_cache[key] = sincos_pos_embed(w, h, self.hidden_dim, self.pe_temperature).to(device)
        return self._pos_cache[key]

    def forward(self, feats):
        assert len(feats) == len(self.in_channels)
        proj = [p(f) for p, f in zip(self.input_proj, feats)]
        if self.num_encoder_layers > 0:
            for i, lvl in enumerate(self.use_encoder_idx):
                b, h, w, c = proj[lvl].shape
                tokens = proj[lvl].reshape(b, h * w, c)          # NHWC is already token-major
                mem = self.encoder[i](tokens, self._pos(w, h, tokens.device))
                proj[lvl] = mem.reshape(b, h, w, c)

        n = len(self.in_channels)
        inner = [proj[-1]]
        for idx in range(n - 1, 0, -1):
            hi = self.lateral_convs[n - 1 - idx](inner[0])
            inner[0] = hi
            fused = K.cat([K.upsample_nearest2x(hi), proj[idx - 1]])
            inner.insert(0, self.fpn_blocks[n - 1 - idx](fused))
        outs = [inner[0]]
        for idx in range(n - 1):
            down = self.downsample_convs[idx][0](outs[-1])
            outs.append(self.pan_blocks[idx](K.cat([down, inner[idx + 1]])))
        return outs
