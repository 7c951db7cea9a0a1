"""Parameter containers + fused conv units shared by backbone / encoder / decoder.

Activations are NHWC fp32 (``[B, H, W, C]``) everywhere inside the graph; the
reference's NCHW only exists at the image input and in the state-dict weight
layout ``[Co, Ci/g, kh, kw]``.  ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.Linear``
instances are used purely as *parameter containers* (identical state-dict keys,
dtypes and default initialisers as the reference); their ``forward`` is never
called — all math goes through the kernel table ``K``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .kernels import K


class FrozenBN(nn.Module):
    """Buffers-only BatchNorm (reference: arch/common.py:29-70)."""

    def __init__(self, n, eps=1e-5):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.eps = eps
        self.num_features = n

    def _load_from_state_dict(self, state_dict, prefix, *args):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, *args)


class Affine(nn.Module):
    """Scalar learnable affine after the activation (hgnetv2.py:25-32)."""

    def __init__(self):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor([1.0]))
        self.bias = nn.Parameter(torch.tensor([0.0]))


class ConvUnit(nn.Module):
    """conv (no bias) -> BatchNorm -> [+pre_add] -> act -> [LAB] -> [+post_add], NHWC.

    Covers the reference's ConvBNAct (hgnetv2.py:35-80), ConvNormLayer(_fuse)
    (hybrid_encoder.py:22-99) and the conv+norm ``input_proj`` pairs.  ``norm_name``
    selects the attribute the BN container is registered under ("bn"/"norm") so
    state-dict keys are identical.
    """

    def __init__(self, cin, cout, k, stride=1, groups=1, pad=None, act=None, lab=False,
                 norm_name="bn", frozen_norm=False):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, padding=(k - 1) // 2, groups=groups, bias=False)
        setattr(self, norm_name, FrozenBN(cout) if frozen_norm else nn.BatchNorm2d(cout))
        self._norm_name = norm_name
        if lab and act is not None:
            self.lab = Affine()
        p = (k - 1) // 2
        self.pad = (p, p, p, p) if pad is None else tuple(pad)  # top, left, bottom, right
        self.stride, self.groups, self.act, self.k = stride, groups, act, k

    def _bn(self):
        return getattr(self, self._norm_name)

    def fused_kernel_bias(self):
        """Conv weight / bias with the BatchNorm's running statistics folded in (hybrid_encoder.py:66-80)."""
        bn = self._bn()
        t = bn.weight / (bn.running_var + bn.eps).sqrt()
        return (self.conv.weight * t.reshape(-1, 1, 1, 1)).detach(), (bn.bias - bn.running_mean * t).detach()

    def convert_to_deploy(self):
        """Inference re-parameterisation (ConvNormLayer_fuse.convert_to_deploy, hybrid_encoder.py:47-63): conv and
        BatchNorm become ONE biased conv (`conv_bn_fused`, the reference's name); the forward pass is then a single
        kernel launch with bias, activation and the LAB scalars in its epilogue.  Applied to every conv unit of the graph
        (the reference folds its encoder layers only; the arithmetic is the same eval-mode affine map)."""
        if hasattr(self, "conv_bn_fused"):
            return
        w, b = self.fused_kernel_bias()
        c = self.conv
        fused = nn.Conv2d(c.in_channels, c.out_channels, c.kernel_size, c.stride, c.padding, groups=c.groups, bias=True)
        fused.weight.data, fused.bias.data = w.contiguous(), b.contiguous()
        fused.requires_grad_(False)
        lab = getattr(self, "lab", None)
        self._lab = None if lab is None else (float(lab.scale), float(lab.bias))
        self.conv_bn_fused = fused
        del self.conv
        delattr(self, self._norm_name)

    def forward(self, x, pre_add=None, post_add=None, tap=False, out=None):
        if hasattr(self, "conv_bn_fused"):
            assert out is None
            f = self.conv_bn_fused
            y = K.conv_bias_act(x, f.weight, f.bias, self.stride, self.pad, self.groups, self.act, self._lab, pre_add, post_add)
            return (y, x) if tap else y
        bn = self._bn()
        frozen = isinstance(bn, FrozenBN)
        lab = getattr(self, "lab", None)
        return K.conv_bn_act(
            x, self.conv.weight, self.stride, self.pad, self.groups,
            bn.weight, bn.bias, bn.running_mean, bn.running_var,
            None if frozen else bn.num_batches_tracked,
            training=self.training and not frozen, momentum=0.1, eps=bn.eps, act=self.act,
            lab_scale=None if lab is None else lab.scale,
            lab_bias=None if lab is None else lab.bias,
            pre_add=pre_add, post_add=post_add, tap=tap, **({} if out is None else {"out": out}))


class MLP(nn.Module):
    """Linear stack with ReLU between layers (dfine_decoder.py:33-46); keys ``layers.i``."""

    def __init__(self, din, dh, dout, n):
        super().__init__()
        dims = [din] + [dh] * (n - 1) + [dout]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        last = len(self.layers) - 1
        for i, lin in enumerate(self.layers):
            x = K.linear(x, lin.weight, lin.bias, act=None if i == last else "relu")
        return x


def mha(container: nn.MultiheadAttention, qk_in, v_in, mask=None):
    """nn.MultiheadAttention semantics (packed in_proj rows [q|k|v], scale on q, bool mask
    True = -inf, dropout 0) with q = k = ``qk_in`` and v = ``v_in``.  The head-averaged
    attention weights the reference materialises and drops (need_weights=True default,
    hybrid_encoder.py:277, dfine_decoder.py:239) are not computed."""
    d = container.embed_dim
    w, b = container.in_proj_weight, container.in_proj_bias
    qk = K.linear(qk_in, w[: 2 * d], b[: 2 * d])          # packed [.., q | k]
    v = K.linear(v_in, w[2 * d:], b[2 * d:])
    o = K.attention(qk, v, container.num_heads, mask)
    return K.linear(o, container.out_proj.weight, container.out_proj.bias)
