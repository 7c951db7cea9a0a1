"""Flat-arena fused optimizer (csrc/optim.cu: clip + AdamW + EMA + zero_grad) against the reference's call
sequence on torch CPU ops: clip_grad_norm_(0.1) -> torch.optim.AdamW.step -> zero_grad -> ModelEMA.update
(/root/reference/src/dl/train.py:512-535, 62-73).  Floating-point kernel: tolerance 5e-6 of each tensor's max (measured 2.0e-6)."""
import copy
import math

import pytest
import torch
import torch.nn as nn

from tests.util import check_close

pytestmark = pytest.mark.gpu


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = nn.Sequential(nn.Conv2d(3, 7, 3), nn.BatchNorm2d(7))
        self.encoder = nn.Linear(13, 5)
        self.decoder = nn.Linear(5, 3)
        self.head = nn.Parameter(torch.randn(33))


def _groups(model):
    from custom_d_fine_b200.model import param_groups
    return param_groups(model, 2e-5, 1.5e-4)


def test_fused_adamw_ema_matches_torch(cuda_ops):
    from custom_d_fine_b200.optim import FusedAdamW
    from custom_d_fine_b200.train import ModelEMA
    torch.manual_seed(0)
    ref = _Toy()
    dev = copy.deepcopy(ref).cuda()
    ema_ref, ema_dev = ModelEMA(ref, 0.9998), ModelEMA(dev, 0.9998)
    opt_ref = torch.optim.AdamW(_groups(ref), lr=1.5e-4, betas=(0.9, 0.999), weight_decay=1.25e-4)
    opt_dev = FusedAdamW(_groups(dev), lr=1.5e-4, betas=(0.9, 0.999), weight_decay=1.25e-4, max_norm=0.1)
    opt_dev.attach_ema(ema_dev, dev)
    g = torch.Generator().manual_seed(1)
    for it in range(1, 6):
        for gi, grp in enumerate(opt_ref.param_groups):          # a scheduler changing lr between steps
            grp["lr"] = opt_dev.param_groups[gi]["lr"] = grp["initial_lr"] * (1 + 0.1 * it)
        scale = 10.0 if it % 2 else 1e-3                         # clipped and un-clipped steps
        for (n, p), (_, q) in zip(ref.named_parameters(), dev.named_parameters()):
            gr = torch.randn(p.shape, generator=g) * scale
            p.grad = gr.clone()
            q.grad.copy_(gr)                                     # gradients live in the flat arena
        ref.backbone[1].running_mean.add_(0.01 * it)             # a floating-point buffer the EMA must track
        dev.backbone[1].running_mean.add_(0.01 * it)
        total = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.1)
        opt_ref.step()
        opt_ref.zero_grad()
        ema_ref.update(it, ref)
        opt_dev.prepare(ema_ref.ema_scheduler(it))
        opt_dev.step()
        torch.cuda.synchronize()
        check_close("grad norm", opt_dev.grad_norm().float().cpu(), total.reshape(1), 1e-5)
        for (n, p), (_, q) in zip(ref.named_parameters(), dev.named_parameters()):
            check_close(f"param {n} step {it}", q, p, 5e-6)
            assert float(q.grad.abs().max()) == 0.0, "step() must leave zeroed gradients"
        es, ed = ema_ref.model.state_dict(), ema_dev.model.state_dict()
        for k in es:
            if es[k].dtype.is_floating_point:
                check_close(f"ema {k} step {it}", ed[k], es[k], 5e-6)
    assert math.isfinite(float(opt_dev.grad_norm()))
    # the AdamW kernel also maintains the operand planes of the forward GEMMs
    from custom_d_fine_b200 import cuda_ops as co
    for a in opt_dev._arenas:
        if a is not None and a["planes"] is not None:
            hi, lo = a["planes"][0], a["planes"][1]
            if hi.dtype == torch.float16:      # 3xFP16: fp16 parts of w * 2^8, 22 significand bits together
                back = (hi.double() + lo.double()) / co._F16_WSCALE
                err = (back - a["p"].double()).abs()
                # 2^-22 relative, or half the smallest fp16 subnormal (2^-25) in the scaled domain for tiny weights
                assert bool((err <= 2.0 ** -21 * a["p"].double().abs() + 2.0 ** -25 / co._F16_WSCALE).all())
                assert torch.equal(hi, (a["p"] * co._F16_WSCALE).half())
            else:                              # 3xTF32: hi is tf32-representable, hi + lo is the parameter
                assert torch.equal(hi + lo, a["p"])
                assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0


def test_fused_adamw_state_dict_round_trip(cuda_ops):
    """The flat-arena optimizer's checkpoint is torch.optim.AdamW's own layout: a stock AdamW resumes from it, and a fresh
    FusedAdamW loading it continues exactly where the first one stopped (moments, step count, bias correction)."""
    from custom_d_fine_b200.optim import FusedAdamW
    torch.manual_seed(0)
    base = _Toy().cuda()
    g = torch.Generator().manual_seed(4)
    grads = [[torch.randn(p.shape, generator=g).cuda() for p in base.parameters()] for _ in range(4)]

    def run(model, opt, steps):
        for gs in steps:
            for p, gr in zip(model.parameters(), gs):
                p.grad.copy_(gr)
            opt.prepare(None)
            opt.step()

    m1 = copy.deepcopy(base)
    o1 = FusedAdamW(_groups(m1), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, max_norm=0.1)
    run(m1, o1, grads)
    m2 = copy.deepcopy(base)
    o2 = FusedAdamW(_groups(m2), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, max_norm=0.1)
    run(m2, o2, grads[:2])
    sd, weights = o2.state_dict(), copy.deepcopy(m2.state_dict())
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    # a stock AdamW accepts the same dict
    m_t = copy.deepcopy(base)
    m_t.load_state_dict(weights)
    o_t = torch.optim.AdamW(_groups(m_t), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2)
    o_t.load_state_dict({k: v for k, v in sd.items() if k != "fused"})
    assert float(o_t.state[next(iter(m_t.parameters()))]["step"]) == 2.0
    m3 = copy.deepcopy(base)
    m3.load_state_dict(weights)
    o3 = FusedAdamW(_groups(m3), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, max_norm=0.1)
    o3.load_state_dict(sd)
    run(m3, o3, grads[2:])
    torch.cuda.synchronize()
    for (n, p), (_, q) in zip(m1.named_parameters(), m3.named_parameters()):
        check_close(f"resumed param {n}", q, p, 1e-6)
