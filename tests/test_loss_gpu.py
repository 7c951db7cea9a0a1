"""Criterion kernels (csrc/loss.cu: VFL, L1 + GIoU, FGL + DDF over every loss head in ~6 launches) through the C ABI
against the torch restatement of the reference criterion evaluated on the same device tensors: every loss scalar and
the gradients with respect to the stacked logits / boxes / corner logits.  (The arithmetic itself is also pinned on CPU
by tests/test_loss_math_cpu.py, and the whole step against the oracle by tests/test_model_gpu.py.)"""
import pytest
import torch

from custom_d_fine_b200 import loss_desc as ld
from custom_d_fine_b200.model import build_loss, build_model
from tests.golden.common import seeded_fill, synthetic_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,sizes", [("s", (10, 7)), ("s", (3, 0)), ("m", (10, 10))])
def test_loss_kernels_match_torch_criterion(cuda_ops, size, sizes):
    torch.manual_seed(0)
    model = build_model(size, 80, False, "cuda", img_size=(320, 320))
    seeded_fill(model, 5)
    model.train()
    x, targets = synthetic_batch(2, 320, 320, seed=77, T=sizes)
    x = x.cuda()
    targets = [{k: v.cuda() for k, v in t.items()} for t in targets]
    crit = build_loss(size, 80, 0.0, False)
    torch.manual_seed(3)
    out = model(x, targets=targets)
    raw, tg = crit.match(out, targets)
    plan = crit.plan(out, targets, raw)
    table, counts = plan.table.cuda(), plan.counts.cuda()
    full = out["_stacked"]["full"]
    L = full["logits"].shape[0]
    n_dn = full["n_dn"]
    res = {}
    for which in ("torch", "kernel"):
        leaf = {k: full[k].detach().clone().requires_grad_(True) for k in ("logits", "boxes", "corners", "pre_logits", "pre_boxes")}
        enc_l = out["enc_aux_outputs"][0]["pred_logits"].detach().clone().requires_grad_(True)
        enc_b = out["enc_aux_outputs"][0]["pred_boxes"].detach().clone().requires_grad_(True)
        o2 = {"up": out["up"], "reg_scale": out["reg_scale"], "dn_meta": out["dn_meta"], "dn_outputs": out["dn_outputs"],
              "pre_outputs": {"pred_logits": leaf["pre_logits"][:, n_dn:], "pred_boxes": leaf["pre_boxes"][:, n_dn:]},
              "dn_pre_outputs": {"pred_logits": leaf["pre_logits"][:, :n_dn], "pred_boxes": leaf["pre_boxes"][:, :n_dn]},
              "enc_aux_outputs": [{"pred_logits": enc_l, "pred_boxes": enc_b}],
              "_stacked": {"logits": leaf["logits"][:, :, n_dn:], "boxes": leaf["boxes"][:, :, n_dn:],
                           "corners": leaf["corners"][:, :, n_dn:], "refs": full["refs"][:, :, n_dn:],
                           "dn_logits": leaf["logits"][:, :, :n_dn], "dn_boxes": leaf["boxes"][:, :, :n_dn],
                           "dn_corners": leaf["corners"][:, :, :n_dn], "dn_refs": full["refs"][:, :, :n_dn],
                           "full": dict(leaf, refs=full["refs"], n_dn=n_dn)}}
        crit._clear_cache()
        fn = crit._families_torch if which == "torch" else crit._families_kernel
        A, DN = fn(o2, tg, table, counts, plan)
        g = torch.Generator().manual_seed(9)
        total = 0
        for fam in A + DN:
            total = total + (fam * (torch.rand(fam.shape, generator=g) + 0.5).cuda()).sum()
        total.backward()
        torch.cuda.synchronize()
        res[which] = (A, DN, dict(leaf, enc_logits=enc_l, enc_boxes=enc_b))
    (A0, D0, g0), (A1, D1, g1) = res["torch"], res["kernel"]
    names = ("vfl", "l1", "giou", "fgl", "ddf")
    for grp, (a, b) in (("A", (A0, A1)), ("DN", (D0, D1))):
        for n, u, v in zip(names, a, b):
            assert torch.allclose(v, u.detach(), rtol=1e-4, atol=1e-6), (grp, n, v, u)
    for k in g0:
        r, got = g0[k].grad, g1[k].grad
        e = float((got - r).norm() / r.norm().clamp_min(1e-20))
        assert e < 1e-4, (k, e)
