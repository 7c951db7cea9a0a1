"""The criterion kernels' arithmetic (custom_d_fine_b200/csrc/loss_math.cuh — the very functions the CUDA kernels of
loss.cu call) compiled for the HOST and driven serially (tests/host_harness/loss_host.cpp), against the torch
restatement of the reference criterion (`DFINECriterion._families_torch`, itself pinned to the real reference's loss dict
by tests/test_oracle_cpu.py): every VFL / L1 / GIoU / FGL / DDF scalar of a D-FINE-s train step and the gradients with
respect to the stacked logits, boxes and corner logits.  Runs without a GPU."""
import ctypes
import subprocess
from pathlib import Path

import pytest
import torch

from custom_d_fine_b200 import kernels, loss_desc as ld
from custom_d_fine_b200.decoder import weighting_function
from custom_d_fine_b200.model import build_loss, build_model
from tests.golden.common import seeded_fill, synthetic_batch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("h") / "libloss_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so),
                           str(ROOT / "tests" / "host_harness" / "loss_host.cpp")])
    lib = ctypes.CDLL(str(so))
    assert lib.loss_host_desc_size() == ctypes.sizeof(ld.LossDesc)
    return lib


@pytest.mark.parametrize("sizes", [(10, 7), (3, 0)])
def test_kernel_arithmetic_matches_torch_criterion(harness, oracle_ops, sizes):
    torch.manual_seed(0)
    model = build_model("s", 80, False, "cpu", img_size=(320, 320))
    seeded_fill(model, 5)
    model.train()
    x, targets = synthetic_batch(2, 320, 320, seed=77, T=sizes)
    crit = build_loss("s", 80, 0.0, False)
    with kernels.use(oracle_ops):
        torch.manual_seed(3)
        out = model(x, targets=targets)
        raw, tg = crit.match(out, targets)
        plan = crit.plan(out, targets, raw)
    table, counts = plan.table.clone(), plan.counts.clone()
    full = out["_stacked"]["full"]
    L = full["logits"].shape[0]
    # detached leaf copies of the unsplit stacks so that both sides differentiate with respect to the same tensors
    leaf = {k: full[k].detach().clone().requires_grad_(True) for k in ("logits", "boxes", "corners", "pre_logits", "pre_boxes")}
    enc_l = out["enc_aux_outputs"][0]["pred_logits"].detach().clone().requires_grad_(True)
    enc_b = out["enc_aux_outputs"][0]["pred_boxes"].detach().clone().requires_grad_(True)
    n_dn = full["n_dn"]
    o2 = {"up": out["up"], "reg_scale": out["reg_scale"], "dn_meta": out["dn_meta"], "dn_outputs": out["dn_outputs"],
          "pre_outputs": {"pred_logits": leaf["pre_logits"][:, n_dn:], "pred_boxes": leaf["pre_boxes"][:, n_dn:]},
          "dn_pre_outputs": {"pred_logits": leaf["pre_logits"][:, :n_dn], "pred_boxes": leaf["pre_boxes"][:, :n_dn]},
          "enc_aux_outputs": [{"pred_logits": enc_l, "pred_boxes": enc_b}],
          "_stacked": {"logits": leaf["logits"][:, :, n_dn:], "boxes": leaf["boxes"][:, :, n_dn:],
                       "corners": leaf["corners"][:, :, n_dn:], "refs": full["refs"][:, :, n_dn:],
                       "dn_logits": leaf["logits"][:, :, :n_dn], "dn_boxes": leaf["boxes"][:, :, :n_dn],
                       "dn_corners": leaf["corners"][:, :, :n_dn], "dn_refs": full["refs"][:, :, :n_dn]}}
    crit._clear_cache()
    A, DN = crit._families_torch(o2, tg, table, counts, plan)
    H = L + 2
    # upstream gradients: a different random weight per scalar
    g = torch.Generator().manual_seed(9)
    gout = torch.rand(6 * H + 4 * L, generator=g) + 0.5
    gv, gl, gg, gf, gd = ld.split_out(gout, L)
    perm = torch.tensor([L - 1] + list(range(L - 1)) + [L, L + 1])          # head order -> plan order
    total = (A[0] * gv[0][perm]).sum() + (A[1] * gl[0][perm]).sum() + (A[2] * gg[0][perm]).sum() + \
        (A[3] * gf[0]).sum() + (A[4] * gd[0][:L - 1]).sum()
    total = total + (DN[0] * gv[1][:L + 1]).sum() + (DN[1] * gl[1][:L + 1]).sum() + (DN[2] * gg[1][:L + 1]).sum() + \
        (DN[3] * gf[1]).sum() + (DN[4] * gd[1][:L - 1]).sum()
    total.backward()

    t = dict(logits=leaf["logits"].detach().contiguous(), boxes=leaf["boxes"].detach().contiguous(),
             corners=leaf["corners"].detach().contiguous(), ref0=full["refs"][0].detach().contiguous().float(),
             pre_logits=leaf["pre_logits"].detach().contiguous(), pre_boxes=leaf["pre_boxes"].detach().contiguous(),
             enc_logits=enc_l.detach().contiguous(), enc_boxes=enc_b.detach().contiguous(), table=table.contiguous(),
             labels=tg[0].contiguous(), tboxes=tg[1].contiguous().float(), counts=counts.contiguous().float(),
             project=weighting_function(32, out["up"], out["reg_scale"]).detach().contiguous().float(),
             reg_scale=out["reg_scale"].detach().reshape(-1).float().contiguous())
    meta = dict(n_dn=n_dn, n_layer=plan.n_layer, go_cap=plan.go_cap, n_dn_entries=plan.n_dn,
                dn_groups=out["dn_meta"]["dn_num_group"], alpha=crit.alpha, gamma=crit.gamma, T=5.0)
    d = ld.build(t, meta)
    res = torch.zeros(ld.out_count(L))
    grads = {k: torch.zeros_like(t[k]) for k in ("logits", "pre_logits", "enc_logits", "boxes", "pre_boxes", "enc_boxes", "corners")}
    P = lambda v: ctypes.c_void_p(v.data_ptr())   # noqa: E731
    rc = harness.loss_host_run(ctypes.byref(d), P(res), P(gout.contiguous()), P(grads["logits"]), P(grads["pre_logits"]),
                               P(grads["enc_logits"]), P(grads["boxes"]), P(grads["pre_boxes"]), P(grads["enc_boxes"]),
                               P(grads["corners"]))
    assert rc == 0
    vfl, l1, gi, fgl, ddf = ld.split_out(res, L)
    want = {"vfl": A[0], "l1": A[1], "giou": A[2], "fgl": A[3], "ddf": A[4]}
    got = {"vfl": vfl[0][perm], "l1": l1[0][perm], "giou": gi[0][perm], "fgl": fgl[0], "ddf": ddf[0][:L - 1]}
    for k in want:
        assert torch.allclose(got[k], want[k].detach(), rtol=2e-5, atol=1e-6), (k, got[k], want[k])
    want = {"vfl": DN[0], "l1": DN[1], "giou": DN[2], "fgl": DN[3], "ddf": DN[4]}
    got = {"vfl": vfl[1][:L + 1], "l1": l1[1][:L + 1], "giou": gi[1][:L + 1], "fgl": fgl[1], "ddf": ddf[1][:L - 1]}
    for k in want:
        assert torch.allclose(got[k], want[k].detach(), rtol=2e-5, atol=1e-6), ("dn", k, got[k], want[k])
    ref_g = {"logits": leaf["logits"].grad, "boxes": leaf["boxes"].grad, "corners": leaf["corners"].grad,
             "pre_logits": leaf["pre_logits"].grad, "pre_boxes": leaf["pre_boxes"].grad, "enc_logits": enc_l.grad,
             "enc_boxes": enc_b.grad}
    for k, r in ref_g.items():
        e = float((grads[k] - r).norm() / r.norm().clamp_min(1e-20))
        assert e < 2e-5, (k, e)
