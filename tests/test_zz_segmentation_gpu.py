"""Segmentation rows through the CUDA table, against the CPU oracle.

The host graph, the oracle ops and their parity with the REAL reference are pinned on CPU
(tests/test_oracle_cpu.py::test_seg_*).  The CUDA side: the plain-conv autograd binding over the library's conv kernels,
the matcher kernel's additive mask-cost input (`dfine_matcher_extra`), GroupNorm / bilinear resize / mask product as
torch device ops.  Both tests passed on a B200 in the last GPU call of round 1 (tools/gpu_trip38.sh).
"""
import pytest
import torch

from custom_d_fine_b200 import kernels
from custom_d_fine_b200.model import build_loss, build_model
from tests.golden.common import rect_masks, seeded_fill, synthetic_batch
from tests.util import check_rows_up_to_order

pytestmark = pytest.mark.gpu


def test_matcher_extra_cost_matches_oracle(cuda_ops, oracle_ops):
    g = torch.Generator().manual_seed(5)
    NL, B, Q, C, sizes = 3, 2, 60, 80, [7, 4]
    logits = [torch.randn(B, Q, C, generator=g) for _ in range(NL)]
    boxes = [torch.rand(B, Q, 4, generator=g) * 0.5 + 0.2 for _ in range(NL)]
    targets = [{"labels": torch.randint(0, C, (n,), generator=g), "boxes": torch.rand(n, 4, generator=g) * 0.4 + 0.2}
               for n in sizes]
    extra = torch.rand(NL, Q * sum(sizes), generator=g) * 3.0
    want = oracle_ops.match(logits, boxes, targets, extra_cost=extra)
    got = cuda_ops.match([t.cuda() for t in logits], [t.cuda() for t in boxes],
                         [{k: v.cuda() for k, v in t.items()} for t in targets], extra_cost=extra.cuda())
    for l in range(NL):
        for b in range(B):
            assert got[l][b][0].tolist() == want[l][b][0].tolist() and got[l][b][1].tolist() == want[l][b][1].tolist()


def test_segmentation_step_matches_cpu_oracle(cuda_ops, oracle_ops):
    hw, seed = 320, 2
    x, targets = synthetic_batch(2, hw, hw, seed=1234 + seed)
    for t in targets:
        t["masks"] = rect_masks(t["boxes"], hw, hw)
    runs = {}
    for dev in ("cpu", "cuda"):
        torch.manual_seed(0)
        model = build_model("s", 80, True, dev, img_size=(hw, hw))
        seeded_fill(model, seed)
        model.train()
        xs = x.to(dev)
        tg = [{k: v.to(dev) for k, v in t.items()} for t in targets]
        crit = build_loss("s", 80, 0.0, True)
        torch.manual_seed(7)
        from tests.test_model_gpu import _host_rng
        with _host_rng():
            if dev == "cpu":
                with kernels.use(oracle_ops):
                    out = model(xs, targets=tg)
                    losses = crit(out, tg)
                    sum(losses.values()).backward()
            else:
                out = model(xs, targets=tg)
                losses = crit(out, tg)
                sum(losses.values()).backward()
                torch.cuda.synchronize()
        runs[dev] = (model, out, losses)
    (m0, o0, l0), (m1, o1, l1) = runs["cpu"], runs["cuda"]
    assert list(l0.keys()) == list(l1.keys())
    for k in l0:
        a, b = float(l1[k]), float(l0[k])
        assert abs(a - b) <= 3e-3 * max(abs(b), 1e-2), (k, a, b)
    both = torch.cat([o1["pred_logits"], o1["pred_boxes"]], -1)
    both_ref = torch.cat([o0["pred_logits"], o0["pred_boxes"]], -1)
    check_rows_up_to_order("seg: pred_logits|pred_boxes", both, both_ref, 1e-3, 1.0)
    d = (o1["dn_pred_masks"].cpu() - o0["dn_pred_masks"]).abs().max() / o0["dn_pred_masks"].abs().max()
    assert float(d) <= 1e-3, float(d)


def test_segmentation_graph_replay_matches_eager(cuda_ops):
    """Segmentation batches (mask matching cost, mask losses, MaskDecoder) replayed as CUDA graphs against the eager step:
    same weights, batch and generator state -> the loss trajectories agree (atomics-order noise only)."""
    from custom_d_fine_b200.model import build_optimizer
    from custom_d_fine_b200.train import GraphedTrainStep, ModelEMA, TrainStep
    x, targets = synthetic_batch(2, 320, 320, seed=5, T=(6, 3))
    for t in targets:
        t["masks"] = rect_masks(t["boxes"], 320, 320)
    x = x.cuda()
    targets = [{k: v.cuda() for k, v in t.items()} for t in targets]
    traj = {}
    for name, cls in (("eager", TrainStep), ("graph", GraphedTrainStep)):
        torch.manual_seed(0)
        model = build_model("s", 80, True, "cuda", img_size=(320, 320))
        seeded_fill(model, 3)
        model.train()
        opt = build_optimizer(model, lr=1e-4, backbone_lr=1e-5, betas=(0.9, 0.999), weight_decay=1e-4, base_lr=1e-4)
        step = cls(model, build_loss("s", 80, 0.0, True), opt, ema=ModelEMA(model, 0.9998), clip_max_norm=0.1)
        torch.manual_seed(11)
        torch.cuda.manual_seed(11)
        traj[name] = [float(step(x, targets)[0]) for _ in range(7)]
        if name == "graph":
            assert step._graphs, "no CUDA graph was captured for the segmentation batch"
    for a, b in zip(traj["eager"], traj["graph"]):
        assert abs(a - b) <= 2e-2 * abs(a), traj


def test_groupnorm_resize_maskdot_kernels_match_oracle(cuda_ops, oracle_ops):
    """Segmentation-head kernels (csrc/seg.cu, io.cu, the grouped tcgen05 mask product) against the oracle's torch ops,
    forward and backward."""
    from tests.util import run_both
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 20, 24, 64, generator=g, requires_grad=True)
    w = (torch.rand(64, generator=g) + 0.5).requires_grad_(True)
    b = (torch.randn(64, generator=g) * 0.1).requires_grad_(True)
    for act in (None, "relu"):
        run_both(lambda K, x, w, b: K.group_norm(x, 16 if act is None else 8, w, b, 1e-5, act=act), cuda_ops, oracle_ops,
                 [x, w, b], 2e-5, 2e-4)
    for size in ((40, 48), (33, 31), (20, 24)):
        run_both(lambda K, x: K.resize_bilinear(x, size), cuda_ops, oracle_ops, [x], 2e-6, 2e-5)
    e = torch.randn(3, 300, 64, generator=g, requires_grad=True)
    f = torch.randn(3, 20, 28, 64, generator=g, requires_grad=True)
    run_both(lambda K, e, f: K.mask_dot(e, f), cuda_ops, oracle_ops, [e, f], 2e-5, 2e-3, grad_metric="l2")


def test_mask_cost_kernel_matches_torch_formulation(cuda_ops):
    """The fused mask-cost kernel against the reference formulation in torch ops (matcher.py:19-71, 175-237) on the same
    device tensors: GT masks with ragged counts incl. an image without targets."""
    from custom_d_fine_b200.matcher import HungarianMatcher
    g = torch.Generator().manual_seed(2)
    B, Q, Hm, Wm, sizes = 3, 300, 24, 20, [5, 0, 19]
    outs = [{"pred_logits": torch.randn(B, Q, 80, generator=g).cuda(), "pred_boxes": torch.rand(B, Q, 4, generator=g).cuda(),
             "pred_masks": (torch.randn(B, Hm, Wm, Q, generator=g) * 3).cuda().permute(0, 3, 1, 2)} for _ in range(2)]
    targets = [{"labels": torch.zeros(n, dtype=torch.int64).cuda(), "boxes": torch.rand(n, 4, generator=g).cuda(),
                "masks": (torch.rand(n, 2 * Hm, 2 * Wm, generator=g) > 0.6).to(torch.uint8).cuda()} for n in sizes]
    m = HungarianMatcher({"cost_class": 2, "cost_bbox": 5, "cost_giou": 2, "cost_mask": 1, "cost_mask_dice": 1},
                         use_focal_loss=True, alpha=0.25, gamma=2.0)
    got = m.mask_cost(outs, targets)
    saved = type(cuda_ops).mask_cost_layer
    try:
        del type(cuda_ops).mask_cost_layer          # force the torch formulation
        want = m.mask_cost(outs, targets)
    finally:
        type(cuda_ops).mask_cost_layer = saved
    assert got.shape == want.shape
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 2e-5, err


def test_mask_loss_kernel_matches_torch_formulation(cuda_ops, monkeypatch):
    """The fused cropped BCE + Dice kernels (csrc/seg.cu) against the torch formulation of dfine_criterion.py:335-450 in
    DFINECriterion.loss_masks: values and the gradient of the matched mask logits, incl. boxes that touch the borders and
    a degenerate (sub-pixel) box."""
    from types import SimpleNamespace
    from custom_d_fine_b200.criterion import DFINECriterion
    g = torch.Generator().manual_seed(4)
    B, Q, Hm, Wm, T, M = 2, 50, 40, 36, 9, 23
    boxes = torch.rand(T, 4, generator=g) * 0.5 + 0.2
    boxes[0] = torch.tensor([0.02, 0.5, 0.2, 0.9])          # clipped on the left
    boxes[1] = torch.tensor([0.5, 0.99, 0.6, 0.3])          # clipped at the bottom
    boxes[2] = torch.tensor([0.31, 0.47, 0.004, 0.003])     # smaller than a pixel
    tg = (None, boxes.cuda(), (torch.rand(T, 2 * Hm, 2 * Wm, generator=g) > 0.5).to(torch.uint8).cuda())
    S = SimpleNamespace(n=M, v=None, t=torch.randint(0, T, (M,), generator=g).cuda(), b=None, q=None)
    out = {"pred_masks": torch.zeros(B, Q, Hm, Wm, device="cuda")}
    crit = DFINECriterion.__new__(DFINECriterion)
    res = {}
    for mode in ("kernel", "torch"):
        monkeypatch.setenv("DFINE_MASK_LOSS", mode)
        src = (torch.randn(M, Hm, Wm, generator=torch.Generator().manual_seed(5)) * 3).cuda().requires_grad_(True)
        d = crit.loss_masks(out, S, tg, 1.0, src=src)
        (d["loss_mask_bce"] * 1.7 + d["loss_mask_dice"] * 0.6).backward()
        res[mode] = (d["loss_mask_bce"].item(), d["loss_mask_dice"].item(), src.grad.clone())
    k, t = res["kernel"], res["torch"]
    assert abs(k[0] - t[0]) < 1e-5 * max(1.0, abs(t[0])) and abs(k[1] - t[1]) < 1e-5
    assert float((k[2] - t[2]).abs().max()) < 1e-6 + 1e-5 * float(t[2].abs().max())


def test_mask_loss_pixel_major_matches_row_major(cuda_ops):
    """The pixel-major loss kernels (logits [B,Hm,Wm,R] as the tcgen05 mask product writes them, padding rows marked -1)
    against the row-major kernels (which the previous test pins to the torch formulation): values and gradients."""
    from custom_d_fine_b200 import cuda_ops as co
    g = torch.Generator().manual_seed(7)
    B, R, Hm, Wm, T = 3, 44, 40, 36, 9
    boxes = (torch.rand(T, 4, generator=g) * 0.5 + 0.2).cuda()
    boxes[0] = torch.tensor([0.02, 0.5, 0.2, 0.9])
    boxes[1] = torch.tensor([0.31, 0.47, 0.004, 0.003])
    gt = torch.rand(T, Hm, Wm, generator=g).cuda()
    t_pad = torch.randint(-1, T, (B * R,), generator=g).cuda()
    t_pad[:5] = -1
    valid = t_pad >= 0
    pred = (torch.randn(B, Hm, Wm, R, generator=g) * 3).cuda().requires_grad_(True)
    w1, w2 = torch.rand(B * R, generator=g).cuda(), torch.rand(B * R, generator=g).cuda()
    bce, dice = co._MaskLossPM.apply(pred, gt, t_pad, boxes)
    ((bce * w1).sum() + (dice * w2).sum()).backward()
    rows = pred.detach().permute(0, 3, 1, 2).reshape(B * R, Hm, Wm)[valid].clone().requires_grad_(True)
    bce_r, dice_r = co._MaskLossRows.apply(rows, gt, t_pad[valid], boxes)
    ((bce_r * w1[valid]).sum() + (dice_r * w2[valid]).sum()).backward()
    assert float(bce[~valid].abs().max()) == 0.0 and float(dice[~valid].abs().max()) == 0.0
    assert float((bce[valid] - bce_r).abs().max()) < 1e-5 * float(bce_r.abs().max())
    assert float((dice[valid] - dice_r).abs().max()) < 1e-5
    gp = pred.grad.permute(0, 3, 1, 2).reshape(B * R, Hm, Wm)
    assert float(gp[~valid].abs().max()) == 0.0
    assert float((gp[valid] - rows.grad).abs().max()) < 1e-6 + 1e-5 * float(rows.grad.abs().max())
