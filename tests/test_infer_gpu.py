"""Inference path on the GPU: eval-mode and deploy()-mode model outputs against the golden
fixture produced by the REAL reference's deployed model (tests/golden/make_golden_eval.py), the device post-processor
against the reference's post-processing arithmetic (dl/export.py:59-100), input preparation (uint8 -> float, resize,
BGR->RGB) against torch, and the top-k selection kernel against torch.topk."""
from pathlib import Path

import pytest
import torch

from custom_d_fine_b200.model import build_model
from custom_d_fine_b200.postprocess import DFINEPostProcessor, multiscale_resize, prepare_inputs
from tests.golden.common import seeded_fill, synthetic_batch
from tests.util import check_close, check_rows_up_to_order

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("case", [0, 1], ids=["detect", "segment"])
@pytest.mark.parametrize("deploy", [False, True], ids=["eval", "deploy"])
def test_eval_and_deploy_outputs_match_reference_fixture(cuda_ops, case, deploy):
    fix = torch.load(GOLD / "eval_s_320.pt", weights_only=False)[case]
    torch.manual_seed(0)
    model = build_model(fix["size"], 80, fix["seg"], "cuda", img_size=(fix["hw"], fix["hw"]))
    seeded_fill(model, fix["seed"])
    model.eval()
    if deploy:
        n_before = sum(1 for _ in model.modules())
        model.deploy()
        assert sum(1 for _ in model.modules()) < n_before, "deploy() must fold / drop modules"
    x, _ = synthetic_batch(fix["B"], fix["hw"], fix["hw"], seed=1234 + fix["seed"])
    with torch.no_grad():
        out = model(x.cuda())
    torch.cuda.synchronize()
    assert sorted(out.keys()) == fix["keys"]
    both = torch.cat([out["pred_logits"], out["pred_boxes"]], -1)
    both_ref = torch.cat([fix["pred_logits"], fix["pred_boxes"]], -1)
    check_rows_up_to_order(f"{case}/deploy={deploy}: pred_logits|pred_boxes", both, both_ref, 1e-3, 1.0)
    if fix["seg"]:
        # masks of the first queries, paired through the row order of the logits
        ref_m = fix["pred_masks_q0_6"]                                     # [B, 6, Hm, Wm]: the reference's first six queries
        for b in range(ref_m.shape[0]):
            d = torch.cdist(both[b].double().cpu(), both_ref[b].double(), p=float("inf")).argmin(0)[:6]
            check_close("pred_masks (sigmoid)", out["pred_masks"][b, d].cpu(), ref_m[b], 2e-3)


def test_postprocessor_matches_reference_arithmetic(cuda_ops, oracle_ops):
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(4, 300, 80, generator=g) * 2
    boxes = torch.rand(4, 300, 4, generator=g) * 0.5 + 0.2
    masks = torch.rand(4, 300, 8, 8, generator=g)
    from custom_d_fine_b200 import kernels
    with kernels.use(oracle_ops):
        want = DFINEPostProcessor(80)({"pred_logits": logits, "pred_boxes": boxes, "pred_masks": masks}, 640, 480)
    got = DFINEPostProcessor(80)({"pred_logits": logits.cuda(), "pred_boxes": boxes.cuda(), "pred_masks": masks.cuda()}, 640, 480)
    # equal scores (distinct logits may share one fp32 sigmoid value) may come in either order: compare the sorted scores,
    # and labels / boxes / masks position by position wherever the score is unique in its row
    check_close("scores", got[2].cpu(), want[2], 1e-6)
    s = want[2]
    uniq = torch.ones_like(s, dtype=torch.bool)
    uniq[:, 1:] &= s[:, 1:] != s[:, :-1]
    uniq[:, :-1] &= s[:, :-1] != s[:, 1:]
    assert float(uniq.float().mean()) > 0.9
    assert torch.equal(got[0].cpu()[uniq], want[0][uniq]), "labels"
    assert torch.equal(got[1].cpu()[uniq], want[1][uniq]), "boxes (integer-rounded pixels)"
    assert torch.equal(got[3].cpu()[uniq], want[3][uniq]), "masks"


@pytest.mark.parametrize("B,L,C,k", [(3, 8400, 80, 300), (2, 33600, 80, 300), (2, 1000, 1, 300), (1, 300, 5, 300)])
def test_topk_kernel_matches_torch(cuda_ops, B, L, C, k):
    g = torch.Generator().manual_seed(L + C)
    logits = torch.randn(B, L, C, generator=g).cuda()
    logits[0, 5] = logits[0, 17]                      # an exact tie: the lower token index wins
    idx = cuda_ops.select_topk(logits, k)
    sc = logits.max(-1).values
    vals, ref = torch.topk(sc, k, dim=-1)
    assert torch.equal(sc.gather(1, idx), vals), "selected scores, descending"
    for b in range(B):
        assert sorted(idx[b].tolist()) == sorted(ref[b].tolist()) or float(vals[b, -1]) == float(torch.topk(sc[b], k + 1)[0][-1])
    tie = (idx[0] == 5).nonzero()
    if len(tie):
        assert int(idx[0, int(tie[0]) + 1]) == 17


def test_input_preparation_matches_torch(cuda_ops, oracle_ops):
    g = torch.Generator().manual_seed(8)
    img = torch.randint(0, 256, (2, 375, 500, 3), generator=g, dtype=torch.uint8)
    for size in (None, (640, 640), (320, 416)):
        want = oracle_ops.preprocess_u8(img, size, 1.0 / 255.0, True).permute(0, 3, 1, 2)
        got = prepare_inputs(img.cuda(), size, bgr=True)
        assert got.shape == want.shape
        check_close(f"prepare_inputs {size}", got.cpu(), want, 2e-6)
    x = torch.rand(2, 3, 64, 96, generator=g)
    want = torch.nn.functional.interpolate(x, size=(128, 160), mode="bilinear", align_corners=False)
    check_close("multiscale resize", multiscale_resize(x.cuda(), (128, 160)).cpu(), want, 2e-6)
