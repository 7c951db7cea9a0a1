"""The drop-in surface (`src.d_fine`, `src.dl.train`) keeps the reference's names and signatures; the Trainer's
host logic (construction order, scheduler, EMA, checkpoint format) runs here on CPU with the oracle provider."""
import inspect
import sys
from pathlib import Path

import pytest
import torch

REF = Path("/root/reference")


def _ref_module(name):
    if not REF.exists():
        pytest.skip("reference checkout not present (GPU box)")
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(REF))
    try:
        import importlib
        return importlib.import_module(name)
    finally:
        sys.path.remove(str(REF))
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_builder_signatures_match_reference():
    ref = _ref_module("src.d_fine.dfine")
    import src.d_fine.dfine as ours
    for fn in ("build_model", "build_loss", "build_optimizer"):
        assert list(inspect.signature(getattr(ours, fn)).parameters) == \
            list(inspect.signature(getattr(ref, fn)).parameters), fn
    assert list(inspect.signature(ours.DFINE.forward).parameters) == \
        list(inspect.signature(ref.DFINE.forward).parameters)


def test_matcher_and_dist_surface_match_reference():
    ref_m = _ref_module("src.d_fine.matcher")
    ref_d = _ref_module("src.d_fine.dist_utils")
    import src.d_fine.dist_utils as du
    import src.d_fine.matcher as m
    assert list(inspect.signature(m.HungarianMatcher.__init__).parameters) == \
        list(inspect.signature(ref_m.HungarianMatcher.__init__).parameters)
    assert list(inspect.signature(m.HungarianMatcher.forward).parameters) == \
        list(inspect.signature(ref_m.HungarianMatcher.forward).parameters)
    for fn in ("init_distributed_mode", "get_rank", "get_world_size", "is_main_process", "get_local_rank",
               "broadcast_scalar", "reduce_dict", "synchronize", "cleanup_distributed",
               "is_dist_available_and_initialized"):
        assert hasattr(ref_d, fn) and hasattr(du, fn), fn


def test_trainer_runs_the_reference_loop_on_cpu(tmp_path):
    from custom_d_fine_b200 import kernels
    from oracle.torch_ops import OracleOps
    from src.dl.train import SyntheticLoader, Trainer, load_cfg
    cfg = load_cfg(None, ["model_name=n", "train.device=cpu", "train.batch_size=2", "train.img_size=[320,320]",
                          "train.epochs=1", f"train.path_to_save={tmp_path}", "train.cuda_graphs=False"])
    loader = SyntheticLoader(2, (320, 320), steps=2, targets_per_image=3)
    tr = Trainer(cfg, train_loader=loader)
    assert isinstance(tr.scheduler, torch.optim.lr_scheduler.OneCycleLR)
    lr0 = tr.optimizer.param_groups[3]["lr"]
    with kernels.use(OracleOps()):
        hist = tr.train()
    assert len(hist) == 1 and torch.isfinite(torch.tensor(hist[0]["loss"]))
    assert tr.optimizer.param_groups[3]["lr"] != lr0, "OneCycleLR must advance once per optimizer step"
    ckpt = torch.load(tmp_path / "last.pt", weights_only=True)
    assert set(ckpt) == set(tr.model.state_dict()), "last.pt is a bare state_dict with the reference's keys"


def test_resume_continues_the_same_trajectory(tmp_path):
    """2 epochs in one run == 1 epoch + checkpoint + restart + 1 epoch: model, EMA, optimizer moments, scheduler and the
    generators all come back from `resume.pt` (the reference cannot resume; its `last.pt` format is kept beside it)."""
    from custom_d_fine_b200 import kernels
    from oracle.torch_ops import OracleOps
    from src.dl.train import SyntheticLoader, Trainer, load_cfg

    def make(epochs, out):
        torch.manual_seed(0)
        cfg = load_cfg(None, ["model_name=n", "train.device=cpu", "train.batch_size=2", "train.img_size=[320,320]",
                              f"train.epochs={epochs}", f"train.path_to_save={out}", "train.cuda_graphs=False"])
        return Trainer(cfg, train_loader=SyntheticLoader(2, (320, 320), steps=2, targets_per_image=3))

    with kernels.use(OracleOps()):
        full = make(2, tmp_path / "a")
        full.train()
        first = make(2, tmp_path / "b")
        first.epochs = 1
        first.train()
        again = make(2, tmp_path / "c")
        assert again.resume(tmp_path / "b" / "resume.pt") == 2
        again.train()
    for (k, a), (_, b) in zip(full.model.state_dict().items(), again.model.state_dict().items()):
        if a.dtype.is_floating_point:
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), k
    for (k, a), (_, b) in zip(full.ema_model.model.state_dict().items(), again.ema_model.model.state_dict().items()):
        if a.dtype.is_floating_point:
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), ("ema", k)
    assert full.optimizer.param_groups[3]["lr"] == again.optimizer.param_groups[3]["lr"]


def test_go_union_fast_paths_equal_the_reference_formulation():
    """The host index planning between the two CUDA graphs of a step evaluates the reference's GO union
    (dfine_criterion.py:570-591: torch.unique(pairs, dim=0) + unstable argsort of the counts + first pair per query)
    on scalar keys / numpy; both fast paths must reproduce the reference formulation pair for pair, in order."""
    import numpy as np
    import torch
    from custom_d_fine_b200.criterion import DFINECriterion, IndexPlan

    def ref_union(q, t):
        ind = torch.cat([q[:, None], t[:, None]], 1)
        unique, counts = torch.unique(ind, return_counts=True, dim=0)
        order = torch.argsort(counts, descending=True)
        seen = {}
        for r, c in unique[order].tolist():
            if r not in seen:
                seen[r] = c
        return list(seen.keys()), list(seen.values())

    rng = np.random.default_rng(0)
    for _ in range(500):
        T, S = int(rng.integers(1, 30)), int(rng.integers(1, 8))
        Q = int(rng.integers(T, 40))                      # few queries: many queries matched to several targets
        oq = np.stack([np.sort(rng.choice(Q, T, replace=False)) for _ in range(S)])
        ot = np.stack([rng.permutation(T) for _ in range(S)])
        q, t = torch.from_numpy(oq.reshape(-1)), torch.from_numpy(ot.reshape(-1))
        want = ref_union(q, t)
        got = DFINECriterion._go_union(q, t)
        assert want[0] == got[0].tolist() and want[1] == got[1].tolist()
        plan = IndexPlan([T], Q, S, None, 0)
        (hq, ht), = DFINECriterion.go_indices_host(oq, ot, plan)
        assert want[0] == hq.tolist() and want[1] == ht.tolist()
    # whole-batch pass (ragged sizes, an image without targets, an image with more targets than queries) == image by image
    for _ in range(100):
        Q, S = int(rng.integers(4, 24)), int(rng.integers(1, 7))
        sizes = [int(rng.integers(0, 30)) for _ in range(int(rng.integers(1, 7)))] + [0]
        rng.shuffle(sizes)
        plan = IndexPlan(sizes, Q, S, None, 0)
        oq, ot = np.zeros((S, sum(sizes)), np.int64), np.zeros((S, sum(sizes)), np.int64)
        for b, (T, n) in enumerate(zip(sizes, plan.per_img)):
            o = int(plan.offs[b])
            for k in range(S):
                oq[k, o:o + n] = np.sort(rng.choice(Q, n, replace=False))
                ot[k, o:o + n] = rng.permutation(T)[:n]
        got = DFINECriterion.go_indices_host(oq, ot, plan)
        for b, n in enumerate(plan.per_img):
            o = int(plan.offs[b])
            q, t = torch.from_numpy(oq[:, o:o + n].reshape(-1)), torch.from_numpy(ot[:, o:o + n].reshape(-1))
            want = DFINECriterion._go_union(q, t) if n else (torch.zeros(0), torch.zeros(0))
            assert want[0].tolist() == got[b][0].tolist() and want[1].tolist() == got[b][1].tolist()
    # table filling: vectorised path == per-image path
    B, S, T, Q = 5, 6, 7, 300
    oq = np.stack([np.concatenate([np.sort(rng.choice(Q, T, replace=False)) for _ in range(B)]) for _ in range(S)])
    ot = np.stack([np.concatenate([rng.permutation(T) for _ in range(B)]) for _ in range(S)])
    p1, p2 = IndexPlan([T] * B, Q, S, None, 0), IndexPlan([T] * B, Q, S, None, 0)
    lists = p1.fill(oq, ot)
    p2.fill(oq, ot, want_lists=False)
    assert torch.equal(p1.table, p2.table)
    n1 = p1.fill_go(DFINECriterion.go_indices(lists[0], lists[1:]))
    n2 = p2.fill_go(DFINECriterion.go_indices_host(oq, ot, p2))
    assert n1 == n2 and torch.equal(p1.table, p2.table)
