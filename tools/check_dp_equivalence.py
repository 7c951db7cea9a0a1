"""Data-parallel check, run under torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_dp_equivalence.py

Trains D-FINE-s for a few steps on per-rank data twice from the same seed — once with the gradient all-reduce of the
encoder / decoder arenas overlapped with the backbone's backward pass (the split backward of train.TrainStep), once with
the plain all-reduce after the whole backward — in eager and in CUDA-graph mode, and compares the parameter checksums of
the runs and of the ranks.  Also times both variants."""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from custom_d_fine_b200 import dist as du  # noqa: E402
from custom_d_fine_b200.model import build_loss, build_model, build_optimizer  # noqa: E402
from custom_d_fine_b200.train import GraphedTrainStep, ModelEMA, TrainStep  # noqa: E402
from tests.golden.common import seeded_fill, synthetic_batch  # noqa: E402

du.init_distributed_mode()
rank, world = du.get_rank(), du.get_world_size()
dev = torch.device("cuda", du.get_local_rank())
x, targets = synthetic_batch(4, 320, 320, seed=100 + rank, T=(5, 3, 7, 2))
x = x.to(dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in targets]


def grads_after_one_step(split):
    """All-reduced gradient arenas of ONE eager step (captured just before the optimizer kernels consume them)."""
    os.environ["DFINE_SPLIT_BWD"] = "1" if split else "0"
    torch.manual_seed(0)
    model = build_model("s", 80, False, dev, img_size=(320, 320))
    seeded_fill(model, 3)
    model.train()
    opt = build_optimizer(model, lr=1e-4, backbone_lr=1e-5, betas=(0.9, 0.999), weight_decay=1e-4, base_lr=1e-4)
    step = TrainStep(model, build_loss("s", 80, 0.0, False), opt, ema=ModelEMA(model, 0.9998), clip_max_norm=0.1)
    got = {}
    orig = opt.step

    def spy(*a, **k):
        torch.cuda.synchronize()
        got["g"] = [x["g"].clone() for x in opt._arenas if x is not None]
        return orig(*a, **k)

    opt.step = spy
    torch.manual_seed(11 + rank)
    torch.cuda.manual_seed(11 + rank)
    step(x, targets)
    torch.cuda.synchronize()
    return got["g"]


def run(cls, split, steps=8):
    os.environ["DFINE_SPLIT_BWD"] = "1" if split else "0"
    torch.manual_seed(0)
    model = build_model("s", 80, False, dev, img_size=(320, 320))
    seeded_fill(model, 3)
    model.train()
    opt = build_optimizer(model, lr=1e-4, backbone_lr=1e-5, betas=(0.9, 0.999), weight_decay=1e-4, base_lr=1e-4)
    step = cls(model, build_loss("s", 80, 0.0, False), opt, ema=ModelEMA(model, 0.9998), clip_max_norm=0.1)
    torch.manual_seed(11 + rank)
    torch.cuda.manual_seed(11 + rank)
    losses = []
    for i in range(steps):
        if i == steps - 3:
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
        loss, _ = step(x, targets)
        losses.append(float(loss))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    assert all(torch.equal(both[0], b) for b in both), f"ranks diverged ({cls.__name__}, split={split})"
    return losses, float(chk), dt


# one-step gradients: plain twice (the run-to-run noise floor of this ill-conditioned seeded setup: float atomics in the
# weight-gradient / BN reductions reorder between runs) and split once — the split must sit inside that floor
ga, ga2, gb = grads_after_one_step(False), grads_after_one_step(False), grads_after_one_step(True)
for i, (u, u2, v) in enumerate(zip(ga, ga2, gb)):
    nrm = u.double().norm().clamp_min(1e-30)
    floor = float((u - u2).double().norm() / nrm)
    e = float((u - v).double().norm() / nrm)
    if rank == 0:
        print(f"gradient arena {i}: split-backward vs plain all-reduce, relative L2 {e:.2e}; plain vs plain (noise floor) "
              f"{floor:.2e} ({u.numel()} elements)")
    assert e < max(1e-4, 5 * floor), f"gradient arena {i}: split vs plain all-reduce differ by {e:.2e} (noise floor {floor:.2e})"
res = {}
for cls in (TrainStep, GraphedTrainStep):
    for split in (False, True):
        res[(cls.__name__, split)] = run(cls, split)
if rank == 0:
    for k, (losses, chk, dt) in res.items():
        print(k, "last loss %.5f" % losses[-1], "checksum %.6f" % chk, "ms/step %.2f" % (dt * 1e3))
    for cls in ("TrainStep", "GraphedTrainStep"):
        a, b = res[(cls, False)], res[(cls, True)]
        rel = abs(a[1] - b[1]) / max(abs(a[1]), 1e-9)
        # (8 steps of this deliberately ill-conditioned seeded setup amplify the atomics-order noise of a step to 1e-3 on
        #  the loss — the eager and the graph-replayed runs differ from each other by as much; the sharp check is the
        #  one-step gradient comparison above)
        assert rel < 1e-4 and abs(a[0][-1] - b[0][-1]) <= 5e-3 * abs(a[0][-1]), (cls, a[1], b[1], a[0][-1], b[0][-1])
    print("data-parallel equivalence ok")
du.cleanup_distributed()
