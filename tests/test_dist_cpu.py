"""World-size-2 `gloo` runs of the data-parallel host logic (one process per rank, 127.0.0.1 rendezvous):
the gradient all-reduce used on the flat arenas, the criterion's single 2-float normaliser all-reduce
(dfine_criterion.py:639-652) and the reference-style DDP wrap on the CPU oracle provider."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from custom_d_fine_b200 import dist as du
    from custom_d_fine_b200 import kernels
    from custom_d_fine_b200.model import build_loss, build_model, build_optimizer
    from custom_d_fine_b200.train import ModelEMA, TrainStep
    from oracle.torch_ops import OracleOps
    from tests.golden.common import seeded_fill, synthetic_batch
    du.init_distributed_mode()
    try:
        assert du.get_world_size() == world and du.get_rank() == rank
        # (1) flat-arena gradient averaging
        g = torch.full((8,), float(rank + 1))
        du.allreduce_mean_(g)
        assert torch.allclose(g, torch.full((8,), (1 + world) / 2))
        assert du.broadcast_scalar(3.0 + rank) == 3.0
        # (2) one train step of D-FINE-n per rank on different data, reference-style DDP wrap
        torch.manual_seed(0)
        model = build_model("n", 80, False, "cpu", img_size=(320, 320))
        seeded_fill(model, 0)
        model.train()
        sizes = (3, 5) if rank == 0 else (0, 2)            # 10 targets over 2 ranks -> num_boxes = 5
        x, targets = synthetic_batch(2, 320, 320, seed=100 + rank, T=sizes)
        net = du.wrap_ddp(model)
        crit = build_loss("n", 80, 0.0, False)
        opt = build_optimizer(model, lr=1e-4, backbone_lr=1e-5, betas=(0.9, 0.999), weight_decay=1e-4, base_lr=1e-4)
        step = TrainStep(net, crit, opt, ema=ModelEMA(model, 0.9998), clip_max_norm=0.1)
        with kernels.use(OracleOps()):
            loss, _ = step(x, targets)
        assert torch.isfinite(loss)
        assert abs(float(crit.last_plan.counts[1]) - 5.0) < 1e-6, crit.last_plan.counts   # world-averaged num_boxes
        chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        assert torch.equal(both[0], both[1]), "ranks diverged after the averaged-gradient step"
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        du.cleanup_distributed()


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
