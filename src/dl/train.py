"""Drop-in for the training entry point /root/reference/src/dl/train.py: ``ModelEMA`` (52-73), ``Trainer``
(77-658: construction order seed -> loader -> model -> EMA -> loss -> optimizer -> OneCycleLR; ``train()`` hot loop
537-608; ``save_model`` 476-503) and ``main`` (661-748).

Scope: the train step.  The reference's dataset / augmentation pipeline, validator, wandb and
visualisation are out of scope; ``Trainer`` therefore takes any iterable of ``(images float32 [B,3,H,W] in [0,1],
targets list[dict(labels int64 [T], boxes float32 [T,4] cxcywh)], paths)`` batches — the reference loader's
collate format (dataset.py:651-656) — and falls back to a seeded synthetic loader of the same format.
hydra / OmegaConf are not required: ``cfg`` may be a nested dict, a namespace, or an OmegaConf object; ``main``
reads ``config.yaml`` with PyYAML and applies ``key=value`` overrides like the reference CLI.

    python -m src.dl.train model_name=m train.batch_size=16 train.epochs=1
    torchrun --nproc_per_node=8 -m src.dl.train train.ddp.enabled=True
"""
from __future__ import annotations

import sys
import time
from pathlib import Path
from types import SimpleNamespace

import torch
from torch.optim.lr_scheduler import OneCycleLR

from custom_d_fine_b200 import dist as dist_utils
from custom_d_fine_b200.model import build_loss, build_model, build_optimizer
from custom_d_fine_b200.train import DevicePrefetcher, GraphedTrainStep, ModelEMA, TrainStep  # noqa: F401

DEFAULTS = {
    "model_name": "m", "task": "detect", "exp": "b200",
    "train": {
        "device": "cuda", "seed": 42, "epochs": 1, "batch_size": 16, "img_size": [640, 640], "b_accum_steps": 1,
        "clip_max_norm": 0.1, "use_ema": True, "ema_momentum": 0.9998, "label_smoothing": 0.0,
        "base_lr": 1.5e-4, "backbone_lr": 2e-5, "betas": [0.9, 0.999], "weight_decay": 1.25e-4,
        "use_scheduler": True, "cycler_pct_start": 0.1, "amp_enabled": False, "pretrained_model_path": None,
        "label_to_name": {i: str(i) for i in range(80)}, "path_to_save": "output/b200", "ddp": {"enabled": False},
        "synthetic_steps_per_epoch": 50, "targets_per_image": 10, "cuda_graphs": True, "resume_from": None,
    },
}


def _ns(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: (_ns(v) if isinstance(v, dict) and k != "label_to_name" else v) for k, v in d.items()})
    return d


def _merge(base, over):
    out = dict(base)
    for k, v in (over or {}).items():
        out[k] = _merge(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else v
    return out


def load_cfg(path=None, overrides=()):
    """config.yaml (+ ``a.b=c`` overrides) -> namespace with the reference's field names (config.yaml:1-162).
    Only the keys the train step reads are interpreted; ``${...}`` interpolations of unrelated keys are left as text."""
    import yaml
    cfg = dict(DEFAULTS)
    if path and Path(path).exists():
        cfg = _merge(cfg, yaml.safe_load(Path(path).read_text()) or {})
    for ov in overrides:
        key, _, val = ov.partition("=")
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = yaml.safe_load(val)
    name = cfg["model_name"]
    tr = cfg["train"]
    for k in ("base_lr", "backbone_lr"):           # config.yaml:77-78 resolves these through train.lrs.<model_name>
        if isinstance(tr.get(k), str) and "lrs" in tr:
            tr[k] = tr["lrs"][name][k]
    return _ns(cfg)


class SyntheticLoader:
    """Batches in the reference collate format (dataset.py:651-656) from a seeded generator."""

    def __init__(self, batch_size, img_size, steps, targets_per_image=10, num_classes=80, seed=1234):
        self.bs, self.hw, self.steps, self.t, self.nc, self.seed = batch_size, tuple(img_size), steps, targets_per_image, num_classes, seed

    def __len__(self):
        return self.steps

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.steps):
            x = torch.rand(self.bs, 3, *self.hw, generator=g)
            targets = []
            for _b in range(self.bs):
                cxcy = torch.rand(self.t, 2, generator=g) * 0.6 + 0.2
                wh = torch.rand(self.t, 2, generator=g) * 0.25 + 0.05
                targets.append({"labels": torch.randint(0, self.nc, (self.t,), generator=g),
                                "boxes": torch.cat([cxcy, wh], 1)})
            yield x, targets, None


class Trainer:
    def __init__(self, cfg, train_loader=None):
        if isinstance(cfg, dict):
            cfg = _ns(_merge(DEFAULTS, cfg))
        self.cfg = cfg
        tr = cfg.train
        self.distributed = bool(getattr(getattr(tr, "ddp", None), "enabled", False)) and \
            dist_utils.is_dist_available_and_initialized()
        self.rank, self.world_size = dist_utils.get_rank(), dist_utils.get_world_size()
        self.is_main = self.rank == 0
        if self.distributed and torch.cuda.is_available():
            self.local_rank = dist_utils.get_local_rank()
            self.device = torch.device("cuda", self.local_rank)
        else:
            self.local_rank = 0
            self.device = torch.device(tr.device)
        if getattr(tr, "amp_enabled", False):
            raise NotImplementedError("the B200 path computes in tf32/fp32 (the reference's amp_enabled=False branch, "
                                      "train.py:577-581); fp16 autocast + GradScaler is not part of it")
        self.epochs = tr.epochs
        self.clip_max_norm = tr.clip_max_norm
        self.b_accum_steps = max(tr.b_accum_steps, 1)
        self.num_labels = len(tr.label_to_name)
        self.task = cfg.task
        self.path_to_save = Path(tr.path_to_save)
        enable_mask_head = self.task == "segment"

        seed = tr.seed + self.rank if self.distributed else tr.seed
        torch.manual_seed(seed)
        if train_loader is None:
            train_loader = SyntheticLoader(tr.batch_size, tr.img_size, tr.synthetic_steps_per_epoch,
                                           tr.targets_per_image, self.num_labels, seed=1234 + self.rank)
        self.train_loader = train_loader

        self.model = build_model(cfg.model_name, self.num_labels, enable_mask_head, str(self.device),
                                 img_size=tuple(tr.img_size), pretrained_model_path=tr.pretrained_model_path)
        self.ema_model = ModelEMA(self.model, tr.ema_momentum) if tr.use_ema else None
        self.loss_fn = build_loss(cfg.model_name, self.num_labels, label_smoothing=tr.label_smoothing,
                                  enable_mask_head=enable_mask_head)
        self.optimizer = build_optimizer(self.model, lr=tr.base_lr, backbone_lr=tr.backbone_lr, betas=tuple(tr.betas),
                                         weight_decay=tr.weight_decay, base_lr=tr.base_lr)
        self.scheduler = None
        if tr.use_scheduler:
            max_lr = tr.base_lr * 2
            if cfg.model_name in ["l", "x"] or enable_mask_head:
                max_lr = [tr.backbone_lr * 2, tr.backbone_lr * 2, tr.base_lr * 2, tr.base_lr * 2]
            self.scheduler = OneCycleLR(self.optimizer, max_lr=max_lr, epochs=tr.epochs,
                                        steps_per_epoch=max(len(self.train_loader) // self.b_accum_steps, 1),
                                        pct_start=tr.cycler_pct_start, cycle_momentum=False)
        # gradients of data-parallel ranks are averaged on the optimizer's flat arenas (no DDP wrapper needed)
        cls = GraphedTrainStep if getattr(tr, "cuda_graphs", True) else TrainStep
        self.step = cls(self.model, self.loss_fn, self.optimizer, scheduler=self.scheduler, ema=self.ema_model,
                        clip_max_norm=self.clip_max_norm, accum_steps=self.b_accum_steps)

    def optimizer_step(self, step_scheduler: bool = True):
        self.step.optimizer_step(step_scheduler)

    def save_model(self, name="last.pt"):
        """Bare state_dict of the EMA weights if enabled, else the model's (train.py:476-503)."""
        if not self.is_main:
            return
        self.path_to_save.mkdir(parents=True, exist_ok=True)
        m = self.ema_model.model if self.ema_model is not None else self.model
        torch.save({k: v.detach().cpu().clone() for k, v in m.state_dict().items()}, self.path_to_save / name)

    # ---- resume -----------------------------------------------------------------------
    # The reference only ever writes bare state dicts (`last.pt`, `model.pt`, train.py:476-503) and cannot resume.  Those
    # files keep their format (Torch_model / export load them); next to them `resume.pt` holds everything a restart
    # needs: raw model weights, the EMA weights and its step counter, the optimizer in torch.optim.AdamW's own layout
    # (optim.FusedAdamW.state_dict), the scheduler, the epoch and the generators' states.
    def save_checkpoint(self, epoch, name="resume.pt"):
        if not self.is_main:
            return
        self.path_to_save.mkdir(parents=True, exist_ok=True)
        cpu = lambda sd: {k: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for k, v in sd.items()}  # noqa: E731
        ck = {"format": "custom_d_fine_b200.resume.v1", "epoch": epoch, "model": cpu(self.model.state_dict()),
              "ema": cpu(self.ema_model.model.state_dict()) if self.ema_model is not None else None,
              "ema_iter": self.step.ema_iter, "batch_idx": self.step.batch_idx,
              "optimizer": self.optimizer.state_dict(),
              "scheduler": self.scheduler.state_dict() if self.scheduler is not None else None,
              "rng": {"cpu": torch.get_rng_state(),
                      "cuda": torch.cuda.get_rng_state(self.device) if self.device.type == "cuda" else None}}
        tmp = self.path_to_save / (name + ".tmp")
        torch.save(ck, tmp)
        tmp.replace(self.path_to_save / name)         # atomic: a crash while writing never corrupts the last checkpoint

    def resume(self, path):
        """Restore a `save_checkpoint` file; returns the epoch to continue from (the saved one + 1)."""
        ck = torch.load(path, map_location="cpu", weights_only=False)
        assert ck.get("format") == "custom_d_fine_b200.resume.v1", "not a resume checkpoint"
        self.model.load_state_dict(ck["model"])
        if self.ema_model is not None and ck["ema"] is not None:
            self.ema_model.model.load_state_dict(ck["ema"])
        self.optimizer.load_state_dict(ck["optimizer"])
        if self.scheduler is not None and ck["scheduler"] is not None:
            self.scheduler.load_state_dict(ck["scheduler"])
        self.step.ema_iter, self.step.batch_idx = ck["ema_iter"], ck["batch_idx"]
        torch.set_rng_state(ck["rng"]["cpu"])
        if self.device.type == "cuda" and ck["rng"]["cuda"] is not None:
            torch.cuda.set_rng_state(ck["rng"]["cuda"], self.device)
        if self.device.type == "cuda":
            from custom_d_fine_b200 import cuda_ops
            cuda_ops.weights_changed()            # re-laid weight copies / operand planes follow the loaded parameters
            if hasattr(self.optimizer, "_register_planes"):
                self.optimizer._register_planes()
        self.start_epoch = int(ck["epoch"]) + 1
        return self.start_epoch

    def train(self):
        history = []
        for epoch in range(getattr(self, "start_epoch", 1), self.epochs + 1):
            self.model.train()
            t0, n_img, losses = time.perf_counter(), 0, []
            # host->device copies of batch k+1 run on a side stream while batch k trains (train.py:558-565)
            for inputs, targets, _ in DevicePrefetcher(self.train_loader, self.device, self.num_labels):
                loss, _ = self.step(inputs, targets)
                losses.append(loss)
                n_img += inputs.shape[0]
            mean_loss = float(torch.stack(losses).mean()) if losses else float("nan")     # one sync per epoch
            dt = time.perf_counter() - t0
            history.append({"epoch": epoch, "loss": mean_loss, "images_per_s": n_img * self.world_size / dt})
            if self.is_main:
                print(f"epoch {epoch}: loss {mean_loss:.4f}, {history[-1]['images_per_s']:.1f} img/s")
            self.save_model("last.pt")
            self.save_checkpoint(epoch)
        return history


def main(cfg=None):
    if cfg is None:
        args = [a for a in sys.argv[1:] if "=" in a]
        cfg = load_cfg(Path(__file__).resolve().parents[2] / "config.yaml", args)
    ddp = bool(getattr(getattr(cfg.train, "ddp", None), "enabled", False))
    if ddp:
        dist_utils.init_distributed_mode()
    try:
        trainer = Trainer(cfg)
        resume = getattr(cfg.train, "resume_from", None)
        if resume:
            trainer.resume(resume)
        return trainer.train()
    finally:
        if ddp:
            dist_utils.cleanup_distributed()


if __name__ == "__main__":
    main()
