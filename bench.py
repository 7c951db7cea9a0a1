"""Benchmark of the D-FINE train-step hot path (BASELINE.json: images/s at 640x640, D-FINE-m, batch 16 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = model forward (with CDN), criterion (on-device Hungarian matcher), backward, gradient clip,
AdamW, EMA — the body of the reference's hot loop (train.py:550-586, 512-535) — on one synthetic batch
(torch.rand images, 10 boxes per image, 80 classes).  Prints ONE JSON line on rank 0.

  value        images/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e          images/s through the public step API with HOST inputs: pinned-memory H2D of images and
               targets every step and a D2H read of the loss inside the timed region
  roofline     the kernel family with the largest share of the step (the tcgen05 conv / linear kernel), summed
               algorithmic bytes / summed launch durations (CUDA events on the launching stream) vs the measured HBM
               peak; the weight-gradient kernel and MSDeformableAttention fwd / bwd are listed under "others"
  cpu_baseline the oracle port (this repo's host graph driven by oracle/torch_ops.py — the reference's
               arithmetic restated on torch CPU ops) on the box's host cores, bounded sample (N=1, rank 0)

--impl reference times that same CPU port with all host threads (the reference is pure Python with
unshipped dependencies and no installable package) on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NUM_CLASSES, T_PER_IMG = 80, 10
# BASELINE.json configs that fit one GPU (the per-GPU half of the data-parallel ones): name -> (size, img, batch/GPU, masks)
WORKLOADS = {
    "m640": ("m", 640, 16, False, 4),       # configs[1] / [2]: the headline
    "lseg640": ("l", 640, 8, True, 2),      # configs[3]: D-FINE-l + mask head
    "x1280": ("x", 1280, 4, False, 1),      # configs[4]: per-GPU share of the 8-GPU run
}   # last entry: images per step of the bounded CPU-arm sample
MODEL, HW, SEG = "m", 640, False
DTYPES = {
    "hf3": "fp32 storage + fp32 accumulate; forward GEMMs on 3xFP16 split operands (22 significand bits, kind::f16 tensor cores), "
           "gradient GEMMs on tf32 operands",
    "tc3": "fp32 storage + fp32 accumulate; forward GEMMs on 3xTF32 split operands, gradient GEMMs on tf32 operands",
    "tc": "fp32 storage + fp32 accumulate; tf32 tensor-core operands",
    "simt": "fp32 (CUDA cores)",
}
CPU_SAMPLE_BATCH = 4


def select_workload(name, batch=None):
    global MODEL, HW, SEG, CPU_SAMPLE_BATCH
    MODEL, HW, b, SEG, CPU_SAMPLE_BATCH = WORKLOADS[name]
    return batch or b


def synthetic(batch, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, HW, HW, generator=g)
    labels = torch.randint(0, NUM_CLASSES, (batch, T_PER_IMG), generator=g)
    cxcy = torch.rand(batch, T_PER_IMG, 2, generator=g) * 0.6 + 0.2
    wh = torch.rand(batch, T_PER_IMG, 2, generator=g) * 0.25 + 0.05
    boxes = torch.cat([cxcy, wh], -1)
    if pin:
        x, labels, boxes = x.pin_memory(), labels.pin_memory(), boxes.pin_memory()
    return x, labels, boxes


def rect_masks(boxes, hw):
    """uint8 [B, T, hw, hw] filled GT rectangles (the segment config's targets)."""
    B, T, _ = boxes.shape
    ys = torch.arange(hw).view(1, 1, hw, 1)
    xs = torch.arange(hw).view(1, 1, 1, hw)
    x0 = ((boxes[..., 0] - boxes[..., 2] / 2) * hw).round().view(B, T, 1, 1)
    x1 = ((boxes[..., 0] + boxes[..., 2] / 2) * hw).round().view(B, T, 1, 1)
    y0 = ((boxes[..., 1] - boxes[..., 3] / 2) * hw).round().view(B, T, 1, 1)
    y1 = ((boxes[..., 1] + boxes[..., 3] / 2) * hw).round().view(B, T, 1, 1)
    return ((xs >= x0) & (xs < x1) & (ys >= y0) & (ys < y1)).to(torch.uint8)


def to_targets(labels, boxes, masks=None):
    out = [{"labels": labels[i], "boxes": boxes[i]} for i in range(labels.shape[0])]
    if masks is not None:
        for i, t in enumerate(out):
            t["masks"] = masks[i]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_step(device, world, local_rank, eager=False):
    from custom_d_fine_b200 import dist as dist_utils
    from custom_d_fine_b200.model import build_loss, build_model, build_optimizer
    from custom_d_fine_b200.train import GraphedTrainStep, ModelEMA, TrainStep
    torch.manual_seed(0)
    model = build_model(MODEL, NUM_CLASSES, SEG, device, img_size=(HW, HW))
    # non-zero heads so every loss term (incl. DDF, zero at fresh init) does real work
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2 and float(p.abs().max()) == 0.0:
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))
    model.train()
    ema = ModelEMA(model, 0.9998)
    net = model      # data-parallel ranks all-reduce the optimizer's flat gradient arenas (no DDP wrapper)
    loss_fn = build_loss(MODEL, NUM_CLASSES, 0.0, SEG)
    opt = build_optimizer(model, lr=1.5e-4, backbone_lr=2e-5, betas=(0.9, 0.999), weight_decay=1.25e-4, base_lr=1.5e-4)
    cls = TrainStep if eager else GraphedTrainStep
    step = cls(net, loss_fn, opt, scheduler=None, ema=ema, clip_max_norm=0.1)
    # an eager twin over the same model / optimizer: used only to bracket single kernels with CUDA events
    step.eager_twin = TrainStep(net, loss_fn, opt, scheduler=None, ema=ema, clip_max_norm=0.1)
    return step


def run_ours(args):
    from custom_d_fine_b200 import cuda_ops
    from custom_d_fine_b200 import dist as dist_utils
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist_utils.init_distributed_mode()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    B = args.batch
    step = build_step(device, world, local_rank)
    hx, hl, hb = synthetic(B, 1234 + rank, pin=True)
    hm = rect_masks(hb, HW).pin_memory() if SEG else None
    dx, dl, db = hx.to(device), hl.to(device), hb.to(device)
    dtargets = to_targets(dl, db, hm.to(device) if SEG else None)

    def sync_all():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        sync_all()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        sync_all()
        ms = torch.tensor([s.elapsed_time(e)], device=device)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        step(dx, dtargets)

    class HostBatches:
        """The public input path (src/dl/train.Trainer): pinned host batches through train.DevicePrefetcher — every
        step's images and targets are copied host->device inside the timed region, on a side stream that overlaps
        the previous step."""

        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __iter__(self):
            for _ in range(self.n):
                yield hx, to_targets(hl, hb, hm), None

    e2e_walls = []

    def run_e2e(n):
        from custom_d_fine_b200.train import DevicePrefetcher
        last = None
        t_prev = time.perf_counter()
        for x, tg, _ in DevicePrefetcher(HostBatches(n), device):
            loss, _ = step(x, tg)
            last = float(loss.item())          # D2H read of the step's result, every step
            t_now = time.perf_counter()
            e2e_walls.append((t_now - t_prev) * 1e3)
            t_prev = t_now
        return last

    def h2d_probe():
        """Pinned host -> device bandwidth of this box for the step's image batch (context for the e2e number)."""
        dst = torch.empty_like(dx)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            dst.copy_(hx, non_blocking=True)
            e.record()
            torch.cuda.synchronize()
            best = max(best, hx.numel() * 4 / (s.elapsed_time(e) * 1e-3) / 1e9)
        return best

    graphed = hasattr(step, "_graphs")
    n_warm = max(args.warmup, 3) + (step.eager_steps + 1 if graphed else 0)   # + eager steps and the capture step
    for _ in range(n_warm):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = cuda_ops.counters.launches
    ms_total = timed(step_resident, args.steps)
    launches = cuda_ops.counters.launches - l0
    # single-kernel durations: the same K steps issued eagerly with CUDA events around the launches of the
    # kernel families that dominate the step (events on the launching stream; durations and algorithmic bytes
    # are summed per family: the conv kernel runs ~300 launches of 49 shapes per step)
    cuda_ops.counters.watch = ("conv_tc", "conv_wgrad_tc", "msda_fwd", "msda_bwd", "bn_apply", "bn_bwd_reduce", "bn_bwd_apply")
    cuda_ops.counters.timed = {}
    cuda_ops.wgrad_stream.disabled = True      # one stream: a kernel's event pair must not time a concurrent kernel too
    timed(lambda: step.eager_twin(dx, dtargets), args.steps)
    cuda_ops.wgrad_stream.disabled = False
    cuda_ops.counters.watch = ()
    kern, split = {}, {}
    pk0 = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    # tf32 tensor peak = half the measured dense bf16 rate (kind::tf32 issues at half the kind::f16 rate);
    # ridge point of a launch, in FLOP per algorithmic fp32 byte
    tf32_peak = float(pk0.get("bf16_tflops_sustained", 1394.5)) / 2.0
    tf32_measured = False
    tfile = ROOT / "profiles" / "r2_tf32_peak.json"
    if tfile.exists():       # cuBLAS tf32 8192^3 measured on this pool's B200 (tools/measure_tf32_peak.py)
        try:
            tf32_peak = float(json.loads(tfile.read_text())["tf32_tflops_sustained"])
            tf32_measured = True
        except (KeyError, ValueError):
            pass
    ridge = tf32_peak * 1e12 / (float(pk0.get("hbm_gbs", 6650.0)) * 1e9)
    if args.dump_launches:       # per-shape table of the timed families (tools: where the conv time goes)
        rows = {}
        for name, recs in cuda_ops.counters.timed.items():
            for r in recs:
                a = rows.setdefault((name, r[4]), [0, 0.0, 0, 0])
                a[0] += 1; a[1] += r[0].elapsed_time(r[1]); a[2] += r[2]; a[3] += r[3]
        with open(args.dump_launches, "w") as f:
            f.write("| family | shape | launches/step | us/launch | ms/step | GB/s | TFLOP/s |\n|---|---|---:|---:|---:|---:|---:|\n")
            for (name, tag), (n, ms, nb, fl) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| {name} | {tag} | {n / args.steps:.1f} | {ms / n * 1e3:.1f} | {ms / args.steps:.3f} | "
                        f"{nb / ms / 1e6:.0f} | {fl / ms / 1e9:.1f} |\n")
    for name, recs in cuda_ops.counters.timed.items():
        durs = [r[0].elapsed_time(r[1]) for r in recs]
        kern[name] = (sum(durs) / len(durs), sum(r[2] for r in recs) / len(recs), len(durs), sum(durs) / args.steps)
        if name in ("conv_tc", "conv_wgrad_tc"):
            # per-launch classification against the ridge: which roofline bounds the launch
            l_hbm = [(d, r[2]) for d, r in zip(durs, recs) if r[3] / max(r[2], 1) < ridge]
            l_tc = [(d, r[3]) for d, r in zip(durs, recs) if r[3] / max(r[2], 1) >= ridge]
            split[name] = {
                "ridge_flop_per_byte": round(ridge, 1),
                "hbm_bound": {"launches_per_step": len(l_hbm) // args.steps, "ms_per_step": round(sum(d for d, _ in l_hbm) / args.steps, 3),
                              "GB/s": round(sum(b for _, b in l_hbm) / max(sum(d for d, _ in l_hbm), 1e-9) / 1e6, 1)} if l_hbm else None,
                "tensor_bound": {"launches_per_step": len(l_tc) // args.steps, "ms_per_step": round(sum(d for d, _ in l_tc) / args.steps, 3),
                                 "TFLOP/s": round(sum(f for _, f in l_tc) / max(sum(d for d, _ in l_tc), 1e-9) / 1e9, 1),
                                 "tf32_peak_TFLOP/s": tf32_peak} if l_tc else None}
            if split[name]["hbm_bound"]:
                split[name]["hbm_bound"]["frac"] = round(split[name]["hbm_bound"]["GB/s"] / float(pk0.get("hbm_gbs", 6650.0)), 4)
            if split[name]["tensor_bound"]:
                split[name]["tensor_bound"]["frac"] = round(split[name]["tensor_bound"]["TFLOP/s"] / tf32_peak, 4)
    run_e2e(2)
    e2e_walls.clear()
    ms_e2e = timed(lambda: run_e2e(args.steps), 1)
    h2d_gbs = h2d_probe()
    mode = "cuda-graph replay (3 graphs/step)" if graphed and step._graphs else "eager launches"
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    roof = None
    if kern:
        def entry(k):
            ms, nbytes, n, per_step = kern[k]
            ach = nbytes / (ms * 1e-3) / 1e9
            return {"avg_launch_ms": round(ms, 4), "GB/s": round(ach, 1), "frac": round(ach / hbm_peak, 4),
                    "launches_per_step": n // args.steps, "ms_per_step": round(per_step, 3),
                    "algorithmic_bytes_per_launch": int(nbytes)}
        name = max(kern, key=lambda k: kern[k][3])        # the family with the largest share of the step
        e = entry(name)
        # DRAM traffic per launch of that family from the committed ncu pass (tools/gpu_trip_final.sh ncu_traffic)
        traffic, tr_file = None, ROOT / "profiles" / "r2_traffic.json"
        if not tr_file.exists():
            tr_file = ROOT / "profiles" / "r1_traffic.json"
        if tr_file.exists() and args.config == "m640":
            tr = json.loads(tr_file.read_text()).get(name)
            traffic = int(tr["dram_bytes_per_launch"]) if tr else None
        roof = {"kernel": name, "bound": "hbm", "achieved": e["GB/s"], "peak": hbm_peak, "unit": "GB/s",
                "frac": e["frac"], "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": e["avg_launch_ms"],
                "launches_per_step": e["launches_per_step"], "ms_per_step": e["ms_per_step"],
                "algorithmic_bytes_per_launch": e["algorithmic_bytes_per_launch"],
                "definition": "sum of algorithmic bytes (in + out + weights, fp32) over the family's launches / sum of "
                              "their CUDA-event durations",
                "timed_in": "an eager pass of the same K steps, one stream, CUDA events on the launching stream; a ~40 us "
                            "spin kernel queued before each start event keeps host launch latency out of the interval",
                "by_bound": split.get(name),
                "others": {k: dict(entry(k), **({"by_bound": split[k]} if k in split else {})) for k in kern if k != name}}
    cpu = cpu_baseline(steps=2, warmup=1) if world == 1 and not args.no_cpu_baseline else None
    tf32_src = "measured cuBLAS tf32 8192^3 (profiles/r2_tf32_peak.json)" if tf32_measured else "assumed = measured bf16 sustained / 2"
    if roof is not None:
        roof["tf32_peak_source"] = tf32_src
    imgs = B * world * args.steps
    line = {
        "metric": metric_name(), "value": round(imgs / (ms_total * 1e-3), 2),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPES.get(cuda_ops.get_gemm_mode(), cuda_ops.get_gemm_mode()), "data": "synthetic",
        "config": {"workload": workload_text(B), "name": args.config, "weights": "seeded default init with randomised zero-"
                   "initialised heads (the COCO checkpoint of SURVEY 8d is used by the parity tests; perf-neutral)",
                   "global_batch": B * world, "parallelism": f"dp{world}", "launch_mode": mode, "gemm_mode": cuda_ops.get_gemm_mode(),
                   "l2": "per-step working set (activations + grads > 10 GB) exceeds the 126 MB L2"},
        "e2e": {"value": round(imgs / (ms_e2e * 1e-3), 2), "unit": "images/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": int(hx.numel() * 4 + hl.numel() * 8 + hb.numel() * 4 + (hm.numel() if SEG else 0)),
                "d2h_bytes_per_step": 4, "h2d_pinned_GB/s": round(h2d_gbs, 1),
                "step_wall_ms": [round(v, 1) for v in e2e_walls]},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
    }
    if graphed and step.host_gap_ms() is not None:
        line["config"]["host_gap_ms"] = round(step.host_gap_ms(), 3)   # device idle between graphs A and B (index planning)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def metric_name():
    return f"images/sec ({HW}x{HW}) D-FINE-{MODEL}{'-seg' if SEG else ''} train step"


def workload_text(B):
    return (f"D-FINE-{MODEL} {'segment' if SEG else 'detect'} train step (fwd + criterion + bwd + clip + AdamW + EMA), "
            f"batch {B}/GPU, {HW}x{HW}, {T_PER_IMG} boxes/img{' + filled-rectangle masks' if SEG else ''}, COCO-80 classes, Lq=500")


# ------------------------------------------------------------------------------------------------ CPU arms
REF_DIR = ROOT / "baseline" / "_ref"


def _reference_available():
    return (REF_DIR / "MANIFEST.json").exists() and (REF_DIR / "src" / "d_fine" / "dfine.py").exists()


def reference_step_fn(batch):
    """The UNMODIFIED reference (`baseline/_ref/src/d_fine`, copied byte for byte by tools/install_reference.py; see its
    MANIFEST.json) on CPU: its own build_model / build_loss / build_optimizer and the body of its hot loop
    (train.py:571-581 forward, criterion, backward; 512-535 clip_grad_norm_ + AdamW.step + zero_grad) in fp32
    (`amp_enabled=False`).  The reference's ModelEMA lives in src/dl/train.py, which cannot be imported here (hydra,
    albumentations, torchmetrics are absent): the EMA update is left out of the CPU arm, in the reference's favour."""
    # the reference's package is called `src`, like this repo's drop-in shims: make its tree the only `src` visible
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    keep = [q for q in sys.path if q not in ("", str(ROOT), str(ROOT) + "/")]
    sys.path[:] = [str(REF_DIR)] + keep
    try:
        import logging
        logging.disable(logging.WARNING)
        try:
            from loguru import logger
            logger.remove()
        except Exception:  # noqa: BLE001
            pass
        from src.d_fine.dfine import build_loss, build_model, build_optimizer
    finally:
        sys.path[:] = [str(ROOT)] + [q for q in sys.path if q != str(REF_DIR)]
    torch.manual_seed(0)
    model = build_model(MODEL, NUM_CLASSES, SEG, "cpu", img_size=(HW, HW))
    model.train()
    loss_fn = build_loss(MODEL, NUM_CLASSES, 0.0, SEG)
    opt = build_optimizer(model, lr=1.5e-4, backbone_lr=2e-5, betas=(0.9, 0.999), weight_decay=1.25e-4, base_lr=1.5e-4)
    x, l, b = synthetic(batch, 1234)
    targets = to_targets(l, b, rect_masks(b, HW) if SEG else None)
    for t in targets:
        t["orig_size"] = torch.tensor([HW, HW])

    def step():
        out = model(x, targets=targets)
        loss = sum(loss_fn(out, targets).values())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
        opt.step()
        opt.zero_grad()
        return float(loss.detach())

    return step


def port_step_fn(batch):
    """Fallback when baseline/_ref is absent: this repo's host graph with the oracle provider (the reference's
    arithmetic restated on torch CPU ops)."""
    from custom_d_fine_b200 import kernels
    from custom_d_fine_b200.model import build_loss, build_model
    from oracle.torch_ops import OracleOps
    torch.manual_seed(0)
    model = build_model(MODEL, NUM_CLASSES, SEG, "cpu", img_size=(HW, HW))
    model.train()
    loss_fn = build_loss(MODEL, NUM_CLASSES, 0.0, SEG)
    x, l, b = synthetic(batch, 1234)
    targets = to_targets(l, b, rect_masks(b, HW) if SEG else None)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
    ops = OracleOps()

    def step():
        with kernels.use(ops):
            out = model(x, targets=targets)
            loss = sum(loss_fn(out, targets).values())
            loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
        opt.step()
        opt.zero_grad()
        return float(loss.detach())

    return step


def cpu_arm(steps, warmup, batch):
    """(images/s, cores, kind, sample text) of the CPU arm: the unmodified reference when baseline/_ref travelled."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "reference" if _reference_available() else "port"
    step = (reference_step_fn if kind == "reference" else port_step_fn)(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    what = ("the unmodified reference src/d_fine (baseline/_ref, fp32, no EMA)" if kind == "reference"
            else "the oracle port (this repo's host graph over oracle/torch_ops.py)")
    sample = (f"{steps} train steps of batch {batch} (bounded sample of the workload: D-FINE-{MODEL}{'-seg' if SEG else ''}, "
              f"{HW}x{HW}, fwd + criterion + bwd + clip + AdamW) after {warmup} warm-up, {what}, torch CPU, {cores} threads")
    return batch * steps / dt, dt / steps, cores, kind, sample


def cpu_baseline(steps, warmup):
    v, _, cores, kind, sample = cpu_arm(steps, warmup, CPU_SAMPLE_BATCH)
    return {"value": round(v, 3), "unit": "images/s", "cores": cores, "kind": kind, "sample": sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, s_per_step, cores, kind, sample = cpu_arm(args.steps, args.warmup, CPU_SAMPLE_BATCH)
    v = round(v, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    print(json.dumps({
        "impl": "reference", "metric": metric_name(), "value": v, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(s_per_step * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": workload_text(args.batch) + f" (CPU arm: each step = {CPU_SAMPLE_BATCH} images of it)",
                   "name": args.config, "parallelism": "cpu", "rank0_only": True},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="m640", choices=sorted(WORKLOADS),
                    help="BASELINE.json workload: m640 (headline, configs 1/2), lseg640 (config 3), x1280 (config 4)")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-launches", default=None, help="write the per-shape table of the timed kernel families here")
    args = ap.parse_args()
    args.batch = select_workload(args.config, args.batch)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
