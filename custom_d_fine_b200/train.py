"""Train-step mirror of /root/reference/src/dl/train.py: ``ModelEMA`` (52-73) and the body of
``Trainer.train`` for one batch (550-586) + ``optimizer_step`` (512-535), fp32 path
(``amp_enabled=False``, train.py:577-581).  Data loading, evaluation, logging and checkpoint
writing are out of scope; ``src/dl/train.py`` wires this into the
reference's ``Trainer`` / ``main`` names.
"""
from __future__ import annotations

import math
import os
from copy import deepcopy

import torch
from torch.nn.parallel import DistributedDataParallel as DDP


def _begin_step(inputs, counters=None):
    """One memset for every per-layer fp64 reduction buffer of the step (cuda_ops.zero_pool), one add for every
    BatchNorm step counter."""
    if inputs.is_cuda:
        from . import cuda_ops
        cuda_ops.zero_pool.begin_step(inputs.device)
        cuda_ops.wgrad_stream.begin(inputs.device)
        if counters is not None:
            counters.add_(1)


def _after_forward():
    """The weight-gradient stream prepared the backward's re-laid weights during the forward: rejoin it."""
    from . import cuda_ops
    cuda_ops.wgrad_stream.sync_main()


def _end_step():
    from . import cuda_ops
    cuda_ops.zero_pool.end_step()
    cuda_ops.wgrad_stream.join()     # the weight-gradient stream rejoins the step before the optimizer reads the arenas


class DevicePrefetcher:
    """Wraps a loader of host batches ``(images, targets, extra)`` (the reference's collate format,
    dataset.py:651-656) and yields device batches whose pinned host->device copies were issued on a side stream
    while the previous step was computing — what the reference gets from ``pin_memory`` + ``non_blocking`` copies
    (train.py:558-565, dataset.py:570-580) only if the copy does not serialise with the step.  One batch ahead.

    The copies land in TWO persistent staging slots per device (allocated once per tensor shape, shared by every
    prefetcher of the process): no per-step device allocation, no allocator stalls on the first batch of an epoch.  A
    slot is rewritten only after the step that consumed it has been enqueued AND has finished on the device (an event
    recorded on the consumer's stream when it asks for the next batch).  Labels outside [0, num_classes) are rejected
    here, where they are still host tensors (the device matcher clamps, it cannot raise)."""

    _pool = {}      # device index -> {"stream", "slots": [ {key: tensor} x 2 ], "done": [event | None] x 2}

    def __init__(self, loader, device, num_classes=None):
        self.loader, self.device, self.num_classes = loader, torch.device(device), num_classes
        self.state = None
        if self.device.type == "cuda":
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            st = DevicePrefetcher._pool.get(idx)
            if st is None:
                st = DevicePrefetcher._pool[idx] = {"stream": torch.cuda.Stream(self.device), "slots": [{}, {}],
                                                    "done": [None, None]}
            self.state = st

    def __len__(self):
        return len(self.loader)

    def _check(self, targets):
        if self.num_classes is None:
            return
        for t in targets:
            lab = t.get("labels")
            if lab is not None and not lab.is_cuda and lab.numel() and (int(lab.min()) < 0 or int(lab.max()) >= self.num_classes):
                raise IndexError(f"target label outside [0, {self.num_classes}) (the reference raises at matcher.py:150)")

    def _into(self, slot, key, host):
        buf = slot.get(key)
        if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
            buf = slot[key] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
        buf.copy_(host, non_blocking=True)
        return buf

    def _stage(self, batch, k):
        x, targets, extra = batch
        self._check(targets)
        if self.state is None:
            return x.to(self.device), [{n: v.to(self.device) for n, v in t.items() if torch.is_tensor(v)} for t in targets], extra, None
        st = self.state
        slot, stream = st["slots"][k % 2], st["stream"]
        with torch.cuda.stream(stream):
            if st["done"][k % 2] is not None:
                stream.wait_event(st["done"][k % 2])          # the step that read this slot has finished
            xd = self._into(slot, "x", x)
            td = [{n: self._into(slot, (i, n), v) for n, v in t.items() if torch.is_tensor(v)} for i, t in enumerate(targets)]
            ev = torch.cuda.Event()
            ev.record(stream)
        return xd, td, extra, ev

    def __iter__(self):
        it = iter(self.loader)
        try:
            k = 0
            nxt = self._stage(next(it), k)
        except StopIteration:
            return
        while nxt is not None:
            x, targets, extra, ev = nxt
            cur = torch.cuda.current_stream(self.device) if ev is not None else None
            if ev is not None:
                cur.wait_event(ev)
            try:
                nxt = self._stage(next(it), k + 1)      # next batch's copies overlap this batch's step
            except StopIteration:
                nxt = None
            yield x, targets, extra
            if ev is not None:                          # the consumer has enqueued its step on `cur`: mark the slot's release
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(self.device))
                self.state["done"][k % 2] = done
            k += 1


class ModelEMA:
    """EMA of every floating-point state entry, momentum m*(1-exp(-it/2000)) (train.py:52-73).
    The reference walks ~920 tensors with two tiny kernels each; here the entries are updated with
    multi-tensor launches, and the momentum lives in a device scalar so the update can sit inside a
    replayed CUDA graph (``set_momentum`` outside the graph, ``apply`` inside)."""

    def __init__(self, student, ema_momentum):
        if isinstance(student, DDP):
            student = student.module
        self.model = deepcopy(student).eval()
        for p in self.model.parameters():
            p.requires_grad_(False)
        self.ema_momentum = ema_momentum
        self._pairs = None
        self._m = self._om = None

    def ema_scheduler(self, x):
        return self.ema_momentum * (1 - math.exp(-x / 2000))

    def _build_pairs(self, student):
        src = student.state_dict()
        ema, stu = [], []
        for name, p in self.model.state_dict().items():
            if p.dtype.is_floating_point:
                ema.append(p)
                stu.append(src[name].detach())
        self._pairs = (ema, stu)
        dev = ema[0].device
        self._m = torch.zeros((), device=dev)
        self._om = torch.ones((), device=dev)

    def set_momentum(self, iters, student):
        if isinstance(student, DDP):
            student = student.module
        if self._pairs is None:
            self._build_pairs(student)
        m = self.ema_scheduler(iters)
        self._m.fill_(m)
        self._om.fill_(1.0 - m)

    @torch.no_grad()
    def apply(self):
        ema, stu = self._pairs
        torch._foreach_mul_(ema, self._m)                      # p *= m
        torch._foreach_add_(ema, torch._foreach_mul(stu, self._om))   # p += (1-m) * student

    @torch.no_grad()
    def update(self, iters, student):
        self.set_momentum(iters, student)
        self.apply()


class TrainStep:
    """One optimisation step on one batch: forward, criterion, backward, (gradient all-reduce), clip, AdamW,
    scheduler, zero_grad, EMA — the reference's hot loop body (train.py:550-586 + 512-535), eager launches.

    With the flat-arena optimizer (optim.FusedAdamW, every CUDA run) clip + AdamW + EMA + zero_grad are three
    fused launches per group and data-parallel ranks all-reduce the flat gradient arenas; with a stock torch
    optimizer (CPU host-logic tests, or a DDP-wrapped model) the reference's call sequence is issued as is."""

    def __init__(self, model, loss_fn, optimizer, scheduler=None, ema=None, clip_max_norm=0.1, accum_steps=1):
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.scheduler, self.ema, self.clip_max_norm = scheduler, ema, clip_max_norm
        self.accum_steps, self.ema_iter, self.batch_idx = accum_steps, 0, 0
        self.fused = getattr(optimizer, "fused_step", False)
        self._counters = None
        m = model.module if isinstance(model, DDP) else model
        if next(m.parameters()).is_cuda:
            from . import cuda_ops
            self._counters = cuda_ops.register_step_counters(m)
        if self.fused:
            optimizer.max_norm = float(clip_max_norm or 0.0)
            if ema is not None:
                optimizer.attach_ema(ema, model)
            optimizer.broadcast_state(0)

    def _host_prepare(self):
        """Host half of optimizer_step (never captured): counters and the per-step scalars."""
        m = None
        if self.ema is not None:
            self.ema_iter += 1
            m = self.ema.ema_scheduler(self.ema_iter)
        if self.fused:
            self.optimizer.prepare(m)
        elif self.ema is not None:
            self.ema.set_momentum(self.ema_iter, self.model)

    def _apply_grads(self):
        """clip -> AdamW -> zero_grad -> EMA blend: the device part of optimizer_step (graph-capturable)."""
        if self.fused:
            self.optimizer.step()
            return
        if self.clip_max_norm:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip_max_norm)
        self.optimizer.step()
        self.optimizer.zero_grad()
        if self.ema is not None:
            self.ema.apply()

    def optimizer_step(self, step_scheduler=True):
        self._host_prepare()
        if self.fused:
            self.optimizer.allreduce_grads()
        self._apply_grads()
        if step_scheduler and self.scheduler is not None:
            self.scheduler.step()

    # ---- data-parallel overlap: the backward pass in two halves --------------------------------------------------
    # The optimizer's groups 2 / 3 (encoder + decoder) have their complete gradients once the backward pass reaches
    # the backbone's outputs; groups 0 / 1 (backbone) only at the very end.  With more than one rank the graph is cut at
    # the backbone outputs: half 1 = losses + decoder + encoder backward, then the all-reduce of arenas 2 / 3 is
    # issued on a communication stream and overlaps half 2 = the backbone's backward; arenas 0 / 1 follow.
    def _split(self):
        from . import dist as dist_utils
        import os
        return self.fused and dist_utils.get_world_size() > 1 and not isinstance(self.model, DDP) \
            and hasattr(self.model, "backbone") and os.environ.get("DFINE_SPLIT_BWD", "1") != "0"

    def _forward(self, inputs, targets):
        """(model output, cut): cut = None, or (backbone outputs, their detached leaves feeding the encoder)."""
        if not self._split():
            return self.model(inputs, targets=targets), None
        m = self.model
        feats = m.backbone(inputs.permute(0, 2, 3, 1))
        leaves = [f.detach().requires_grad_(True) for f in feats]
        return m.decoder(m.encoder(leaves), targets), (feats, leaves)

    def _comm_stream(self, device):
        if getattr(self, "_comm", None) is None:
            # (above the step's stream, which is above the weight-gradient stream: the all-reduce kernels get their few SMs
            #  as soon as they are launched instead of queueing behind compute CTAs — 2 GPUs, same box: 36.24 ms per step
            #  against 36.56 at the step stream's priority and 36.48 at the weight-gradient stream's)
            prio = -2 if os.environ.get("DFINE_STREAM_PRIO", "1") != "0" else 0
            self._comm = torch.cuda.Stream(device, priority=int(os.environ.get("DFINE_COMM_PRIO", prio)))
        return self._comm

    def _backward_half2(self, cut):
        feats, leaves = cut
        keep = [(f, l.grad) for f, l in zip(feats, leaves) if l.grad is not None]
        torch.autograd.backward([f for f, _ in keep], [g for _, g in keep])

    def __call__(self, inputs, targets):
        """inputs float32 [B,3,H,W] on the device; targets list of dicts (labels int64 [T], boxes [T,4])."""
        _begin_step(inputs, self._counters if self.model.training else None)
        output, cut = self._forward(inputs, targets)
        _after_forward()
        loss_dict = self.loss_fn(output, targets)
        loss = sum(loss_dict.values()) / self.accum_steps
        loss.backward()
        step_now = (self.batch_idx + 1) % self.accum_steps == 0
        if cut is not None:
            from . import cuda_ops
            bb, rest = self.optimizer.backbone_groups()
            cuda_ops.wgrad_stream.sync_main()           # encoder / decoder weight gradients are complete
            if step_now:
                self.optimizer.allreduce_grads(rest, self._comm_stream(inputs.device))
            self._backward_half2(cut)
        _end_step()
        self.batch_idx += 1
        if step_now:
            if cut is not None:
                bb, rest = self.optimizer.backbone_groups()
                self._host_prepare()
                self.optimizer.allreduce_grads(bb, self._comm_stream(inputs.device))
                torch.cuda.current_stream().wait_stream(self._comm)
                self._apply_grads()
                if self.scheduler is not None:
                    self.scheduler.step()
            else:
                self.optimizer_step()
        return loss.detach(), loss_dict


class GraphedTrainStep(TrainStep):
    """The same step replayed as CUDA graphs around the two host-side pieces of a step:

        graph A : model forward (+CDN) + the one-launch Hungarian matcher
        host    : D2H of the [n_layers, sumT] index table, GO union, normalisers, one pinned H2D
        graph B : every loss term + backward (gradients accumulate into the flat arenas)
        NCCL    : all-reduce of the flat gradient arenas (data-parallel runs only; issued between the graphs)
        graph C : gradient clip + AdamW + EMA + zero_grad (flat-arena optimizer)

    The reference's step is ~9 k kernel launches driven by Python; replaying it removes the host from the
    critical path.  Graphs are keyed by (input shape, targets' sizes); the first ``eager_steps`` calls of a
    key run eagerly (they are real training steps and double as the warm-up CUDA graphs require), then
    the key is captured.  Learning rate, weight decay, step count and EMA momentum are read by graph C from a
    device table refreshed before each replay, so an LR scheduler keeps working.  Constraint: no gradient
    accumulation (accum_steps == 1) and the flat-arena optimizer; anything else runs the eager step."""

    def __init__(self, *args, eager_steps=3, max_graphs=8, max_keys=1024, **kw):
        super().__init__(*args, **kw)
        from collections import OrderedDict
        self.eager_steps = eager_steps
        # Every per-key cache is bounded.  With a real loader the per-image target counts rarely repeat: such keys run
        # the eager step (a correct, slower path), are remembered in an LRU of `max_keys` counters, and only a key seen
        # more than `eager_steps` times is captured; at most `max_graphs` captured keys (each owns a private memory pool
        # for its forward + backward) are kept, least recently replayed evicted first.
        self.max_graphs, self.max_keys = max_graphs, max_keys
        self._seen = OrderedDict()
        self._graphs = OrderedDict()
        self._plans = {}
        self._static, self._static_seen = OrderedDict(), {}      # backbone + encoder graphs per input shape (hybrid steps)
        # Warm-up steps and captures share ONE side stream: autograd's AccumulateGrad nodes remember the stream
        # they were created on, and a node created on the legacy default stream cannot be used under capture.
        # ... at a HIGHER priority than the weight-gradient stream (kernel nodes keep their capture stream's priority):
        # when SMs free up, the critical chain of data gradients is scheduled ahead of the queued weight-gradient CTAs
        # instead of waiting behind them.  DFINE_STREAM_PRIO=0: equal priorities.
        prio = -1 if os.environ.get("DFINE_STREAM_PRIO", "1") != "0" else 0
        self._side = torch.cuda.Stream(priority=prio) if next(self.model.parameters()).is_cuda else None

    def host_gap_ms(self):
        """Device idle time between graph A and graph B of the last replayed step (host index planning)."""
        ev = getattr(self, "_last_gap", None)
        if ev is None:
            return None
        ev[1].synchronize()
        return ev[0].elapsed_time(ev[1])

    def _can_graph(self):
        return (self.accum_steps == 1 and self.fused and not isinstance(self.model, DDP)
                and next(self.model.parameters()).is_cuda)

    def __call__(self, inputs, targets):
        if not self._can_graph():
            return super().__call__(inputs, targets)
        key = (tuple(inputs.shape), tuple(int(t["labels"].shape[0]) for t in targets),
               tuple(tuple(t["masks"].shape) if t.get("masks") is not None else None for t in targets))
        g = self._graphs.get(key)
        if g is not None:
            self._graphs.move_to_end(key)
        else:
            n = self._seen.pop(key, 0)
            self._seen[key] = n + 1                        # (re-inserted at the recent end)
            while len(self._seen) > self.max_keys:
                old, _ = self._seen.popitem(last=False)
                self._plans.pop(old, None)
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                if n < self.eager_steps or key not in self._plans:     # (a plan evicted with its counter: one more eager step)
                    res = self._uncaptured_step(inputs, targets)
                    if n >= self.eager_steps - 1:          # the index plan of the LAST eager step is reused to capture
                        self._plans[key] = self.loss_fn.last_plan
                    eager = True
                else:
                    g = self._capture(inputs, targets, self._plans.pop(key))
                    self._graphs[key] = g
                    self._seen.pop(key, None)
                    while len(self._graphs) > self.max_graphs:      # drop the least recently replayed key and its pool
                        self._graphs.popitem(last=False)
                    eager = False
            cur.wait_stream(self._side)
            if eager:
                return res
        return self._replay(g, inputs, targets)

    # ---- steps whose key is not captured: the target-independent half of the step still replays ---------------------
    # With a real loader the per-image target counts (part of the graph key: the denoising groups, the matcher and the
    # index table take their shapes from them) rarely repeat, and a fully eager step is host-bound (~2250 launches from
    # Python: 134 ms against 35 ms replayed for D-FINE-m).  Backbone + encoder see only the images, so their forward and
    # their backward are captured ONCE per input shape (the autograd graph is cut at the encoder's outputs, like the
    # data-parallel split of TrainStep._forward); the decoder, the matcher, the criterion and their backward run eagerly
    # between the two replays.  DFINE_HYBRID_GRAPH=0 keeps such steps fully eager.
    def _uncaptured_step(self, inputs, targets):
        import os
        m = self.model
        ok = (os.environ.get("DFINE_HYBRID_GRAPH", "1") != "0" and hasattr(m, "backbone") and hasattr(m, "encoder")
              and hasattr(m, "decoder") and m.training)
        if not ok:
            return TrainStep.__call__(self, inputs, targets)
        skey = tuple(inputs.shape)
        st = self._static.get(skey)
        if st is None:
            n = self._static_seen.get(skey, 0)
            self._static_seen[skey] = n + 1
            if n < self.eager_steps:                   # eager steps double as the warm-up CUDA graphs need
                return TrainStep.__call__(self, inputs, targets)
            st = self._static[skey] = self._capture_static(inputs)
            while len(self._static) > 2:               # (each owns the activations of a backbone + encoder pass)
                self._static.pop(next(iter(self._static)))
        return self._hybrid_step(st, inputs, targets)

    def _capture_static(self, inputs):
        from . import cuda_ops
        m, dev = self.model, inputs.device
        st = {"x": inputs.clone()}
        pool, saved = cuda_ops._ZeroPool(), cuda_ops.zero_pool
        cuda_ops.weights_changed()          # every cached weight re-layout is stale: the captures re-create them as graph nodes
        torch.cuda.synchronize()
        gF, gB = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        try:
            cuda_ops.zero_pool = pool       # the captured half owns its fp64 scratch (zeroed by a captured memset)
            with torch.cuda.graph(gF, stream=self._side):
                pool.begin_step(dev)
                cuda_ops.wgrad_stream.begin(dev)
                feats = list(m.encoder(m.backbone(st["x"].permute(0, 2, 3, 1))))
                cuda_ops.wgrad_stream.sync_main()
            st["gbuf"] = [torch.zeros_like(f) for f in feats]
            with torch.cuda.graph(gB, pool=gF.pool(), stream=self._side):
                pool.active = True
                torch.autograd.backward(feats, st["gbuf"])
                cuda_ops.wgrad_stream.sync_main()
        finally:
            pool.end_step()
            cuda_ops.zero_pool = saved
            cuda_ops.wgrad_stream.join()
        st.update(gF=gF, gB=gB, feats=feats)
        return st

    def _hybrid_step(self, st, inputs, targets):
        from . import cuda_ops
        st["x"].copy_(inputs, non_blocking=True)
        _begin_step(inputs, self._counters)
        st["gF"].replay()
        leaves = [f.detach().requires_grad_(True) for f in st["feats"]]
        output = self.model.decoder(leaves, targets)
        _after_forward()
        loss_dict = self.loss_fn(output, targets)
        loss = sum(loss_dict.values())
        loss.backward()
        cuda_ops.wgrad_stream.sync_main()
        for b, l in zip(st["gbuf"], leaves):
            if l.grad is None:
                b.zero_()
            else:
                b.copy_(l.grad)
        st["gB"].replay()
        _end_step()
        self.batch_idx += 1
        self.optimizer_step()
        cuda_ops.weights_changed()
        return loss.detach(), loss_dict

    def _capture(self, inputs, targets, plan):
        from . import cuda_ops
        g = {}
        dev = inputs.device
        g["x"] = inputs.clone()
        g["targets"] = [{k: t[k].clone() for k in ("labels", "boxes", "masks") if t.get(k) is not None} for t in targets]
        crit = self.loss_fn
        # capture records launches without running them, so the index table used while capturing graph B is
        # the (same-shaped) plan of the last eager step of this key; replays refill it before graph B runs
        g["plan"] = plan
        n0 = cuda_ops.counters.launches
        g["table"] = plan.table.to(dev)
        g["counts"] = plan.counts.to(dev)
        self.optimizer.zero_grad()
        torch.cuda.synchronize()
        gA = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gA, stream=self._side):
            _begin_step(g["x"], self._counters)
            out, cut = self._forward(g["x"], g["targets"])
            raw, tg = crit.match(out, g["targets"])
            _after_forward()
        gB = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gB, pool=gA.pool(), stream=self._side):
            loss_dict = crit.compute(out, tg, g["table"], g["counts"], plan)
            loss = sum(loss_dict.values())
            loss.backward()
            if cut is None:
                _end_step()
            else:
                cuda_ops.wgrad_stream.sync_main()      # (a capture must end with its forked streams joined)
            g["loss"] = loss.detach()
            g["loss_dict"] = {k: v.detach() for k, v in loss_dict.items()}
        g["gB2"] = None
        if cut is not None:       # second half of the backward pass: the backbone, overlapped with the first all-reduce
            gB2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gB2, pool=gA.pool(), stream=self._side):
                self._backward_half2(cut)
                _end_step()
            g["gB2"] = gB2
        gC = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gC, pool=gA.pool(), stream=self._side):
            self._apply_grads()
        g.update(gA=gA, gB=gB, gC=gC, out=out, raw=raw, launches=cuda_ops.counters.launches - n0)
        return g

    def _replay(self, g, inputs, targets):
        g["x"].copy_(inputs, non_blocking=True)
        for s, t in zip(g["targets"], targets):
            for k, v in s.items():
                v.copy_(t[k], non_blocking=True)
        self._host_prepare()
        g["gA"].replay()
        ev = g.setdefault("gap_events", (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
        ev[0].record()
        # (Enqueueing graph B behind a stream memory-wait while the host plans was tried: no gain where the planning
        #  takes 0.6 ms, and a ~3800-node graph submitted to a stream that cannot drain blocks the submitting host
        #  thread — which is the thread that would raise the flag.  The plain order stays.)
        plan = self.loss_fn.plan(g["out"], g["targets"], g["raw"], g["plan"], local_counts=True)   # syncs on the matcher D2H
        self._replay_tail(g, inputs, plan, ev)
        if self.scheduler is not None:
            self.scheduler.step()
        from . import cuda_ops
        cuda_ops.weights_changed()                       # graph C rewrote the parameters
        cuda_ops.counters.launches += g["launches"]      # library kernels replayed by the graphs
        self.batch_idx += 1
        # the graph's output buffers are overwritten by every replay: hand out copies (one stacked clone), so that a
        # caller may keep the losses of several steps (src/dl/train.Trainer averages them per epoch)
        vals = torch.stack([g["loss"]] + list(g["loss_dict"].values())).clone()
        return vals[0], dict(zip(g["loss_dict"].keys(), vals[1:].unbind(0)))

    def _replay_tail(self, g, inputs, plan, ev):
        """Everything the device runs after the index table: enqueued only (no host wait in here)."""
        g["table"].copy_(plan.table, non_blocking=True)
        g["counts"].copy_(plan.counts, non_blocking=True)
        self.loss_fn.finish_counts(g["counts"])          # 2-float all-reduce + clamp on the device: no host wait
        ev[1].record()                                   # ev[0] -> ev[1] = device idle time while the host plans
        self._last_gap = ev
        g["gB"].replay()
        if g["gB2"] is None:
            self.optimizer.allreduce_grads()
        else:
            bb, rest = self.optimizer.backbone_groups()
            comm = self._comm_stream(inputs.device)
            self.optimizer.allreduce_grads(rest, comm)   # encoder / decoder arenas: overlaps the backbone's backward
            g["gB2"].replay()
            self.optimizer.allreduce_grads(bb, comm)
            torch.cuda.current_stream().wait_stream(comm)
        g["gC"].replay()
