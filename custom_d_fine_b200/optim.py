"""Flat-arena optimizer: gradient clip + AdamW + EMA + zero_grad as three kernel launches per parameter group.

Drop-in for what ``build_optimizer`` returns in the reference (/root/reference/src/d_fine/dfine.py:87-124: AdamW
over four parameter groups) plus the device half of ``Trainer.optimizer_step`` (train.py:512-535:
``clip_grad_norm_`` -> ``optimizer.step`` -> ``zero_grad``) and ``ModelEMA.update`` (train.py:62-73).

B200-first layout: every trainable parameter of a group is re-homed (``p.data`` becomes a view) into one flat
fp32 arena; gradients (``p.grad`` views), both Adam moments and the EMA copy use arenas with the same element
order.  One step is ``sumsq`` (global grad norm) + ``adamw_ema`` per group + one ``ema_blend`` over the
floating-point buffers — HBM-bound passes of 36 B/parameter instead of ~2 k tiny kernels.  All per-step scalars
(lr, weight decay, step count, EMA momentum) are read from a small device table refreshed by one pinned H2D
copy, so the launches replay unchanged from a CUDA graph while a scheduler keeps mutating ``param_groups``.

Known divergence from torch.optim.AdamW: the kernel updates every arena element, so a parameter that received no
gradient in a step (the mask head on a batch without masks) still sees weight decay and its decaying momentum, where
torch skips parameters whose ``.grad`` is None.  No BASELINE config has such a step.

Data-parallel runs all-reduce the flat gradient arenas (one NCCL call per group over NVLink) in
``allreduce_grads`` — the only collective on the gradient path.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import dist as dist_utils


def _pad4(n):
    # 8 elements: float4 access for the fp32 arenas AND 16-byte TMA alignment of the matching fp16 operand planes
    return (n + 7) // 8 * 8


def _arena_view(arena, off, t):
    """View of ``arena`` with ``t``'s shape.  Dense conv weights [Co,Ci,kh,kw] are stored channels-last
    ([Co,kh,kw,Ci] in memory): that is the K-major layout the implicit-GEMM kernels read, so neither the forward
    weight operand nor the weight gradient ever needs a re-layout copy."""
    n = t.numel()
    if t.dim() == 4 and t.shape[1] > 1 and t.shape[2] * t.shape[3] > 1:
        co, ci, kh, kw = t.shape
        return arena[off:off + n].view(co, kh, kw, ci).permute(0, 3, 1, 2)
    return arena[off:off + n].view(t.shape)


def flatten_into_arena(tensors, device=None):
    """Re-home ``tensors`` (fp32) into one contiguous arena; each tensor's ``.data`` becomes a view whose offset
    is a multiple of 4 floats (16-byte aligned: TMA / float4 access).  Returns (arena, offsets)."""
    offs, total = [], 0
    for t in tensors:
        offs.append(total)
        total += _pad4(t.numel())
    device = device or (tensors[0].device if tensors else "cpu")
    arena = torch.zeros(max(total, 4), dtype=torch.float32, device=device)
    for t, o in zip(tensors, offs):
        view = _arena_view(arena, o, t)
        view.copy_(t.detach())
        t.data = view
    return arena, offs


class FusedAdamW(torch.optim.Optimizer):
    """AdamW with the reference's four groups, fused with gradient clipping, EMA and zero_grad (CUDA only)."""

    fused_step = True

    def __init__(self, param_groups, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(param_groups, defaults)
        self.max_norm = float(max_norm)
        self._t = 0
        self._ema = None
        self._ema_m = 0.0
        self._arenas = []
        dev = None
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.requires_grad]
            if not ps:
                self._arenas.append(None)
                continue
            dev = ps[0].device
            if not ps[0].is_cuda:
                raise RuntimeError("FusedAdamW needs CUDA parameters (there is no CPU compute path)")
            pflat, offs = flatten_into_arena(ps)
            gflat = torch.zeros_like(pflat)
            from . import cuda_ops
            for p, o in zip(ps, offs):
                p.grad = _arena_view(gflat, o, p)          # same layout as the parameter: kernels accumulate in place
                cuda_ops.register_direct_grad(p)           # backward kernels may add straight into this view
            a = dict(params=ps, offs=offs, p=pflat, g=gflat, m=torch.zeros_like(pflat), v=torch.zeros_like(pflat),
                     ema=None, planes=None)
            if any(p.dim() >= 2 for p in ps):
                # operand planes (hi | lo) of the forward GEMMs in the arena's element order, rewritten by the AdamW
                # kernel itself: fp16 planes of w * 2^8 for the 3xFP16 mode, fp32 tf32-split planes for 3xTF32
                from . import cuda_ops
                half = cuda_ops.get_gemm_mode() == "hf3"
                a["planes"] = torch.empty((2, pflat.numel()), dtype=torch.float16 if half else torch.float32,
                                          device=pflat.device)
            self._arenas.append(a)
        self._dev = dev
        n = len(self.param_groups)
        self._hyper = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        self._hyper_host = [torch.zeros((n, 4), dtype=torch.float32).pin_memory() for _ in range(4)]
        self._gnorm = torch.zeros(1, dtype=torch.float64, device=dev)
        self._buf_src = self._buf_ema = None
        self._mom = torch.zeros(1, dtype=torch.float32, device=dev)
        self._register_planes()

    def _register_planes(self):
        """Split every arena once and tell the kernel table where each dense weight's tf32 hi / lo planes live."""
        from . import cuda_ops
        for a in self._arenas:
            if a is None or a["planes"] is None:
                continue
            hi, lo = a["planes"][0], a["planes"][1]
            half = hi.dtype == torch.float16
            if half:
                cuda_ops._check(cuda_ops.lib().dfine_f16_split_flat(cuda_ops._p(a["p"]), cuda_ops._p(hi), cuda_ops._p(lo),
                                                                    ctypes.c_long(a["p"].numel()),
                                                                    ctypes.c_float(cuda_ops._F16_WSCALE), cuda_ops._stream()),
                                "f16_split_flat")
            else:
                cuda_ops._check(cuda_ops.lib().dfine_tf32_split(cuda_ops._p(a["p"]), cuda_ops._p(hi), cuda_ops._p(lo),
                                                                ctypes.c_long(a["p"].numel()), cuda_ops._stream()),
                                "tf32_split")
            for p, o in zip(a["params"], a["offs"]):
                if p.dim() >= 2:
                    n = p.numel()
                    rows = p.shape[0]
                    # (fp16 planes are read by TMA with every tap's channel run starting on a 16-byte boundary: the
                    #  arena order qualifies when the channel count is a multiple of 8; other layers split per use)
                    if half and (p.shape[1] % 8 != 0 or o % 8 != 0):
                        continue
                    cuda_ops.register_weight_planes(p, hi[o:o + n].view(rows, n // rows), lo[o:o + n].view(rows, n // rows))

    # ---- wiring -------------------------------------------------------------------------------
    def attach_ema(self, ema, student):
        """Lay the EMA model's parameters / floating-point buffers out in arenas matching the student's."""
        from torch.nn.parallel import DistributedDataParallel as DDP
        if self._ema is ema:
            return
        if isinstance(student, DDP):
            student = student.module
        e_params = dict(ema.model.named_parameters())
        s_names = {id(p): n for n, p in student.named_parameters()}
        for a in self._arenas:
            if a is None:
                continue
            eps_ = [e_params[s_names[id(p)]] for p in a["params"]]
            a["ema"], offs = flatten_into_arena(eps_)
            assert offs == a["offs"]
        # floating-point entries the optimizer does not own (buffers, frozen parameters): plain EMA blend
        owned = {id(p) for a in self._arenas if a is not None for p in a["params"]}
        s_state, e_state = student.state_dict(keep_vars=True), ema.model.state_dict(keep_vars=True)
        src, dst = [], []
        for k, v in s_state.items():
            if v.dtype == torch.float32 and id(v) not in owned and k in e_state:
                src.append(v)
                dst.append(e_state[k])
        # aliases (decoder.up / decoder.decoder.up ...) appear twice in the state dict: keep one
        seen, s2, d2 = set(), [], []
        for s, d in zip(src, dst):
            if id(s) in seen:
                continue
            seen.add(id(s))
            s2.append(s)
            d2.append(d)
        if s2:
            self._buf_src, _ = flatten_into_arena(s2)
            self._buf_ema, _ = flatten_into_arena(d2)
        self._ema = ema

    def broadcast_state(self, src=0):
        """Rank-``src`` parameters and buffers to every rank (what the reference's DDP wrap does at construction)."""
        if dist_utils.get_world_size() < 2:
            return
        import torch.distributed as dist
        for a in self._arenas:
            if a is not None:
                dist.broadcast(a["p"], src)
                if a["ema"] is not None:
                    dist.broadcast(a["ema"], src)
        if self._buf_src is not None:
            dist.broadcast(self._buf_src, src)
            dist.broadcast(self._buf_ema, src)
        self._register_planes()      # the arenas were overwritten behind the parameters' version counters

    # ---- checkpoint format: torch.optim.AdamW's (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``) ------------
    def state_dict(self):
        """The Adam moments live in flat arenas; they are exposed here per parameter in torch.optim.AdamW's own
        layout (logical parameter shapes, cloned), so a checkpoint written by this optimizer resumes in either this
        class or a stock ``torch.optim.AdamW`` over the same groups (the reference's optimizer, dfine.py:87-124)."""
        self.state.clear()
        for a in self._arenas:
            if a is None:
                continue
            for p, o in zip(a["params"], a["offs"]):
                self.state[p] = {"step": torch.tensor(float(self._t)),
                                 "exp_avg": _arena_view(a["m"], o, p).detach().clone(),
                                 "exp_avg_sq": _arena_view(a["v"], o, p).detach().clone()}
        sd = super().state_dict()
        self.state.clear()
        sd["fused"] = {"t": self._t, "ema_momentum": self._ema_m}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        fused = state_dict.pop("fused", None)
        super().load_state_dict(state_dict)            # param_groups (lr, betas, ...) + per-parameter state tensors
        t = None
        with torch.no_grad():
            for a in self._arenas:
                if a is None:
                    continue
                for p, o in zip(a["params"], a["offs"]):
                    st = self.state.get(p)
                    if not st:
                        continue
                    _arena_view(a["m"], o, p).copy_(st["exp_avg"])
                    _arena_view(a["v"], o, p).copy_(st["exp_avg_sq"])
                    t = int(st["step"]) if t is None else t
        self.state.clear()
        if fused is not None:
            self._t, self._ema_m = int(fused["t"]), float(fused["ema_momentum"])
        elif t is not None:
            self._t = t
        self._register_planes()

    def allreduce_grads(self, groups=None, stream=None):
        """Average the flat gradient arenas over ranks (NCCL over NVLink; the step's only gradient collective).
        ``groups``: indices of the parameter groups to reduce (default all).  ``stream``: issue the collectives on this
        side stream after it has waited for the current one — the caller joins it before the optimizer kernels
        (train.GraphedTrainStep reduces the encoder / decoder arenas while the backbone's backward is still running)."""
        if dist_utils.get_world_size() < 2:
            return
        idx = range(len(self._arenas)) if groups is None else groups
        if stream is None:
            for i in idx:
                if self._arenas[i] is not None:
                    dist_utils.allreduce_mean_(self._arenas[i]["g"])
            return
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            for i in idx:
                if self._arenas[i] is not None:
                    dist_utils.allreduce_mean_(self._arenas[i]["g"])

    def backbone_groups(self):
        """(backbone group indices, other group indices) by the reference's grouping (dfine.py:87-124: groups 0 / 1 hold the
        backbone's weights / norms, 2 / 3 the encoder's and decoder's)."""
        n = len(self._arenas)
        return [i for i in (0, 1) if i < n], [i for i in range(2, n)]

    # ---- step ----------------------------------------------------------------------------------
    def prepare(self, ema_momentum=None):
        """Host half of a step (NOT captured): bump the step count and refresh the device scalar table from
        ``param_groups`` (a scheduler may have changed lr) with one pinned, stream-ordered H2D copy."""
        self._t += 1
        if ema_momentum is not None:
            self._ema_m = float(ema_momentum)
        h = self._hyper_host[self._t % len(self._hyper_host)]
        for i, g in enumerate(self.param_groups):
            h[i, 0], h[i, 1], h[i, 2], h[i, 3] = g["lr"], g["weight_decay"], float(self._t), self._ema_m
        self._hyper.copy_(h, non_blocking=True)

    @torch.no_grad()
    def step(self, closure=None):
        """Device half (graph-capturable): grad-norm, then clip + AdamW + EMA + zero_grad per group."""
        from .cuda_ops import _F16_WSCALE as _wscale
        from .cuda_ops import _check, _p, _stream, lib, weights_changed
        L = lib()
        weights_changed()          # parameters are rewritten through raw pointers: invalidate re-laid weight copies
        use_clip = self.max_norm > 0
        if use_clip:
            self._gnorm.zero_()
            for a in self._arenas:
                if a is not None:
                    _check(L.dfine_sumsq(_p(a["g"]), ctypes.c_long(a["g"].numel()), _p(self._gnorm), _stream()), "sumsq")
        for i, (a, g) in enumerate(zip(self._arenas, self.param_groups)):
            if a is None:
                continue
            b1, b2 = g["betas"]
            _check(L.dfine_adamw_ema(_p(a["p"]), _p(a["g"]), _p(a["m"]), _p(a["v"]), _p(a["ema"]),
                                     ctypes.c_long(a["p"].numel()), ctypes.c_void_p(self._hyper[i].data_ptr()),
                                     _p(self._gnorm) if use_clip else None, ctypes.c_float(self.max_norm),
                                     ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(g["eps"]), 1,
                                     _p(a["planes"][0]) if a["planes"] is not None else None,
                                     _p(a["planes"][1]) if a["planes"] is not None else None,
                                     1 if (a["planes"] is not None and a["planes"].dtype == torch.float16) else 0,
                                     ctypes.c_float(_wscale), _stream()),
                   "adamw_ema")
        if self._buf_src is not None:
            _check(L.dfine_ema_blend(_p(self._buf_ema), _p(self._buf_src), ctypes.c_long(self._buf_src.numel()),
                                     ctypes.c_void_p(self._hyper[0, 3:].data_ptr()), _stream()), "ema_blend")

    def zero_grad(self, set_to_none=False):
        """Gradients are zeroed by ``step``; an explicit call zeroes the arenas (never detaches the views)."""
        for a in self._arenas:
            if a is not None:
                a["g"].zero_()

    def grad_norm(self):
        """sqrt of the squared norm accumulated by the last ``step`` (device tensor)."""
        return self._gnorm.sqrt()


def ema_momentum(base, iters):
    """train.py:57-60: momentum * (1 - exp(-iters / 2000))."""
    return base * (1 - math.exp(-iters / 2000))
