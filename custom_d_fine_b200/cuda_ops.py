"""The product's only compute provider: the kernel table backed by ``csrc/libdfine_sm100.so``.

Every method of :class:`CudaOps` is an entry of the kernel table ``K`` used by the host graph.
Tensors, streams and autograd bookkeeping are PyTorch's; all arithmetic of the ops listed in
SURVEY §8a runs in the hand-written sm_100a kernels reached through the C ABI declared in
``include/dfine_sm100.h`` (ctypes — no torch types cross the boundary).  There is no CPU path:
a missing library, a non-CUDA tensor or a non-zero return code raises.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from ctypes import c_float, c_int, c_long, c_void_p
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libdfine_sm100.so"
_lib = None

ACT = {None: 0, "relu": 1, "silu": 2, "gelu": 3}


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing — build it with `python -m custom_d_fine_b200.build` "
                "(there is no fallback compute path)")
        L = ctypes.CDLL(str(_LIB_PATH))
        for name, (ret, args) in abi_prototypes().items():
            fn = getattr(L, name)          # AttributeError = header/library mismatch: fail loudly
            fn.restype, fn.argtypes = ret, args
        _lib = L
    return _lib


_HEADER = Path(__file__).resolve().parent.parent / "include" / "dfine_sm100.h"


def abi_prototypes(header: Path = _HEADER):
    """Parse ``include/dfine_sm100.h`` -> {name: (restype, [argtypes])} for ctypes."""
    import re

    def ctype(decl):
        decl = decl.strip()
        if "*" in decl:
            return c_void_p
        return {"int": c_int, "long": c_long, "float": c_float, "double": ctypes.c_double}[decl.rsplit(" ", 1)[0]]

    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    protos = {}
    for m in re.finditer(r"(const char\*|int|long)\s+(dfine_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.groups()
        restype = {"const char*": ctypes.c_char_p, "int": c_int, "long": c_long}[ret]
        arglist = [] if args.strip() == "void" else [ctype(a) for a in args.split(",")]
        protos[name] = (restype, arglist)
    return protos


class _Counters:
    """Launch accounting for bench.py: every C-ABI call that enqueues kernels bumps ``launches``;
    ``timed`` (kernel-family name -> list of (start, end, algorithmic_bytes)) is filled only while
    ``watch`` names that family (CUDA events on the launching stream)."""
    launches = 0
    watch = ()
    timed = {}


counters = _Counters()


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"libdfine_sm100 {what} failed (rc={rc}): {lib().dfine_last_error().decode()}")
    counters.launches += 1


_SPIN_CYCLES = 80_000      # ~40 us at 1.97 GHz


class _timed:
    def __init__(self, name, nbytes, flops=0, tag=""):
        self.on = name in counters.watch
        self.name, self.nbytes, self.flops, self.tag = name, nbytes, flops, tag

    def __enter__(self):
        if self.on:
            self.s, self.e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # In an eager step the device is idle when a launch arrives, so an event recorded before the (ctypes) call
            # would also time the host's 10-20 us of argument marshalling.  A short spin kernel queued first keeps the
            # stream busy until start event, kernel and stop event are all enqueued: the pair then brackets the kernel.
            if hasattr(torch.cuda, "_sleep"):
                torch.cuda._sleep(_SPIN_CYCLES)
            self.s.record()

    def __exit__(self, *exc):
        if self.on:
            self.e.record()
            counters.timed.setdefault(self.name, []).append((self.s, self.e, self.nbytes, self.flops, self.tag))


def _p(t):
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


class _ZeroPool:
    """Pre-zeroed fp64 scratch for the per-layer BatchNorm statistics / backward reductions.  The reference's
    equivalent is implicit in ATen; here ~270 tiny `torch.zeros` launches per step become ONE memset: the train
    step calls ``begin_step`` (also under CUDA-graph capture, where the memset is captured once), ops take
    slices with ``take``; without an active step (unit tests calling ops directly) ``take`` allocates."""

    def __init__(self):
        self.buf, self.off, self.active = None, 0, False

    def begin_step(self, device, n_doubles=1 << 19):
        if self.buf is None or self.buf.device != device or self.buf.numel() < n_doubles:
            self.buf = torch.empty(n_doubles, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.off, self.active = 0, True

    def end_step(self):
        self.active = False

    def take(self, n, device):
        n_al = (n + 1) // 2 * 2          # keep 16-byte alignment
        if not self.active or self.buf is None or self.buf.device != device or self.off + n_al > self.buf.numel():
            return torch.zeros(n, dtype=torch.float64, device=device)
        out = self.buf[self.off:self.off + n]
        self.off += n_al
        return out


zero_pool = _ZeroPool()


_scalar_cache = {}


_ones = {}


def _ones_cache(n, device):
    key = (n, str(device))
    t = _ones.get(key)
    if t is None:
        t = _ones[key] = torch.ones(n, device=device, dtype=torch.float32)
    return t


def _as_dev_scalar(v, device):
    """A [1] fp32 device tensor for a kernel that reads the scalar from memory (tensors pass through; python
    numbers are uploaded once per value and cached — a pageable H2D copy is not graph-capturable)."""
    if isinstance(v, torch.Tensor):
        return v
    key = (float(v), str(device))
    t = _scalar_cache.get(key)
    if t is None:
        t = _scalar_cache[key] = torch.full((1,), float(v), device=device, dtype=torch.float32)
    return t


class _WgradStream:
    """Weight gradients on a second stream.  Nothing in the backward pass consumes a weight gradient (the kernels
    accumulate straight into the optimizer's flat arenas), so each conv / linear weight-gradient launch is forked
    off the backward stream right after its operands exist and the chain of data gradients continues without
    waiting for it; the step joins the stream once, after backward (train._end_step).  Under CUDA-graph capture the
    fork / join become graph edges.  Operand tensors are kept alive until the join: the caching allocator must not
    hand their memory to a later main-stream allocation while the side stream still reads it.
    Active only between train._begin_step and train._end_step (DFINE_WGRAD_STREAM=0 disables it)."""

    def __init__(self):
        self.stream, self.on, self.used, self.keep = None, False, False, []
        self.in_step = False       # between train._begin_step and train._end_step (also when the stream itself is off)
        self.disabled = False      # bench.py serialises the step while it brackets single kernels with events

    def begin(self, device):
        _msda_shared.clear()
        self.in_step = True
        if self.disabled or os.environ.get("DFINE_WGRAD_STREAM", "1") == "0":
            return
        if self.stream is None or self.stream.device != device:
            self.stream = torch.cuda.Stream(device)
        self.on = True

    def run(self, fn, *keep):
        if not self.on:
            fn()
            return
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            fn()
        self.keep.extend(keep)
        self.used = True

    def sync_main(self):
        """The main stream waits for everything queued here so far (end of the forward pass: the re-laid weights
        prepared for the backward; also required before a CUDA-graph capture that contains a fork ends)."""
        if self.used:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.keep.clear()
        self.used = False

    def join(self):
        self.sync_main()
        self.on = self.in_step = False


wgrad_stream = _WgradStream()


# BatchNorm `num_batches_tracked` buffers re-homed into one int64 arena by train.TrainStep: one add per step instead of
# one tiny launch per BatchNorm layer (133 for D-FINE-m).  data_ptr -> True.
_step_counters = {}


def register_step_counters(model):
    """Re-home every BatchNorm step counter of `model` (CUDA) into one int64 arena; returns the arena (or None)."""
    mods = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d) and m.num_batches_tracked is not None
            and m.num_batches_tracked.is_cuda]
    if not mods:
        return None
    arena = torch.stack([m.num_batches_tracked.detach().reshape(()) for m in mods]).contiguous()
    for i, m in enumerate(mods):
        m.num_batches_tracked.data = arena[i]
        _step_counters[arena[i].data_ptr()] = True
    return arena


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("custom_d_fine_b200 ops need CUDA tensors (no CPU path)")


def _pad_ld(C):
    """Pixel stride for a freshly allocated NHWC activation of C channels.  Padding the 12 / 24 / 48-channel stem
    tensors to 128-byte rows was measured (profiles/README.md): the tcgen05 kernels on them did not get faster (their
    cost is TMA's per-row zero fill of the channels past Cin, not the row alignment) while the BatchNorm and copy
    kernels got slower, so activations stay dense.  The kernels keep their pixel-stride parameters."""
    return C


def _alloc_nhwc(B, H, W, C, device):
    ld = _pad_ld(C)
    if ld == C:
        return torch.empty((B, H, W, C), device=device, dtype=torch.float32), C
    return torch.empty((B, H, W, ld), device=device, dtype=torch.float32)[..., :C], ld


def _rows(x):
    """View ``x`` [..., C] as rows with one uniform row stride; returns (tensor, n_rows, ld)."""
    C = x.shape[-1]
    if x.stride(-1) != 1 and C > 1:
        x = x.contiguous()
    n = x.numel() // C if C else 0
    if x.dim() == 1:
        return x, 1, C
    ld = x.stride(-2)
    exp = ld
    ok = ld >= C
    for d in range(x.dim() - 2, -1, -1):
        if x.shape[d] != 1 and x.stride(d) != exp:
            ok = False
            break
        exp *= x.shape[d]
    if not ok or (x.data_ptr() % 16) or (ld % 4):
        x = x.contiguous()
        ld = C
    return x, n, ld


# Dense conv / linear shapes the tcgen05 kernels accept run on tensor cores.  Modes (env DFINE_GEMM / set_gemm_mode):
#   "hf3"  (default) forward GEMMs as error-compensated 3xFP16 — see the entry below; the parity mode of the tensor-core
#          path since round 2 (measured closer to the fp32 oracle than "tc3" on the headline network, 6.5 % faster);
#   "tc3"  forward GEMMs as error-compensated 3xTF32 (fp32-class accuracy), data / weight gradients as plain kind::tf32;
#   "tch"  forward GEMMs as a hybrid split: a_hi*w_hi on kind::tf32, the cross terms a_lo*w_hi + a_hi*w_lo on bf16
#          copies through kind::f16 (3xTF32-class accuracy for 2/3 of its tensor time), gradients as in "tc3";
#   "bf3"  forward GEMMs as error-compensated 3xBF16 (two bf16 parts per operand = 16 mantissa bits, three
#          kind::f16 MMAs at twice the tf32 rate; ~1e-5 relative), gradients as in "tc3";
#   "hf3"  forward GEMMs as error-compensated 3xFP16 (two fp16 parts per operand = 22 significand bits like 3xTF32, three
#          kind::f16 MMAs at twice the tf32 rate, weights pre-scaled by 2^8), gradients as in "tc3";
#   "tc"   plain kind::tf32 everywhere (the precision class of the reference's cuDNN convolutions on GPU);
#   "simt" fp32 CUDA-core kernels of the same library (strict-fp32 parity runs and kernel bring-up).
_MODE = os.environ.get("DFINE_GEMM", "hf3")
# Gradient-routing aliases (`tap`): fold the gradient-accumulation add of a tensor with two consumers into the data-
# gradient kernel's epilogue.  Measured on B200 (profiles/README.md): the extra epilogue read makes the persistent dgrad
# kernels epilogue-bound and the step 0.8 ms SLOWER than the separate add kernels -> off by default, kept for A/B.
_TAP = os.environ.get("DFINE_TAP", "0") == "1"


def set_gemm_mode(mode: str) -> None:
    global _MODE
    if mode not in ("tc", "tc3", "tch", "bf3", "hf3", "simt"):
        raise ValueError(mode)
    _MODE = mode


def get_gemm_mode() -> str:
    return _MODE


def _tc_ok(Cin, Cout, k, stride, pad, ldx, ldy):
    if _MODE == "simt":
        return False
    t, l, b, r = pad
    return bool(lib().dfine_conv_tc_supported(Cin, Cout, k, k, stride, t, l, b, r, c_long(ldx), c_long(ldy)))


def _conv2x2_ok(geom, ldx, ldy):
    """The stem's kernel-2 convs go to the direct fp32 kernels of csrc/stem.cu (DFINE_CONV2X2=0: tensor-core path)."""
    B, H, W, Cin, OH, OW, Cout, k, stride, pad = geom
    if k != 2 or _MODE == "simt" or not _CONV2X2:
        return False
    return bool(lib().dfine_conv2x2_supported(Cin, Cout, k, k, stride, pad[0], pad[1], pad[2], pad[3], c_long(ldx), c_long(ldy)))


_CONV2X2 = os.environ.get("DFINE_CONV2X2", "1") != "0"
_TC_TS = os.environ.get("DFINE_TC_TS", "1") != "0"
_FOLD_FROZEN = os.environ.get("DFINE_FOLD_FROZEN_BN", "1") != "0"   # frozen / eval BatchNorm + ReLU in the conv epilogue
_MASK_PM = os.environ.get("DFINE_MASK_PM", "1") != "0"      # matched-mask logits through the tcgen05 mask product
_taps_cache = {}


def _taps(key, make):
    arr = _taps_cache.get(key)
    if arr is None:
        flat = [int(v) for t in make() for v in t]
        arr = ((c_int * len(flat))(*flat), len(flat) // 3)
        _taps_cache[key] = arr
    return arr


def _tc_launch(x, ldx, H, W, Cin, w, w_lo, ldw, bias, y, ldy, B, OH, OW, Cout, YH, YW, os_, oo, in_stride, taps, act,
               stats, what, bf16_planes=None, res=None, ldres=0, plane_stride=0, lab=None, grouped=(0, 0, 0), ch_scale=None,
               bn_fin=None):
    """grouped = (total weight rows, per-image row offset, per-image column offset): per-image weights (the mask product).
    bn_fin = (counter, bn_w, bn_b, running_mean, running_var, mean, invstd, scale, shift, momentum, eps): the train-mode
    BatchNorm finalize runs in the kernel's tail (3xFP16 tensor-memory kernel only)."""
    arr, n = taps
    # algorithmic bytes: input pixels + output pixels + weights, each touched once, fp32
    nbytes = 4 * (B * min(H * W, OH * OW * in_stride * in_stride) * Cin + B * OH * OW * Cout + Cout * n * Cin)
    if bn_fin is not None and len(bn_fin) > 11:      # the fused normalise pass: conv output read again, y (+ residual) moved
        nbytes += 4 * B * OH * OW * Cout * (2 + (bn_fin[13] is not None))
    with _timed("conv_tc", nbytes, 2 * B * OH * OW * Cout * n * Cin,
                f"{what} {Cin}->{Cout} taps{n} s{in_stride} {OH}x{OW} B{B}{' planes' if bf16_planes is not None else ''}"):
        if bf16_planes is not None and w is not None:      # hybrid: tf32 hi plane + bf16 cross-term planes
            _check(lib().dfine_conv_tc_hybrid(_p(x), _p(w), _p(bf16_planes), _p(bias), _p(y), _p(stats), B, H, W, Cin,
                                              c_long(ldx), OH, OW, Cout, c_long(ldy), YH, YW, os_[0], os_[1], oo[0],
                                              oo[1], in_stride, n, arr, c_long(ldw), c_long(bf16_planes.shape[-1]),
                                              act, _stream()), what)
        elif bn_fin is not None:
            assert bf16_planes is not None and bf16_planes.dtype == torch.float16 and bias is None and act == 0 \
                and lab is None and ch_scale is None and os_ == (1, 1) and oo == (0, 0) and (YH, YW) == (OH, OW)
            cnt, bw, bb, rm, rv, mean, invstd, scale, shift, momentum, eps = bn_fin[:11]
            y_out, ld_out, post, ld_post, lab_s, lab_b, bn_act = bn_fin[11:] if len(bn_fin) > 11 else (None, 0, None, 0, None, None, 0)
            _check(lib().dfine_conv_tc_f16x3_bn(_p(x), _p(bf16_planes), _p(y), _p(stats), _p(cnt), _p(bw), _p(bb), _p(rm),
                                                _p(rv), _p(mean), _p(invstd), _p(scale), _p(shift), c_float(momentum),
                                                c_float(eps), B, H, W, Cin, c_long(ldx), OH, OW, Cout, c_long(ldy),
                                                in_stride, n, arr, c_long(bf16_planes.shape[-1]),
                                                c_float(1.0 / _F16_WSCALE), c_long(plane_stride), _p(y_out), c_long(ld_out),
                                                _p(post), c_long(ld_post), _p(lab_s), _p(lab_b), bn_act, _stream()), what)
        elif bf16_planes is not None and bf16_planes.dtype == torch.float16:
            _check(lib().dfine_conv_tc_f16x3(_p(x), _p(bf16_planes), _p(bias), _p(y), _p(stats), B, H, W, Cin,
                                             c_long(ldx), OH, OW, Cout, c_long(ldy), YH, YW, os_[0], os_[1], oo[0],
                                             oo[1], in_stride, n, arr, c_long(bf16_planes.shape[-1]), act,
                                             c_float(1.0 / _F16_WSCALE), c_long(plane_stride), 0 if lab is None else 1,
                                             c_float(1.0 if lab is None else lab[0]), c_float(0.0 if lab is None else lab[1]),
                                             grouped[0], grouped[1], _p(ch_scale), _stream()), what)
        elif bf16_planes is not None:
            assert ch_scale is None
            _check(lib().dfine_conv_tc_bf16x3(_p(x), _p(bf16_planes), _p(bias), _p(y), _p(stats), B, H, W, Cin,
                                              c_long(ldx), OH, OW, Cout, c_long(ldy), YH, YW, os_[0], os_[1], oo[0],
                                              oo[1], in_stride, n, arr, c_long(bf16_planes.shape[-1]), act, _stream()),
                   what)
        else:
            _check(lib().dfine_conv_tc(_p(x), _p(w), _p(w_lo), _p(bias), _p(y), _p(stats), B, H, W, Cin, c_long(ldx),
                                       OH, OW, Cout, c_long(ldy), YH, YW, os_[0], os_[1], oo[0], oo[1], in_stride, n,
                                       arr, c_long(ldw), act, _p(res), c_long(ldres), grouped[0], grouped[1], grouped[2],
                                       _stream()), what)


def _split_tf32(w2d):
    """(hi, lo) of a re-laid weight matrix for the 3xTF32 forward (hi = RN tf32, lo = w - hi)."""
    planes = torch.empty((2,) + tuple(w2d.shape), device=w2d.device, dtype=torch.float32)   # adjacent: one TMA box
    hi, lo = planes[0], planes[1]
    _check(lib().dfine_tf32_split(_p(w2d), _p(hi), _p(lo), c_long(w2d.numel()), _stream()), "tf32_split")
    return hi, lo


_F16_WSCALE = 256.0      # power of two applied to the weights of the 3xFP16 mode (see dfine_conv_tc_f16x3)


def _split_f16(w2d, taps, Cin):
    """fp16 (hi, lo) planes [2, rows, taps * Cin_p] of w * 2^8 for the 3xFP16 forward (tap runs padded to 8 channels)."""
    rows = w2d.shape[0]
    cin_p = (Cin + 7) // 8 * 8
    planes = torch.empty((2, rows, taps * cin_p), device=w2d.device, dtype=torch.float16)
    _check(lib().dfine_f16_split(_p(w2d), c_long(w2d.shape[1]), _p(planes), c_long(rows), taps, Cin, cin_p,
                                 c_float(_F16_WSCALE), _stream()), "f16_split")
    return planes


def _split_bf16(w2d, taps, Cin, mode=0):
    """bf16 (hi, lo) planes [2, rows, taps * Cin_p] of a re-laid weight matrix [rows, taps * Cin] for the 3xBF16
    forward; Cin_p = Cin rounded up to 8 elements (every tap starts on a 16-byte boundary for TMA; pads are zero)."""
    rows = w2d.shape[0]
    cin_p = (Cin + 7) // 8 * 8
    planes = torch.empty((2, rows, taps * cin_p), device=w2d.device, dtype=torch.bfloat16)
    _check(lib().dfine_bf16_split(_p(w2d), c_long(w2d.shape[1]), _p(planes), c_long(rows), taps, Cin, cin_p, mode,
                                  _stream()), "bf16_split")
    return planes


# ------------------------------------------------------------------------------------------------
# raw launchers (no autograd)
# ------------------------------------------------------------------------------------------------
def _conv_fwd(x, ldx, weight, wkey, bias, y, ldy, geom, act, stats=None, lab=None, ch_scale=None, bn_fin=None):
    """geom = (B,H,W,Cin,OH,OW,Cout,k,stride,pad4).  ``weight`` is the parameter ([Cout,Cin,k,k] conv or [N,K]
    linear); ``wkey`` its cache getter.  Returns True if the tensor-core kernel ran (it fuses the BN statistics).
    ``bn_fin`` (see _tc_launch) may only be passed when ``_bn_fin_ok`` holds for the layer."""
    B, H, W, Cin, OH, OW, Cout, k, stride, pad = geom
    assert lab is None or _MODE == "hf3", "the fused LAB epilogue exists on the 3xFP16 path only"
    assert ch_scale is None or (_MODE == "hf3" and _TC_TS), "the per-channel epilogue scale exists on the tensor-memory kernel only"
    if bias is None and act == 0 and lab is None and ch_scale is None and _conv2x2_ok(geom, ldx, ldy):
        # the stem's 2x2 convs: direct fp32 kernel (csrc/stem.cu), BN statistics fused
        wt = wkey("w2f", lambda: weight.permute(2, 3, 1, 0).reshape(4 * Cin, Cout).contiguous())
        _check(lib().dfine_conv2x2(_p(x), c_long(ldx), _p(wt), _p(y), c_long(ldy), _p(stats), B, H, W, Cin, Cout, 0, _stream()),
               "conv2x2")
        return True
    if _tc_ok(Cin, Cout, k, stride, pad, ldx, ldy):
        K = k * k * Cin
        wr = wkey("wr", lambda: weight.reshape(Cout, Cin, k, k).permute(0, 2, 3, 1).reshape(Cout, K).contiguous())
        w_hi, w_lo, planes, plane_stride = wr, None, None, 0
        if _MODE == "tc3":
            pl = _planes_of(weight)
            w_hi, w_lo = pl if pl is not None else wkey("wr3", lambda: _split_tf32(wr))
        elif _MODE == "bf3":
            planes = wkey("wrb", lambda: _split_bf16(wr, k * k, Cin))
            w_hi = None
        elif _MODE == "hf3":
            pl = _planes_of(weight) if Cin % 8 == 0 else None
            if pl is not None:       # fp16 planes kept current by the AdamW kernel (two arenas a fixed distance apart)
                planes, plane_stride = pl[0], (pl[1].data_ptr() - pl[0].data_ptr()) // 2
            else:
                planes = wkey("wrf", lambda: _split_f16(wr, k * k, Cin))
            w_hi = None
        elif _MODE == "tch":
            w_hi, _ = wkey("wr3", lambda: _split_tf32(wr))
            planes = wkey("wrh", lambda: _split_bf16(wr, k * k, Cin, mode=1))
        cs = (Cin + 7) // 8 * 8 if _MODE in ("bf3", "hf3") else Cin        # channel run of one tap in the weight matrix
        taps = _taps(("f", k, pad[0], pad[1], cs),
                     lambda: [(kh - pad[0], kw - pad[1], (kh * k + kw) * cs) for kh in range(k) for kw in range(k)])
        _tc_launch(x, ldx, H, W, Cin, w_hi, w_lo, K, bias, y, ldy, B, OH, OW, Cout, OH, OW, (1, 1), (0, 0), stride,
                   taps, act, stats, "conv_fwd_tc", planes, plane_stride=plane_stride, lab=lab, ch_scale=ch_scale,
                   bn_fin=bn_fin)
        return True
    assert lab is None and ch_scale is None and bn_fin is None, "fused LAB / scale / finalize epilogues need a tensor-core geometry"
    wr = wkey("wr", lambda: weight.reshape(Cout, Cin, k, k).permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous())
    if bias is None and act == 0 and _MODE != "simt" and \
            lib().dfine_stem_conv_supported(Cin, Cout, k, k, stride, pad[0], pad[1], pad[2], pad[3], c_long(ldy)):
        _check(lib().dfine_stem_conv_fwd(_p(x), c_long(ldx), _p(wr), _p(y), c_long(ldy), B, H, W, Cout, _stream()),
               "stem_conv_fwd")
        return False
    _check(lib().dfine_conv_fwd_simt(_p(x), _p(wr), _p(bias), _p(y), B, H, W, Cin, OH, OW, Cout, k, k, stride,
                                     pad[0], pad[1], c_long(ldx), c_long(ldy), act, _stream()), "conv_fwd_simt")
    return False


_BN_FIN = os.environ.get("DFINE_BN_FIN", "1") != "0"     # train-mode BatchNorm finalize in the conv kernel's tail
# ... and (opt-in, DFINE_BN_APPLY=1) the normalise / activation pass after a grid-wide wait.  Measured on B200 (profiles/
# README.md): 148 CTAs cannot match the bandwidth of the standalone full-occupancy bn_apply kernel — the D-FINE-m step is
# 2.3 ms slower with it on every layer and 0.1 ms slower restricted to conv outputs <= 32 MB — so it stays off.
_BN_APPLY = os.environ.get("DFINE_BN_APPLY", "0") == "1"
_BN_APPLY_MAX = int(float(os.environ.get("DFINE_BN_APPLY_MAX_MB", "1e9")) * (1 << 20))   # ... for conv outputs up to this size


def _bn_fin_ok(geom, ldx, ldy):
    """The layer's forward runs on tc_fwd_ts (3xFP16, activation planes in tensor memory), whose last CTA can finalize."""
    B, H, W, Cin, OH, OW, Cout, k, stride, pad = geom
    return _BN_FIN and _MODE == "hf3" and _TC_TS and not _conv2x2_ok(geom, ldx, ldy) and _tc_ok(Cin, Cout, k, stride, pad, ldx, ldy)


def _prefetch_dgrad_weight(weight, geom, ldx, ldy):
    """Forward-time, on the weight-gradient stream: the [Cin, taps*Cout] weight copy the data gradient of this layer
    will read.  It depends on the weights only, so it overlaps the forward pass instead of sitting in the backward
    chain (one small strided copy per layer and step: the parameters change every step)."""
    if not wgrad_stream.on or not torch.is_grad_enabled():
        return
    B, H, W, Cin, OH, OW, Cout, k, stride, pad = geom
    if not _tc_ok(Cout, Cin, k, stride, pad, ldy, ldx):
        return
    K = k * k * Cout
    wkey = _wcache.getter(weight)
    wgrad_stream.run(lambda: wkey("wd", lambda: weight.reshape(Cout, Cin, k, k).permute(1, 2, 3, 0).reshape(Cin, K).contiguous()))


def _conv_dgrad(dy, ldy, weight, wkey, dx, ldx, geom, res=None):
    """dx[B,H,W,Cin] (pixel stride ldx) from dy[B,OH,OW,Cout]: the forward kernel run on dy with transposed taps;
    a stride-2 conv's data gradient is one launch per output-pixel parity.  ``res`` ([B,H,W,Cin], possibly a
    channel slice with a larger pixel stride): added to dx — in the kernel epilogue on the stride-1 tensor-core path."""
    B, H, W, Cin, OH, OW, Cout, k, stride, pad = geom
    if _conv2x2_ok(geom, ldx, ldy) and lib().dfine_conv2x2_supported(Cout, Cin, 2, 2, 1, 0, 0, 1, 1, c_long(ldy), c_long(ldx)):
        # wt[(2i + j)][co][ci] = w[co][ci][1 - i][1 - j]: the same correlation on dy, zero border at the top / left
        wt = wkey("w2d", lambda: weight.flip(2, 3).permute(2, 3, 0, 1).reshape(4 * Cout, Cin).contiguous())
        _check(lib().dfine_conv2x2(_p(dy), c_long(ldy), _p(wt), _p(dx), c_long(ldx), None, B, H, W, Cout, Cin, -1, _stream()),
               "conv2x2(dgrad)")
        if res is not None:
            dx.add_(res)
        return
    if res is not None:
        ok = (res.dim() == 4 and res.stride(3) == 1 and res.stride(2) % 4 == 0 and res.stride(2) >= Cin
              and res.stride(1) == W * res.stride(2) and res.stride(0) == H * W * res.stride(2)
              and res.data_ptr() % 16 == 0)
        if not (ok and stride == 1 and _tc_ok(Cout, Cin, k, stride, pad, ldy, ldx)):
            _conv_dgrad(dy, ldy, weight, wkey, dx, ldx, geom)
            dx.add_(res)
            return
    if _tc_ok(Cout, Cin, k, stride, pad, ldy, ldx):
        K = k * k * Cout
        wd = wkey("wd", lambda: weight.reshape(Cout, Cin, k, k).permute(1, 2, 3, 0).reshape(Cin, K).contiguous())
        if stride == 1:
            taps = _taps(("d1", k, pad[0], pad[1], Cout),
                         lambda: [(pad[0] - kh, pad[1] - kw, (kh * k + kw) * Cout) for kh in range(k) for kw in range(k)])
            _tc_launch(dy, ldy, OH, OW, Cout, wd, None, K, None, dx, ldx, B, H, W, Cin, H, W, (1, 1), (0, 0), 1, taps,
                       0, None, "conv_dgrad_tc", None, res, res.stride(2) if res is not None else 0)
            return
        for ph in range(2):
            for pw in range(2):
                hp, wp = (H - ph + 1) // 2, (W - pw + 1) // 2
                if hp <= 0 or wp <= 0:
                    continue
                tl = [((ph + pad[0] - kh) // 2, (pw + pad[1] - kw) // 2, (kh * k + kw) * Cout)
                      for kh in range(k) for kw in range(k)
                      if (ph + pad[0] - kh) % 2 == 0 and (pw + pad[1] - kw) % 2 == 0]
                if not tl:
                    dx[:, ph::2, pw::2].zero_()
                    continue
                taps = _taps(("d2", k, pad[0], pad[1], Cout, ph, pw), lambda: tl)
                _tc_launch(dy, ldy, OH, OW, Cout, wd, None, K, None, dx, ldx, B, hp, wp, Cin, H, W, (2, 2), (ph, pw), 1,
                           taps, 0, None, "conv_dgrad_tc_s2")
        return
    wr = wkey("wr", lambda: weight.reshape(Cout, Cin, k, k).permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous())
    _check(lib().dfine_conv_dgrad_simt(_p(dy), _p(wr), _p(dx), B, H, W, Cin, OH, OW, Cout, k, k, stride, pad[0],
                                       pad[1], c_long(ldx), c_long(ldy), _stream()), "conv_dgrad_simt")


# Parameters whose ``.grad`` is a view of a flat-arena optimizer's gradient arena (optim.FusedAdamW registers them):
# id(param) -> weakref.  ONLY these take the direct-accumulation path below.  Any other parameter — a stock torch
# optimizer, a DDP wrap with gradient_as_bucket_view (whose reducer must see the AccumulateGrad hook fire),
# torch.autograd.grad, user hooks on weights — gets its gradient returned to autograd as usual.
_direct_grads = {}


def register_direct_grad(p):
    _direct_grads[id(p)] = weakref.ref(p)


def _grad_dst(p, kind):
    """The parameter's pre-zeroed ``.grad`` viewed in the layout a kernel accumulates into, or None.
    With the flat-arena optimizer (optim.FusedAdamW) gradients live in an arena zeroed by the optimizer step, so
    backward kernels add straight into it (no temporary, no AccumulateGrad add kernel); a parameter that is not
    registered by such an optimizer takes the returned-gradient path."""
    if p is None or not p.is_leaf or not p.requires_grad:
        return None
    ent = _direct_grads.get(id(p))
    if ent is None or ent() is not p:
        return None
    g = p.grad
    if g is None or not g.is_cuda or g.dtype != torch.float32:
        return None
    if kind == "conv":          # [Co,Ci,kh,kw] stored channels-last -> [Co,kh,kw,Ci] contiguous
        v = g.permute(0, 2, 3, 1)
        return v if v.is_contiguous() else None
    return g if g.is_contiguous() else None


def _conv_wgrad(dy, ldy, x, ldx, geom, dst=None):
    B, H, W, Cin, OH, OW, Cout, k, stride, pad = geom
    dwr = dst if dst is not None else torch.zeros((Cout, k, k, Cin), device=dy.device, dtype=torch.float32)
    if _tc_ok(Cin, Cout, k, stride, pad, ldx, ldy):
        nbytes = 4 * (B * H * W * Cin + B * OH * OW * Cout + 2 * Cout * k * k * Cin)
        with _timed("conv_wgrad_tc", nbytes, 2 * B * OH * OW * Cout * k * k * Cin,
                    f"wgrad {Cin}->{Cout} k{k} s{stride} {OH}x{OW} B{B}"):
            _check(lib().dfine_conv_wgrad_tc(_p(dy), _p(x), _p(dwr), B, H, W, Cin, OH, OW, Cout, k, k, stride, pad[0],
                                             pad[1], c_long(ldx), c_long(ldy), _stream()), "conv_wgrad_tc")
    elif _MODE != "simt" and lib().dfine_stem_conv_supported(Cin, Cout, k, k, stride, pad[0], pad[1], pad[2], pad[3],
                                                              c_long(ldy)):
        _check(lib().dfine_stem_conv_wgrad(_p(dy), c_long(ldy), _p(x), c_long(ldx), _p(dwr), B, H, W, Cout, _stream()),
               "stem_conv_wgrad")
    else:
        _check(lib().dfine_conv_wgrad_simt(_p(dy), _p(x), _p(dwr), B, H, W, Cin, OH, OW, Cout, k, k, stride, pad[0],
                                           pad[1], c_long(ldx), c_long(ldy), _stream()), "conv_wgrad_simt")
    return dwr


class _WCache:
    """Re-laid weight copies, keyed by parameter identity + version + the optimizer epoch (the flat-arena
    optimizer updates parameters through raw pointers, which does not bump ``_version``)."""

    def __init__(self):
        self.d = {}
        self.epoch = 0

    def get(self, w, kind, make):
        key = (id(w), kind)
        ent = self.d.get(key)
        ver = (w._version, self.epoch)
        if ent is not None and ent[0] == ver and ent[2]() is w:      # id() can be recycled: check identity
            return ent[1]
        with torch.no_grad():
            val = make()
        self.d[key] = (ver, val, weakref.ref(w))
        return val

    def getter(self, w):
        return lambda kind, make: self.get(w, kind, make)


_wcache = _WCache()


# tf32 hi / lo planes kept by the flat-arena optimizer (optim.FusedAdamW): id(param) -> (weakref, version, hi, lo).
# The AdamW kernel rewrites them with the parameters; a parameter modified any other way (``_version`` moved: a
# checkpoint load, an in-place init) is re-split on its next use.
_arena_planes = {}


def register_weight_planes(p, hi, lo):
    _arena_planes[id(p)] = (weakref.ref(p), p._version, hi, lo)


def _planes_of(weight):
    ent = _arena_planes.get(id(weight))
    if ent is None or ent[0]() is not weight:
        return None
    _, ver, hi, lo = ent
    if hi.dtype != (torch.float16 if _MODE == "hf3" else torch.float32):
        return None                  # planes of another operand mode (the mode was switched after the optimizer was built)
    if ver != weight._version:       # rewritten outside the optimizer kernel: refresh this weight's planes
        src = weight.detach()
        src = src.permute(0, 2, 3, 1) if src.dim() == 4 else src
        with torch.no_grad():
            flat = src.reshape(hi.shape)
            if flat.data_ptr() != weight.data_ptr():      # not the channels-last arena view any more
                return None
            if hi.dtype == torch.float16:
                if hi.numel() % 4:
                    return None
                _check(lib().dfine_f16_split_flat(_p(flat), _p(hi), _p(lo), c_long(hi.numel()), c_float(_F16_WSCALE),
                                                  _stream()), "f16_split_flat")
            else:
                _check(lib().dfine_tf32_split(_p(flat), _p(hi), _p(lo), c_long(hi.numel()), _stream()), "tf32_split")
        _arena_planes[id(weight)] = (ent[0], weight._version, hi, lo)
    return hi, lo


def weights_changed():
    """Called by the optimizer after it rewrote parameters in place."""
    _wcache.epoch += 1


# ------------------------------------------------------------------------------------------------
# conv + BatchNorm + act (+LAB, +adds)
# ------------------------------------------------------------------------------------------------
class _ConvBnAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bn_w, bn_b, lab_s, lab_b, pre_add, post_add, running_mean, running_var, cfg, out=None):
        stride, pad, groups, training, momentum, eps, act, frozen, tap = cfg
        _req_cuda(x, weight)
        x_in = x
        if x.stride(-1) != 1 or x.dim() != 4:
            x = x.contiguous()
        B, H, W, Cin = x.shape
        ok = x.stride(2) >= Cin and x.stride(1) == W * x.stride(2) and x.stride(0) == H * W * x.stride(2) \
            and x.stride(2) % 4 == 0 and x.data_ptr() % 16 == 0
        if not ok:
            x = x.contiguous()
        ldx = x.stride(2)
        Cout, _, k, _ = weight.shape
        pt, pl, pb, pr = pad
        OH = (H + pt + pb - k) // stride + 1
        OW = (W + pl + pr - k) // stride + 1
        geom = (B, H, W, Cin, OH, OW, Cout, k, stride, pad)
        dev = x.device
        M = B * OH * OW
        # BatchNorm with FIXED statistics (FrozenBatchNorm2d of the l / x backbones, eval mode) followed by ReLU / nothing:
        # folded into the conv epilogue, y = act(conv * scale[c] + shift[c]) — no separate pass, no pre-activation tensor;
        # the backward reads act' off the output (ReLU: y > 0).
        ctx.folded = (_FOLD_FROZEN and not training and groups == 1 and pre_add is None and post_add is None
                      and lab_s is None and act in (None, "relu") and _MODE == "hf3" and _TC_TS
                      and not ctx.needs_input_grad[2] and not ctx.needs_input_grad[3]
                      and _tc_ok(Cin, Cout, k, stride, pad, ldx, Cout) and not _conv2x2_ok(geom, ldx, Cout))
        if ctx.folded:
            scale, shift = _frozen_scale_shift(bn_w, bn_b, running_mean, running_var, eps)
            if out is not None:
                assert tuple(out.shape) == (B, OH, OW, Cout) and out.stride(3) == 1 and out.stride(2) % 4 == 0 and \
                    out.stride(1) == OW * out.stride(2) and out.stride(0) == OH * OW * out.stride(2) and out.data_ptr() % 16 == 0
                y, ldy_out = out.as_strided(out.shape, out.stride(), out.storage_offset()), out.stride(2)
            else:
                y, ldy_out = _alloc_nhwc(B, OH, OW, Cout, dev)
            _conv_fwd(x, ldx, weight, _wcache.getter(weight), shift, y, ldy_out, geom, ACT[act], None, ch_scale=scale)
            if ctx.needs_input_grad[0]:
                _prefetch_dgrad_weight(weight, geom, ldx, Cout)
            ctx.save_for_backward(x, weight, y, scale)
            ctx.geom, ctx.ldx, ctx.cfg, ctx.ldy = geom, ldx, cfg, ldy_out
            ctx.has_post = False
            return (y, x_in) if tap else y
        conv_out = torch.empty((B, OH, OW, Cout), device=dev, dtype=torch.float32)
        need_stats = training
        depthwise = groups > 1
        # train-mode finalize (mean / invstd / scale / shift / running statistics) in the conv kernel's last CTA: the
        # statistics buffer carries one extra zeroed slot, the retirement ticket
        fin = need_stats and not depthwise and _bn_fin_ok(geom, ldx, Cout)
        stats = zero_pool.take(2 * Cout + (1 if fin else 0), dev) if need_stats else None   # (+1: ticket and flag, 2 x u32)
        if depthwise:
            assert groups == Cin == Cout and pt == pl == pb == pr, "only depthwise grouped convs are on the path"
            if ldx != Cin:
                x = x.contiguous()
                ldx = Cin
            wt = _wcache.get(weight, "dw", lambda: weight.reshape(Cout, k * k).t().contiguous())
            _check(lib().dfine_dwconv_fwd(_p(x), _p(wt), _p(conv_out), B, H, W, Cin, k, stride, pt, _stream()),
                   "dwconv_fwd")
            fused_stats = False
        scale = torch.empty(Cout, device=dev, dtype=torch.float32)
        shift = torch.empty(Cout, device=dev, dtype=torch.float32)
        mean = invstd = None
        if pre_add is not None:
            pre_add = pre_add.contiguous()
        ld_post = Cout
        if post_add is not None:      # a channel slice of a concatenation buffer is read in place through its row stride
            if (post_add.dim() == 4 and post_add.stride(3) == 1 and post_add.stride(2) % 4 == 0 and post_add.stride(2) >= Cout
                    and post_add.stride(1) == OW * post_add.stride(2) and post_add.stride(0) == OH * OW * post_add.stride(2)
                    and post_add.data_ptr() % 16 == 0):
                ld_post = post_add.stride(2)
            else:
                post_add = post_add.contiguous()
        if out is not None:
            # the caller's channel slice of a concatenation buffer (concat by channel slice: no copy kernel later); written
            # through the raw pointer and returned as an alias, so autograd sees no in-place operation
            assert tuple(out.shape) == (B, OH, OW, Cout) and out.stride(3) == 1 and out.stride(2) % 4 == 0 and \
                out.stride(1) == OW * out.stride(2) and out.stride(0) == OH * OW * out.stride(2) and out.data_ptr() % 16 == 0
            y, ldy_out = out.as_strided(out.shape, out.stride(), out.storage_offset()), out.stride(2)
        else:
            y, ldy_out = _alloc_nhwc(B, OH, OW, Cout, dev)
        # ... and the normalise / activation / LAB / residual pass too: after a grid-wide wait for the finalize every CTA
        # of the conv kernel normalises the tiles it wrote (no bn_apply launch; its reads are L2 hits)
        fin_apply = fin and _BN_APPLY and pre_add is None and 4 * M * Cout <= _BN_APPLY_MAX
        if fin:
            mean = torch.empty(Cout, device=dev, dtype=torch.float32)
            invstd = torch.empty(Cout, device=dev, dtype=torch.float32)
            _conv_fwd(x, ldx, weight, _wcache.getter(weight), None, conv_out, Cout, geom, 0, stats,
                      bn_fin=(stats[2 * Cout:], bn_w, bn_b, running_mean, running_var, mean, invstd, scale, shift, momentum, eps)
                      + ((y, ldy_out, post_add, ld_post, lab_s, lab_b, ACT[act]) if fin_apply else ()))
            fused_stats = True
            if ctx.needs_input_grad[0]:
                _prefetch_dgrad_weight(weight, geom, ldx, Cout)
        elif not depthwise:
            fused_stats = _conv_fwd(x, ldx, weight, _wcache.getter(weight), None, conv_out, Cout, geom, 0, stats)
            if ctx.needs_input_grad[0]:
                _prefetch_dgrad_weight(weight, geom, ldx, Cout)
        if need_stats and not fused_stats:
            _check(lib().dfine_bn_stats(_p(conv_out), _p(stats), c_long(M), Cout, _stream()), "bn_stats")
        # (a fused finalize+apply launch was measured SLOWER: every CTA re-derives the scale / shift table in fp64 and
        #  synchronises before its first load — 20.5 us against 12.3 + 4.8 us per layer, profiles/README.md)
        if fin:
            pass
        elif training:
            mean = torch.empty(Cout, device=dev, dtype=torch.float32)
            invstd = torch.empty(Cout, device=dev, dtype=torch.float32)
            _check(lib().dfine_bn_finalize(_p(stats), _p(bn_w), _p(bn_b), _p(running_mean), _p(running_var), _p(mean),
                                           _p(invstd), _p(scale), _p(shift), c_long(M), Cout, c_float(momentum),
                                           c_float(eps), _stream()), "bn_finalize")
        else:
            _check(lib().dfine_bn_fold(_p(bn_w), _p(bn_b), _p(running_mean), _p(running_var), _p(scale), _p(shift),
                                       Cout, c_float(eps), _stream()), "bn_fold")
        if not fin_apply:
            with _timed("bn_apply", 4 * M * Cout * (2 + (pre_add is not None) + (post_add is not None)), 0, f"bn_apply M{M} C{Cout}"):
                _check(lib().dfine_bn_apply(_p(conv_out), _p(scale), _p(shift), _p(pre_add), _p(post_add), _p(lab_s),
                                            _p(lab_b), _p(y), c_long(M), Cout, ACT[act], c_long(ldy_out), c_long(ld_post), _stream()),
                       "bn_apply")
        ctx.save_for_backward(x, weight, conv_out, scale, shift, mean, invstd, pre_add, lab_s, lab_b, bn_w)
        ctx.geom, ctx.ldx, ctx.cfg = geom, ldx, cfg
        ctx.has_post = post_add is not None
        ctx.bn_b_ref = bn_b
        if tap:
            # second output = the input itself: a consumer that reads x besides this conv (the channel concat of an
            # HG block) takes this alias instead, so its gradient arrives HERE (dtap) and is added to the data
            # gradient in the dgrad kernel's epilogue — no separate gradient-accumulation kernel
            return y, x_in
        return y

    @staticmethod
    def backward(ctx, dy, dtap=None):
        if ctx.folded:
            return _ConvBnAct._backward_folded(ctx, dy, dtap)
        x, weight, conv_out, scale, shift, mean, invstd, pre_add, lab_s, lab_b, bn_w = ctx.saved_tensors
        stride, pad, groups, training, momentum, eps, act, frozen, tap = ctx.cfg
        B, H, W, Cin, OH, OW, Cout, k, _, _ = ctx.geom
        # dy is often a channel slice of a concatenation's gradient: consumed in place through its row stride
        ld_dy = dy.stride(2) if dy.dim() == 4 else 0
        if not (dy.dim() == 4 and dy.stride(3) == 1 and ld_dy >= Cout and ld_dy % 4 == 0 and dy.stride(1) == OW * ld_dy
                and dy.stride(0) == OH * OW * ld_dy and dy.data_ptr() % 16 == 0):
            dy = dy.contiguous()
            ld_dy = Cout
        M = B * OH * OW
        dev = dy.device
        bn_train = training and mean is not None
        need_red = bn_train or lab_s is not None
        red = zero_pool.take(2 * Cout + 2, dev) if need_red else None
        if need_red:
            # (eval-mode BN with LAB still needs the LAB scalar gradients; mean/invstd unused then)
            m_ = mean if mean is not None else shift
            i_ = invstd if invstd is not None else scale
            with _timed("bn_bwd_reduce", 4 * M * Cout * (2 + (pre_add is not None)), 0, f"bn_bwd_reduce M{M} C{Cout} ld{ld_dy}"):
                _check(lib().dfine_bn_bwd_reduce(_p(dy), _p(conv_out), _p(scale), _p(shift), _p(m_), _p(i_), _p(pre_add),
                                                 _p(lab_s), _p(red), c_long(M), Cout, ACT[act], c_long(ld_dy), _stream()),
                       "bn_bwd_reduce")
        dconv, ld_dc = (torch.empty_like(conv_out), Cout) if groups > 1 else _alloc_nhwc(B, OH, OW, Cout, dev)
        dpre = torch.empty_like(conv_out) if (pre_add is not None and ctx.needs_input_grad[6]) else None
        # parameter gradients of BN / LAB: accumulated by the apply kernel straight into the .grad arenas
        bn_direct = lab_direct = None
        if bn_train and bn_w is not None and ctx.needs_input_grad[2]:
            gw_, gb_ = _grad_dst(bn_w, "flat"), _grad_dst(ctx.bn_b_ref, "flat")
            if gw_ is not None and gb_ is not None:
                bn_direct = (gw_, gb_)
        if lab_s is not None:
            gs_, gl_ = _grad_dst(lab_s, "flat"), _grad_dst(lab_b, "flat")
            if gs_ is not None and gl_ is not None:
                lab_direct = (gs_, gl_)
        with _timed("bn_bwd_apply", 4 * M * Cout * (3 + (pre_add is not None) + (dpre is not None)), 0,
                    f"bn_bwd_apply M{M} C{Cout} ld{ld_dy}"):
            _check(lib().dfine_bn_bwd_apply(_p(dy), _p(conv_out), _p(scale), _p(shift), _p(mean), _p(invstd), _p(pre_add),
                                            _p(lab_s), _p(red), _p(dconv), _p(dpre), c_long(M), Cout, ACT[act],
                                            1 if bn_train else 0, _p(bn_direct[0]) if bn_direct else None,
                                            _p(bn_direct[1]) if bn_direct else None,
                                            _p(lab_direct[0]) if lab_direct else None,
                                            _p(lab_direct[1]) if lab_direct else None, c_long(ld_dy), c_long(ld_dc),
                                            _stream()), "bn_bwd_apply")
        g_bn_w = g_bn_b = g_lab_s = g_lab_b = None
        if red is not None:
            redf = None
            if bn_train and bn_w is not None and ctx.needs_input_grad[2] and bn_direct is None:
                redf = red.float()
                g_bn_w, g_bn_b = redf[Cout:2 * Cout], redf[:Cout]
            elif not bn_train and bn_w is not None and ctx.needs_input_grad[2]:
                raise RuntimeError("eval-mode BatchNorm affine gradients are not on the training path")
            if lab_s is not None and lab_direct is None:
                redf = red.float() if redf is None else redf
                g_lab_s, g_lab_b = redf[2 * Cout:2 * Cout + 1], redf[2 * Cout + 1:2 * Cout + 2]
        g_x = g_w = None
        ldx = ctx.ldx
        if groups > 1:
            wt = _wcache.get(weight, "dw", lambda: weight.reshape(Cout, k * k).t().contiguous())
            if ctx.needs_input_grad[0]:
                g_x = torch.empty((B, H, W, Cin), device=dev, dtype=torch.float32)
                _check(lib().dfine_dwconv_bwd_data(_p(dconv), _p(wt), _p(g_x), B, H, W, Cin, k, stride, pad[0],
                                                   _stream()), "dwconv_bwd_data")
                if dtap is not None:
                    g_x.add_(dtap)
            if ctx.needs_input_grad[1]:
                dwt = torch.zeros((k * k, Cout), device=dev, dtype=torch.float32)
                _check(lib().dfine_dwconv_bwd_weight(_p(dconv), _p(x), _p(dwt), B, H, W, Cin, k, stride, pad[0],
                                                     _stream()), "dwconv_bwd_weight")
                g_w = dwt.t().reshape(Cout, 1, k, k)
        else:
            if ctx.needs_input_grad[1]:       # forked first: it then overlaps this layer's data gradient too
                dst = _grad_dst(weight, "conv")
                if dst is not None:      # accumulated in place on the weight-gradient stream; autograd gets None
                    geom_ = ctx.geom
                    wgrad_stream.run(lambda: _conv_wgrad(dconv, ld_dc, x, ldx, geom_, dst), dconv, x)
                else:
                    g_w = _conv_wgrad(dconv, ld_dc, x, ldx, ctx.geom).permute(0, 3, 1, 2)
            if ctx.needs_input_grad[0]:
                g_x = torch.empty((B, H, W, Cin), device=dev, dtype=torch.float32)
                _conv_dgrad(dconv, ld_dc, weight, _wcache.getter(weight), g_x, Cin, ctx.geom, dtap)
        g_post = (dy if dy.is_contiguous() else dy.contiguous()) if ctx.has_post else None
        return g_x, g_w, g_bn_w, g_bn_b, g_lab_s, g_lab_b, dpre, g_post, None, None, None, None

    @staticmethod
    def _backward_folded(ctx, dy, dtap):
        """Backward of the folded (fixed-statistics BatchNorm + ReLU in the conv epilogue) forward."""
        x, weight, y, scale = ctx.saved_tensors
        stride, pad, groups, training, momentum, eps, act, frozen, tap = ctx.cfg
        B, H, W, Cin, OH, OW, Cout, k, _, _ = ctx.geom
        ld_dy = dy.stride(2) if dy.dim() == 4 else 0
        if not (dy.dim() == 4 and dy.stride(3) == 1 and ld_dy >= Cout and ld_dy % 4 == 0 and dy.stride(1) == OW * ld_dy
                and dy.stride(0) == OH * OW * ld_dy and dy.data_ptr() % 16 == 0):
            dy = dy.contiguous()
            ld_dy = Cout
        dev = dy.device
        dconv, ld_dc = _alloc_nhwc(B, OH, OW, Cout, dev)
        _check(lib().dfine_frozen_bn_bwd(_p(dy), _p(y), _p(scale), _p(dconv), c_long(B * OH * OW), Cout, ACT[act],
                                         c_long(ld_dy), c_long(ctx.ldy), c_long(ld_dc), _stream()), "frozen_bn_bwd")
        g_x = g_w = None
        ldx = ctx.ldx
        if ctx.needs_input_grad[1]:
            dst = _grad_dst(weight, "conv")
            if dst is not None:
                geom_ = ctx.geom
                wgrad_stream.run(lambda: _conv_wgrad(dconv, ld_dc, x, ldx, geom_, dst), dconv, x)
            else:
                g_w = _conv_wgrad(dconv, ld_dc, x, ldx, ctx.geom).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[0]:
            g_x = torch.empty((B, H, W, Cin), device=dev, dtype=torch.float32)
            _conv_dgrad(dconv, ld_dc, weight, _wcache.getter(weight), g_x, Cin, ctx.geom, dtap)
        return g_x, g_w, None, None, None, None, None, None, None, None, None, None


def _frozen_scale_shift(bn_w, bn_b, running_mean, running_var, eps):
    """scale / shift of a BatchNorm evaluated with its running statistics, cached while none of the four tensors changes
    (FrozenBatchNorm2d never does: the 130 per-step fold launches of the D-FINE-x backbone run once)."""
    key = (id(bn_w), id(running_mean))
    ver = (bn_w._version, bn_b._version, running_mean._version, running_var._version, _wcache.epoch if bn_w.requires_grad else -1)
    hit = _fold_cache.get(key)
    if hit is not None and hit[0] == ver and hit[3]() is bn_w:
        return hit[1], hit[2]
    C = bn_w.numel()
    scale = torch.empty(C, device=bn_w.device, dtype=torch.float32)
    shift = torch.empty(C, device=bn_w.device, dtype=torch.float32)
    _check(lib().dfine_bn_fold(_p(bn_w), _p(bn_b), _p(running_mean), _p(running_var), _p(scale), _p(shift), C, c_float(eps),
                               _stream()), "bn_fold")
    if len(_fold_cache) > 4096:
        _fold_cache.clear()
    _fold_cache[key] = (ver, scale, shift, weakref.ref(bn_w))
    return scale, shift


_fold_cache = {}


class _CatAlias(torch.autograd.Function):
    """The channel concatenation of tensors that were WRITTEN into consecutive channel slices of `buf` by their
    producers: forward returns the buffer itself (no copy kernel), backward hands every producer its slice of the
    gradient as a view."""

    @staticmethod
    def forward(ctx, buf, *parts):
        ctx.splits = [p.shape[-1] for p in parts]
        return buf.as_strided(buf.shape, buf.stride(), buf.storage_offset())

    @staticmethod
    def backward(ctx, g):
        outs, o = [], 0
        for n in ctx.splits:
            outs.append(g[..., o:o + n])
            o += n
        return (None, *outs)


class _CopyInto(torch.autograd.Function):
    """x copied into a channel slice of a concatenation buffer (for a member whose producer could not write there)."""

    @staticmethod
    def forward(ctx, x, out):
        out.data.copy_(x)
        return out.as_strided(out.shape, out.stride(), out.storage_offset())

    @staticmethod
    def backward(ctx, g):
        return g, None


# ------------------------------------------------------------------------------------------------
# plain convolution (no norm): the MaskDecoder's lateral / fusion / up convs, which are followed by GroupNorm
# ------------------------------------------------------------------------------------------------
class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, pad):
        _req_cuda(x, weight)
        if x.stride(-1) != 1 or x.dim() != 4:
            x = x.contiguous()
        B, H, W, Cin = x.shape
        ok = x.stride(2) >= Cin and x.stride(1) == W * x.stride(2) and x.stride(0) == H * W * x.stride(2) \
            and x.stride(2) % 4 == 0 and x.data_ptr() % 16 == 0
        if not ok:
            x = x.contiguous()
        ldx = x.stride(2)
        Cout, _, k, _ = weight.shape
        pt, pl, pb, pr = pad
        OH = (H + pt + pb - k) // stride + 1
        OW = (W + pl + pr - k) // stride + 1
        geom = (B, H, W, Cin, OH, OW, Cout, k, stride, tuple(pad))
        y = torch.empty((B, OH, OW, Cout), device=x.device, dtype=torch.float32)
        _conv_fwd(x, ldx, weight, _wcache.getter(weight), None, y, Cout, geom, 0, None)
        if ctx.needs_input_grad[0]:
            _prefetch_dgrad_weight(weight, geom, ldx, Cout)
        ctx.save_for_backward(x, weight)
        ctx.geom, ctx.ldx = geom, ldx
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        B, H, W, Cin, OH, OW, Cout, k, stride, pad = ctx.geom
        dy = dy.contiguous()
        geom, ldx = ctx.geom, ctx.ldx
        g_x = g_w = None
        if ctx.needs_input_grad[1]:
            dst = _grad_dst(weight, "conv")
            if dst is not None:
                wgrad_stream.run(lambda: _conv_wgrad(dy, Cout, x, ldx, geom, dst), dy, x)
            else:
                g_w = _conv_wgrad(dy, Cout, x, ldx, geom).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[0]:
            g_x = torch.empty((B, H, W, Cin), device=dy.device, dtype=torch.float32)
            _conv_dgrad(dy, Cout, weight, _wcache.getter(weight), g_x, Cin, geom)
        return g_x, g_w, None, None


# ------------------------------------------------------------------------------------------------
# linear (+bias, +act)
# ------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        _req_cuda(x, w)
        N, Kd = w.shape
        x2, M, ldx = _rows(x)
        out = torch.empty(x.shape[:-1] + (N,), device=x.device, dtype=torch.float32)
        geom = (1, 1, M, Kd, 1, M, N, 1, 1, (0, 0, 0, 0))
        fused_act = act if act in (None, "relu") else None
        _conv_fwd(x2, ldx, w, _wcache.getter(w), b, out, N, geom, ACT[fused_act])
        if ctx.needs_input_grad[0]:
            _prefetch_dgrad_weight(w, geom, ldx, N)
        saved_z = None
        if act is not None and fused_act is None:
            saved_z = out
            out = torch.empty_like(saved_z)
            _check(lib().dfine_act_fwd(_p(saved_z), _p(out), c_long(out.numel()), ACT[act], _stream()), "act_fwd")
        ctx.save_for_backward(x2, w, saved_z if saved_z is not None else (out if act == "relu" else None))
        ctx.meta = (M, ldx, geom, act, b is not None, x.shape)
        ctx.bias_ref = b
        return out

    @staticmethod
    def backward(ctx, dy):
        x2, w, z = ctx.saved_tensors
        M, ldx, geom, act, has_bias, xshape = ctx.meta
        N, Kd = w.shape
        dy = dy.contiguous()
        if act is not None:
            dz = torch.empty_like(dy)
            _check(lib().dfine_act_bwd(_p(dy), _p(z), _p(dz), c_long(dy.numel()), ACT[act], _stream()), "act_bwd")
            dy = dz
        g_x = g_w = g_b = None
        if ctx.needs_input_grad[1]:
            dst = _grad_dst(w, "flat")
            if dst is not None:
                dst4 = dst.view(N, 1, 1, Kd)
                wgrad_stream.run(lambda: _conv_wgrad(dy, N, x2, ldx, geom, dst4), dy, x2)
            else:
                g_w = _conv_wgrad(dy, N, x2, ldx, geom).reshape(N, Kd)
        if has_bias and ctx.needs_input_grad[2]:
            b_ = ctx.bias_ref
            dst = _grad_dst(b_, "flat") if b_ is not None else None
            if dst is None:
                g_b = dst = torch.zeros(N, device=dy.device, dtype=torch.float32)
                _check(lib().dfine_colsum(_p(dy), _p(dst), c_long(M), N, c_long(N), _stream()), "colsum")
            else:
                dstb = dst
                wgrad_stream.run(lambda: _check(lib().dfine_colsum(_p(dy), _p(dstb), c_long(M), N, c_long(N), _stream()),
                                                "colsum"), dy)
        if ctx.needs_input_grad[0]:
            g_x = torch.empty(xshape, device=dy.device, dtype=torch.float32)
            _conv_dgrad(dy, N, w, _wcache.getter(w), g_x, Kd, geom)
        return g_x, g_w, g_b, None


# ------------------------------------------------------------------------------------------------
# layernorm (+ fused residual)
# ------------------------------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, w, b, eps):
        _req_cuda(x, w)
        x = x.contiguous()
        res = res.contiguous() if res is not None else None
        D = x.shape[-1]
        rows = x.numel() // D
        y = torch.empty_like(x)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        _check(lib().dfine_layernorm_fwd(_p(x), _p(res), _p(w), _p(b), _p(y), _p(mean), _p(rstd), c_long(rows), D,
                                         c_float(eps), _stream()), "layernorm_fwd")
        ctx.save_for_backward(x, res, w, mean, rstd)
        ctx.b_ref = b
        return y

    @staticmethod
    def backward(ctx, dy):
        x, res, w, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        D = x.shape[-1]
        rows = x.numel() // D
        dx = torch.empty_like(x)
        gw_, gb_ = _grad_dst(w, "flat"), _grad_dst(ctx.b_ref, "flat")
        direct = gw_ is not None and gb_ is not None
        if not direct:
            dwb = torch.zeros(2, D, device=x.device, dtype=torch.float32)
            gw_, gb_ = dwb[0], dwb[1]
        _check(lib().dfine_layernorm_bwd(_p(dy), _p(x), _p(res), _p(w), _p(mean), _p(rstd), _p(dx), _p(gw_),
                                         _p(gb_), c_long(rows), D, _stream()), "layernorm_bwd")
        return dx, (dx if res is not None else None), None if direct else gw_, None if direct else gb_, None


# ------------------------------------------------------------------------------------------------
# attention core
# ------------------------------------------------------------------------------------------------
class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qk, v, heads, mask):
        _req_cuda(qk, v)
        qk, v = qk.contiguous(), v.contiguous()
        B, S, D = v.shape
        hd = D // heads
        o = torch.empty_like(v)
        lse = torch.empty((B, heads, S), device=v.device, dtype=torch.float32)
        m8 = None
        if mask is not None:
            m8 = mask.to(torch.uint8).contiguous()
        scale = hd ** -0.5
        q, k = qk, qk[..., D:]
        _check(lib().dfine_attn_fwd(_p(q), c_long(2 * D), c_void_p(qk.data_ptr() + 4 * D), c_long(2 * D), _p(v),
                                    c_long(D), _p(m8), _p(o), c_long(D), _p(lse), B, S, heads, hd, c_float(scale),
                                    _stream()), "attn_fwd")
        ctx.save_for_backward(qk, v, o, lse, m8)
        ctx.meta = (B, S, D, heads, hd, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        qk, v, o, lse, m8 = ctx.saved_tensors
        B, S, D, heads, hd, scale = ctx.meta
        do = do.contiguous()
        dqk = torch.empty_like(qk)
        dv = torch.empty_like(v)
        dsum = torch.empty_like(lse)
        _check(lib().dfine_attn_bwd(_p(qk), c_long(2 * D), c_void_p(qk.data_ptr() + 4 * D), c_long(2 * D), _p(v),
                                    c_long(D), _p(m8), _p(o), c_long(D), _p(do), c_long(D), _p(lse), _p(dsum),
                                    _p(dqk), c_long(2 * D), c_void_p(dqk.data_ptr() + 4 * D), c_long(2 * D), _p(dv),
                                    c_long(D), B, S, heads, hd, c_float(scale), _stream()), "attn_bwd")
        return dqk, dv, None, None


# ------------------------------------------------------------------------------------------------
# multi-scale deformable attention
# ------------------------------------------------------------------------------------------------
# d(memory) of the decoder's deformable-attention layers.  Every layer gathers from the SAME memory tensor, so autograd
# would receive one zero-filled [B, L, 256] gradient per layer and add them (4 fills + 3 adds of 137 MB each at batch
# 16).  The layers scatter with red.add anyway: they share ONE buffer per memory tensor — the first backward to run
# zero-fills it, the others accumulate, and only the last one hands it to autograd (the others return None).
_msda_shared = {}


class _Msda(torch.autograd.Function):
    @staticmethod
    def forward(ctx, memory, proj, ref, pscale, shapes, points, heads, n_off, offset_scale):
        _req_cuda(memory, proj, ref)
        ctx.share_key = None
        if memory.requires_grad and torch.is_grad_enabled() and wgrad_stream.in_step:   # inside a train step only
            ctx.share_key = (memory.data_ptr(), memory._version, tuple(memory.shape))
            _msda_shared.setdefault(ctx.share_key, [0, None])[0] += 1
        memory, proj = memory.contiguous(), proj.contiguous()
        ref = ref.detach().contiguous().float()
        B, L, D = memory.shape
        Q = proj.shape[1]
        hd = D // heads
        hw = (c_int * (2 * len(shapes)))(*[int(v) for s in shapes for v in s])
        pts = (c_int * len(points))(*[int(p) for p in points])
        ld = proj.shape[-1]
        out = torch.empty((B, Q, D), device=memory.device, dtype=torch.float32)
        # algorithmic bytes: value + offsets + logits + ref read, out written
        nbytes = 4 * (memory.numel() + proj.numel() + ref.numel() + out.numel())
        with _timed("msda_fwd", nbytes):
          _check(lib().dfine_msda_fwd(_p(memory), _p(proj), c_long(ld), c_void_p(proj.data_ptr() + 4 * n_off),
                                    c_long(ld), _p(ref), _p(pscale), _p(out), B, Q, L, heads, hd, len(shapes), hw,
                                    pts, c_float(offset_scale), _stream()), "msda_fwd")
        ctx.save_for_backward(memory, proj, ref, pscale)
        ctx.meta = (B, Q, L, D, heads, hd, [tuple(s) for s in shapes], list(points), n_off, offset_scale)
        return out

    @staticmethod
    def backward(ctx, gout):
        memory, proj, ref, pscale = ctx.saved_tensors
        B, Q, L, D, heads, hd, shapes, points, n_off, offset_scale = ctx.meta
        gout = gout.contiguous()
        hw = (c_int * (2 * len(shapes)))(*[int(v) for s in shapes for v in s])
        pts = (c_int * len(points))(*[int(p) for p in points])
        ld = proj.shape[-1]
        ent = _msda_shared.get(ctx.share_key) if ctx.share_key is not None else None
        if ent is not None:
            if ent[1] is None:
                ent[1] = torch.zeros_like(memory)
            gmem = ent[1]
            ent[0] -= 1
            last = ent[0] == 0
            if last:
                del _msda_shared[ctx.share_key]
        else:
            gmem, last = torch.zeros_like(memory), True
        gproj = torch.empty_like(proj)
        # value/offsets/logits/ref/gout read, gvalue read-modify-write (2x), goff/glogit written
        nbytes = 4 * (memory.numel() + proj.numel() + ref.numel() + gout.numel() + 2 * gmem.numel() + gproj.numel())
        with _timed("msda_bwd", nbytes):
          _check(lib().dfine_msda_bwd(_p(memory), _p(proj), c_long(ld), c_void_p(proj.data_ptr() + 4 * n_off),
                                    c_long(ld), _p(ref), _p(pscale), _p(gout), _p(gmem), _p(gproj), c_long(ld),
                                    c_void_p(gproj.data_ptr() + 4 * n_off), c_long(ld), B, Q, L, heads, hd,
                                    len(shapes), hw, pts, c_float(offset_scale), _stream()), "msda_bwd")
        return (gmem if last else None), gproj, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# FDR head: Integral + distance2bbox + LQE statistics in one pass over pred_corners
# ------------------------------------------------------------------------------------------------
class _FdrHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, corners, ref, project, reg_scale, k, want_box, want_stat):
        _req_cuda(corners, ref, project, reg_scale)
        corners = corners.contiguous()
        NB = project.shape[0]
        rows = corners.numel() // (4 * NB)
        ref = ref.detach().contiguous().float()
        project = project.detach().contiguous().float()
        rs = reg_scale.detach().reshape(-1).float().contiguous()      # device scalar: no host read, graph-capturable
        lead = tuple(corners.shape[:-1])
        box = torch.empty(lead + (4,), device=corners.device, dtype=torch.float32) if want_box else None
        stat = torch.empty(lead + (4 * (k + 1),), device=corners.device, dtype=torch.float32) if want_stat else None
        _check(lib().dfine_fdr_head_fwd(_p(corners), _p(ref), _p(project), _p(rs), _p(box), _p(stat), c_long(rows), NB,
                                        k, _stream()), "fdr_head_fwd")
        ctx.save_for_backward(corners, ref, project, rs)
        ctx.meta = (rows, NB, k)
        return box, stat

    @staticmethod
    def backward(ctx, dbox, dstat):
        corners, ref, project, rs = ctx.saved_tensors
        rows, NB, k = ctx.meta
        dbox = dbox.contiguous() if dbox is not None else None
        dstat = dstat.contiguous() if dstat is not None else None
        dcorners = torch.empty_like(corners)
        _check(lib().dfine_fdr_head_bwd(_p(corners), _p(ref), _p(project), _p(rs), _p(dbox), _p(dstat), _p(dcorners),
                                        c_long(rows), NB, k, _stream()), "fdr_head_bwd")
        return dcorners, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# criterion: every loss head of the step (VFL, L1 + GIoU, FGL + DDF) in one autograd node over the stacked tensors
# ------------------------------------------------------------------------------------------------
class _Criterion(torch.autograd.Function):
    """Inputs: the UNSPLIT decoder stacks logits [L,B,n_dn+Q,C], boxes [L,B,n_dn+Q,4], corners [L,B,n_dn+Q,4*NB], the
    pre head's [B,n_dn+Q,*] and the encoder head's [B,Q,*] tensors, plus the criterion's plan table / normalisers.
    Outputs: the unweighted loss scalars as five small tensors (loss_desc.split_out).  Backward writes one gradient per
    input tensor (logits / corners completely, boxes on matched rows of a zero fill)."""

    @staticmethod
    def forward(ctx, logits, boxes, corners, pre_logits, pre_boxes, enc_logits, enc_boxes, ref0, table, counts, labels,
                tboxes, project, reg_scale, meta):
        from . import loss_desc as ld
        _req_cuda(logits, boxes, corners, table)
        t = dict(logits=logits.contiguous(), boxes=boxes.contiguous(), corners=corners.contiguous(),
                 ref0=ref0.detach().contiguous().float(), pre_logits=pre_logits.contiguous(),
                 pre_boxes=pre_boxes.contiguous(), enc_logits=enc_logits.contiguous(), enc_boxes=enc_boxes.contiguous(),
                 table=table.contiguous(), labels=labels.contiguous(), tboxes=tboxes.contiguous().float(),
                 counts=counts.contiguous().float(), project=project.detach().contiguous().float(),
                 reg_scale=reg_scale.detach().reshape(-1).float().contiguous())
        assert t["table"].dtype == torch.int64 and t["labels"].dtype == torch.int64
        d = ld.build(t, meta)
        L = d.L
        dev = logits.device
        ws = torch.empty(int(lib().dfine_loss_workspace_bytes(L, d.B, d.Q, d.n_dn)), dtype=torch.uint8, device=dev)
        out = torch.empty(ld.out_count(L), dtype=torch.float32, device=dev)
        ref = ctypes.byref(d)
        st = _stream()
        _check(lib().dfine_loss_prepare(ref, _p(ws), st), "loss_prepare")
        _check(lib().dfine_loss_vfl_fwd(ref, _p(ws), st), "loss_vfl_fwd")
        _check(lib().dfine_loss_box_fwd(ref, _p(ws), st), "loss_box_fwd")
        _check(lib().dfine_loss_fgl_ddf_fwd(ref, _p(ws), st), "loss_fgl_ddf_fwd")
        _check(lib().dfine_loss_finalize(ref, _p(ws), _p(out), st), "loss_finalize")
        ctx.t, ctx.meta, ctx.ws, ctx.out = t, dict(meta), ws, out
        return tuple(v.clone() for v in ld.split_out(out, L))

    @staticmethod
    def backward(ctx, g_vfl, g_l1, g_gi, g_fgl, g_ddf):
        from . import loss_desc as ld
        t, ws, out = ctx.t, ctx.ws, ctx.out
        d = ld.build(t, ctx.meta)
        L, H = d.L, d.L + 2
        dev = out.device
        parts = []
        for g, n in ((g_vfl, 2 * H), (g_l1, 2 * H), (g_gi, 2 * H), (g_fgl, 2 * L), (g_ddf, 2 * L)):
            parts.append(torch.zeros(n, device=dev) if g is None else g.reshape(-1).float())
        gout = torch.cat(parts).contiguous()
        dlogits, dpre_l, denc_l = torch.empty_like(t["logits"]), torch.empty_like(t["pre_logits"]), torch.empty_like(t["enc_logits"])
        dboxes, dpre_b, denc_b = torch.zeros_like(t["boxes"]), torch.zeros_like(t["pre_boxes"]), torch.zeros_like(t["enc_boxes"])
        dcorners = torch.empty_like(t["corners"])
        ref = ctypes.byref(d)
        st = _stream()
        _check(lib().dfine_loss_vfl_bwd(ref, _p(ws), _p(gout), _p(dlogits), _p(dpre_l), _p(denc_l), st), "loss_vfl_bwd")
        _check(lib().dfine_loss_box_bwd(ref, _p(ws), _p(gout), _p(dboxes), _p(dpre_b), _p(denc_b), st), "loss_box_bwd")
        _check(lib().dfine_loss_fgl_ddf_bwd(ref, _p(ws), _p(out), _p(gout), _p(dcorners), st), "loss_fgl_ddf_bwd")
        return (dlogits, dboxes, dcorners, dpre_l, dpre_b, denc_l, denc_b) + (None,) * 8


# ------------------------------------------------------------------------------------------------
# segmentation head: GroupNorm(+ReLU), bilinear resize, mask product
# ------------------------------------------------------------------------------------------------
class _GroupNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, groups, eps, relu):
        _req_cuda(x, w, b)
        x = x.contiguous()
        B, H, W, C = x.shape
        y = torch.empty_like(x)
        stats = zero_pool.take(2 * B * groups, x.device)
        mean = torch.empty(B * groups, device=x.device, dtype=torch.float32)
        rstd = torch.empty(B * groups, device=x.device, dtype=torch.float32)
        wc, bc = w.detach().contiguous(), b.detach().contiguous()
        _check(lib().dfine_groupnorm_fwd(_p(x), _p(wc), _p(bc), _p(y), _p(stats), _p(mean), _p(rstd), B, c_long(H * W), C,
                                         groups, c_float(eps), 1 if relu else 0, _stream()), "groupnorm_fwd")
        ctx.save_for_backward(x, w, b, mean, rstd)
        ctx.meta = (groups, relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b, mean, rstd = ctx.saved_tensors
        groups, relu = ctx.meta
        B, H, W, C = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        gw, gb = _grad_dst(w, "flat"), _grad_dst(b, "flat")
        direct = gw is not None and gb is not None
        if not direct:
            gw, gb = torch.zeros(C, device=x.device), torch.zeros(C, device=x.device)
        red = zero_pool.take(2 * B * groups, x.device)
        _check(lib().dfine_groupnorm_bwd(_p(dy), _p(x), _p(w.detach().contiguous()), _p(b.detach().contiguous()), _p(mean),
                                         _p(rstd), _p(dx), _p(gw), _p(gb), _p(red), B, c_long(H * W), C, groups,
                                         1 if relu else 0, _stream()), "groupnorm_bwd")
        return dx, None if direct else gw, None if direct else gb, None, None, None


class _ResizeBilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size):
        _req_cuda(x)
        x = x.contiguous()
        B, Hs, Ws, C = x.shape
        H, W = size
        y = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
        _check(lib().dfine_resize_bilinear_f32(_p(x), _p(y), B, Hs, Ws, H, W, C, _stream()), "resize_bilinear_f32")
        ctx.shape = (B, Hs, Ws, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, Hs, Ws, C, H, W = ctx.shape
        dy = dy.contiguous()
        dx = torch.empty((B, Hs, Ws, C), device=dy.device, dtype=torch.float32)
        _check(lib().dfine_resize_bilinear_bwd(_p(dy), _p(dx), B, Hs, Ws, H, W, C, _stream()), "resize_bilinear_bwd")
        return dx, None


class _MaskDot(torch.autograd.Function):
    """out[b, h, w, q] = sum_c feat[b, h, w, c] * embed[b, q, c]: one tcgen05 implicit-GEMM launch over all images with
    per-image weights (the embeddings of image b are rows [b*Q, (b+1)*Q) of one stacked weight matrix)."""

    @staticmethod
    def forward(ctx, embed, feat):
        _req_cuda(embed, feat)
        embed, feat = embed.contiguous(), feat.contiguous()
        B, Q, C = embed.shape
        _, H, W, _ = feat.shape
        assert Q % 4 == 0 and C % 8 == 0, "mask product: Q % 4 and C % 8"
        out = torch.empty((B, H, W, Q), device=feat.device, dtype=torch.float32)
        e2 = embed.reshape(B * Q, C)
        taps = _taps(("f", 1, 0, 0, C), lambda: [(0, 0, 0)])
        if _MODE == "hf3":
            planes = _split_f16(e2, 1, C)
            _tc_launch(feat, C, H, W, C, None, None, C, None, out, Q, B, H, W, Q, H, W, (1, 1), (0, 0), 1, taps, 0, None,
                       "mask_dot", planes, grouped=(B * Q, Q, 0))
        else:
            _tc_launch(feat, C, H, W, C, e2, None, C, None, out, Q, B, H, W, Q, H, W, (1, 1), (0, 0), 1, taps, 0, None,
                       "mask_dot", grouped=(B * Q, Q, 0))
        ctx.save_for_backward(embed, feat)
        return out

    @staticmethod
    def backward(ctx, dout):
        embed, feat = ctx.saved_tensors
        B, Q, C = embed.shape
        _, H, W, _ = feat.shape
        dout = dout.contiguous()
        dfeat = dembed = None
        if ctx.needs_input_grad[1]:
            # dfeat[b] = dout[b] [HW, Q] x embed[b] [Q, C]: the same kernel on dout with the stacked transposed embeddings
            # [C, B*Q]; image b reads weight columns [b*Q, (b+1)*Q)
            wd = embed.reshape(B * Q, C).t().contiguous()
            dfeat = torch.empty_like(feat)
            taps = _taps(("f", 1, 0, 0, Q), lambda: [(0, 0, 0)])
            _tc_launch(dout, Q, H, W, Q, wd, None, B * Q, None, dfeat, C, B, H, W, C, H, W, (1, 1), (0, 0), 1, taps, 0, None,
                       "mask_dot_dgrad", grouped=(C, 0, Q))
        if ctx.needs_input_grad[0]:
            dembed = torch.zeros((B, Q, 1, 1, C), device=feat.device, dtype=torch.float32)
            for b in range(B):
                _conv_wgrad(dout[b:b + 1], Q, feat[b:b + 1], C, (1, H, W, C, H, W, Q, 1, 1, (0, 0, 0, 0)), dembed[b])
            dembed = dembed.reshape(B, Q, C)
        return dembed, dfeat


class _MaskLossRows(torch.autograd.Function):
    """Cropped BCE + cropped Dice per matched mask (dfine_criterion.py:335-450): one kernel forward, one backward."""

    @staticmethod
    def forward(ctx, pred, gt, t_idx, tboxes):
        pred, gt = pred.contiguous(), gt.contiguous()
        t_idx, tboxes = t_idx.contiguous(), tboxes.contiguous().float()
        M, Hm, Wm = pred.shape
        bce = torch.empty(M, device=pred.device, dtype=torch.float32)
        dice = torch.empty(M, device=pred.device, dtype=torch.float32)
        sums = torch.empty((M, 4), device=pred.device, dtype=torch.float32)
        _check(lib().dfine_mask_loss_fwd(_p(pred), _p(gt), _p(t_idx), _p(tboxes), _p(bce), _p(dice), _p(sums), M, Hm, Wm,
                                         _stream()), "mask_loss_fwd")
        ctx.save_for_backward(pred, gt, t_idx, tboxes, sums)
        return bce, dice

    @staticmethod
    def backward(ctx, g_bce, g_dice):
        pred, gt, t_idx, tboxes, sums = ctx.saved_tensors
        M, Hm, Wm = pred.shape
        z = lambda g: torch.zeros(M, device=pred.device) if g is None else g.contiguous().float()   # noqa: E731
        dpred = torch.empty_like(pred)
        _check(lib().dfine_mask_loss_bwd(_p(pred), _p(gt), _p(t_idx), _p(tboxes), _p(sums), _p(z(g_bce)), _p(z(g_dice)),
                                         _p(dpred), M, Hm, Wm, _stream()), "mask_loss_bwd")
        return dpred, None, None, None


class _MaskLossPM(torch.autograd.Function):
    """The same two losses on PIXEL-MAJOR logits pred [B,Hm,Wm,R] (what the tcgen05 mask product writes); t_idx [B*R],
    negative on padding rows.  -> (bce_row [B*R], dice_row [B*R])."""

    @staticmethod
    def forward(ctx, pred, gt, t_idx, tboxes):
        pred, gt = pred.contiguous(), gt.contiguous()
        t_idx, tboxes = t_idx.contiguous(), tboxes.contiguous().float()
        B, Hm, Wm, R = pred.shape
        bce = torch.empty(B * R, device=pred.device, dtype=torch.float32)
        dice = torch.empty(B * R, device=pred.device, dtype=torch.float32)
        sums = torch.zeros((B * R, 4), device=pred.device, dtype=torch.float32)
        _check(lib().dfine_mask_loss_pm_fwd(_p(pred), _p(gt), _p(t_idx), _p(tboxes), _p(bce), _p(dice), _p(sums), B, R, Hm, Wm,
                                            _stream()), "mask_loss_pm_fwd")
        ctx.save_for_backward(pred, gt, t_idx, tboxes, sums)
        return bce, dice

    @staticmethod
    def backward(ctx, g_bce, g_dice):
        pred, gt, t_idx, tboxes, sums = ctx.saved_tensors
        B, Hm, Wm, R = pred.shape
        z = lambda g: torch.zeros(B * R, device=pred.device) if g is None else g.contiguous().float()   # noqa: E731
        dpred = torch.empty_like(pred)
        _check(lib().dfine_mask_loss_pm_bwd(_p(pred), _p(gt), _p(t_idx), _p(tboxes), _p(sums), _p(z(g_bce)), _p(z(g_dice)),
                                            _p(dpred), B, R, Hm, Wm, _stream()), "mask_loss_pm_bwd")
        return dpred, None, None, None


class _RowsTimesFeat(torch.autograd.Function):
    """out[m] = rows[m] . feat[b(m)] over the pixels, rows stacked image-major (`totals[b]` rows of image b): one product
    per image forward, two backward, ONE d(feat) tensor."""

    @staticmethod
    def forward(ctx, rows, feat, totals):
        B, Hm, Wm, C = feat.shape
        f2 = feat.reshape(B, Hm * Wm, C)
        out = torch.empty((rows.shape[0], Hm * Wm), device=feat.device, dtype=torch.float32)
        o = 0
        for b, n in enumerate(totals):
            if n:
                torch.matmul(rows[o:o + n], f2[b].t(), out=out[o:o + n])
                o += n
        ctx.save_for_backward(rows, feat)
        ctx.totals = totals
        return out

    @staticmethod
    def backward(ctx, dout):
        rows, feat = ctx.saved_tensors
        B, Hm, Wm, C = feat.shape
        f2 = feat.reshape(B, Hm * Wm, C)
        dout = dout.contiguous()
        drows = torch.empty_like(rows)
        dfeat = torch.empty_like(f2)
        o = 0
        for b, n in enumerate(ctx.totals):
            if n:
                torch.matmul(dout[o:o + n], f2[b], out=drows[o:o + n])
                torch.matmul(dout[o:o + n].t(), rows[o:o + n], out=dfeat[b])
                o += n
            else:
                dfeat[b].zero_()
        return drows, dfeat.reshape(feat.shape), None


# ------------------------------------------------------------------------------------------------
# decoder gate: sigmoid(g[:, :D]) * x1 + sigmoid(g[:, D:]) * x2
# ------------------------------------------------------------------------------------------------
class _GateMix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, x1, x2):
        _req_cuda(g, x1, x2)
        g, x1, x2 = g.contiguous(), x1.contiguous(), x2.contiguous()
        D = x1.shape[-1]
        rows = x1.numel() // D
        out = torch.empty_like(x1)
        _check(lib().dfine_gate_mix_fwd(_p(g), _p(x1), _p(x2), _p(out), c_long(rows), D, _stream()), "gate_mix_fwd")
        ctx.save_for_backward(g, x1, x2)
        return out

    @staticmethod
    def backward(ctx, dout):
        g, x1, x2 = ctx.saved_tensors
        D = x1.shape[-1]
        rows = x1.numel() // D
        dout = dout.contiguous()
        dg, dx1, dx2 = torch.empty_like(g), torch.empty_like(x1), torch.empty_like(x2)
        _check(lib().dfine_gate_mix_bwd(_p(dout), _p(g), _p(x1), _p(x2), _p(dg), _p(dx1), _p(dx2), c_long(rows), D,
                                        _stream()), "gate_mix_bwd")
        return dg, dx1, dx2


# ------------------------------------------------------------------------------------------------
# small spatial ops
# ------------------------------------------------------------------------------------------------
def _nhwc_ld(x):
    """(tensor, pixel stride) of an NHWC activation that is dense or a channel-prefix view of a padded buffer."""
    B, H, W, C = x.shape
    ld = x.stride(2)
    if x.stride(3) == 1 and ld >= C and ld % 4 == 0 and x.stride(1) == W * ld and x.stride(0) == H * W * ld \
            and x.data_ptr() % 16 == 0 and (ld == C or C % 4 == 0):
        return x, ld
    return x.contiguous(), C


class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _req_cuda(x)
        x, ldx = _nhwc_ld(x)
        B, H, W, C = x.shape
        y = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
        _check(lib().dfine_maxpool2x2_fwd(_p(x), c_long(ldx), _p(y), B, H, W, C, _stream()), "maxpool_fwd")
        ctx.save_for_backward(x)
        ctx.ldx = ldx
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        B, H, W, C = x.shape
        dx = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
        _check(lib().dfine_maxpool2x2_bwd(_p(x), c_long(ctx.ldx), _p(dy.contiguous()), _p(dx), B, H, W, C, _stream()),
               "maxpool_bwd")
        return dx


class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _req_cuda(x)
        x = x.contiguous()
        B, H, W, C = x.shape
        y = torch.empty((B, 2 * H, 2 * W, C), device=x.device, dtype=torch.float32)
        _check(lib().dfine_upsample2x_fwd(_p(x), _p(y), B, H, W, C, _stream()), "upsample_fwd")
        ctx.shape = (B, H, W, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C = ctx.shape
        dx = torch.empty((B, H, W, C), device=dy.device, dtype=torch.float32)
        _check(lib().dfine_upsample2x_bwd(_p(dy.contiguous()), _p(dx), B, H, W, C, _stream()), "upsample_bwd")
        return dx


# ------------------------------------------------------------------------------------------------
# the kernel table
# ------------------------------------------------------------------------------------------------
class CudaOps:
    name = "libdfine_sm100"

    def __init__(self):
        lib()
        if not torch.cuda.is_available():
            raise RuntimeError("custom_d_fine_b200 needs a CUDA device (sm_100a); there is no CPU compute path")
        _check(lib().dfine_check_device(), "device check")

    # ---- conv / norm ----
    def conv_bn_act(self, x, w, stride, pad, groups, bn_w, bn_b, running_mean, running_var, num_batches_tracked,
                    training, momentum=0.1, eps=1e-5, act=None, lab_scale=None, lab_bias=None, pre_add=None,
                    post_add=None, tap=False, out=None):
        frozen = num_batches_tracked is None
        if training and num_batches_tracked is not None and not (
                wgrad_stream.in_step and num_batches_tracked.data_ptr() in _step_counters):
            num_batches_tracked.add_(1)      # (inside a train step the step bumps all registered counters at once)
        # (the 3-channel image convolution has its own direct kernels, csrc/stem.cu)
        if groups == 1 and ((x.shape[-1] % 4 and x.shape[-1] != 3) or w.shape[0] % 4):
            assert out is None
            return self._conv_bn_act_padded(x, w, stride, pad, bn_w, bn_b, running_mean, running_var, training,
                                            momentum, eps, act, lab_scale, lab_bias, pre_add, post_add, frozen, tap)
        # tap: also return the input as a second output (see _ConvBnAct.forward); plain pass-through without autograd
        tap_ag = bool(tap) and _TAP and torch.is_grad_enabled() and x.requires_grad
        cfg = (stride, tuple(pad), groups, bool(training), float(momentum), float(eps), act, frozen, tap_ag)
        res = _ConvBnAct.apply(x, w, bn_w, bn_b, lab_scale, lab_bias, pre_add, post_add, running_mean, running_var, cfg, out)
        if tap and not tap_ag:
            return res, x
        return res

    # ---- concat by channel slice ----
    def concat_buffer(self, like, H, W, channels):
        """An empty [B,H,W,channels] NHWC buffer whose channel slices are handed to producers as `out=`."""
        return torch.empty((like.shape[0], H, W, channels), device=like.device, dtype=torch.float32)

    def cat_alias(self, buf, parts):
        return _CatAlias.apply(buf, *parts)

    def copy_into(self, x, out):
        return _CopyInto.apply(x, out)

    def _conv_bn_act_padded(self, x, w, stride, pad, bn_w, bn_b, running_mean, running_var, training, momentum, eps,
                            act, lab_scale, lab_bias, pre_add, post_add, frozen, tap):
        """Channel counts that are not multiples of 4 (D-FINE-n: the 21-channel CSP layers of expansion 0.34 and their
        298-channel concat, hybrid_encoder.py:181-239): the kernels' 16-byte channel granularity is met by padding the
        channel STRIDE, not by changing the math — input channels, weight rows / columns and the BatchNorm vectors are
        zero-extended to the next multiple of 4 (gamma = beta = 0 on the pad channels, so they stay exactly zero
        through BatchNorm, the activation and every gradient), the same kernels run on the padded tensors and the
        result is the channel-prefix view.  Parameters keep their reference shapes (state-dict identical); the pad /
        slice are autograd ops, so parameter gradients arrive through the ordinary accumulation path."""
        F = torch.nn.functional
        Cin, Cout = x.shape[-1], w.shape[0]
        ci, co = (-Cin) % 4, (-Cout) % 4
        xp = F.pad(x, (0, ci)) if ci else x
        wp = F.pad(w, (0, 0, 0, 0, 0, ci, 0, co)) if (ci or co) else w
        vec = (lambda t, v=0.0: F.pad(t, (0, co), value=v)) if co else (lambda t, v=0.0: t)
        rm, rv = vec(running_mean), vec(running_var, 1.0)
        padc = (lambda t: None if t is None else (F.pad(t, (0, co)) if co else t))
        cfg = (stride, tuple(pad), 1, bool(training), float(momentum), float(eps), act, frozen, False)
        out = _ConvBnAct.apply(xp, wp, vec(bn_w), vec(bn_b), lab_scale, lab_bias, padc(pre_add), padc(post_add), rm, rv,
                               cfg)
        if training and co:
            with torch.no_grad():
                running_mean.copy_(rm[:Cout])
                running_var.copy_(rv[:Cout])
        y = out[..., :Cout] if co else out
        return (y, x) if tap else y

    @torch.no_grad()
    def conv_bias_act(self, x, w, b, stride, pad, groups, act=None, lab=None, pre_add=None, post_add=None):
        """Deploy-mode conv unit (inference only, no autograd): conv with the BatchNorm folded into weight / bias
        (hybrid_encoder.py:47-80, 123-137), activation, LAB scalars `lab = (scale, bias)` — ONE kernel on the tensor-core
        path (bias + act + LAB in the epilogue); depthwise / ragged geometries and fused adds run conv + one apply pass."""
        _req_cuda(x, w)
        if x.stride(-1) != 1 or x.dim() != 4:
            x = x.contiguous()
        B, H, W, Cin = x.shape
        ok = x.stride(2) >= Cin and x.stride(1) == W * x.stride(2) and x.stride(0) == H * W * x.stride(2) \
            and x.stride(2) % 4 == 0 and x.data_ptr() % 16 == 0
        if not ok:
            x = x.contiguous()
        ldx = x.stride(2)
        Cout, _, k, _ = w.shape
        pt, pl, pb, pr = pad
        OH = (H + pt + pb - k) // stride + 1
        OW = (W + pl + pr - k) // stride + 1
        geom = (B, H, W, Cin, OH, OW, Cout, k, stride, tuple(pad))
        y = torch.empty((B, OH, OW, Cout), device=x.device, dtype=torch.float32)
        fused = (groups == 1 and pre_add is None and post_add is None and _MODE == "hf3" and Cin % 4 == 0 and Cout % 4 == 0
                 and act in (None, "relu", "silu") and _tc_ok(Cin, Cout, k, stride, pad, ldx, Cout))
        if fused:
            _conv_fwd(x, ldx, w, _wcache.getter(w), b, y, Cout, geom, ACT[act], None, lab)
            return y
        if groups > 1:
            assert groups == Cin == Cout and pt == pl == pb == pr
            if ldx != Cin:
                x, ldx = x.contiguous(), Cin
            wt = _wcache.get(w, "dw", lambda: w.reshape(Cout, k * k).t().contiguous())
            _check(lib().dfine_dwconv_fwd(_p(x), _p(wt), _p(y), B, H, W, Cin, k, stride, pt, _stream()), "dwconv_fwd")
        elif Cin % 4 or Cout % 4:
            if Cin == 3:
                _conv_fwd(x, ldx, w, _wcache.getter(w), None, y, Cout, geom, 0, None)
            else:   # D-FINE-n's 21-channel layers: zero-extended channel strides (see _conv_bn_act_padded)
                F = torch.nn.functional
                ci, co = (-Cin) % 4, (-Cout) % 4
                yp = self.conv_bias_act(F.pad(x, (0, ci)), F.pad(w, (0, 0, 0, 0, 0, ci, 0, co)), F.pad(b, (0, co)), stride,
                                        pad, 1, act, lab, None if pre_add is None else F.pad(pre_add, (0, co)),
                                        None if post_add is None else F.pad(post_add, (0, co)))
                return yp[..., :Cout]
        else:
            _conv_fwd(x, ldx, w, _wcache.getter(w), None, y, Cout, geom, 0, None)
        one = _ones_cache(Cout, x.device)
        out = torch.empty_like(y)
        ls = None if lab is None else _as_dev_scalar(lab[0], x.device)
        lb = None if lab is None else _as_dev_scalar(lab[1], x.device)
        _check(lib().dfine_bn_apply(_p(y), _p(one), _p(b), _p(None if pre_add is None else pre_add.contiguous()),
                                    _p(None if post_add is None else post_add.contiguous()), _p(ls), _p(lb), _p(out),
                                    c_long(B * OH * OW), Cout, ACT[act], c_long(Cout), c_long(Cout), _stream()), "bn_apply")
        return out

    def maxpool2x2_s1_padbr(self, x):
        return _MaxPool.apply(x)

    def upsample_nearest2x(self, x):
        return _Upsample2x.apply(x)

    def cat(self, xs, dim=-1):
        return torch.cat(list(xs), dim)

    # ---- dense ----
    def linear(self, x, w, b=None, act=None):
        return _Linear.apply(x, w, b, act)

    def layernorm(self, x, w, b, eps=1e-5, residual=None):
        return _LayerNorm.apply(x, residual, w, b, eps)

    def attention(self, qk, v, heads, mask=None):
        return _Attention.apply(qk, v, heads, mask)

    def msda(self, memory, spatial_shapes, points, heads, proj, n_off, ref, pscale, offset_scale=0.5):
        return _Msda.apply(memory, proj, ref, pscale, spatial_shapes, points, heads, n_off, offset_scale)

    # ---- decoder glue ----
    def gate_mix(self, g, x1, x2):
        return _GateMix.apply(g, x1, x2)

    @torch.no_grad()
    def select_topk(self, logits, k):
        """Indices int64 [B,k] of the k tokens with the largest row maximum of logits [B,L,C], descending."""
        _req_cuda(logits)
        logits = logits.contiguous().float()
        B, L, C = logits.shape
        scores = torch.empty((B, L), device=logits.device, dtype=torch.float32)
        idx = torch.empty((B, k), device=logits.device, dtype=torch.int64)
        _check(lib().dfine_topk_rowmax(_p(logits), _p(scores), _p(idx), B, L, C, int(k), _stream()), "topk_rowmax")
        return idx

    def fdr_decode(self, corners, ref, project, reg_scale):
        return _FdrHead.apply(corners, ref, project, _as_dev_scalar(reg_scale, corners.device), 0, True, False)[0]

    def lqe_stat(self, corners, k, reg_max):
        one = _as_dev_scalar(1.0, corners.device)
        dummy = torch.zeros(reg_max + 1, device=corners.device, dtype=torch.float32)
        return _FdrHead.apply(corners, corners.new_zeros(corners.shape[:-1] + (4,)), dummy, one, k, False, True)[1]

    def fdr_head(self, corners, ref, project, reg_scale, k):
        """(boxes, LQE statistics) of one decoder layer from one pass over pred_corners."""
        return _FdrHead.apply(corners, ref, project, _as_dev_scalar(reg_scale, corners.device), k, True, True)

    # ---- segmentation head (parity-checked on B200 by tests/test_zz_segmentation_gpu.py) ----
    # The three convolutions go through the library's conv kernels; GroupNorm, the bilinear resizes and the
    # [B,Q,C] x [B,HW,C] mask product are composed from torch device ops for now (their own kernels are next).
    def conv2d(self, x, w, stride=1, pad=(0, 0, 0, 0), groups=1):
        if groups != 1:
            raise NotImplementedError("plain grouped convolutions are not on any path")
        return _Conv.apply(x, w, int(stride), tuple(int(v) for v in pad))

    def group_norm(self, x, groups, w, b, eps=1e-5, act=None):
        if act not in (None, "relu"):
            raise NotImplementedError(act)
        return _GroupNorm.apply(x, w, b, int(groups), float(eps), act == "relu")

    def resize_bilinear(self, x, size):
        return _ResizeBilinear.apply(x, (int(size[0]), int(size[1])))

    def mask_dot(self, embed, feat_nhwc):
        """[B,Q,C] x [B,Hm,Wm,C] -> mask logits [B,Q,Hm,Wm] — returned as a VIEW of a pixel-major [B,Hm,Wm,Q] buffer (the
        layout the implicit-GEMM kernel writes and the mask-cost kernel reads)."""
        if _MODE == "simt":          # strict-fp32 parity mode: no tensor-core product
            return torch.einsum("bqc,bhwc->bqhw", embed, feat_nhwc)
        return _MaskDot.apply(embed, feat_nhwc).permute(0, 3, 1, 2)

    def _multi_plan(self, pers, device):
        """Index tensors of the image-major stacking used by mask_losses_multi, cached per tuple of per-head per-image
        counts (host-known: they are part of the CUDA-graph key): perm[i] = head-major row of image-major row i, and the
        [n_heads, sumM] averaging matrix (1 / M_h on head h's rows)."""
        key = (tuple(tuple(p) for p in pers), device)
        hit = self._multi_cache.get(key)
        if hit is None:
            B = len(pers[0])
            base, o = [], 0
            for per in pers:                      # head-major offsets: head h's rows are sorted by image
                offs = [o]
                for n in per:
                    offs.append(offs[-1] + n)
                base.append(offs)
                o = offs[-1]
            perm, head_of, totals = [], [], []
            for b in range(B):
                n_b = 0
                for h, per in enumerate(pers):
                    perm.extend(range(base[h][b], base[h][b] + per[b]))
                    head_of.extend([h] * per[b])
                    n_b += per[b]
                totals.append(n_b)
            avg = torch.zeros((len(pers), max(len(perm), 1)), dtype=torch.float32)
            for i, h in enumerate(head_of):
                avg[h, i] = 1.0 / sum(pers[h])
            # padded image-major stacking for the pixel-major path: Rmax rows per image, the pads select an appended zero
            # row / target -1 and carry weight 0 in the averaging matrix
            n_all = len(perm)
            rmax = (max(totals + [1]) + 3) // 4 * 4
            perm_pad, avg_pad, o = [], torch.zeros((len(pers), B * rmax), dtype=torch.float32), 0
            for b in range(B):
                perm_pad += perm[o:o + totals[b]] + [n_all] * (rmax - totals[b])
                avg_pad[:, b * rmax:b * rmax + totals[b]] = avg[:, o:o + totals[b]]
                o += totals[b]
            hit = (torch.tensor(perm, dtype=torch.int64).to(device), avg.to(device), tuple(totals),
                   torch.tensor(perm_pad, dtype=torch.int64).to(device), avg_pad.to(device), rmax)
            if len(self._multi_cache) >= 64:
                self._multi_cache.pop(next(iter(self._multi_cache)))
            self._multi_cache[key] = hit
        return hit

    def mask_losses_multi(self, feat_nhwc, heads, gt_resized, tboxes):
        """Cropped BCE / Dice mask losses (dfine_criterion.py:504-556) of SEVERAL heads at once, evaluated on the matched
        (image, query) pairs only: heads = list of (embed [B,Q,C], b_idx, q_idx, t_idx, per-image host counts; pairs sorted
        by image).  The matched rows of every head are stacked image-major (one index_select), their logits are ONE
        product per image with that image's mask features, the loss kernels run once over all rows and the per-head
        means are one small matrix product: the unmatched queries' [B,Q,Hm,Wm] logits never enter the autograd graph and
        no per-head slice of the stacked logits does either (104 slice backwards were 30 ms of full-size fills + adds).
        -> (bce [n_heads], dice [n_heads])."""
        B, Hm, Wm, C = feat_nhwc.shape
        perm, avg, totals, perm_pad, avg_pad, rmax = self._multi_plan([per for *_, per in heads], feat_nhwc.device)
        if perm.numel() == 0:
            z = feat_nhwc.new_zeros(len(heads))
            return z, z
        rows_hm = torch.cat([e.reshape(-1, C).index_select(0, bi * e.shape[1] + qi) for e, bi, qi, _, _ in heads])
        if _MASK_PM and C % 8 == 0 and _MODE != "simt":
            # pixel-major: the logits of the stacked rows are ONE grouped tcgen05 product over all images (per-image
            # "weights" = that image's rows, padded to Rmax), its gradients the grouped data-gradient / per-image
            # weight-gradient kernels of _MaskDot, and the loss kernels read that layout directly — no library sgemm
            rows_pad = torch.cat([rows_hm, rows_hm.new_zeros((1, C))]).index_select(0, perm_pad).view(B, rmax, C)
            t_pad = torch.cat([ti for _, _, _, ti, _ in heads] + [perm_pad.new_full((1,), -1)]).index_select(0, perm_pad)
            pred = _MaskDot.apply(rows_pad, feat_nhwc)                                                   # [B,Hm,Wm,Rmax]
            bce_rows, dice_rows = _MaskLossPM.apply(pred, gt_resized, t_pad, tboxes)
            both = avg_pad @ torch.stack((bce_rows, dice_rows), 1)
            return both[:, 0], both[:, 1]
        t_all = torch.cat([ti for _, _, _, ti, _ in heads]).index_select(0, perm)
        out = _RowsTimesFeat.apply(rows_hm.index_select(0, perm), feat_nhwc, totals)                      # [sumM, HW]
        bce_rows, dice_rows = _MaskLossRows.apply(out.view(-1, Hm, Wm), gt_resized, t_all, tboxes)
        both = avg @ torch.stack((bce_rows, dice_rows), 1)                                                # [n_heads, 2]
        return both[:, 0], both[:, 1]

    def mask_loss_rows(self, pred, gt_resized, t_idx, tboxes):
        """(bce_row [M], dice_row [M]) of matched mask logits pred [M,Hm,Wm] against gt_resized[t_idx] inside the GT boxes."""
        return _MaskLossRows.apply(pred, gt_resized, t_idx, tboxes)

    @torch.no_grad()
    def mask_cost_layer(self, pred_masks, gt, gsum, toff_dev, sizes, alpha, gamma, w_dice, w_mask):
        """Mask term of the matching cost of one layer (matcher.py:175-237): pred_masks [B,Q,Hm,Wm] (a view of the
        pixel-major buffer), gt [sumT, Hm*Wm] resized GT masks, gsum their sums -> [Q*sumT] cost blocks."""
        B, Q, Hm, Wm = pred_masks.shape
        pm = pred_masks.permute(0, 2, 3, 1)
        if not pm.is_contiguous():
            pm = pm.contiguous()
        sumT, Tmax = sum(sizes), max(sizes)
        extra = torch.zeros(Q * sumT, device=pm.device, dtype=torch.float32)
        ws = torch.empty(2 * Q * (sumT + B), device=pm.device, dtype=torch.float32)
        _check(lib().dfine_mask_cost(_p(pm), _p(gt), _p(gsum), _p(toff_dev), _p(extra), _p(ws), B, c_long(Hm * Wm), Q, sumT,
                                     Tmax, c_float(alpha), c_float(gamma), c_float(w_dice), c_float(w_mask), _stream()),
               "mask_cost")
        return extra

    def toff_device(self, sizes, device):
        key = (tuple(sizes), device)
        toff = self._toff_cache.get(key)
        if toff is None:
            offs = [0]
            for sz in sizes:
                offs.append(offs[-1] + sz)
            toff = torch.tensor(offs, dtype=torch.int32).to(device)
            if len(self._toff_cache) >= 256:
                self._toff_cache.pop(next(iter(self._toff_cache)))
            self._toff_cache[key] = toff
        return toff

    # ---- input side / post-processing (the callers either side of the hot path) ----
    @torch.no_grad()
    def preprocess_u8(self, images_u8, size=None, mul=1.0 / 255.0, swap_rb=True):
        _req_cuda(images_u8)
        assert images_u8.dtype == torch.uint8 and images_u8.dim() == 4
        x = images_u8.contiguous()
        B, Hs, Ws, C = x.shape
        H, W = (Hs, Ws) if size is None else (int(size[0]), int(size[1]))
        out = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
        _check(lib().dfine_preprocess_u8(_p(x), _p(out), B, Hs, Ws, H, W, C, c_float(mul), 1 if swap_rb else 0, _stream()),
               "preprocess_u8")
        return out

    @torch.no_grad()
    def resize_images(self, x_nhwc, size):
        _req_cuda(x_nhwc)
        x = x_nhwc.contiguous().float()
        B, Hs, Ws, C = x.shape
        H, W = int(size[0]), int(size[1])
        out = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
        _check(lib().dfine_resize_bilinear_f32(_p(x), _p(out), B, Hs, Ws, H, W, C, _stream()), "resize_bilinear_f32")
        return out

    @torch.no_grad()
    def postprocess(self, logits, boxes, k, height, width, to_round=True):
        """(labels int64 [B,k], boxes xyxy [B,k,4] in input pixels, scores [B,k], query index int64 [B,k])."""
        _req_cuda(logits, boxes)
        logits, boxes = logits.contiguous().float(), boxes.contiguous().float()
        B, Q, C = logits.shape
        dev = logits.device
        ws = torch.empty((B, Q * C), device=dev, dtype=torch.float32)
        idx = torch.empty((B, k), device=dev, dtype=torch.int64)
        _check(lib().dfine_topk_rowmax(_p(logits), _p(ws), _p(idx), B, Q * C, 1, int(k), _stream()), "topk_rowmax")
        labels = torch.empty((B, k), device=dev, dtype=torch.int64)
        qidx = torch.empty((B, k), device=dev, dtype=torch.int64)
        out_boxes = torch.empty((B, k, 4), device=dev, dtype=torch.float32)
        scores = torch.empty((B, k), device=dev, dtype=torch.float32)
        _check(lib().dfine_postprocess(_p(logits), _p(boxes), _p(idx), _p(labels), _p(out_boxes), _p(scores), _p(qidx), B, Q,
                                       C, int(k), c_float(height), c_float(width), 1 if to_round else 0, _stream()),
               "postprocess")
        return labels, out_boxes, scores, qidx

    # ---- criterion ----
    def criterion_sets(self, full, enc_logits, enc_boxes, table, counts, labels, tboxes, project, reg_scale, meta):
        """All VFL / L1 / GIoU / FGL / DDF terms of a step from the unsplit decoder stacks (loss.cu); returns
        (vfl [2,L+2], l1, giou, fgl [2,L], ddf [2,L]) — group 0 = matching queries, 1 = denoising; heads = layers, pre, enc."""
        return _Criterion.apply(full["logits"], full["boxes"], full["corners"], full["pre_logits"], full["pre_boxes"],
                                enc_logits, enc_boxes, full["refs"][0], table, counts, labels, tboxes, project, reg_scale,
                                meta)

    # ---- matcher ----
    @torch.no_grad()
    def match_device(self, logits, boxes, labels, tboxes, toff_dev, sumT, Tmax, alpha, gamma, w_class, w_bbox,
                     w_giou, want_cost=False, extra_cost=None):
        """logits [NL,B,Q,C], boxes [NL,B,Q,4] contiguous; returns device int64 [NL,sumT] x2 (+ cost)."""
        NL, B, Q, C = logits.shape
        dev = logits.device
        out_q = torch.zeros((NL, sumT), device=dev, dtype=torch.int64)
        out_t = torch.zeros((NL, sumT), device=dev, dtype=torch.int64)
        cost = torch.zeros((NL, Q * sumT), device=dev, dtype=torch.float32) if want_cost else None
        ws_bytes = lib().dfine_matcher_workspace_bytes(NL, B, Q, Tmax)
        ws = torch.empty(ws_bytes // 4, device=dev, dtype=torch.float32) if ws_bytes else None
        if extra_cost is not None:      # segmentation: mask cost blocks [NL, Q*sumT] (matcher.HungarianMatcher.mask_cost)
            extra_cost = extra_cost.detach().float().contiguous()
            assert tuple(extra_cost.shape) == (NL, Q * sumT), (tuple(extra_cost.shape), NL, Q, sumT)
            _check(lib().dfine_matcher_extra(_p(logits), _p(boxes), _p(labels), _p(tboxes), _p(toff_dev), _p(extra_cost),
                                             _p(out_q), _p(out_t), _p(cost), _p(ws), NL, B, Q, C, sumT, Tmax,
                                             c_float(alpha), c_float(gamma), c_float(w_class), c_float(w_bbox),
                                             c_float(w_giou), _stream()), "matcher")
            return out_q, out_t, cost
        _check(lib().dfine_matcher(_p(logits), _p(boxes), _p(labels), _p(tboxes), _p(toff_dev), _p(out_q), _p(out_t),
                                   _p(cost), _p(ws), NL, B, Q, C, sumT, Tmax, c_float(alpha), c_float(gamma),
                                   c_float(w_class), c_float(w_bbox), c_float(w_giou), _stream()), "matcher")
        return out_q, out_t, cost

    _toff_cache = {}
    _multi_cache = {}

    @torch.no_grad()
    def match_raw(self, logits_list, boxes_list, targets, alpha=0.25, gamma=2.0, w_class=2.0, w_bbox=5.0,
                  w_giou=2.0, extra_cost=None):
        """Launch only (graph-capturable): returns device int64 (out_q, out_t) of shape [n_layers, sumT]."""
        logits = torch.stack([l.detach().float() for l in logits_list]).contiguous()
        boxes = torch.stack([b.detach().float() for b in boxes_list]).contiguous()
        _req_cuda(logits, boxes)
        dev = logits.device
        sizes = tuple(int(t["labels"].shape[0]) for t in targets)
        sumT, Tmax = sum(sizes), max(sizes) if sizes else 0
        if sumT == 0:
            z = torch.zeros((logits.shape[0], 0), device=dev, dtype=torch.int64)
            return z, z
        key = (sizes, dev)
        toff = self._toff_cache.get(key)
        if toff is None:   # depends on the targets' sizes only; built once (a pageable H2D is not capturable)
            offs = [0]
            for sz in sizes:
                offs.append(offs[-1] + sz)
            toff = torch.tensor(offs, dtype=torch.int32).to(dev)
            if len(self._toff_cache) >= 256:          # bounded: per-image target counts rarely repeat with a real loader
                self._toff_cache.pop(next(iter(self._toff_cache)))
            self._toff_cache[key] = toff
        labels = torch.cat([t["labels"] for t in targets]).to(dev, torch.int64).contiguous()
        tboxes = torch.cat([t["boxes"] for t in targets]).to(dev, torch.float32).contiguous()
        out_q, out_t, _ = self.match_device(logits, boxes, labels, tboxes, toff, sumT, Tmax, alpha, gamma, w_class,
                                            w_bbox, w_giou, extra_cost=extra_cost)
        return out_q, out_t

    @staticmethod
    def match_raw_to_host(raw, plan):
        both = torch.stack(list(raw)).cpu().numpy()        # the step's single matcher D2H (+ sync)
        return both[0], both[1]

    @torch.no_grad()
    def match(self, logits_list, boxes_list, targets, alpha=0.25, gamma=2.0, w_class=2.0, w_bbox=5.0, w_giou=2.0,
              extra_cost=None):
        logits = torch.stack([l.detach().float() for l in logits_list]).contiguous()
        boxes = torch.stack([b.detach().float() for b in boxes_list]).contiguous()
        _req_cuda(logits, boxes)
        NL, B, Q, C = logits.shape
        sizes = [int(t["labels"].shape[0]) for t in targets]
        sumT, Tmax = sum(sizes), max(sizes) if sizes else 0
        empty = (torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64))
        if sumT == 0:
            return [[empty for _ in range(B)] for _ in range(NL)]
        dev = logits.device
        labels = torch.cat([t["labels"] for t in targets]).to(dev, torch.int64).contiguous()
        tboxes = torch.cat([t["boxes"] for t in targets]).to(dev, torch.float32).contiguous()
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        toff = torch.tensor(offs, dtype=torch.int32).to(dev, non_blocking=True)
        out_q, out_t, _ = self.match_device(logits, boxes, labels, tboxes, toff, sumT, Tmax, alpha, gamma, w_class,
                                            w_bbox, w_giou, extra_cost=extra_cost)
        both = torch.stack([out_q, out_t]).cpu()        # the step's single matcher D2H
        res = []
        for l in range(NL):
            per = []
            for b in range(B):
                n = min(Q, sizes[b])
                if n == 0:
                    per.append(empty)
                else:
                    per.append((both[0, l, offs[b]:offs[b] + n].clone(), both[1, l, offs[b]:offs[b] + n].clone()))
            res.append(per)
        return res
