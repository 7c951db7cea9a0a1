"""GPU parity tests: every kernel-table op of the CUDA path against the CPU oracle (oracle/torch_ops.py)
on the same seeded inputs, called through the C ABI (ctypes) exactly as the model calls them.

Tolerances: fp32 CUDA-core kernels 3e-5 relative to the tensor's max magnitude; tcgen05 kernels 3e-3: in the
default "tc3" mode the forward GEMM is 3xTF32 (measured 3e-7..8e-6, profiles/README.md) but data / weight
gradients use plain kind::tf32 operands (10-bit mantissa, fp32 accumulation) — the operand precision cuDNN
uses for the reference's convolutions on GPU (torch.backends.cudnn.allow_tf32 defaults to True).
"""
import math

import pytest
import torch

from tests.util import check_close, run_both

pytestmark = pytest.mark.gpu

F32 = 3e-5
TF32 = 3e-3


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _conv_case(cuda_ops, oracle_ops, B, H, W, Cin, Cout, k, stride, pad, groups, act, lab, pre, post, training, tol,
               seed=0, slice_in=0):
    g = _g(seed)
    xfull = torch.randn(B, H, W, Cin + slice_in, generator=g)
    w = torch.randn(Cout, Cin // groups, k, k, generator=g) * (1.0 / math.sqrt(k * k * Cin / groups))
    bn_w = torch.rand(Cout, generator=g) + 0.5
    bn_b = torch.randn(Cout, generator=g) * 0.1
    rm = torch.randn(Cout, generator=g) * 0.1
    rv = torch.rand(Cout, generator=g) + 0.5
    ls = torch.tensor([1.3]) if lab else None
    lb = torch.tensor([-0.2]) if lab else None
    pt, pl, pb, pr = pad
    OH = (H + pt + pb - k) // stride + 1
    OW = (W + pl + pr - k) // stride + 1
    pre_t = torch.randn(B, OH, OW, Cout, generator=g) if pre else None
    post_t = torch.randn(B, OH, OW, Cout, generator=g) if post else None
    # eval-mode BatchNorm (frozen backbones of l/x, inference) has no affine gradient on the path
    ins = [xfull.requires_grad_(), w.requires_grad_(), bn_w.requires_grad_(training), bn_b.requires_grad_(training)]
    if lab:
        ins += [ls.requires_grad_(), lb.requires_grad_()]
    if pre:
        ins.append(pre_t.requires_grad_())
    if post:
        ins.append(post_t.requires_grad_())
    state = {}

    def fn(K, x, w, bw, bb, *rest):
        rest = list(rest)
        l_s = rest.pop(0) if lab else None
        l_b = rest.pop(0) if lab else None
        p1 = rest.pop(0) if pre else None
        p2 = rest.pop(0) if post else None
        dev = x.device
        r_m, r_v = rm.clone().to(dev), rv.clone().to(dev)
        nbt = torch.zeros((), dtype=torch.int64, device=dev)
        xin = x[..., slice_in:] if slice_in else x
        y = K.conv_bn_act(xin, w, stride, pad, groups, bw, bb, r_m, r_v, nbt, training=training, momentum=0.1,
                          eps=1e-5, act=act, lab_scale=l_s, lab_bias=l_b, pre_add=p1, post_add=p2)
        state[str(dev.type)] = (r_m.cpu(), r_v.cpu(), int(nbt))
        return y

    # tf32 forward differences flip a few ReLU masks (pre-activations within ~1e-3 of zero); each flip moves
    # the affected gradient entries by O(1), so tf32+ReLU gradients are compared in the L2 norm
    kink = tol == TF32 and act == "relu"
    errs = run_both(fn, cuda_ops, oracle_ops, ins, tol, 8e-2 if kink else tol * 3, seed,
                    grad_metric="l2" if kink else "max")
    if training:
        check_close("running_mean", state["cuda"][0], state["cpu"][0], tol)
        check_close("running_var", state["cuda"][1], state["cpu"][1], tol)
        assert state["cuda"][2] == state["cpu"][2] == 1
    return errs


CONV_CASES = [
    # name, B,H,W,Cin,Cout,k,stride,pad,groups,act,lab,pre,post,training,tol
    ("stem1_3x3s2", 2, 64, 64, 3, 24, 3, 2, (1, 1, 1, 1), 1, "relu", True, False, False, True, F32),
    ("stem2a_2x2_padbr", 2, 32, 32, 24, 12, 2, 1, (0, 0, 1, 1), 1, "relu", True, False, False, True, TF32),
    ("stem3_3x3s2", 2, 32, 32, 48, 24, 3, 2, (1, 1, 1, 1), 1, "relu", True, False, False, True, TF32),
    ("dw3x3s2", 2, 40, 40, 96, 96, 3, 2, (1, 1, 1, 1), 96, None, False, False, False, True, F32),
    ("dw3x3s2_odd", 2, 37, 41, 32, 32, 3, 2, (1, 1, 1, 1), 32, None, False, False, False, True, F32),
    ("dw5x5", 2, 20, 20, 128, 128, 5, 1, (2, 2, 2, 2), 128, "relu", True, False, False, True, F32),
    ("pw1x1_tc", 2, 40, 40, 160, 48, 1, 1, (0, 0, 0, 0), 1, "relu", True, False, False, True, TF32),
    ("pw1x1_tc_big", 2, 20, 20, 896, 384, 1, 1, (0, 0, 0, 0), 1, "relu", True, False, True, True, TF32),
    ("c3x3_tc_32", 2, 40, 40, 32, 32, 3, 1, (1, 1, 1, 1), 1, "relu", True, False, False, True, TF32),
    ("c3x3_tc_128", 2, 20, 20, 128, 128, 3, 1, (1, 1, 1, 1), 1, "silu", False, False, False, True, TF32),
    ("c3x3_tc_128_odd", 1, 23, 37, 128, 128, 3, 1, (1, 1, 1, 1), 1, "silu", False, False, False, True, TF32),
    ("repvgg_pre_add", 2, 20, 20, 128, 128, 1, 1, (0, 0, 0, 0), 1, "silu", False, True, False, True, TF32),
    ("proj_eval", 2, 20, 20, 384, 256, 1, 1, (0, 0, 0, 0), 1, None, False, False, False, False, TF32),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_bn_act(cuda_ops, oracle_ops, case):
    _conv_case(cuda_ops, oracle_ops, *case[1:])


def test_conv_channel_slice_input(cuda_ops, oracle_ops):
    # RepNCSPELAN4 feeds the second half of cv1's output to cv2 (hybrid_encoder.py:203-206): strided view
    _conv_case(cuda_ops, oracle_ops, 2, 20, 20, 128, 64, 1, 1, (0, 0, 0, 0), 1, "silu", False, False, False, True,
               TF32, slice_in=128)


LINEAR_CASES = [
    ("ffn_relu", (3, 50, 256), 1024, "relu", TF32),
    ("ffn_gelu", (3, 50, 256), 1024, "gelu", TF32),
    ("score_head", (2, 37, 256), 80, None, TF32),
    ("bbox_head_4", (2, 37, 256), 4, None, TF32),
    ("corners_132", (2, 37, 256), 132, None, TF32),
    ("lqe_out_1", (2, 37, 64), 1, None, F32),
    ("lqe_in_20", (2, 37, 20), 64, "relu", TF32),
    ("qpos_4", (2, 37, 4), 512, "relu", TF32),
    ("big", (1, 8400, 256), 256, None, TF32),
]


@pytest.mark.parametrize("case", LINEAR_CASES, ids=[c[0] for c in LINEAR_CASES])
def test_linear(cuda_ops, oracle_ops, case):
    _, xs, n, act, tol = case
    g = _g(1)
    x = torch.randn(*xs, generator=g).requires_grad_()
    w = (torch.randn(n, xs[-1], generator=g) / math.sqrt(xs[-1])).requires_grad_()
    b = torch.randn(n, generator=g).requires_grad_()
    kink = tol == TF32 and act == "relu"
    run_both(lambda K, x, w, b: K.linear(x, w, b, act=act), cuda_ops, oracle_ops, [x, w, b], tol,
             8e-2 if kink else tol * 3, grad_metric="l2" if kink else "max")


def test_layernorm_residual(cuda_ops, oracle_ops):
    g = _g(2)
    x = torch.randn(4, 77, 256, generator=g).requires_grad_()
    r = torch.randn(4, 77, 256, generator=g).requires_grad_()
    w = (torch.rand(256, generator=g) + 0.5).requires_grad_()
    b = torch.randn(256, generator=g).requires_grad_()
    run_both(lambda K, x, r, w, b: K.layernorm(x, w, b, 1e-5, residual=r), cuda_ops, oracle_ops, [x, r, w, b], F32,
             1e-4)
    run_both(lambda K, x, w, b: K.layernorm(x, w, b, 1e-5), cuda_ops, oracle_ops, [x, w, b], F32, 1e-4)


@pytest.mark.parametrize("S,D,heads,masked", [(400, 256, 8, False), (500, 256, 8, True), (130, 128, 8, True),
                                               (67, 384, 8, False)])
def test_attention(cuda_ops, oracle_ops, S, D, heads, masked):
    g = _g(3)
    qk = torch.randn(2, S, 2 * D, generator=g).requires_grad_()
    v = torch.randn(2, S, D, generator=g).requires_grad_()
    mask = None
    if masked:  # CDN-style block mask: first n_dn rows/cols are denoising groups
        n_dn, grp = (S // 5) * 2, 10
        idx = torch.arange(S)
        is_dn = idx < n_dn
        gid = idx // grp
        mask = is_dn[None, :] & (~is_dn[:, None] | (gid[:, None] != gid[None, :]))

    def fn(K, qk, v):
        return K.attention(qk, v, heads, None if mask is None else mask.to(qk.device))

    # forward: 3xTF32 logits and P*V (fp32-class).  backward: logits recomputed in 3xTF32, the four gradient
    # products (dP, dQ, dK, dV) are single round-to-nearest tf32 MMAs like every other gradient GEMM of the library
    run_both(fn, cuda_ops, oracle_ops, [qk, v], 5e-5, TF32)


@pytest.mark.parametrize("D,shapes,points", [(256, [(80, 80), (40, 40), (20, 20)], [3, 6, 3]),
                                              (128, [(40, 40), (20, 20)], [6, 6]),
                                              (256, [(13, 17), (7, 9), (4, 5)], [3, 6, 3])])
def test_msda(cuda_ops, oracle_ops, D, shapes, points):
    g = _g(4)
    heads, B, Q = 8, 2, 150
    L, P = sum(h * w for h, w in shapes), sum(points)
    mem = torch.randn(B, L, D, generator=g).requires_grad_()
    proj = torch.randn(B, Q, heads * P * 3, generator=g)
    proj[..., : heads * P * 2] *= 3.0          # offsets large enough to hit the borders
    proj.requires_grad_()
    ref = torch.rand(B, Q, 4, generator=g)
    ref[..., 2:] = ref[..., 2:] * 0.4 + 0.02
    ref[0, 0] = torch.tensor([0.0, 1.0, 0.5, 0.5])   # corner query: samples fall outside the map
    pscale = torch.tensor([1.0 / n for n in points for _ in range(n)])

    def fn(K, mem, proj):
        dev = mem.device
        return K.msda(mem, shapes, points, heads, proj, heads * P * 2, ref.to(dev), pscale.to(dev), 0.5)

    run_both(fn, cuda_ops, oracle_ops, [mem, proj], F32, 2e-4)


@pytest.mark.parametrize("which", ["both", "decode", "stat"])
def test_fdr_head(cuda_ops, oracle_ops, which):
    """Integral + distance2bbox + LQE statistics kernel (fdr.cu) against the reference's op chain
    (dfine_decoder.py:291-313, arch/utils.py:119-188) on the same corner logits, forward and d(pred_corners)."""
    from custom_d_fine_b200.decoder import weighting_function
    g = _g(11)
    corners = (torch.randn(2, 37, 4 * 33, generator=g) * 2.0).requires_grad_()
    ref = torch.rand(2, 37, 4, generator=g) * 0.5 + 0.2
    reg_scale = torch.tensor([4.0])
    project = weighting_function(32, torch.tensor([0.5]), reg_scale)

    def fn(K, c):
        dev = c.device
        if which == "decode":
            return K.fdr_decode(c, ref.to(dev), project.to(dev), reg_scale.to(dev))
        if which == "stat":
            return K.lqe_stat(c, 4, 32)
        box, stat = K.fdr_head(c, ref.to(dev), project.to(dev), reg_scale.to(dev), 4)
        return torch.cat([box, stat], -1)

    run_both(fn, cuda_ops, oracle_ops, [corners], F32, F32)


def test_maxpool_upsample(cuda_ops, oracle_ops):
    g = _g(5)
    x = torch.randn(2, 17, 19, 24, generator=g)
    x[0, :, :, 0] = -1.0   # all-negative plane: the zero padding wins at the border
    x = x.requires_grad_()
    run_both(lambda K, x: K.maxpool2x2_s1_padbr(x), cuda_ops, oracle_ops, [x], 0.0, 1e-6)
    y = torch.randn(2, 10, 10, 256, generator=g).requires_grad_()
    run_both(lambda K, y: K.upsample_nearest2x(y), cuda_ops, oracle_ops, [y], 0.0, 1e-6)


TC_CASES = [
    # B, H, W, Cin, Cout, k, stride, pad(t,l,b,r)
    (2, 40, 40, 128, 128, 3, 1, (1, 1, 1, 1)),
    (1, 1, 4000, 256, 512, 1, 1, (0, 0, 0, 0)),       # nn.Linear shape
    (2, 80, 80, 64, 64, 3, 1, (1, 1, 1, 1)),
    (3, 20, 20, 1280, 384, 1, 1, (0, 0, 0, 0)),
    (1, 1, 999, 20, 64, 1, 1, (0, 0, 0, 0)),          # ragged K (LQE MLP)
    (2, 64, 64, 48, 24, 3, 2, (1, 1, 1, 1)),          # stem3: stride 2 through the TMA element strides
    (2, 33, 47, 48, 24, 3, 2, (1, 1, 1, 1)),          # odd sizes, stride 2
    (2, 40, 40, 24, 12, 2, 1, (0, 0, 1, 1)),          # stem2a: 2x2, bottom/right padding, Cout 12
    (2, 40, 40, 12, 24, 2, 1, (0, 0, 1, 1)),          # stem2b: Cin 12
    (2, 37, 131, 32, 16, 2, 1, (0, 0, 1, 1)),         # l / x stem2a, ragged tile edges (direct 2x2 kernels, csrc/stem.cu)
    (1, 19, 70, 16, 32, 2, 1, (0, 0, 1, 1)),          # l / x stem2b
    (2, 24, 65, 16, 8, 2, 1, (0, 0, 1, 1)),           # n stem2a
    (2, 24, 65, 8, 16, 2, 1, (0, 0, 1, 1)),           # n stem2b
    (1, 1, 700, 4, 512, 1, 1, (0, 0, 0, 0)),          # query_pos_head layer 0: K = 4
    (1, 1, 700, 256, 4, 1, 1, (0, 0, 0, 0)),          # bbox head: N = 4
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "x".join(str(v) for v in c[:7]))
def test_tc_matches_simt(cuda_ops, case):
    """tcgen05 kernels (plain tf32 and 3xTF32 forward, tf32 dgrad / wgrad) against the fp32 CUDA-core kernels of
    the same library on identical device data, through the host launchers the autograd functions use."""
    from custom_d_fine_b200 import cuda_ops as co
    B, H, W, Cin, Cout, k, stride, pad = case
    g = _g(6)
    OH = (H + pad[0] + pad[2] - k) // stride + 1
    OW = (W + pad[1] + pad[3] - k) // stride + 1
    geom = (B, H, W, Cin, OH, OW, Cout, k, stride, pad)
    x = torch.randn(B, H, W, Cin, generator=g).cuda()
    weight = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(k * k * Cin)).cuda()
    dy = torch.randn(B, OH, OW, Cout, generator=g).cuda()
    M = B * OH * OW
    res = {}
    prev = co.get_gemm_mode()
    try:
        for mode in ("simt", "tc", "tc3", "tch", "bf3", "hf3"):
            co.set_gemm_mode(mode)
            cache = co._WCache()
            y = torch.zeros(B, OH, OW, Cout).cuda()
            stats = torch.zeros(2 * Cout, dtype=torch.float64).cuda()
            used_tc = co._conv_fwd(x, Cin, weight, cache.getter(weight), None, y, Cout, geom, 0, stats)
            assert used_tc == (mode != "simt"), "the tensor-core path must take every TC_CASES geometry"
            dx = torch.full((B, H, W, Cin), float("nan")).cuda()
            co._conv_dgrad(dy, Cout, weight, cache.getter(weight), dx, Cin, geom)
            dw = co._conv_wgrad(dy, Cout, x, Cin, geom)
            if mode == "tc":     # residual folded into the data-gradient epilogue (a channel slice: pixel stride Cin + 8)
                rfull = torch.randn(B, H, W, Cin + 8, generator=g).cuda()
                dxr = torch.full((B, H, W, Cin), float("nan")).cuda()
                co._conv_dgrad(dy, Cout, weight, cache.getter(weight), dxr, Cin, geom, rfull[..., 4:4 + Cin])
                check_close(f"dgrad + residual {case}", dxr, dx + rfull[..., 4:4 + Cin], 1e-6)
            torch.cuda.synchronize()
            res[mode] = (y, stats, dx, dw)
    finally:
        co.set_gemm_mode(prev)
    y0, _, dx0, dw0 = res["simt"]
    for mode, tol in (("tc", TF32), ("tc3", 2e-5), ("tch", 2e-5), ("bf3", 5e-5), ("hf3", 2e-5)):   # bf3: 16 mantissa bits per operand
        y, stats, dx, dw = res[mode]
        check_close(f"{mode} fwd {case}", y, y0, tol)
        check_close(f"{mode} fused stats sum", stats[:Cout].float(), y0.reshape(M, Cout).sum(0), max(tol, 1e-4) * 5)
        check_close(f"{mode} fused stats sumsq", stats[Cout:].float(), (y0.reshape(M, Cout) ** 2).sum(0), max(tol, 1e-4))
        check_close(f"{mode} dgrad {case}", dx, dx0, TF32)
        check_close(f"{mode} wgrad {case}", dw, dw0, TF32)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 40, 40, 160, 48, 1, "relu", True, True), (2, 20, 20, 128, 128, 3, "silu", False, False),
                                   (1, 23, 37, 128, 256, 1, None, False, True)],
                         ids=["pw1x1_lab_post", "c3x3_silu", "odd_256"])
def test_bn_finalize_in_conv_tail_equals_separate_launches(cuda_ops, shape):
    """Train-mode conv + BatchNorm: the finalize in the conv kernel's last CTA (default), the opt-in in-kernel apply pass
    behind a grid-wide wait and the separate bn_finalize / bn_apply launches are the SAME arithmetic on the same
    statistics — outputs, saved statistics (through the backward pass) and running statistics must agree to rounding."""
    from custom_d_fine_b200 import cuda_ops as co
    B, H, W, Cin, Cout, k, act, lab, post = shape
    g = _g(11)
    dev = torch.device("cuda", 0)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(k * k * Cin)).to(dev)
    bw, bb = (torch.rand(Cout, generator=g) + 0.5).to(dev), (torch.randn(Cout, generator=g) * 0.1).to(dev)
    ls = torch.tensor([1.3], device=dev) if lab else None
    lb = torch.tensor([-0.2], device=dev) if lab else None
    p2 = torch.randn(B, H, W, Cout, generator=g).to(dev) if post else None
    dy = torch.randn(B, H, W, Cout, generator=g).to(dev)
    pad = ((k - 1) // 2,) * 4

    def run(fin, apply_):
        old = co._BN_FIN, co._BN_APPLY
        co._BN_FIN, co._BN_APPLY = fin, apply_
        try:
            xs, ws = x.clone().requires_grad_(), w.clone().requires_grad_()
            bws, bbs = bw.clone().requires_grad_(), bb.clone().requires_grad_()
            rm, rv = torch.zeros(Cout, device=dev), torch.ones(Cout, device=dev)
            nbt = torch.zeros((), dtype=torch.int64, device=dev)
            y = cuda_ops.conv_bn_act(xs, ws, 1, pad, 1, bws, bbs, rm, rv, nbt, training=True, momentum=0.1, eps=1e-5,
                                     act=act, lab_scale=ls, lab_bias=lb, pre_add=None, post_add=p2)
            y.backward(dy)
            torch.cuda.synchronize()
            return [t.detach().clone() for t in (y, rm, rv, xs.grad, ws.grad, bws.grad, bbs.grad)]
        finally:
            co._BN_FIN, co._BN_APPLY = old

    base = run(False, False)
    for mode in ((True, False), (True, True)):
        got = run(*mode)
        for name, a, b in zip(("y", "running_mean", "running_var", "dx", "dw", "dgamma", "dbeta"), got, base):
            scale = float(b.abs().max()) + 1e-12
            # (the gradient kernels accumulate split-K partial sums with fp32 atomics: run-to-run noise of ~1e-6)
            tol = 2e-6 if name in ("y", "running_mean", "running_var") else 3e-5
            assert float((a - b).abs().max()) <= tol * scale, (mode, name, float((a - b).abs().max()), scale)


@pytest.mark.gpu
def test_tc_trace_stamps_every_cta(cuda_ops):
    """dfine_tc_trace: every CTA of the forward kernel writes ordered phase timestamps; a null buffer switches it off."""
    import ctypes
    from custom_d_fine_b200 import cuda_ops as co
    dev = torch.device("cuda", 0)
    B, H, W, Cin, Cout = 4, 40, 40, 128, 128
    x, w = torch.randn(B, H, W, Cin, device=dev), torch.randn(Cout, Cin, 1, 1, device=dev) * 0.05
    y = torch.empty(B, H, W, Cout, device=dev)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    buf = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
    cache = co._WCache()
    co.lib().dfine_tc_trace(ctypes.c_void_p(buf.data_ptr()))
    try:
        co._conv_fwd(x, Cin, w, cache.getter(w), None, y, Cout, (B, H, W, Cin, H, W, Cout, 1, 1, (0, 0, 0, 0)), 0, stats)
    finally:
        co.lib().dfine_tc_trace(ctypes.c_void_p(0))
    torch.cuda.synchronize()
    t = buf.cpu().view(256, 16)
    live = t[:, 0] > 0
    assert int(live.sum()) >= 1
    for a, b in ((0, 1), (1, 2), (2, 3), (3, 6), (6, 7), (7, 9), (9, 10)):
        assert bool((t[live, a] <= t[live, b]).all()), (a, b)
    before = buf.clone()
    co._conv_fwd(x, Cin, w, cache.getter(w), None, y, Cout, (B, H, W, Cin, H, W, Cout, 1, 1, (0, 0, 0, 0)), 0, stats)
    torch.cuda.synchronize()
    assert torch.equal(before, buf)
