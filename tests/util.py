"""Shared helpers for the parity tests (CUDA path vs the CPU oracle on the same seeded inputs)."""
import torch


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp_min(1e-12)
    return ((a - b).abs().max() / denom).item()


def check_close(name, got, want, tol):
    assert got.shape == want.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    e = rel_err(got, want)
    assert e <= tol, f"{name}: max|diff|/max|ref| = {e:.3e} > {tol:.1e}"
    return e


def check_l2(name, got, want, tol):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    e = ((got - want).norm() / want.norm().clamp_min(1e-12)).item()
    assert e <= tol, f"{name}: |diff|_2/|ref|_2 = {e:.3e} > {tol:.1e}"
    return e


def run_both(fn, cuda_ops, oracle_ops, inputs, tol_fwd, tol_bwd, seed=0, grad_metric="max"):
    """``fn(K, *tensors)`` -> tensor.  Runs on CPU with the oracle and on cuda:0 with the CUDA table,
    compares the output and the gradients of every floating-point input that requires grad."""
    cpu_in = [t.detach().clone().requires_grad_(t.requires_grad) if isinstance(t, torch.Tensor) else t for t in inputs]
    gpu_in = [t.detach().cuda().requires_grad_(t.requires_grad) if isinstance(t, torch.Tensor) else t for t in inputs]
    y_ref = fn(oracle_ops, *cpu_in)
    y = fn(cuda_ops, *gpu_in)
    errs = {"out": check_close("output", y, y_ref, tol_fwd)}
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs):
        g = torch.Generator().manual_seed(seed + 99)
        go = torch.randn(y_ref.shape, generator=g)
        y_ref.backward(go)
        y.backward(go.cuda())
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip(gpu_in, cpu_in)):
            if isinstance(b, torch.Tensor) and b.requires_grad:
                assert a.grad is not None, f"input {i}: no gradient from the CUDA path"
                chk = check_close if grad_metric == "max" else check_l2
                errs[f"grad{i}"] = chk(f"grad of input {i}", a.grad, b.grad, tol_bwd)
    return errs


def check_rows_up_to_order(name, got, want, tol, min_frac=1.0):
    """[B, Q, C] tensors whose rows (queries) may be permuted per image: the top-k query selection orders
    near-tied encoder scores differently under any floating-point re-association, which permutes
    neighbouring queries without changing the set.  Every row must have a distinct partner within tol."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = want.abs().max().clamp_min(1e-12)
    worst = 0.0
    for b in range(want.shape[0]):
        dist = torch.cdist(got[b], want[b], p=float("inf")) / scale
        vals, idx = dist.min(1)
        if min_frac >= 1.0:
            assert len(set(idx.tolist())) == want.shape[1], f"{name}: image {b}: rows do not pair up one-to-one"
            worst = max(worst, vals.max().item())
        else:
            # reduced-precision runs may swap a few queries at the top-k boundary (rank ~300 of the encoder
            # scores): require min_frac of the rows to have a distinct partner within tol
            good = vals <= tol
            n_ok = len(set(idx[good].tolist()))
            assert n_ok >= min_frac * want.shape[1], f"{name}: image {b}: only {n_ok}/{want.shape[1]} rows pair up"
            worst = max(worst, vals[good].max().item() if good.any() else 0.0)
    assert worst <= tol, f"{name}: max row distance {worst:.3e} > {tol:.1e}"
    return worst
