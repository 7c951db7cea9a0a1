"""D-FINE-m with the reference's COCO checkpoint, one train step on a seeded 640x640 batch: CUDA library (each GEMM
mode) against the CPU oracle.  Prints the error metrics the parity bars are set from."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from custom_d_fine_b200 import cuda_ops as co, kernels  # noqa: E402
from custom_d_fine_b200.model import build_loss, build_model  # noqa: E402
from oracle.torch_ops import OracleOps  # noqa: E402
from tests.golden.common import synthetic_batch  # noqa: E402
from tests.test_model_gpu import _host_rng  # noqa: E402

ckpt = ROOT / "baseline" / "_ref" / "dfine_m_coco.pth"
hw = 640
x, targets = synthetic_batch(2, hw, hw, seed=4321, T=(10, 7))


def run(dev, mode=None):
    if mode:
        co.set_gemm_mode(mode)
    torch.manual_seed(0)
    model = build_model("m", 80, False, dev, img_size=(hw, hw), pretrained_model_path=str(ckpt))
    model.train()
    xs = x.to(dev)
    tg = [{k: v.to(dev) for k, v in t.items()} for t in targets]
    crit = build_loss("m", 80, 0.0, False)
    torch.manual_seed(7)
    with _host_rng():
        if dev == "cpu":
            with kernels.use(OracleOps()):
                out = model(xs, targets=tg)
                losses = crit(out, tg)
        else:
            out = model(xs, targets=tg)
            losses = crit(out, tg)
            torch.cuda.synchronize()
    return out, losses


o0, l0 = run("cpu")
co.CudaOps()
for mode in ("simt", "tc3", "hf3", "tch", "bf3", "tc"):
    o1, l1 = run("cuda", mode)
    line = [mode]
    for key in ("pred_logits", "pred_boxes"):
        a, b = o1[key].detach().double().cpu(), o0[key].detach().double()
        # rows may be permuted by the top-k: pair each row with its nearest reference row
        best = []
        for i in range(a.shape[0]):
            d = torch.cdist(torch.cat([o1["pred_logits"][i], o1["pred_boxes"][i]], -1).double().cpu(),
                            torch.cat([o0["pred_logits"][i], o0["pred_boxes"][i]], -1).double(), p=float("inf"))
            best.append(d.argmin(1))
        bp = torch.stack([b[i][best[i]] for i in range(a.shape[0])])
        diff = (a - bp)
        line.append(f"{key}: max|d|={float(diff.abs().max()):.3e} max|ref|={float(bp.abs().max()):.3e} "
                    f"relL2={float(diff.norm() / bp.norm()):.3e} mean|d|={float(diff.abs().mean()):.3e}")
    lossd = max(abs(float(l1[k]) - float(l0[k])) / max(abs(float(l0[k])), 1e-2) for k in l0)
    line.append(f"worst loss term rel diff {lossd:.3e}; total {float(sum(l1.values())):.5f} vs {float(sum(l0.values())):.5f}")
    print(" | ".join(line), flush=True)
