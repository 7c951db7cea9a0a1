"""Headline-configuration fixture: the UNMODIFIED reference (/root/reference), D-FINE-m with its own COCO checkpoint
(pretrained/dfine_m_coco.pth), one train step's forward + criterion at 640x640 on the batch of
tests/test_model_gpu.py::test_m_with_pretrained_weights_matches_cpu_oracle.  Losses, indices and a slice of the outputs
only (the checkpoint itself is not redistributed).  Build container only:  python tests/golden/make_golden_m.py"""
import importlib.util
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path = [p for p in sys.path if Path(p or ".").resolve() != HERE.parents[1]]   # see make_golden_seg.py
sys.path.insert(0, "/root/reference")
from src.d_fine.dfine import build_loss, build_model  # noqa: E402

_spec = importlib.util.spec_from_file_location("golden_common", HERE / "common.py")
_common = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_common)
assert "/root/reference" in build_model.__code__.co_filename

if __name__ == "__main__":
    torch.set_num_threads(8)
    hw = 640
    torch.manual_seed(0)
    model = build_model("m", 80, False, "cpu", img_size=(hw, hw),
                        pretrained_model_path="/root/reference/pretrained/dfine_m_coco.pth")
    model.train()
    x, targets = _common.synthetic_batch(2, hw, hw, seed=4321, T=(10, 7))
    torch.manual_seed(7)
    out = model(x, targets=targets)
    crit = build_loss("m", 80, 0.0, False)
    losses = crit(out, targets)
    with torch.no_grad():
        idx = crit.matcher({k: v for k, v in out.items() if "aux" not in k}, targets)["indices"]
    fix = {"hw": hw, "losses": {k: float(v) for k, v in losses.items()},
           "indices": [(i.clone(), j.clone()) for i, j in idx],
           "pred_logits": out["pred_logits"].detach(), "pred_boxes": out["pred_boxes"].detach()}
    torch.save(fix, HERE / "model_m_pretrained_640.pt")
    print(len(fix["losses"]), float(sum(losses.values())), (HERE / "model_m_pretrained_640.pt").stat().st_size)
