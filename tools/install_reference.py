"""Recipe that puts the UNMODIFIED reference hot path where the GPU box can run it.

    python tools/install_reference.py            # build container only (needs /root/reference)

The reference (ArgoHA/custom_d_fine) is pure Python with no setup.py / pyproject, so `pip install --target` has
nothing to install (recorded in DESIGN.md).  Its train-step half — `src/d_fine/**` (dfine.py, configs.py, matcher.py,
dfine_criterion.py, dist_utils.py, utils.py, arch/*) — imports only torch, torchvision, scipy, numpy and loguru, all of
which are in the image, so it runs as is.  This script copies those files byte for byte into the git-ignored
`baseline/_ref/src/d_fine/` (not gpurun-ignored: the copy travels to the GPU box with the snapshot) together with the
D-FINE-m COCO checkpoint, and writes `baseline/_ref/MANIFEST.json` (sha256 of every file) so that `bench.py --impl
reference`, `cpu_baseline` and `tools/ref_on_gpu.py` can state that what they time is the unmodified reference.
Nothing under `baseline/_ref/` is tracked by git; no reference source enters the repository history.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"


def sha(p: Path) -> str:
    h = hashlib.sha256()
    with p.open("rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def install(verbose: bool = True) -> bool:
    if not (REF / "src" / "d_fine").is_dir():
        if verbose:
            print("install_reference: /root/reference absent (GPU box): using the prebuilt baseline/_ref", file=sys.stderr)
        return (DST / "MANIFEST.json").exists()
    files = {}
    for src in sorted((REF / "src" / "d_fine").rglob("*.py")):
        rel = src.relative_to(REF)
        out = DST / rel
        out.parent.mkdir(parents=True, exist_ok=True)
        if not out.exists() or sha(out) != sha(src):
            shutil.copyfile(src, out)
        files[str(rel)] = sha(out)
    ck = REF / "pretrained" / "dfine_m_coco.pth"
    if ck.exists():
        out = DST / "dfine_m_coco.pth"
        if not out.exists() or out.stat().st_size != ck.stat().st_size:
            shutil.copyfile(ck, out)
        files["dfine_m_coco.pth"] = sha(out)
    commit = None
    sm = REF / ".SUBMODULES.json"
    if sm.exists():
        try:
            commit = json.loads(sm.read_text()).get("commit")
        except Exception:  # noqa: BLE001
            commit = None
    (DST / "MANIFEST.json").write_text(json.dumps({"source": "ArgoHA/custom_d_fine", "commit": commit, "files": files},
                                                  indent=1))
    if verbose:
        print(f"install_reference: {len(files)} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
