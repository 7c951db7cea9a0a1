"""Builds ``csrc/libdfine_sm100.so`` (sm_100a only) with nvcc — in-tree, no JIT cache.

``python -m custom_d_fine_b200.build`` or ``build()`` from ``__graft_entry__``.  nvcc
cross-compiles without a GPU.  Objects are rebuilt only when their source (or a header) is
newer.  ``matcher.cu`` is compiled with ``-fmad=false`` so that the fp32 cost arithmetic keeps
the reference's operation-by-operation rounding (matcher.py:135-172).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libdfine_sm100.so"
SOURCES = ["lib.cu", "msda.cu", "matcher.cu", "norm_act.cu", "spatial.cu", "conv_simt.cu", "attention.cu", "attention_mma.cu",
           "gemm_tc.cu", "fdr.cu", "stem.cu", "optim.cu", "loss.cu", "select.cu", "io.cu", "seg.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]
PER_FILE = {"matcher.cu": ["-fmad=false"]}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((CSRC.parents[1] / "include").glob("dfine_loss_desc.h"))
    srcs = [s for s in SOURCES if (CSRC / s).exists()]
    objs, jobs = [], []
    for s in srcs:
        src, obj = CSRC / s, CSRC / (Path(s).stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, *headers]):
            jobs.append([nvcc, *ARCH, *COMMON, *PER_FILE.get(s, []), "-c", str(src), "-o", str(obj)])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([nvcc, *ARCH, "-shared", "-o", str(LIB), *map(str, objs)])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
