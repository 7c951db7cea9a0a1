"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/lsap.c (see its header)."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "liboracle_lsap.so"
_lib = None


def build(force=False):
    src = _HERE / "lsap.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-D_GNU_SOURCE", "-o", str(_SO), str(src), "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
        _lib.lsap_solve.restype = ctypes.c_int
        _lib.lsap_solve.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def lsap(cost):
    """(row_ind, col_ind) int64 arrays with scipy.optimize.linear_sum_assignment semantics."""
    c = np.ascontiguousarray(np.asarray(cost, dtype=np.float64))
    if c.ndim != 2:
        raise ValueError("expected a matrix")
    nr, nc = c.shape
    n = min(nr, nc)
    rows = np.zeros(n, dtype=np.int64)
    cols = np.zeros(n, dtype=np.int64)
    rc = _load().lsap_solve(nr, nc, c.ctypes.data, rows.ctypes.data, cols.ctypes.data)
    if rc == -1:
        raise ValueError("cost matrix is infeasible")
    if rc == -2:
        raise ValueError("matrix contains invalid numeric entries")
    return rows, cols
