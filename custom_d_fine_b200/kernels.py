"""Kernel table ``K`` used by the host graph.

``K.<op>(...)`` resolves to the active provider.  The product has exactly one
provider — :class:`custom_d_fine_b200.cuda_ops.CudaOps`, which calls the sm_100a
C-ABI library and raises if it is missing.  ``use(provider)`` exists so that the
test-suite can drive the *host* logic (graph wiring, state-dict layout, criterion
bookkeeping) with the CPU oracle on a GPU-less box; nothing in the package ever
installs another provider by itself.
"""
from __future__ import annotations

import contextlib
import threading

_lock = threading.Lock()
_active = None


def _default():
    from . import cuda_ops  # noqa: WPS433  (deliberately lazy: needs the built .so)

    return cuda_ops.CudaOps()


def get():
    global _active
    if _active is None:
        with _lock:
            if _active is None:
                _active = _default()
    return _active


@contextlib.contextmanager
def use(provider):
    """Temporarily install ``provider`` (tests only)."""
    global _active
    prev = _active
    _active = provider
    try:
        yield provider
    finally:
        _active = prev


class _Proxy:
    def __getattr__(self, name):
        return getattr(get(), name)


K = _Proxy()
