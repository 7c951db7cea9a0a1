"""Drop-in for /root/reference/src/d_fine/dfine.py:19-124 — same names and signatures; the modules behind them run
on the sm_100a CUDA library (custom_d_fine_b200)."""
from custom_d_fine_b200.model import DFINE, build_loss, build_model, build_optimizer  # noqa: F401

__all__ = ["DFINE", "build_model", "build_loss", "build_optimizer"]
