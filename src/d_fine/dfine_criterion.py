"""Drop-in for /root/reference/src/d_fine/dfine_criterion.py:21-864 (same loss keys, order and values)."""
from custom_d_fine_b200.criterion import DFINECriterion  # noqa: F401
