"""Drop-in for /root/reference/src/d_fine/configs.py:1-213 (``base_cfg``, per-size dicts merged into ``models``)."""
from custom_d_fine_b200.specs import base_cfg, merge_configs, models, sizes_cfg  # noqa: F401
