"""Drop-in for /root/reference/src/d_fine/utils.py:92-181 (checkpoint loading with the obj365->coco head remap)."""
from custom_d_fine_b200.model import load_tuning_state  # noqa: F401
