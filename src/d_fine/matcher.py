"""Drop-in for /root/reference/src/d_fine/matcher.py:74-257: ``HungarianMatcher(weight_dict, use_focal_loss, alpha,
gamma).forward(outputs, targets) -> {"indices": [(q_idx int64 CPU ascending, t_idx int64 CPU)]}`` — cost blocks and
the assignment solved on the GPU in one launch."""
from custom_d_fine_b200.matcher import HungarianMatcher  # noqa: F401
