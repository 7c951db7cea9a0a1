"""Drop-in for /root/reference/src/d_fine/dist_utils.py:13-205 (train-step subset; eval gathers are out of scope)."""
from custom_d_fine_b200.dist import (  # noqa: F401
    broadcast_scalar, cleanup_distributed, get_local_rank, get_rank, get_world_size, init_distributed_mode,
    is_dist_available_and_initialized, is_main_process, reduce_dict, synchronize)
