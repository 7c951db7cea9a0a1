"""custom_d_fine_b200 — B200-native D-FINE train/inference hot path.

Host side is Python/PyTorch (tensors, streams, autograd graph, torch.distributed);
every compute op goes through ``kernels.K`` which resolves to the sm_100a CUDA
library ``csrc/libdfine_sm100.so`` (C ABI declared in ``include/dfine_sm100.h``).
There is no CPU compute path in this package: without the CUDA library the ops
raise.  The drop-in surface of the reference lives in ``src/d_fine`` / ``src/dl``.
"""

__version__ = "0.1.0"
