"""linr-pcgc_b200: B200-native overfit + coding hot path of LINR-PCGC."""
__version__ = "0.1.0"

import os as _os

# A training call overlaps its weight-gradient launches with the grad-input chain on a second stream (linr_ctx,
# include/linr_b200.h).  CUDA maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); with the coder /
# decoder worker streams alive two streams can share a queue and serialise (measured: 961 vs 885 ms/step for a GOP).
# Read by the driver when the CUDA context is created, so it has to be set before the first CUDA call.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
