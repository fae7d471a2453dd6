"""`torchac` surface used by the reference (models/module_utils.py:28,38; model_compression/model_size_est.py:482,561),
served by the host range coder of liblinr_b200.so (csrc/rc_host.cpp)."""
from ...rc import decode_float_cdf, encode_float_cdf  # noqa: F401
