"""ctypes binding of liblinr_b200.so (the C ABI declared in include/linr_b200.h).

There is no CPU fallback: if the library is missing the import raises, and every device entry point
raises when handed anything but CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "lib", "liblinr_b200.so")


class LinrError(RuntimeError):
    pass


class Rows(C.Structure):
    """struct linr_rows (include/linr_b200.h)."""
    _fields_ = [("n_rows", C.c_int64), ("ld", C.c_int64), ("d_anchor", C.c_void_p), ("d_mask", C.c_void_p),
                ("d_nbr7", C.c_void_p), ("d_scale", C.c_void_p), ("d_occ", C.c_void_p), ("d_tile_rng", C.c_void_p),
                ("d_pair_cnt", C.c_void_p), ("d_pair_list", C.c_void_p)]


_P, _I64, _I, _F, _SZ = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t
_RP = C.POINTER(Rows)

# name -> (restype, argtypes).  Every symbol include/linr_b200.h declares is listed here.
SIGNATURES = {
    "linr_version": (_I, []),
    "linr_last_error": (C.c_char_p, []),
    "linr_device_info": (_I, [_I, C.POINTER(_I), C.POINTER(_I64)]),
    "linr_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "linr_ctx_destroy": (_I, [_P]),
    "linr_ctx_set_current": (_I, [_P]),
    "linr_ctx_hint_same_params": (_I, [_P]),
    "linr_ctx_bank_calls": (_I64, [_P]),
    "linr_ctx_bank_launches": (_I64, [_P]),
    "linr_side_stream_enable": (_I, [_I]),
    "linr_prof_enable": (_I, [C.c_uint32]),
    "linr_prof_read": (_I, [_I, C.POINTER(C.c_double), C.POINTER(_I64), C.POINTER(_I64)]),
    "linr_prof_classes": (_I, []),
    "linr_prof_name": (C.c_char_p, [_I]),
    "linr_coord_ws_bytes": (_SZ, [_I64]),
    "linr_coord_sort_unique": (_I, [_P, _I64, _I, _P, _P, _P, _SZ, _P]),
    "linr_coord_sort": (_I, [_P, _I64, _I, _P, _P, _SZ, _P]),
    "linr_coord_min_sub": (_I, [_P, _I64, _P, _P, _P]),
    "linr_octree_down": (_I, [_P, _I64, _I, _P, _P, _P, _P, _SZ, _P]),
    "linr_octree_up_count": (_I, [_P, _I64, _P, _P, _SZ, _P]),
    "linr_octree_up_expand": (_I, [_P, _P, _P, _I64, _I64, _I, _P, _P, _SZ, _P]),
    "linr_hash_bytes": (_SZ, [_I64]),
    "linr_hash_build": (_I, [_P, _P, _I64, _P, _I64, _P]),
    "linr_nbr_build": (_I, [_P, _P, _I64, _P, _I64, _P, _P, _I64, _P, _P, _P]),
    "linr_tile_ranges": (_I, [_RP, _P, _P]),
    "linr_pair_lists": (_I, [_RP, _P, _P, _P]),
    "linr_pair_list_order": (None, [_P]),
    "linr_hash_lookup": (_I, [_P, _P, _I64, _P, _I64, _P, _P]),
    "linr_param_count": (_I64, [_I]),
    "linr_param_offsets": (_I, [_I, C.POINTER(_I64), _I]),
    "linr_net_ws_bytes": (_SZ, [_I64, _I]),
    "linr_net_forward": (_I, [_P, _I, _RP, _I, _F, _P, _P, _P, _P, _SZ, _P]),
    "linr_net_backward": (_I, [_P, _I, _RP, _P, _P, _SZ, _P]),
    "linr_net_forward_stages": (_I, [_P, _I, _RP, _I, _I, _I, _I, _F, _P, _P, _P, _P, _SZ, _P]),
    "linr_net_backward_stages": (_I, [_P, _I, _RP, _I, _I, _I, _I, _P, _P, _SZ, _P]),
    "linr_net_ws_offsets": (_I, [_I64, _I, _I, C.POINTER(_I64), C.POINTER(_I64)]),
    "linr_net_decode_begin": (_I, [_P, _I, _RP, _P, _SZ, _P]),
    "linr_net_decode_stage": (_I, [_P, _I, _RP, _I, _P, _P, _P, _SZ, _P]),
    "linr_occ_set_stage": (_I, [_P, _P, _I64, _I, _P]),
    "linr_net_decode_scale": (_I, [_P, _I, _RP, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "linr_spconv27_fwd": (_I, [_P, _I, _P, _P, _P, _I, _RP, _I, _P]),
    "linr_spconv27_bwd_in": (_I, [_P, _I, _P, _P, _I, _RP, _P]),
    "linr_spconv27_bwd_w_ws_bytes": (_SZ, [_I64, _I, _I]),
    "linr_spconv27_bwd_w": (_I, [_P, _I, _P, _I, _RP, _P, _P, _P, _SZ, _P]),
    "linr_adam_fused": (_I, [_P, _P, _P, _P, _I64, _I64, _F, _F, _F, _F, _F, _P]),
    "linr_param_quant": (_I, [_P, _I64, _I, _P, _P, _P, _P]),
    "linr_param_quant16": (_I, [_P, _I64, _I, _P, _P, _P, _P]),
    "linr_rc_encode_binary": (_I64, [_P, _P, _I64, _P, _I64]),
    "linr_rc_decode_binary": (_I, [_P, _P, _I64, _P, _I64]),
    "linr_rc_decode_binary_batch": (_I, [_I, _P, _P, _P, _P, _P, _I]),
    "linr_net_decode_scale_batch": (_I, [_P, _I, _RP, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _SZ, _P]),
    "linr_rc_encode_binary_batch": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _I]),
    "linr_rc_encode_shared": (_I64, [_P, _I, _P, _I64, _P, _I64]),
    "linr_rc_decode_shared": (_I, [_P, _I, _P, _I64, _P, _I64]),
}

_lib = None


def load():
    """Load the shared library (raises if it has not been built: no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise LinrError(f"{SO} is missing: run `python linr-pcgc_b200/build.py` (nvcc, sm_100a). "
                            "linr_pcgc_b200 has no CPU fallback.")
        lib = C.CDLL(SO)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise LinrError(f"{what} failed ({rc}): {load().linr_last_error().decode()}")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL). Refuses host tensors: no silent CPU path."""
    if t is None:
        return None
    if not t.is_cuda:
        raise LinrError("linr_pcgc_b200 device entry points take CUDA tensors only")
    if not t.is_contiguous():
        raise LinrError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
