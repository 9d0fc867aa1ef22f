"""ctypes binding of libb2s.so (the C ABI declared in include/b2s.h).

There is deliberately NO fallback: if the CUDA library is missing or an entry
point reports an error, the call raises.  The product path never routes through
``oracle/`` or any CPU implementation.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libb2s.so")

_P = c_void_p
_SIGNATURES = {
    # name: (restype, [argtypes])
    "b2s_last_error": (c_char_p, []),
    "b2s_version": (c_int32, []),
    "b2s_set_sm_reserve": (None, [c_int32]),
    "b2s_hash": (c_int32, [_P, c_int64, _P, _P]),
    "b2s_kernel_hash": (c_int32, [_P, c_int64, _P, c_int32, _P, _P]),
    "b2s_table_slots": (c_int64, [c_int64]),
    "b2s_table_bytes": (c_size_t, [c_int64]),
    "b2s_table_build": (c_int32, [_P, c_int64, _P, c_size_t, _P]),
    "b2s_table_build_coords": (c_int32, [_P, c_int64, _P, c_size_t, _P]),
    "b2s_table_query": (c_int32, [_P, c_int64, _P, c_int64, _P, _P]),
    "b2s_count": (c_int32, [_P, c_int64, _P, c_int64, _P]),
    "b2s_unique_workspace_bytes": (c_size_t, [c_int64]),
    "b2s_unique_i64": (c_int32, [_P, c_int64, _P, _P, _P, c_size_t, _P]),
    "b2s_downsample_capacity": (c_int64, [c_int64, _P, _P]),
    "b2s_downsample_workspace_bytes": (c_size_t, [c_int64, _P, _P]),
    "b2s_downsample_coords": (c_int32, [_P, c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "b2s_kmap_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "b2s_kmap_build": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int32, _P, _P, _P, _P, _P, _P, c_size_t,
                                 _P]),
    "b2s_kmap_pairs": (c_int32, [_P, c_int32, c_int64, _P, _P, _P, c_size_t, _P]),
    "b2s_kmap_pairs_chunked_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "b2s_kmap_pairs_chunked": (c_int32, [_P, c_int32, c_int64, _P, c_int32, _P, _P, _P, _P, c_size_t, _P]),
    "b2s_conv_wgrad_segments": (c_int32, [c_int32, _P, c_int64, _P, c_int64, c_int32, c_int32, c_int32, _P, _P,
                                          c_int32, c_int32, _P, _P]),
    "b2s_conv_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int32, c_int32, c_int32]),
    "b2s_conv_gather_gemm": (c_int32, [c_int32, _P, c_int64, _P, c_int32, c_int32, c_int32, c_int32,
                                       c_int32, _P, _P, _P, c_int64, _P, _P, _P, c_size_t, _P]),
    "b2s_conv_steps_supported": (c_int32, [c_int32, c_int64, c_int32, c_int32]),
    "b2s_conv_tile_rows": (c_int32, [c_int32, c_int64]),
    "b2s_weight_to_kmajor": (c_int32, [_P, c_int32, c_int32, c_int32, _P, _P]),
    "b2s_weights_refresh": (c_int32, [_P, c_int32, c_int64, _P]),
    "b2s_conv_gather_gemm_steps": (c_int32, [c_int32, _P, c_int64, _P, c_int32, c_int32, c_int32, c_int32,
                                             c_int32, c_int32, _P, _P, _P, _P, c_int32, _P, c_int64, _P, _P, _P,
                                             _P, c_size_t, _P]),
    "b2s_tile_order_key_bits": (c_int32, [_P, c_int32, c_int64, _P, _P, c_int32, _P, _P, _P]),
    "b2s_tile_steps": (c_int32, [_P, c_int32, c_int64, _P, _P, c_int32, _P, _P, _P, _P]),
    "b2s_tile_mask": (c_int32, [_P, c_int32, c_int64, _P, _P]),
    "b2s_tile_order_key": (c_int32, [_P, c_int32, c_int64, _P, _P, c_int32, _P, _P]),
    "b2s_conv_wgrad": (c_int32, [c_int32, _P, c_int64, _P, c_int64, c_int32, c_int32, c_int32, _P, _P,
                                 c_int32, _P, _P, c_size_t, _P]),
    "b2s_voxelize_fwd": (c_int32, [c_int32, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P, _P]),
    "b2s_voxelize_bwd": (c_int32, [c_int32, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P]),
    "b2s_devoxelize_fwd": (c_int32, [c_int32, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P]),
    "b2s_devoxelize_bwd": (c_int32, [c_int32, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P, _P]),
    "b2s_devoxelize_bwd_sorted": (c_int32, [c_int32, _P, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P, _P]),
    "b2s_scatter_max": (c_int32, [c_int32, _P, _P, c_int64, c_int32, c_int64, _P, _P, _P, _P]),
    "b2s_trilinear_map": (c_int32, [_P, c_int64, c_int32, _P, c_int64, _P, _P, _P]),
    "b2s_ti_weights": (c_int32, [_P, c_int64, _P, c_float, _P, _P]),
    "b2s_bn_supported": (c_int32, [c_int32, c_int32]),
    "b2s_bn_forward": (c_int32, [c_int32, _P, _P, c_int64, c_int32, _P, _P, c_float, c_float, _P, _P,
                                 c_int32, _P, _P, _P, _P, _P, _P]),
    "b2s_bn_forward_sums": (c_int32, [c_int32, _P, _P, c_int64, c_int32, _P, _P, c_float, c_float, _P, _P,
                                      c_int32, _P, _P, _P, _P, _P, c_int32, _P]),
    "b2s_bn_stats": (c_int32, [c_int32, _P, c_int64, c_int32, _P, _P]),
    "b2s_bn_backward_reduce": (c_int32, [c_int32, _P, _P, _P, c_int64, c_int32, _P, _P, c_int32, _P, _P, _P]),
    "b2s_bn_backward_apply": (c_int32, [c_int32, _P, _P, _P, c_int64, c_int32, _P, _P, _P, c_int32, _P, _P, _P, _P,
                                        _P, _P]),
    "b2s_bn_backward": (c_int32, [c_int32, _P, _P, _P, c_int64, c_int32, _P, _P, _P, c_int32, _P, _P, _P,
                                  _P]),
    "b2s_map_count": (c_int32, [_P, c_int64, c_int32, c_int32, c_int32, _P, _P]),
    "b2s_denselize_fwd": (c_int32, [_P, _P, _P, c_int64, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "b2s_denselize_bwd": (c_int32, [_P, _P, _P, c_int64, c_int32, c_int32, c_int32, c_int32, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class B2SError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load libb2s.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2SError(
                f"{LIB_PATH} not found: build the CUDA library first "
                "(python -m openpcseg_b200.build, or __graft_entry__.build()). "
                "There is no CPU fallback for the sparse-voxel hot path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)        # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().b2s_last_error().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError(f"libb2s {what}: {msg}")
        raise B2SError(f"libb2s {what} failed (status {rc}): {msg}")
