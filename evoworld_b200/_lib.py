"""ctypes binding of libevoworld_b200.so (the C ABI declared in include/evoworld_b200.h).

Fails loudly: there is no CPU fallback.  `lib()` raises if the shared library is missing and every
call goes through `check()` which raises RuntimeError with evw_last_error() on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os as _os

# EVW_LIB: load another build of the same library (A/B experiments with differently compiled kernels); default = the in-tree build
_LIB_PATH = Path(_os.environ["EVW_LIB"]) if _os.environ.get("EVW_LIB") else Path(__file__).resolve().parent / "_lib" / "libevoworld_b200.so"
_lib = None

c_void_p, c_int, c_i64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/evoworld_b200.h declares
SIGNATURES = {
    "evw_last_error": (C.c_char_p, []),
    "evw_abi_version": (c_int, []),
    "evw_device_info": (c_int, [C.POINTER(c_int), C.POINTER(c_i64), C.POINTER(c_int)]),
    "evw_plucker": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "evw_equi2pers_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "evw_equi2pers_table": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "evw_equi2pers_yaw_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "evw_lift_depth": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "evw_lift_pack_points": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "evw_pack_points": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_i64, c_void_p]),
    "evw_conf_select_workspace": (c_i64, [c_i64]),
    "evw_conf_select": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_float, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "evw_splat_workspace": (c_i64, [c_int, c_int]),
    "evw_splat_workspace_flags": (c_i64, [c_int, c_int, c_int]),
    "evw_set_splat_ctas_per_sm": (None, [c_int]),
    "evw_splat_cubemap_equirect": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_int, c_int, c_float, c_float,
                                           c_void_p, c_int, c_int, c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "evw_splat_faces_u8": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_int, c_float, c_float, c_void_p,
                                   c_void_p, c_i64, c_void_p]),
    "evw_cube_to_equirect_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "evw_gemm_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             C.c_char_p, c_void_p, c_int, c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_int, c_float,
                             c_void_p, c_float, c_float, c_int, c_int, c_void_p, c_void_p]),
    "evw_gemm_f16_gn": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                C.c_char_p, c_void_p, c_int, c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_int, c_float,
                                c_void_p, c_float, c_float, c_int, c_int, c_void_p, c_void_p, c_i64, c_void_p]),
    "evw_upconv2x_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "evw_spatial_attention_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "evw_set_attention_variant": (None, [c_int]),
    "evw_set_gemm_cluster": (None, [c_int]),
    "evw_set_gemm_gn_stats": (None, [c_int]),
    "evw_set_gemm_store_tma": (None, [c_int]),
    "evw_temporal_attention_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_i64, c_int, c_void_p]),
    "evw_group_norm_f16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_i64, c_i64, c_float, c_void_p, c_void_p,
                                   c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "evw_layer_norm_f16": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_float, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "evw_layer_norm_f32": (c_int, [c_void_p, c_i64, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "evw_unet_create": (c_int, [C.POINTER(c_void_p), C.POINTER(c_int), c_int, C.POINTER(c_float), c_int,
                                C.POINTER(C.c_char_p), C.POINTER(c_void_p), c_int, C.POINTER(C.c_char_p),
                                C.POINTER(C.c_double), c_int]),
    "evw_unet_destroy": (c_int, [c_void_p]),
    "evw_unet_workspace_bytes": (c_i64, [c_void_p, c_int, c_int, c_int, c_int]),
    "evw_unet_forward": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                 c_void_p, c_i64, c_void_p]),
    "evw_denoise_step": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_float, c_float,
                                 c_int, c_int, c_int, c_void_p, c_i64, c_void_p]),
    "evw_unet_plan_info": (c_int, [c_void_p, C.POINTER(c_i64), C.POINTER(C.c_double)]),
    "evw_resize_pil_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_int, c_void_p]),
    "evw_small_attention_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "evw_act_f16": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "evw_vae_create": (c_int, [C.POINTER(c_void_p), C.POINTER(c_int), c_int, C.POINTER(C.c_char_p), C.POINTER(c_void_p), c_int,
                               C.POINTER(C.c_char_p), C.POINTER(C.c_double), c_int]),
    "evw_vae_destroy": (c_int, [c_void_p]),
    "evw_vae_workspace_bytes": (c_i64, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "evw_vae_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_i64, c_void_p]),
    "evw_vae_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_i64, c_void_p]),
    "evw_vae_plan_info": (c_int, [c_void_p, c_int, C.POINTER(c_i64), C.POINTER(C.c_double), C.POINTER(c_i64)]),
    "evw_unet_graph_replays": (c_i64, [c_void_p]),
    "evw_unet_gn_fused": (c_i64, [c_void_p]),
    "evw_splat_cube_equirect": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p,
                                        c_int, c_int, c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p]),
    "evw_splat_cube_faces_debug": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_int, c_float, c_float, c_void_p,
                                           c_void_p, c_i64, c_void_p]),
    "evw_splat_faces_debug": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_int, c_float, c_float, c_void_p,
                                      c_void_p, c_i64, c_void_p]),
    "evw_qknorm_rope_f16": (c_int, [c_void_p, c_i64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_float, c_void_p]),
    "evw_bilinear_ac_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "evw_patchify_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, C.POINTER(c_float), C.POINTER(c_float), c_void_p]),
    "evw_relu_inplace_f16": (c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    "evw_adaln_modulate_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "evw_dpt_activate_f32": (c_int, [c_void_p, c_i64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -m evoworld_b200.build` "
                "(or __graft_entry__.build()). evoworld_b200 has no CPU fallback."
            )
        handle = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().evw_last_error().decode(errors="replace")
        raise RuntimeError(f"evoworld_b200 {what} failed with status {status}: {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (must be contiguous) or None."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("evoworld_b200: tensor passed to the C ABI must be contiguous")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"evoworld_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
