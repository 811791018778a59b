"""ctypes binding of libclipself_b200.so (the C ABI declared in include/clipself_b200.h).

The product path has no CPU implementation: if the library is missing or a call fails this module
raises — it never falls back to PyTorch ops.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libclipself_b200.so")

CS_F32, CS_BF16 = 0, 1
EPI_STORE, EPI_QKV_ROPE, EPI_SWIGLU, EPI_TOKENS = 0, 1, 2, 3

vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double


class GemmEpilogue(C.Structure):
    """cs_gemm_epilogue_t"""
    _fields_ = [("mode", C.c_int32), ("out_dtype", C.c_int32), ("out", vp), ("ldo", i64),
                ("bias", vp), ("residual", vp), ("ldr", i64), ("rope_pos", vp), ("rope_freq", vp),
                ("rope_grid", C.c_int32), ("tokens", C.c_int32), ("rope_cols", C.c_int32), ("pos_embed", vp),
                ("alpha", f32), ("reserved", C.c_int32), ("ln_stats", vp), ("ln_c1", vp), ("ln_parts", C.c_int32),
                ("ln_dim", C.c_int32), ("ln_eps", f32), ("reserved2", C.c_int32), ("stats_out", vp),
                ("out2_bf16", vp), ("ldo2", i64)]


# name -> argtypes, exactly the prototypes of include/clipself_b200.h
PROTOTYPES = {
    "cs_abi_version": [],
    "cs_device_info": [C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)],
    "cs_tensor_map_encodes": [],
    "cs_extract_rois": [vp, i32, i32, vp, vp, vp, vp, vp],
    "cs_gather_rows": [vp, vp, i32, i64, vp, vp],
    "cs_roi_align_fwd": [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp],
    "cs_roi_align_bwd": [vp, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp],
    "cs_mask_pool_fwd": [vp, i32, i32, i32, vp, vp, i32, vp, vp],
    "cs_cosine_loss_fwd": [vp, vp, i32, i32, f32, vp, vp, vp],
    "cs_cosine_loss_bwd": [vp, vp, vp, i32, i32, f32, vp, vp, vp],
    "cs_l2norm_fwd": [vp, i64, i32, vp, vp, vp],
    "cs_l2norm_bwd": [vp, vp, vp, i64, i32, vp, vp],
    "cs_crop_workspace_bytes": [i32, i32, i32, i32, C.POINTER(C.c_int64)],
    "cs_crop_resize_normalize": [vp, i32, i32, vp, i32, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), vp, vp,
                                 i64, vp],
    "cs_crop_resize_normalize_batched": [vp, vp, vp, vp, vp, i32, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), vp,
                                         vp, i64, vp],
    "cs_im2col_patches": [vp, i32, i32, i32, i32, vp, i64, vp],
    "cs_resize_bilinear": [vp, i32, i64, i32, i32, i32, i32, vp, vp],
    "cs_fill_cls_rows": [vp, vp, i32, i32, i32, vp, vp],
    "cs_layernorm_fwd": [vp, i32, i64, i64, i32, i32, i32, i32, vp, vp, f32, vp, i64, vp, vp, vp],
    "cs_row_stats_cast": [vp, i64, i64, i32, vp, i64, vp, i32, vp],
    "cs_gemm_bf16": [vp, i64, vp, i64, i64, i32, i32, C.POINTER(GemmEpilogue), vp],
    "cs_gemm_bf16_tn": [vp, i64, vp, i64, i64, i32, i32, C.POINTER(GemmEpilogue), vp],
    "cs_pack_swiglu_weights": [vp, vp, i32, i32, i32, vp, vp, vp, i64, vp, vp],
    "cs_attention_fwd": [vp, i32, i32, i32, f32, vp, vp, vp, vp],
    "cs_attention_cls_fwd": [vp, i32, i32, i32, f32, vp, vp, vp],
    "cs_cast_pad_bf16": [vp, i64, i64, i64, vp, i64, vp],
    "cs_attention_bwd": [vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, vp, vp],
    "cs_cast_transpose_bf16": [vp, i32, i64, i32, i64, vp, i64, vp, i64, vp],
    "cs_layernorm_bwd_dx": [vp, i32, i64, vp, i32, i64, i64, i32, i32, i32, i32, vp, vp, vp, vp, i64, vp, i32,
                            i64, i32, vp],
    "cs_col_reduce": [vp, i32, i64, vp, i32, i64, i64, i32, i32, i32, i32, vp, vp, vp, vp, vp, i64, vp],
    "cs_swiglu_fwd": [vp, i64, i32, i64, vp, i64, i32, vp],
    "cs_swiglu_bwd": [vp, vp, i64, i32, i64, i64, vp, i32, vp],
    "cs_adamw_step": [vp, vp, vp, vp, i64, f64, f64, f64, f64, f64, i32, f64, vp],
}

_lib = None


class TowerCfgC(C.Structure):
    """cs_tower_cfg_t"""
    _fields_ = [(n, C.c_int32) for n in ("image_size", "patch", "width", "heads", "layers", "hidden", "embed_dim", "pt_seq_len")] + \
               [("ln_eps", f32)]


PROTOTYPES.update({
    "cs_pack_weights_bytes": [C.POINTER(TowerCfgC), C.POINTER(i64)],
    "cs_pack_weights_create": [C.POINTER(TowerCfgC), C.POINTER(C.c_char_p), C.POINTER(vp), i32, vp, i64, vp, C.POINTER(vp)],
    "cs_pack_weights_update": [vp, C.POINTER(C.c_char_p), C.POINTER(vp), i32, vp],
    "cs_pack_weights_destroy": [vp],
    "cs_query_workspace": [C.POINTER(TowerCfgC), i32, i32, C.POINTER(i64)],
    "cs_vit_forward_cls": [vp, vp, i32, i32, vp, i64, i32, vp, vp],
    "cs_vit_forward_dense": [vp, vp, i32, i32, i32, vp, vp, i64, i32, vp, vp],
})


class ClipselfB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ClipselfB200Error(
                f"{LIB_PATH} is missing: build it with `python -m clipself_b200.build` "
                "(there is no CPU / PyTorch fallback for this path)")
        l = C.CDLL(LIB_PATH)
        l.cs_last_error.restype = C.c_char_p
        l.cs_last_error.argtypes = []
        for name, args in PROTOTYPES.items():
            fn = getattr(l, name)          # AttributeError if the .so does not export it
            fn.restype = C.c_int64 if name == "cs_tensor_map_encodes" else C.c_int
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise ClipselfB200Error(f"{what} failed (code {rc}): {lib().cs_last_error().decode()}")


# kernels launched by one successful call of each entry point (for bench.py's `gpu_launches`)
KERNELS_PER_CALL = {"cs_crop_resize_normalize": 3, "cs_crop_resize_normalize_batched": 3, "cs_crop_workspace_bytes": 0,
                    "cs_roi_align_fwd": 2, "cs_cosine_loss_fwd": 2, "cs_col_reduce": 2, "cs_attention_bwd": 3,
                    "cs_abi_version": 0, "cs_device_info": 0, "cs_pack_weights_bytes": 0, "cs_query_workspace": 0,
                    "cs_pack_weights_destroy": 0, "cs_pack_weights_create": 0, "cs_pack_weights_update": 0,
                    "cs_vit_forward_cls": 0, "cs_vit_forward_dense": 0}       # tower calls: counted by the caller
launch_count = 0


def call(name: str, *args) -> None:
    global launch_count
    check(getattr(lib(), name)(*args), name)
    launch_count += KERNELS_PER_CALL.get(name, 1)


_device_info = {}


def require_device() -> dict:
    """Raise unless the current CUDA device is an sm_100 part.  The answer is cached per device once it is positive:
    cs_device_info queries the device properties, which costs 10 ms and more per call (it sat in the per-step crop path)."""
    try:
        import torch
        key = torch.cuda.current_device() if torch.cuda.is_available() else -1
    except Exception:           # noqa: BLE001
        key = -1
    if key >= 0 and key in _device_info:
        return _device_info[key]
    sm, n, mem = i32(), i32(), i64()
    call("cs_device_info", C.byref(sm), C.byref(n), C.byref(mem))
    info = dict(sm=sm.value, num_sms=n.value, hbm_bytes=mem.value)
    if key >= 0:
        _device_info[key] = info
    return info
