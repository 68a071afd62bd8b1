"""ctypes binding of libidf_b200.so (see include/idf_b200.h). The library is the only compute path: if it is
missing or a call fails, an exception is raised — there is no eager/PyTorch or CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IDF_B200_LIB") or os.path.join(_HERE, "libidf_b200.so")  # (override: debug builds)

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class NativeError(RuntimeError):
    pass


class NHWC(C.Structure):
    _fields_ = [("ptr", _vp), ("n", _i32), ("h", _i32), ("w", _i32), ("c", _i32),
                ("sn", _i64), ("sh", _i64), ("sw", _i64)]


class IgemmArgs(C.Structure):
    _fields_ = [
        ("a", NHWC * 2), ("taps", _i32 * 2),
        ("w", _vp), ("ldw", _i64), ("N", _i32),
        ("out", _vp), ("ldo", _i64), ("out_f32", _i32),
        ("bias", _vp), ("rowbias", _vp), ("rowbias_idx", _vp), ("rowbias_ld", _i32),
        ("res", _vp), ("ldres", _i64),
        ("vt", _vp), ("vt_col0", _i32), ("vt_ld", _i64),
        ("zero_pad_last", _i32), ("epi_h", _i32), ("epi_w", _i32), ("s2_batch", _i32),
        ("ws", _vp), ("ws_bytes", _i64),
        ("custom_taps", _i32), ("tap_dh", C.c_int8 * 9), ("tap_dw", C.c_int8 * 9), ("force_splits", _i32),
        ("out_up2", _i32), ("out_ph", _i32), ("out_pw", _i32),
        ("w_mn", _i32), ("w_tap_ids", C.c_int8 * 9), ("s2_direct", _i32),
        ("w_batch_row", _i64), ("w_batch_col", _i64),
        ("out_nchw", _vp), ("out_nchw_c", _i32),
        ("gn_mode", _i32), ("gn_groups", _i32), ("gn_silu", _i32), ("gn_eps", _f32),
        ("gn_gamma", _vp), ("gn_beta", _vp), ("gn_out", _vp), ("gn_ldo", _i64), ("gn_ws", _vp), ("gn_ws_bytes", _i64),
    ]


class WgradArgs(C.Structure):
    _fields_ = [("x", NHWC), ("taps", _i32), ("s2_batch", _i32), ("dy", _vp), ("ld_dy", _i64), ("cout", _i32),
                ("grad", _vp), ("accumulate", _i32), ("ws", _vp), ("ws_bytes", _i64)]


class PackJob(C.Structure):
    _fields_ = [("src", _vp), ("src2", _vp), ("dst", _vp), ("n_outer", _i32), ("n_taps", _i32), ("n_inner", _i32),
                ("out_f32", _i32), ("so", _i64), ("si", _i64), ("st", _i64), ("dldo", _i64), ("dldt", _i64)]


# name -> argtypes (the trailing stream argument is appended automatically)
_SIGNATURES = {
    "idf_conv2d_igemm": [C.POINTER(IgemmArgs)],
    "idf_tile_walk_trace": [_i32, _i32, _i32, _i32, _i32, _vp],
    "idf_gn_plan_check": [_i32, _i32, _i32, _i32, _i32, _vp],
    "idf_groupnorm_silu": [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32],
    "idf_groupnorm_silu_rows": [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _i64, _i64],
    "idf_attention_fwd": [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _f32],
    "idf_softmax_rows": [_vp, _i64, _vp, _i64, _i32, _i32, _f32],
    "idf_embed_time_class": [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp],
    "idf_cfg_posterior_step": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                               _i32],
    "idf_cfg_ddim_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i32, _vp, _vp, _i32, _i32],
    "idf_add_noise": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32],
    "idf_vq_argmin": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32],
    "idf_kl_loss_reparam": [_vp, _vp, _vp, _vp, _vp, _i32, _i32],
    "idf_vq_loss_perplexity": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _i64],
    "idf_conv3x3_small_cin": [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32],
    "idf_conv3x3_small_cout": [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32],
    "idf_conv1x1_small_f32": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32],
    "idf_upsample_nearest2x": [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32],
    "idf_space_to_depth2": [_vp, _i64, _vp, _i32, _i32, _i32, _i32],
    "idf_nchw_f32_to_nhwc_bf16": [_vp, _vp, _i64, _i32, _i32, _i32],
    "idf_nhwc_bf16_to_nchw_f32": [_vp, _i64, _vp, _i32, _i32, _i32],
    # training step
    "idf_conv2d_wgrad": [C.POINTER(WgradArgs)],
    "idf_groupnorm_silu_train": [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp],
    "idf_groupnorm_silu_bwd": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32,
                               _i32, _i32, _i32],
    "idf_groupnorm_bwd_finalize": [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp],
    "idf_reduce_rows_f32": [_vp, _i64, _i32, _i32, _vp, _i32],
    "idf_colsum_bf16": [_vp, _i64, _i32, _i32, _i32, _vp, _i64, _vp, _i32],
    "idf_sum2x2_bf16": [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32],
    "idf_depth_to_space2": [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32],
    "idf_zero_last_rowcol": [_vp, _i64, _i32, _i32, _i32, _i32],
    "idf_conv3x3_small_cin_wgrad": [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32],
    "idf_conv3x3_small_cout_bwd": [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32],
    "idf_embed_time_class_train": [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp],
    "idf_embed_time_class_bwd": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _i64],
    "idf_mse_loss_grad": [_vp, _vp, _i64, _f32, _vp, _vp],
    "idf_grad_norm_clip": [_vp, _i64, _f32, _f32, _vp, _vp, _i64],
    "idf_adam_step": [_vp, _vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32, _vp],
    "idf_reparam_add_noise": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32],
    "idf_attention_fwd_train": [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _vp],
    "idf_attention_delta": [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp],
    "idf_attention_fwd_qkv": [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _vp],
    "idf_attention_bwd": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _f32],
    "idf_f32_to_bf16_rows": [_vp, _vp, _i64, _i64, _i32],
    "idf_u8_nhwc_to_f32_nchw": [_vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32],
    "idf_pack_weights": [_vp, _vp, _i32, _i32],
    "idf_rowidx_from_timestep": [_vp, _vp, _i32, _vp, _i32],
}
EXPORTS = sorted(list(_SIGNATURES) + ["idf_last_error", "idf_abi_version", "idf_struct_size"])
ABI_VERSION = 4  # include/idf_b200.h: IDF_B200_ABI_VERSION this binding was written against

_lib = None
launch_count = 0  # kernels-launching C-ABI calls made through this module (bench.py reports it)


def load() -> C.CDLL:
    """Loads the shared library (once). Raises NativeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (or image-diffusion_b200/csrc/build.py). "
            "There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    lib.idf_last_error.restype = C.c_char_p
    lib.idf_last_error.argtypes = []
    lib.idf_abi_version.restype = C.c_int
    lib.idf_abi_version.argtypes = []
    lib.idf_struct_size.restype = C.c_int
    lib.idf_struct_size.argtypes = [C.c_int]
    if lib.idf_abi_version() != ABI_VERSION:
        raise NativeError(f"{LIB_PATH} has ABI version {lib.idf_abi_version()}, this binding expects {ABI_VERSION}: rebuild "
                          "the library (python __graft_entry__.py build)")
    for which, struct in enumerate((NHWC, IgemmArgs, WgradArgs, PackJob)):
        if lib.idf_struct_size(which) != C.sizeof(struct):
            raise NativeError(f"{struct.__name__}: ctypes layout has {C.sizeof(struct)} bytes, the library's struct "
                              f"{lib.idf_struct_size(which)}: native.py and include/idf_b200.h disagree")
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = list(argtypes) + [_vp]
    _lib = lib
    return lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    if t is None:
        return None
    return t.data_ptr()


def call(name: str, *args) -> None:
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args, _stream())
    launch_count += 1
    if rc != 0:
        raise NativeError(f"{name} failed (code {rc}): {lib.idf_last_error().decode()}")


def nhwc_view(t: torch.Tensor | None, n: int = 0, h: int = 0, w: int = 0, c: int = 0, ld: int | None = None) -> NHWC:
    """View over a (rows, ld) bf16 matrix holding an (n, h, w) pixel grid with `c` channels starting at t's pointer."""
    v = NHWC()
    if t is None:
        v.ptr = None
        return v
    ld = int(ld if ld is not None else t.stride(-2))
    v.ptr = t.data_ptr()
    v.n, v.h, v.w, v.c = n, h, w, c
    v.sw, v.sh, v.sn = ld, ld * w, ld * w * h
    return v


def matrix_view(t: torch.Tensor, rows: int, cols: int, ld: int | None = None) -> NHWC:
    return nhwc_view(t, 1, 1, rows, cols, ld)
