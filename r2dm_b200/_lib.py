"""ctypes binding of libr2dm_b200.so (the C ABI declared in include/r2dm_b200.h).

There is no CPU / PyTorch fallback: if the shared library is missing or fails to load, importing
the compute path raises.  Build it with `./build.sh` (or `python -c "import __graft_entry__ as g;
g.build()"`).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# R2DM_LIB_PATH: developer knob for A/B builds of the same C ABI (tools/); the product uses the in-tree .so
LIB_PATH = os.environ.get("R2DM_LIB_PATH") or os.path.join(_HERE, "libr2dm_b200.so")

F32, BF16 = 0, 1


class R2dmConfig(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int),
        ("height", C.c_int),
        ("width", C.c_int),
        ("base_channels", C.c_int),
        ("temb_channels", C.c_int),
        ("channel_multiplier", C.c_int * 4),
        ("num_residual_blocks", C.c_int * 4),
        ("gn_num_groups", C.c_int),
        ("gn_eps", C.c_float),
        ("attn_num_heads", C.c_int),
        ("extra_channels", C.c_int),
        ("residual_scale", C.c_float),
        ("dtype", C.c_int),
    ]


class R2dmPhilox(C.Structure):
    """r2dm_philox (include/r2dm_b200.h): device-side noise stream description."""
    _fields_ = [
        ("seeds", C.c_void_p), ("offsets", C.c_void_p), ("ctr0", C.c_void_p), ("ctr1", C.c_void_p),
        ("mul0", C.c_int), ("mul1", C.c_int), ("offset_per_draw", C.c_uint32), ("threads", C.c_uint32),
    ]


class R2dmPointNetWeights(C.Structure):
    """r2dm_pointnet_weights (include/r2dm_b200.h): BatchNorm-folded fp32 device pointers."""
    _fields_ = [("weight", C.c_void_p * 12), ("bias", C.c_void_p * 12), ("num_classes", C.c_int)]


class R2dmError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_SIGS = {
    "r2dm_last_error": (C.c_char_p, []),
    "r2dm_version": (C.c_int, []),
    "r2dm_create": (C.c_int, [C.POINTER(R2dmConfig), C.POINTER(_P)]),
    "r2dm_destroy": (C.c_int, [_P]),
    "r2dm_weight_arena_bytes": (C.c_size_t, [_P]),
    "r2dm_bind_weight_arena": (C.c_int, [_P, _P, C.c_size_t]),
    "r2dm_load_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, _P]),
    "r2dm_missing_tensors": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    "r2dm_workspace_bytes": (C.c_size_t, [_P, C.c_int]),
    "r2dm_bind_workspace": (C.c_int, [_P, _P, C.c_size_t, C.c_int, _P]),
    "r2dm_film_width": (C.c_int, [_P]),
    "r2dm_cond_embed": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "r2dm_unet_forward": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "r2dm_num_launches": (C.c_int, [_P]),
    "r2dm_profile_forward": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                       C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "r2dm_debug_forward_kinds": (C.c_int, [_P, _P, _P, _P, C.c_uint, _P]),
    "r2dm_sampler_update": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_float,
                                      _P, _P, _P, C.c_int, C.c_size_t, _P]),
    "r2dm_axpby": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_size_t, _P]),
    "r2dm_axpby_table": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_size_t, _P]),
    "r2dm_advance_step": (C.c_int, [_P, C.c_int, _P]),
    "r2dm_sampler_update_philox": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_float, _P, _P,
                                             C.POINTER(R2dmPhilox), C.c_int, C.c_int, C.c_int, C.c_size_t, _P]),
    "r2dm_axpby_table_philox": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(R2dmPhilox), C.c_int,
                                          C.c_int, C.c_size_t, _P]),
    "r2dm_philox_normal": (C.c_int, [_P, C.POINTER(R2dmPhilox), C.c_int, C.c_int, C.c_size_t, _P]),
    "r2dm_lidar_postprocess": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                         C.c_float, _P]),
    "r2dm_render_point_clouds": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "r2dm_bilinear_rasterize": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "r2dm_surface_normal": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "r2dm_bev_histogram": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P]),
    "r2dm_pointnet_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "r2dm_pointnet_features": (C.c_int, [C.c_int, _P, C.POINTER(R2dmPointNetWeights), _P, C.c_int, C.c_int, _P,
                                         C.c_size_t, _P]),
    "r2dm_op_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "r2dm_op_conv": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, C.c_float, _P, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "r2dm_op_gn_conv": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, C.c_float, C.c_int, _P, _P, _P, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "r2dm_op_gn_conv_skip": (C.c_int, [C.c_int, _P, _P, C.c_float, _P, _P, _P, _P, _P, C.c_float, _P, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "r2dm_op_groupnorm": (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_float, C.c_int, _P, C.c_int,
                                    C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "r2dm_op_resample": (C.c_int, [C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P,
                                   C.c_size_t, _P]),
    "r2dm_op_attention": (C.c_int, [C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P,
                                    C.c_size_t, _P]),
    "r2dm_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "r2dm_debug_set_trace": (C.c_int, [_P, C.c_int]),
    "r2dm_debug_set_ktime": (C.c_int, [_P, _P]),
    "r2dm_debug_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), _P]),
}
EXPORTS = tuple(_SIGS)


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises R2dmError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise R2dmError(f"{LIB_PATH} not found: the CUDA extension is not built (run ./build.sh); "
                            "r2dm_b200 has no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = lib().r2dm_last_error().decode()
        raise R2dmError(f"{what or 'r2dm call'} failed ({rc}): {msg}")
    return rc


def ptr(t) -> int:
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return t.data_ptr()


def f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(dtype=torch.float32).contiguous()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
