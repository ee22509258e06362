"""ctypes binding of libosmosis_b200.so (the C ABI declared in include/osmosis_b200.h).

PyTorch is used here only for device memory and streams: tensors are passed as raw `data_ptr()`s and the
current CUDA stream handle.  There is no CPU fallback - a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libosmosis_b200.so")


class OsmError(RuntimeError):
    pass


class UNetConfigC(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("out_channels", C.c_int), ("model_channels", C.c_int),
                ("num_res_blocks", C.c_int), ("num_levels", C.c_int), ("channel_mult", C.c_int * 8),
                ("num_attention_ds", C.c_int), ("attention_ds", C.c_int * 8), ("num_heads", C.c_int),
                ("num_head_channels", C.c_int), ("conv_mode", C.c_int)]


class GuidanceParamsC(C.Structure):
    _fields_ = [("op_kind", C.c_int), ("depth_kind", C.c_int), ("depth_val", C.c_float * 3), ("weight_kind", C.c_int),
                ("weight_depth_kind", C.c_int), ("weight_val", C.c_float * 3), ("eta", C.c_float * 3), ("n_iter", C.c_int),
                ("gamma_avrg", C.c_float), ("gamma_val", C.c_float), ("loss_kind", C.c_int), ("optimizer", C.c_int),
                ("opt_state", C.c_void_p), ("phi_batch", C.c_int)]


_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes); every symbol include/osmosis_b200.h declares
SIGNATURES = {
    "osm_last_error_string": (C.c_char_p, []),
    "osm_abi_version": (_I, []),
    "osm_unet_create": (_I, [C.POINTER(UNetConfigC), C.POINTER(_P)]),
    "osm_unet_destroy": (_I, [_P]),
    "osm_unet_param_count": (_I, [_P]),
    "osm_unet_param_info": (_I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(_I), C.POINTER(_L)]),
    "osm_unet_load_param": (_I, [_P, C.c_char_p, _P, _L, _P]),
    "osm_unet_workspace_bytes": (_L, [_P, _I, _I, _I]),
    "osm_unet_bind": (_I, [_P, _I, _I, _I, _P, _L]),
    "osm_unet_forward": (_I, [_P, _P, _P, _P, _P]),
    "osm_unet_vjp_input": (_I, [_P, _P, _P, _P]),
    "osm_unet_launch_count": (_I, [_P, _I]),
    "osm_unet_forward_flops": (C.c_double, [_P]),
    "osm_unet_profile_ops": (_I, [_P, _I, _P, _I, C.POINTER(_F), C.POINTER(_I), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double), C.POINTER(_I)]),
    "osm_posterior_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "osm_posterior_vjp": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "osm_sampler_update": (_I, [_P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "osm_posterior_fwd_ex": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "osm_posterior_vjp_ex": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _P]),
    "osm_sampler_update_ex": (_I, [_P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "osm_ddim_sample": (_I, [_P, _P, _P, _P, _P, _F, _P, _I, _I, _I, _P]),
    "osm_ps_guidance": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "osm_ddpm_uncond_update": (_I, [_P, _P, _P, _F, _F, _F, _I, _I, _I, _I, _P]),
    "osm_operator_forward": (_I, [_I, _I, C.POINTER(_F), _P, _P, _P, _I, _I, _P]),
    "osm_guidance_phi_loop": (_I, [C.POINTER(GuidanceParamsC), _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "osm_postprocess": (_I, [_I, _I, C.POINTER(_F), _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "osm_minmax_percentile": (_I, [_P, _P, _I, _I, _F, _F, _F, _F, _P]),
    "osm_colormap": (_I, [_P, _P, _P, _I, _I, _P]),
    "osm_preprocess_image": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _P]),
    "osm_degamma": (_I, [_P, _P, _L, _P]),
    "osm_dbg_conv": (_I, [_I, _P, _I, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "osm_dbg_conv_halo": (_I, [_P, _I, _P, _P, _P, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "osm_dbg_conv_halo16": (_I, [_P, _I, _P, _P, _P, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "osm_dbg_conv_f16": (_I, [_P, _I, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "osm_dbg_pack_conv_weight_f16": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "osm_dbg_pack_conv_weight": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "osm_dbg_gn_forward": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P]),
    "osm_dbg_gn_backward": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "osm_dbg_gn_forward_f16": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P]),
    "osm_dbg_gn_backward_f16": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "osm_dbg_attention": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "osm_dbg_attention_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "osm_dbg_conv_stats": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "osm_dbg_conv_stats_f16": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "osm_dbg_attention_flash": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "osm_dbg_attention_flash_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
}

_lib = None


def load():
    """Loads the shared library (building it is `python -m osmosis_diffusion_code_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OsmError(f"{LIB_PATH} is missing - build it with `python -m osmosis_diffusion_code_b200.build`; "
                           "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(code: int):
    if code != 0:
        msg = load().osm_last_error_string().decode()
        raise OsmError(f"libosmosis_b200 error {code}: {msg}")


def ptr(t):
    """Device pointer of a contiguous fp32/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise OsmError("a CUDA device (B200, sm_100a) is required: this package has no CPU fallback")
