"""ctypes binding of libtecogan_b200.so (C ABI: include/tecogan_b200.h).

There is no fallback: if the library is missing, cannot be loaded, or the device is not
sm_100, every entry point raises.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libtecogan_b200.so")   # the in-tree build; nothing in the environment redirects it

AMODE_HALO = 0
AMODE_DX3 = 1
OUT_F32, OUT_F16, OUT_U8 = 0, 1, 2   # tg_gen_clip_step_fmt output formats (include/tecogan_b200.h)
AMODE_FRAME = 2     # whole generator forward as one persistent kernel (tg_gen_forward / tg_gen_clip_forward)

_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_size_t = ctypes.c_size_t
_c_ll = ctypes.c_longlong
_c_float = ctypes.c_float

# name -> (restype, argtypes); mirrors include/tecogan_b200.h one to one
SIGNATURES = {
    "tg_last_error_string": (ctypes.c_char_p, []),
    "tg_version": (_c_int, []),
    "tg_check_device": (_c_int, []),
    "tg_launch_count": (_c_ll, []),
    "tg_profile_begin": (_c_int, []),
    "tg_profile_end": (_c_int, [_c_int, _c_void_p, _c_void_p, _c_void_p]),
    "tg_frame_set_trace": (_c_int, [_c_void_p, _c_size_t]),
    "tg_frame_set_pair": (_c_int, [_c_int]),
    "tg_space_to_depth": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_depth_to_space": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_warp_bilinear": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                  _c_void_p]),
    "tg_upscale4_bilinear": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p]),
    "tg_fused_warp_s2d_concat": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_ll,
                                          _c_ll, _c_void_p]),
    "tg_pack_nchw_to_nhwc64": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_packed_conv_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "tg_pack_weights": (_c_int, [_c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]),
    "tg_conv3x3_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                _c_int, _c_int, _c_void_p]),
    "tg_conv4x4s2_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_convT3x3s2_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                   _c_int, _c_void_p]),
    "tg_conv3x3_out_sigmoid": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int,
                                        _c_void_p]),
    "tg_conv3x3_wgrad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_conv3x3_wgrad_bias": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_convT3x3s2_wgrad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_conv4x4s2_wgrad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_bias_grad": (_c_int, [_c_void_p, _c_void_p, _c_ll, _c_int, _c_void_p]),
    "tg_conv3x3_dgrad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                  _c_void_p]),
    "tg_convT3x3s2_dgrad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                     _c_void_p]),
    "tg_conv4x4s2_dgrad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                    _c_void_p]),
    "tg_gen_param_count": (_c_size_t, [_c_int]),
    "tg_gen_packed_bytes": (_c_size_t, [_c_int]),
    "tg_gen_pack": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p]),
    "tg_gen_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "tg_gen_forward": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int,
                                _c_int, _c_int, _c_int, _c_void_p]),
    "tg_gen_clip_step": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int,
                                  _c_int, _c_int, _c_ll, _c_ll, _c_ll, _c_int, _c_void_p]),
    "tg_gen_clip_step_chained": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int,
                                  _c_int, _c_int, _c_ll, _c_ll, _c_ll, _c_int, _c_void_p]),
    "tg_gen_clip_step_fmt": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_size_t,
                                      _c_int, _c_int, _c_int, _c_ll, _c_ll, _c_void_p]),
    "tg_gen_clip_forward": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int, _c_int,
                                     _c_int, _c_int, _c_int, _c_void_p]),
    "tg_gen_train_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "tg_gen_packed_dgrad_bytes": (_c_size_t, [_c_int]),
    "tg_gen_pack_dgrad": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p]),
    "tg_gen_forward_train": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int, _c_int, _c_int,
                                      _c_void_p]),
    "tg_gen_clip_forward_train": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int, _c_int, _c_int,
                                           _c_int, _c_void_p]),
    "tg_gen_backward": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int, _c_int,
                                 _c_int, _c_void_p]),
    "tg_disc_input_assemble": (_c_int, [_c_void_p, _c_void_p, _c_ll, _c_ll, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                        _c_int, _c_int, _c_int, _c_void_p]),
    "tg_disc_param_count": (_c_size_t, [_c_int, _c_int, _c_int]),
    "tg_disc_packed_bytes": (_c_size_t, [_c_int, _c_int]),
    "tg_disc_pack": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]),
    "tg_disc_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int, _c_int]),
    "tg_disc_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                 _c_int, _c_void_p, _c_size_t, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_disc_forward_groups": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                        _c_int, _c_void_p, _c_size_t, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_disc_backward_groups": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                         _c_size_t, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_disc_packed_dgrad_bytes": (_c_size_t, [_c_int, _c_int]),
    "tg_disc_pack_dgrad": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]),
    "tg_disc_backward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                  _c_size_t, _c_int, _c_int, _c_int, _c_void_p]),
    "tg_workspace_bytes_bn": (_c_size_t, []),
    "tg_bn_stats": (_c_int, [_c_void_p, _c_ll, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                             _c_size_t, _c_void_p]),
    "tg_bn_apply": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_void_p, _c_int, _c_void_p]),
    "tg_bn_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                           _c_size_t, _c_void_p]),
    "tg_grad_check_finite": (_c_int, [_c_void_p, _c_ll, _c_void_p, _c_void_p]),
    "tg_adam_step": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_void_p, _c_float, _c_float, _c_float, _c_void_p,
                              _c_void_p, _c_void_p, _c_void_p]),
    "tg_scaler_update": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_float, _c_float, _c_int, _c_void_p, _c_void_p]),
    "tg_workspace_bytes_gen_forward": (_c_size_t, [_c_int, _c_int, _c_int]),
    "tg_workspace_bytes_gen_train": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "tg_workspace_bytes_disc": (_c_size_t, [_c_int, _c_int, _c_int, _c_int, _c_int]),
}

_lib = None
_device_checked = False


def load():
    """Load the shared library (no CUDA call is made; safe on a GPU-less box)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libtecogan_b200.so not found at {LIB_PATH}; build it with "
                "`python pytorch-tecogan_b200/build.py` — there is no CPU or PyTorch fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def lib():
    """Library handle for compute calls: also verifies once that a B200-class GPU is present."""
    global _device_checked
    l = load()
    if not _device_checked:
        if not torch.cuda.is_available():
            raise RuntimeError("tecogan_b200 needs a CUDA device (sm_100a); none is visible and there is no fallback")
        check(l.tg_check_device())
        _device_checked = True
    return l


def check(rc):
    if rc != 0:
        msg = load().tg_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"libtecogan_b200 error {rc}: {msg}")


def stream_ptr(device=None):
    """current torch stream of `device` (default: the current device) as the void* the C ABI takes."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """device pointer of a tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def require_cuda_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
