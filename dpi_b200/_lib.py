"""ctypes binding of libdpi_b200.so (the C ABI declared in include/dpi_b200.h).

There is deliberately no fallback: if the shared library is missing, `load()` raises with the
build command, and every op that needs it fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DPI_B200_LIB") or os.path.join(_HERE, "libdpi_b200.so")   # override: A/B builds

_lib = None
_lock = threading.Lock()

vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/dpi_b200.h declares
PROTOTYPES = {
    "dpi_last_error": (C.c_char_p, []),
    "dpi_version": (i32, []),
    "dpi_launch_count": (C.c_longlong, []),
    "dpi_flow_layout": (i32, [i32, i32, i32, i32, vp, vp, vp, vp]),
    "dpi_flow_create": (i32, [C.POINTER(vp), i32, i32, i32, i32, i32, vp]),
    "dpi_flow_destroy": (i32, [vp]),
    "dpi_flow_param_count": (i64, [vp]),
    "dpi_flow_bn_count": (i64, [vp]),
    "dpi_flow_param_offset": (i64, [vp, i32, i32]),
    "dpi_flow_bn_offset": (i64, [vp, i32, i32, i32]),
    "dpi_flow_gather_map": (i32, [vp, i32, i32, vp]),
    "dpi_flow_workspace_bytes": (sz, [vp, i32]),
    "dpi_flow_ws_offset": (i64, [vp, i32, i32, i32]),
    "dpi_flow_tc_plan": (i32, [i32, i32, i32, i32, vp]),
    "dpi_conv2d_fm_workspace_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
    "dpi_conv2d_fm_forward": (i32, [i32, i32, i32, i32, i32, i32, f32, vp, i64, vp, vp, vp, vp, i64, vp, sz, vp]),
    "dpi_conv2d_fm_lrelu": (i32, [i32, i32, i32, i32, i32, i32, f32, vp, i64, vp, vp, f32, vp, i64, vp, sz, vp]),
    "dpi_conv2d_fm_affine": (i32, [i32, i32, i32, i32, i32, vp, i64, vp, vp, vp, vp, i64, vp, sz, vp]),
    "dpi_conv2d_fm_wgrad_workspace_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
    "dpi_conv2d_fm_wgrad": (i32, [i32, i32, i32, i32, i32, i32, f32, vp, i64, vp, i64, vp, vp, sz, vp]),
    "dpi_glow_bn_rows": (i32, [i32, i64, i64, vp, vp, vp, vp, vp, i32, f32, vp, vp]),
    "dpi_glow_lu_inverse": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "dpi_glow_affine_reverse": (i32, [i32, i32, i32, vp, i64, vp, i64, vp, vp]),
    "dpi_glow_prior_sample": (i32, [i32, i32, i32, vp, i64, vp, i64, vp, i64, vp, vp]),
    "dpi_glow_unsqueeze": (i32, [i32, i32, i32, i32, vp, vp, vp]),
    "dpi_glow_lu_weight": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "dpi_glow_affine_forward": (i32, [i32, i32, i32, vp, i64, vp, i64, vp, vp]),
    "dpi_glow_prior_forward": (i32, [i32, i32, i32, vp, i64, vp, i64, vp, i64, vp, vp, vp]),
    "dpi_glow_squeeze": (i32, [i32, i32, i32, i32, vp, vp, vp]),
    "dpi_glow_row_affine": (i32, [i32, i64, vp, vp, vp, vp, vp]),
    "dpi_glow_actnorm_reverse_bwd": (i32, [i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "dpi_glow_affine_reverse_bwd": (i32, [i32, i32, i32, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp]),
    "dpi_glow_zeroconv_bwd": (i32, [i32, i64, vp, vp, vp, vp, vp, vp]),
    "dpi_glow_bn_lrelu_bwd": (i32, [i32, i64, vp, vp, vp, f32, f32, vp, vp, vp, vp]),
    "dpi_glow_row_sum": (i32, [i32, i64, vp, f32, vp, vp]),
    "dpi_glow_lu_bwd": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "dpi_glow_prior_sample_bwd": (i32, [i32, i32, i32, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, i64, vp]),
    "dpi_glow_scaled_sum": (i32, [i32, vp, f32, vp, vp]),
    "dpi_glow_logdet_scalar": (i32, [i32, i32, i32, vp, f32, vp, f32, vp, vp]),
    "dpi_flow_run": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
    "dpi_flow_backward": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "dpi_flow_backward_range": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, sz, i32, i32, vp]),
    "dpi_flow_tc_active": (i32, [vp, i32]),
    "dpi_flow_cl_active": (i32, [vp, i32]),
    "dpi_flow_nar_active": (i32, [vp, i32]),
    "dpi_flow_range_ok": (i32, [vp, i32]),
    "dpi_flow_sr_active": (i32, [vp, i32]),
    "dpi_flow_debug_timing": (i32, [vp, i32]),
    "dpi_flow_debug_read": (i32, [vp, vp, vp, i32]),
    "dpi_flow_set_mode": (i32, [vp, i32, i32]),
    "dpi_positivity_forward": (i32, [i32, i32, vp, vp, vp, vp, vp]),
    "dpi_positivity_backward": (i32, [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "dpi_image_priors": (i32, [i32, i32, vp, vp, f32, f32, vp, vp, vp, vp]),
    "dpi_image_head": (i32, [i32, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp]),
    "dpi_step_terms": (i32, [i32, vp, f32, vp, f32, vp, vp, vp, vp, f32, vp, vp]),
    "dpi_reduce_sum": (i32, [i32, vp, vp, vp, vp]),
    "dpi_eht_create": (i32, [C.POINTER(vp), i32, i32, i32, vp, i32, vp, vp, i32, vp]),
    "dpi_eht_destroy": (i32, [vp]),
    "dpi_eht_workspace_bytes": (sz, [vp, i32]),
    "dpi_eht_forward": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
    "dpi_eht_backward": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "dpi_eht_data_grad": (i32, [vp, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp, sz, vp]),
    "dpi_chi2": (i32, [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "dpi_crescent_forward": (i32, [i32, i32, i32, f32, vp, vp, vp]),
    "dpi_crescent_backward": (i32, [i32, i32, i32, f32, vp, vp, vp, vp, vp]),
    "dpi_sigmoid_forward": (i32, [i32, i32, vp, vp, vp, vp]),
    "dpi_sigmoid_backward": (i32, [i32, i32, vp, vp, vp, vp, vp]),
    "dpi_alpha_objective": (i32, [i32, i32, vp, vp, vp, vp, vp, f32, f32, f32, f32, f32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "dpi_fft2c_forward": (i32, [i32, i32, vp, vp, vp]),
    "dpi_row_scale_dot": (i32, [i32, i32, vp, vp, vp, vp, f32, vp, vp]),
    "dpi_complex_mul": (i32, [i64, i32, i32, vp, vp, vp, vp]),
    "dpi_kspace_loss": (i32, [i32, i32, i32, f32, vp, vp, vp, vp, vp, vp]),
    "dpi_fft2c_backward": (i32, [i32, i32, vp, vp, vp]),
    "dpi_mri_data_loss": (i32, [i32, i32, vp, vp, vp, f32, f32, vp, vp, vp, vp]),
    "dpi_clip_adam_workspace_bytes": (sz, [i64]),
    "dpi_grad_sqnorm": (i32, [i64, vp, vp, vp, sz, vp]),
    "dpi_clip_adam": (i32, [i64, vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, f32, vp, f32, vp, vp]),
    "dpi_step_increment": (i32, [vp, vp]),
}


def load():
    """Load the CUDA library; raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"dpi_b200: CUDA library {LIB_PATH} is missing - build it with "
                f"`make -C {os.path.join(_HERE, 'csrc')}` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dpi_last_error()
        raise RuntimeError(f"dpi_b200 {what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().dpi_launch_count())
