"""Building blocks of the Glow flow (reference: generative_model/glow_model.py) on the tcgen05 GEMM machinery.

Only the convolution forward exists so far (SURVEY.md section 8 row a20 is NOT built: see DESIGN.md section 7 for the
plan and oracle/glow_oracle.py for the pinned algorithm). Tensors are FEATURE-major image stacks [channel][b*h*w],
the layout the rest of the flow kernels use, so BatchNorm2d becomes the row reduction the GEMM epilogues already do.
"""
from __future__ import annotations

from typing import Optional

import torch

from dpi_b200 import _lib


def to_feature_major(x: torch.Tensor) -> torch.Tensor:
    """[B, C, h, w] -> [C, B*h*w] (contiguous)."""
    B, C, h, w = x.shape
    return x.permute(1, 0, 2, 3).reshape(C, B * h * w).contiguous()


def from_feature_major(x: torch.Tensor, batch: int, h: int, w: int) -> torch.Tensor:
    """[C, B*h*w] -> [B, C, h, w]."""
    return x.view(x.shape[0], batch, h, w).permute(1, 0, 2, 3).contiguous()


def conv2d_fm(x_fm: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], batch: int, h: int, w: int,
              pad_value: float = 0.0, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Conv2d(c_in, c_out, k, padding=k//2) for k in {1, 3} (glow_model.py:261,264) or, with pad_value=1 and `scale`,
    ZeroConv2d (:246-251: pad with ones, 3x3 conv, times exp(3 scale)), on a feature-major stack. Forward only."""
    if not x_fm.is_cuda:
        raise RuntimeError("dpi_b200 has no CPU path: conv2d_fm needs CUDA tensors")
    lib = _lib.load()
    c_out, c_in, k, k2 = weight.shape
    assert k == k2 and x_fm.shape == (c_in, batch * h * w) and x_fm.dtype == torch.float32
    x_fm, weight = x_fm.contiguous(), weight.contiguous()
    y = torch.empty(c_out, batch * h * w, dtype=torch.float32, device=x_fm.device)
    nbytes = lib.dpi_conv2d_fm_workspace_bytes(batch, h, w, c_in, c_out, k)
    if nbytes == 0:
        raise RuntimeError("dpi_conv2d_fm_workspace_bytes: unsupported shape")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x_fm.device)
    with torch.cuda.device(x_fm.device):
        _lib.check(lib.dpi_conv2d_fm_forward(batch, h, w, c_in, c_out, k, float(pad_value), x_fm.data_ptr(), x_fm.shape[1],
                                             weight.data_ptr(), _lib.ptr(bias), _lib.ptr(scale.reshape(-1).contiguous()
                                                                                          if scale is not None else None),
                                             y.data_ptr(), y.shape[1], ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                   "dpi_conv2d_fm_forward")
    return y


def conv2d_fm_dgrad(g_fm: torch.Tensor, weight: torch.Tensor, batch: int, h: int, w: int) -> torch.Tensor:
    """Gradient of conv2d_fm with respect to its input stack: the same convolution applied to the output gradient with
    the spatially flipped, channel-transposed weight (padding contributes nothing: it is a constant)."""
    wt = weight.flip(2, 3).permute(1, 0, 2, 3).contiguous()
    return conv2d_fm(g_fm, wt, None, batch, h, w, pad_value=0.0)


def conv2d_fm_wgrad(x_fm: torch.Tensor, g_fm: torch.Tensor, ksize: int, batch: int, h: int, w: int,
                    pad_value: float = 0.0) -> torch.Tensor:
    """Gradient with respect to the weight [c_out, c_in, k, k] (for ZeroConv2d pass pad_value=1 and the gradient of the
    un-scaled convolution output)."""
    if not x_fm.is_cuda:
        raise RuntimeError("dpi_b200 has no CPU path: conv2d_fm_wgrad needs CUDA tensors")
    lib = _lib.load()
    c_in, c_out = x_fm.shape[0], g_fm.shape[0]
    x_fm, g_fm = x_fm.contiguous(), g_fm.contiguous()
    gw = torch.empty(c_out, c_in, ksize, ksize, dtype=torch.float32, device=x_fm.device)
    ws = torch.empty(lib.dpi_conv2d_fm_wgrad_workspace_bytes(batch, h, w, c_in, c_out, ksize), dtype=torch.uint8,
                     device=x_fm.device)
    with torch.cuda.device(x_fm.device):
        _lib.check(lib.dpi_conv2d_fm_wgrad(batch, h, w, c_in, c_out, ksize, float(pad_value), x_fm.data_ptr(), x_fm.shape[1],
                                           g_fm.data_ptr(), g_fm.shape[1], gw.data_ptr(), ws.data_ptr(), ws.numel(),
                                           _lib.stream_ptr()), "dpi_conv2d_fm_wgrad")
    return gw
