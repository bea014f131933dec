"""Drop-in for DPItorch/MRI_helpers.py (+ the two script-level helpers of DPI_MRI.py) backed by
libdpi_b200 (sm_100a).

    Loss_kspace_diff(sigma) / Loss_kspace_diff2(sigma) -> func(y_true, y_pred) -> [B]   MRI_helpers.py:18-26
    Loss_l1 / Loss_TSV / Loss_TV(img) -> [B]                                           MRI_helpers.py:28-38
    fft2c_torch(img[B,n,n]) -> kspace[B,n,n,2]  (ortho FFT2 of a real image)            DPI_MRI.py:70-74
    Img_logscale(scale)                                                                DPI_MRI.py:77-85
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from dpi_b200 import _lib
from dpi_b200.interferometry_helpers import Loss_l1, Loss_TSV, Loss_TV, _need_cuda  # same definitions (:28-38)

__all__ = ["Loss_kspace_diff", "Loss_kspace_diff2", "Loss_l1", "Loss_TSV", "Loss_TV", "fft2c_torch",
           "Img_logscale"]


class _Fft2cFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img):
        lib = _lib.load()
        _need_cuda(img, "fft2c_torch")
        x = img.contiguous().float()
        B, n = x.shape[0], x.shape[-1]
        k = torch.empty(B, n, n, 2, dtype=torch.float32, device=x.device)
        _lib.check(lib.dpi_fft2c_forward(B, n, x.data_ptr(), k.data_ptr(), _lib.stream_ptr()), "dpi_fft2c_forward")
        return k

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        g = g.contiguous().float()
        B, n = g.shape[0], g.shape[1]
        gi = torch.empty(B, n, n, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(lib.dpi_fft2c_backward(B, n, g.data_ptr(), gi.data_ptr(), _lib.stream_ptr()),
                       "dpi_fft2c_backward")
        return gi


def fft2c_torch(img):
    """DPI_MRI.py:70-74 (the removed torch.fft(x, 2, normalized=True) == numpy fft2(norm='ortho'))."""
    return _Fft2cFunction.apply(img)


class _KspaceLoss(torch.autograd.Function):
    """mean over (1,2,3) of |d|^p / sigma^p, p = 1 or 2 (MRI_helpers.py:18-26): one kernel forward, one backward."""

    @staticmethod
    def forward(ctx, power, sigma, y_true, y_pred):
        lib = _lib.load()
        _need_cuda(y_pred, "Loss_kspace_diff*")
        yp = y_pred.contiguous().float()
        B = yp.shape[0]
        n = yp[0].numel()
        yt = y_true.to(yp.device).expand_as(yp[0]).contiguous().float().reshape(-1)
        ctx.power, ctx.sigma = int(power), float(sigma)
        ctx.save_for_backward(yt, yp)
        loss = torch.empty(B, dtype=torch.float32, device=yp.device)
        with torch.cuda.device(yp.device):
            _lib.check(lib.dpi_kspace_loss(int(power), B, n, float(sigma), yt.data_ptr(), yp.data_ptr(), loss.data_ptr(),
                                           None, None, _lib.stream_ptr()), "dpi_kspace_loss")
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        yt, yp = ctx.saved_tensors
        B, n = yp.shape[0], yt.numel()
        gl = g.contiguous().float()
        gp = torch.empty_like(yp)
        with torch.cuda.device(yp.device):
            _lib.check(lib.dpi_kspace_loss(ctx.power, B, n, ctx.sigma, yt.data_ptr(), yp.data_ptr(), None, gl.data_ptr(),
                                           gp.data_ptr(), _lib.stream_ptr()), "dpi_kspace_loss")
        return None, None, None, gp


def Loss_kspace_diff(sigma):
    """MRI_helpers.py:18-21 (L1 variant; unused by DPI_MRI.py:134)."""
    def func(y_true, y_pred):
        return _KspaceLoss.apply(1, sigma, y_true, y_pred)
    return func


def Loss_kspace_diff2(sigma):
    """MRI_helpers.py:23-26."""
    def func(y_true, y_pred):
        return _KspaceLoss.apply(2, sigma, y_true, y_pred)
    return func


class Img_logscale(nn.Module):
    """DPI_MRI.py:77-85 / DPI_interferometry.py:74-82."""

    def __init__(self, scale=1):
        super().__init__()
        log_scale = torch.Tensor(np.log(scale) * np.ones(1))
        self.log_scale = nn.Parameter(log_scale)

    def forward(self):
        return self.log_scale
