"""Drop-in for the reference's `geometric_model.SimpleCrescentNuisanceFloor_Param2Img`
(DPItorch/geometric_model.py:237-389): crescent (asymmetric ring + erf disk floor) + n elliptical nuisance
Gaussians rendered from parameters in (0, 1), unit total flux. Forward and backward run as hand-written CUDA
(dpi_b200/csrc/crescent.cu, one CTA per image) through a torch.autograd.Function; there is no CPU fallback.

The constructor's default parameter ranges are compiled into the kernel (r_range [10, 40] as DPIx_interferometry.py
passes); flux_flag=False (data_product 'cphase_logcamp', :214-217) renders unit-flux images, flux_flag=True (the other data
products) adds the total-flux parameter: the same kernels on the re-ordered parameters, times flux (:373-376).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from dpi_b200 import _lib


class _CrescentFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, npix, n_gaussian, fov):
        if not params.is_cuda:
            raise RuntimeError("dpi_b200.geometric_model runs on CUDA only (no CPU fallback); got a CPU tensor")
        lib = _lib.load()
        p = params.contiguous().float()
        B = p.shape[0]
        img = torch.empty(B, npix, npix, dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            _lib.check(lib.dpi_crescent_forward(B, npix, n_gaussian, float(fov), p.data_ptr(), img.data_ptr(),
                                                _lib.stream_ptr()), "dpi_crescent_forward")
        ctx.save_for_backward(p)
        ctx.cfg = (npix, n_gaussian, float(fov))
        return img

    @staticmethod
    def backward(ctx, g_img):
        (p,) = ctx.saved_tensors
        npix, n_gaussian, fov = ctx.cfg
        lib = _lib.load()
        g = g_img.contiguous().float()
        gp = torch.empty_like(p)
        with torch.cuda.device(p.device):
            _lib.check(lib.dpi_crescent_backward(p.shape[0], npix, n_gaussian, fov, p.data_ptr(), g.data_ptr(), None,
                                                 gp.data_ptr(), _lib.stream_ptr()), "dpi_crescent_backward")
        return gp, None, None, None


class _CrescentFluxFunction(torch.autograd.Function):
    """flux_flag=True (geometric_model.py:253-256, :305-312, :373-376): parameters [4 + 6 g shape | total flux | floor |
    crescent flux]; img = flux * R(shape, floor, crescent flux) with R the unit-flux renderer and flux = lo + (hi - lo) p."""

    @staticmethod
    def forward(ctx, params, npix, n_gaussian, fov, lo, hi):
        if not params.is_cuda:
            raise RuntimeError("dpi_b200.geometric_model runs on CUDA only (no CPU fallback); got a CPU tensor")
        lib = _lib.load()
        p = params.contiguous().float()
        B, k = p.shape[0], 4 + 6 * n_gaussian
        pr = torch.cat([p[:, :k], p[:, k + 1:k + 3]], 1).contiguous()        # the unit-flux renderer's parameter order
        flux = (lo + (hi - lo) * p[:, k]).contiguous()
        unit = torch.empty(B, npix, npix, dtype=torch.float32, device=p.device)
        img = torch.empty_like(unit)
        with torch.cuda.device(p.device):
            st = _lib.stream_ptr()
            _lib.check(lib.dpi_crescent_forward(B, npix, n_gaussian, float(fov), pr.data_ptr(), unit.data_ptr(), st),
                       "dpi_crescent_forward")
            _lib.check(lib.dpi_row_scale_dot(B, npix * npix, unit.data_ptr(), flux.data_ptr(), img.data_ptr(), None, 1.0, None,
                                             st), "dpi_row_scale_dot")
        ctx.save_for_backward(pr, flux, unit)
        ctx.cfg = (npix, n_gaussian, float(fov), float(hi - lo))
        return img

    @staticmethod
    def backward(ctx, g_img):
        pr, flux, unit = ctx.saved_tensors
        npix, n_gaussian, fov, span = ctx.cfg
        lib = _lib.load()
        g = g_img.contiguous().float()
        B, k = pr.shape[0], 4 + 6 * n_gaussian
        gpr = torch.empty_like(pr)
        gflux = torch.empty(B, dtype=torch.float32, device=pr.device)
        with torch.cuda.device(pr.device):
            st = _lib.stream_ptr()
            # d img / d pr = flux * d R / d pr (the kernel's per-sample scale); d img / d p_flux = span * R
            _lib.check(lib.dpi_crescent_backward(B, npix, n_gaussian, fov, pr.data_ptr(), g.data_ptr(), flux.data_ptr(),
                                                 gpr.data_ptr(), st), "dpi_crescent_backward")
            _lib.check(lib.dpi_row_scale_dot(B, npix * npix, unit.data_ptr(), None, None, g.data_ptr(), span, gflux.data_ptr(),
                                             st), "dpi_row_scale_dot")
        gp = torch.cat([gpr[:, :k], gflux[:, None], gpr[:, k:]], 1)
        return gp, None, None, None, None, None


class SimpleCrescentNuisanceFloor_Param2Img(nn.Module):
    """Same constructor signature and `forward(params) -> img [B, npix, npix]` as the reference class."""

    def __init__(self, npix, n_gaussian=1, fov=160, r_range=[10.0, 40.0], asym_range=[1e-3, 0.99],
                 width_range=[1.0, 40.0], floor_range=[0.0, 1.0], flux_range=[0.8, 1.2], crescent_flux_range=[1e-3, 2.0],
                 shift_range=[-200.0, 200.0], sigma_range=[1.0, 100.0], gaussian_scale_range=[1e-3, 2.0], flux_flag=False):
        super().__init__()
        defaults = dict(r_range=[10.0, 40.0], asym_range=[1e-3, 0.99], width_range=[1.0, 40.0], floor_range=[0.0, 1.0],
                        crescent_flux_range=[1e-3, 2.0], shift_range=[-200.0, 200.0], sigma_range=[1.0, 100.0],
                        gaussian_scale_range=[1e-3, 2.0])
        given = dict(r_range=r_range, asym_range=asym_range, width_range=width_range, floor_range=floor_range,
                     crescent_flux_range=crescent_flux_range, shift_range=shift_range, sigma_range=sigma_range,
                     gaussian_scale_range=gaussian_scale_range)
        for k, v in defaults.items():
            if [float(x) for x in given[k]] != v:
                raise NotImplementedError(f"dpi_b200: only the default {k}={v} is compiled into the CUDA renderer")
        if not 0 <= n_gaussian <= 4:
            raise NotImplementedError("dpi_b200: 0..4 nuisance Gaussians")
        self.npix, self.n_gaussian, self.fov, self.flux_flag = int(npix), int(n_gaussian), float(fov), bool(flux_flag)
        self.flux_range = [float(flux_range[0]), float(flux_range[1])]
        self.nparams = (5 if flux_flag else 4) + 6 * n_gaussian + 2          # geometric_model.py:253-258
        self.eps = 1e-4

    def compute_features(self, params):
        """Physical features of each sample (geometric_model.py:275-321), same return tuple as the reference with
        flux_flag=False. Offline analysis helper (posterior summaries, DPIx_interferometry.py:492-520): plain
        elementwise torch on whatever device `params` lives on; the renderer kernel recomputes them internally."""
        hf = 0.5 * self.fov
        col = lambda i: params[:, i].unsqueeze(-1).unsqueeze(-1)
        r = 10.0 / hf + col(0) * 30.0 / hf
        sigma = 1.0 / hf + col(1) * 39.0 / hf
        s = 1e-3 + col(2) * (0.99 - 1e-3)
        eta = 181 / 180 * np.pi * (2.0 * col(3) - 1.0)
        scale, xs, ys, sx, sy, th = [], [], [], [], [], []
        for k in range(self.n_gaussian):
            xs.append(-200.0 / hf + col(4 + k * 6) * 400.0 / hf)
            ys.append(-200.0 / hf + col(5 + k * 6) * 400.0 / hf)
            scale.append(1e-3 + col(6 + k * 6) * (2.0 - 1e-3))
            sx.append(1.0 / hf + col(7 + k * 6) * 99.0 / hf)
            sy.append(1.0 / hf + col(8 + k * 6) * 99.0 / hf)
            th.append(181 / 180 * 0.5 * np.pi * col(9 + k * 6))
        if self.flux_flag:                                   # :305-309
            flux = self.flux_range[0] + (self.flux_range[1] - self.flux_range[0]) * col(4 + self.n_gaussian * 6)
            floor = col(5 + self.n_gaussian * 6)
            crescent_flux = 1e-3 + (2.0 - 1e-3) * col(6 + self.n_gaussian * 6)
            return r, sigma, s, eta, flux, crescent_flux, floor, scale, xs, ys, sx, sy, th
        floor = col(4 + self.n_gaussian * 6)
        crescent_flux = 1e-3 + (2.0 - 1e-3) * col(5 + self.n_gaussian * 6)
        return r, sigma, s, eta, crescent_flux, floor, scale, xs, ys, sx, sy, th

    def forward(self, params):
        if self.flux_flag:
            return _CrescentFluxFunction.apply(params, self.npix, self.n_gaussian, self.fov, self.flux_range[0],
                                               self.flux_range[1])
        return _CrescentFunction.apply(params, self.npix, self.n_gaussian, self.fov)

    def to(self, device=None, **kw):      # the reference overrides .to() to move its grids by hand (:384-389)
        return self
