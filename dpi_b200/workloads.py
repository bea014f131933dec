"""Synthetic instances of the BASELINE.json configurations (host setup, numpy/torch-CPU only).

    interferometry_workload(npix=32)  - C1: DPI_interferometry npix=32, n_flow=16, B=64, EHT-like
                                        uv coverage with 938 visibilities, closure phase + log
                                        closure amplitude + MEM/TSV/flux/center/L1 priors
                                        (DPI_interferometry.py:86-172 with default flags)
    mri_workload(npix=64, ratio=4)    - C2: DPI_MRI npix=64, n_flow=16, knee-shaped phantom, 4x mask,
                                        sigma=5e-7, TV prior (DPI_MRI.py:88-147)

Both return plain dicts of CPU tensors / floats that the fused trainers (dpi_b200.trainer), the
drop-in callables and the test oracle all consume, so every implementation sees identical inputs.
"""
from __future__ import annotations

import numpy as np
import torch

from dpi_b200 import synthetic as syn
from dpi_b200.interferometry_helpers import Obs_params_torch


def interferometry_workload(npix=32, fov_uas=160.0, prior_fwhm_uas=50.0, seed=190, n_stations=7, n_scans=100,
                            flux=2.0, logdet=1.0, l1=1.0, tsv=100.0, flux_w=1000.0, center=1.0, mem=1024.0):
    obs = syn.SyntheticObs(n_stations=n_stations, n_scans=n_scans, seed=seed, min_vis=2)
    simim = syn.SimImage(npix, fov_uas)
    dft_mat, ktraj, pulsefac, cp_ind, cp_sign, ca_ind = Obs_params_torch(obs, simim, ttype='direct')
    n_vis, n_cp, n_ca = dft_mat.shape[1], len(cp_ind[0]), len(ca_ind[0])
    rng = np.random.RandomState(seed + 1)
    # ground truth through the same operator in float64, plus thermal-like noise on the closure products
    gt = syn.ground_truth_image(npix, flux)
    Fc = dft_mat[:, :, 0].double().numpy() + 1j * dft_mat[:, :, 1].double().numpy()
    vis = gt.reshape(-1) @ Fc
    ang = np.angle(vis)
    cphase = sum(cp_sign[q].numpy() * ang[cp_ind[q].numpy()] for q in range(3)) * 180 / np.pi
    amp = np.abs(vis)
    logcamp = np.log(amp[ca_ind[0].numpy()]) + np.log(amp[ca_ind[1].numpy()]) \
        - np.log(amp[ca_ind[2].numpy()]) - np.log(amp[ca_ind[3].numpy()])
    sigma_cp = rng.uniform(2.0, 20.0, n_cp)
    sigma_ca = rng.uniform(0.05, 0.5, n_ca)
    cphase_true = cphase + rng.normal(size=n_cp) * sigma_cp
    logcamp_true = logcamp + rng.normal(size=n_ca) * sigma_ca
    prior = syn.gaussian_prior_image(npix, fov_uas, prior_fwhm_uas, flux)
    w = dict(camp=1.0, cphase=n_cp / n_ca, visamp=0.0, l1=l1 * npix * npix / flux, tsv=tsv * npix * npix,
             flux=flux_w, center=center * 1e5 / (npix * npix), mem=mem, logdet=2.0 * logdet / n_ca)
    f32 = torch.float32
    return dict(kind="interferometry", npix=npix, n_vis=n_vis, n_cp=n_cp, n_ca=n_ca, flux=float(flux),
                dft_mat=dft_mat, ktraj=ktraj, pulsefac=pulsefac, cphase_ind=cp_ind, cphase_sign=cp_sign,
                camp_ind=ca_ind,
                cphase_true=torch.tensor(cphase_true, dtype=f32), logcamp_true=torch.tensor(logcamp_true, dtype=f32),
                sigma_cp=torch.tensor(sigma_cp, dtype=f32), sigma_ca=torch.tensor(sigma_ca, dtype=f32),
                prior=torch.tensor(prior, dtype=f32), w=w, log_scale_init=float(np.log(flux / (0.8 * npix * npix))),
                gt=torch.tensor(gt, dtype=f32))


def dpix_workload(npix=64, fov_uas=160.0, seed=77, n_stations=7, n_scans=25, n_gaussian=2, alpha=0.95, logdet=1.0):
    """C3: DPIx_interferometry.py with geometric_model='simple_crescent_floor_nuisance', data_product
    'cphase_logcamp' (flux_flag False -> 6 + 6 g = 18 parameters), M87-like coverage (7 stations, 25 scans),
    alpha-divergence 0.95 (DPIx_interferometry.py:205-326). Targets: a crescent + two Gaussians rendered by
    the same parametric model at fixed parameters, through the same operator, plus closure noise."""
    obs = syn.SyntheticObs(n_stations=n_stations, n_scans=n_scans, seed=seed, min_vis=2)
    simim = syn.SimImage(npix, fov_uas)
    dft_mat, ktraj, pulsefac, cp_ind, cp_sign, ca_ind = Obs_params_torch(obs, simim, ttype='direct')
    n_vis, n_cp, n_ca = dft_mat.shape[1], len(cp_ind[0]), len(ca_ind[0])
    rng = np.random.RandomState(seed + 1)
    gt_params = np.array([[0.45, 0.12, 0.5, 0.3, 0.6, 0.42, 0.2, 0.15, 0.2, 0.3, 0.35, 0.62, 0.15, 0.2, 0.12, 0.7,
                           0.1, 0.5][:6 + 6 * n_gaussian]])
    gt = _render_gt(gt_params, npix, fov_uas, n_gaussian)
    Fc = dft_mat[:, :, 0].double().numpy() + 1j * dft_mat[:, :, 1].double().numpy()
    vis = gt.reshape(-1) @ Fc
    ang, amp = np.angle(vis), np.abs(vis)
    cphase = sum(cp_sign[q].numpy() * ang[cp_ind[q].numpy()] for q in range(3)) * 180 / np.pi
    logcamp = np.log(amp[ca_ind[0].numpy()]) + np.log(amp[ca_ind[1].numpy()]) \
        - np.log(amp[ca_ind[2].numpy()]) - np.log(amp[ca_ind[3].numpy()])
    sigma_cp = rng.uniform(2.0, 20.0, n_cp)
    sigma_ca = rng.uniform(0.05, 0.5, n_ca)
    f32 = torch.float32
    w = dict(camp=1.0, cphase=n_cp / n_ca, logdet=2.0 * logdet / n_ca, scale_factor=1.0 / n_ca)
    return dict(kind="dpix", npix=npix, fov=float(fov_uas), n_gaussian=n_gaussian, nparams=6 + 6 * n_gaussian,
                n_vis=n_vis, n_cp=n_cp, n_ca=n_ca, alpha=float(alpha), dft_mat=dft_mat, cphase_ind=cp_ind,
                cphase_sign=cp_sign, camp_ind=ca_ind,
                cphase_true=torch.tensor(cphase + rng.normal(size=n_cp) * sigma_cp, dtype=f32),
                logcamp_true=torch.tensor(logcamp + rng.normal(size=n_ca) * sigma_ca, dtype=f32),
                sigma_cp=torch.tensor(sigma_cp, dtype=f32), sigma_ca=torch.tensor(sigma_ca, dtype=f32), w=w,
                gt=torch.tensor(gt, dtype=f32))


def _render_gt(params, npix, fov, n_gaussian):
    """Host float64 evaluation of the parametric image (geometric_model.py:333-381) for the synthetic target."""
    from scipy.special import erf
    hf = 0.5 * fov
    gap = 1.0 / npix
    xs = np.arange(-1 + gap, 1, 2 * gap)
    gy, gx = np.meshgrid(-xs, xs, indexing="ij")
    gr, gth = np.sqrt(gx ** 2 + gy ** 2), np.arctan2(gy, gx)
    p = params[0]
    r = 10.0 / hf + p[0] * 30.0 / hf
    sg = 1.0 / hf + p[1] * 39.0 / hf
    s = 1e-3 + p[2] * (0.99 - 1e-3)
    eta = 181 / 180 * np.pi * (2 * p[3] - 1)
    floor = p[4 + 6 * n_gaussian]
    cflux = 1e-3 + (2.0 - 1e-3) * p[5 + 6 * n_gaussian]
    ring = np.exp(-0.5 * (gr - r) ** 2 / sg ** 2) * (1 + s * np.cos(gth - eta))
    disk = 0.5 * (1 + erf((r - gr) / (np.sqrt(2) * sg)))
    img = cflux * ((1 - floor) * ring / (ring.sum() + 1e-4) + floor * disk / (disk.sum() + 1e-4))
    for k in range(n_gaussian):
        x0 = -200.0 / hf + p[4 + 6 * k] * 400.0 / hf
        y0 = -200.0 / hf + p[5 + 6 * k] * 400.0 / hf
        a = 1e-3 + p[6 + 6 * k] * (2.0 - 1e-3)
        sx = 1.0 / hf + p[7 + 6 * k] * 99.0 / hf
        sy = 1.0 / hf + p[8 + 6 * k] * 99.0 / hf
        th = 181 / 180 * 0.5 * np.pi * p[9 + 6 * k]
        xc, yc = gx - x0, gy - y0
        xr, yr = xc * np.cos(th) + yc * np.sin(th), -xc * np.sin(th) + yc * np.cos(th)
        n = np.exp(-0.5 * (xr ** 2 / sx ** 2 + yr ** 2 / sy ** 2)) / (2 * np.pi * sx * sy)
        img = img + a * n / (n.sum() + 1e-4)
    return img / (img.sum() + 1e-4)


def mri_workload(npix=64, ratio=4, sigma=5e-7, seed=0, tv=1e3, l1=0.0, logdet=1.0):
    img = syn.knee_phantom(npix)
    flux = float(img.sum())
    k = np.fft.fft2(img, norm="ortho")
    kspace = np.stack((k.real, k.imag), axis=-1)
    rng = np.random.RandomState(seed)
    kspace = kspace + rng.normal(size=kspace.shape) * sigma          # DPI_MRI.py:105
    mask = syn.mri_mask(npix, ratio, seed)                            # :107-111
    w = dict(kspace=1.0, l1=l1 / flux, tv=tv * npix / flux, logdet=logdet / (0.5 * float(mask.sum())))
    f32 = torch.float32
    return dict(kind="mri", npix=npix, sigma=float(sigma), flux=flux, mask=torch.tensor(mask, dtype=f32),
                kspace_true=torch.tensor(mask * kspace, dtype=f32), mean_mask=float(mask.mean()), w=w,
                log_scale_init=float(np.log(flux / (0.8 * npix * npix))), gt=torch.tensor(img, dtype=f32))
