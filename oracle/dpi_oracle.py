"""CPU oracle for the DPI training step - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state-dict in, tensors out) restatement in plain PyTorch-CPU of the reference's
variational-posterior training step. Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this file; the product path
(`dpi_b200/`) never does and fails loudly when its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors. This oracle is
pinned instead against the *live reference modules* imported in the build container
(`tests/test_oracle_vs_reference.py`, skipped where /root/reference is absent) and against the
fixtures `tests/golden/*.npz` that `tests/golden/make_golden.py` generated from those modules.
The NUFFT branch (torchkbnufft) and ehtim's closure-set selection are third-party and stay
"parity unpinned" (DESIGN.md).

Each function cites the reference lines it follows (paths relative to /root/reference).
Everything is dtype-generic: pass float64 tensors for a high-precision check.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

BN_EPS = 1e-2        # DPItorch/generative_model/realnvpfc_model.py:178,185
BN_MOMENTUM = 0.1    # nn.BatchNorm1d default
LRELU_SLOPE = 0.01   # realnvpfc_model.py:177,184


# ----------------------------------------------------------------------------------------------
# RealNVP (DPItorch/generative_model/realnvpfc_model.py)
# ----------------------------------------------------------------------------------------------
def make_orders(ndim: int, n_flow: int, permute: str = "random") -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """Fixed permutations, realnvpfc_model.py:568-581 (Order_inverse :553-559 is argsort)."""
    orders = []
    for i in range(n_flow):
        if permute == "random":
            orders.append(np.random.RandomState(seed=i).permutation(ndim))
        elif permute == "reverse":
            orders.append(np.arange(ndim - 1, -1, -1))
        else:
            orders.append(np.arange(ndim))
    inverse = []
    for o in orders:
        inv = np.empty(ndim, dtype=o.dtype)
        inv[o] = np.arange(ndim, dtype=o.dtype)
        inverse.append(inv)
    return orders, inverse


def hidden_width(ndim: int, seqfrac: float) -> int:
    return int(ndim / (2 * seqfrac))  # realnvpfc_model.py:176


def init_state(ndim: int, n_flow: int, seqfrac: float = 4, seed: int = 0, randomize: bool = False,
               dtype=torch.float32, batch_norm: bool = True) -> State:
    """A state dict with the reference's keys/shapes.

    randomize=False reproduces the constructor's distributions (N(0,0.05) on net.0/net.3, zero
    ZeroFC, zero ActNorm; realnvpfc_model.py:11-12,131-135,189-193). randomize=True makes every
    tensor non-trivial so parity checks exercise all terms.
    """
    g = torch.Generator().manual_seed(seed)
    H = hidden_width(ndim, seqfrac)
    Da, Dout = ndim - ndim // 2, 2 * (ndim // 2)
    sd: State = {}

    def rn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g, dtype=torch.float64) * std).to(dtype)

    for i in range(n_flow):
        for an in ("actnorm", "actnorm2"):
            p = f"blocks.{i}.{an}."
            sd[p + "loc"] = rn(1, std=0.1) if randomize else torch.zeros(1, dtype=dtype)
            sd[p + "log_scale_inv"] = rn(1, std=0.1) if randomize else torch.zeros(1, dtype=dtype)
            sd[p + "initialized"] = torch.tensor(1 if randomize else 0, dtype=torch.uint8)
        for cn in ("coupling", "coupling2"):  # key order = the reference module's state_dict order
            p = f"blocks.{i}.{cn}.net."

            def bn_layer(idx):
                sd[p + idx + ".weight"] = (1 + rn(H, std=0.1)) if randomize else torch.ones(H, dtype=dtype)
                sd[p + idx + ".bias"] = rn(H, std=0.1) if randomize else torch.zeros(H, dtype=dtype)
                sd[p + idx + ".running_mean"] = torch.zeros(H, dtype=dtype)
                sd[p + idx + ".running_var"] = torch.ones(H, dtype=dtype)
                sd[p + idx + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)

            sd[p + "0.weight"] = rn(H, Da, std=0.05)
            sd[p + "0.bias"] = rn(H, std=0.05) if randomize else torch.zeros(H, dtype=dtype)
            if batch_norm:
                bn_layer("2")
            second = "3" if batch_norm else "2"
            sd[p + second + ".weight"] = rn(H, H, std=0.05)
            sd[p + second + ".bias"] = rn(H, std=0.05) if randomize else torch.zeros(H, dtype=dtype)
            if batch_norm:
                bn_layer("5")
            last = "6" if batch_norm else "4"
            sd[p + last + ".scale"] = rn(Dout, std=0.05) if randomize else torch.zeros(Dout, dtype=dtype)
            sd[p + last + ".fc.weight"] = rn(Dout, H, std=0.02) if randomize else torch.zeros(Dout, H, dtype=dtype)
            sd[p + last + ".fc.bias"] = rn(Dout, std=0.02) if randomize else torch.zeros(Dout, dtype=dtype)
    return sd


def param_keys(sd: State) -> List[str]:
    """Keys of trainable tensors in registration order (= nn.Module.parameters() order)."""
    return [k for k in sd if not (k.endswith("initialized") or k.endswith("running_mean")
                                  or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


def _bn(sd: State, p: str, x: Tensor, train: bool) -> Tensor:
    """nn.BatchNorm1d(H, eps=1e-2): batch statistics + running-stat update in train mode."""
    if train:
        rm, rv = sd[p + "running_mean"], sd[p + "running_var"]
        y = F.batch_norm(x, rm, rv, sd[p + "weight"], sd[p + "bias"], True, BN_MOMENTUM, BN_EPS)
        sd[p + "num_batches_tracked"] += 1
        return y
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"],
                        sd[p + "bias"], False, BN_MOMENTUM, BN_EPS)


def coupling_net(sd: State, p: str, a: Tensor, train: bool) -> Tensor:
    """AffineCoupling.net, realnvpfc_model.py:171-187 (+ ZeroFC :137-141)."""
    bn = (p + "6.scale") in sd
    h = F.leaky_relu(F.linear(a, sd[p + "0.weight"], sd[p + "0.bias"]), LRELU_SLOPE)
    if bn:
        h = _bn(sd, p + "2.", h, train)
        h = F.leaky_relu(F.linear(h, sd[p + "3.weight"], sd[p + "3.bias"]), LRELU_SLOPE)
        h = _bn(sd, p + "5.", h, train)
        last = p + "6."
    else:
        h = F.leaky_relu(F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"]), LRELU_SLOPE)
        last = p + "4."
    out = F.linear(h, sd[last + "fc.weight"], sd[last + "fc.bias"])
    return out * torch.exp(sd[last + "scale"] * 3)


def _split(x: Tensor) -> Tuple[Tensor, Tensor]:
    d = x.shape[1]
    da = d - d // 2  # torch.chunk(2, 1): first chunk gets ceil(d/2)
    return x[:, :da], x[:, da:]


def coupling_reverse(sd: State, p: str, out: Tensor, train: bool) -> Tuple[Tensor, Tensor]:
    """AffineCoupling.reverse, realnvpfc_model.py:240-249."""
    a, b = _split(out)
    log_s0, t = coupling_net(sd, p + "net.", a, train).chunk(2, 1)
    log_s = torch.tanh(log_s0)
    in_b = b / torch.exp(log_s) - t
    return torch.cat([a, in_b], 1), -log_s.sum(1)


def coupling_forward(sd: State, p: str, x: Tensor, train: bool) -> Tuple[Tensor, Tensor]:
    """AffineCoupling.forward, realnvpfc_model.py:222-231."""
    a, b = _split(x)
    log_s0, t = coupling_net(sd, p + "net.", a, train).chunk(2, 1)
    log_s = torch.tanh(log_s0)
    out_b = (b + t) * torch.exp(log_s)
    return torch.cat([a, out_b], 1), log_s.sum(1)


def actnorm_reverse(sd: State, p: str, y: Tensor) -> Tuple[Tensor, Tensor]:
    """ActNorm.reverse, realnvpfc_model.py:49-66 (first call zero-initialises, :23-25)."""
    if int(sd[p + "initialized"]) == 0:
        with torch.no_grad():
            sd[p + "loc"].zero_()
            sd[p + "log_scale_inv"].zero_()
        sd[p + "initialized"].fill_(1)
    lsi = sd[p + "log_scale_inv"]
    return y * torch.exp(lsi) - sd[p + "loc"], y.shape[1] * lsi.sum()


def actnorm_forward(sd: State, p: str, x: Tensor) -> Tuple[Tensor, Tensor]:
    """ActNorm.forward, realnvpfc_model.py:30-47 (first call: global scalar mean/std init :18-28)."""
    if int(sd[p + "initialized"]) == 0:
        with torch.no_grad():
            sd[p + "loc"].copy_(-x.mean().reshape(1))
            sd[p + "log_scale_inv"].copy_(torch.log(x.std().reshape(1) + 1e-6))
        sd[p + "initialized"].fill_(1)
    lsi = sd[p + "log_scale_inv"]
    return (1.0 / torch.exp(lsi)) * (x + sd[p + "loc"]), -x.shape[1] * lsi.sum()


def n_flow_of(sd: State) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


def realnvp_reverse(sd: State, z: Tensor, train: bool = True, permute: str = "random") -> Tuple[Tensor, Tensor]:
    """RealNVP.reverse (:594-603) over Flow.reverse (:457-474). Mutates BN running stats in `sd`."""
    ndim, n_flow = z.shape[1], n_flow_of(sd)
    _, inv_orders = make_orders(ndim, n_flow, permute)
    flip = torch.arange(ndim - 1, -1, -1)
    x, logdet = z, 0
    for i in range(n_flow - 1, -1, -1):
        p = f"blocks.{i}."
        x = x[:, torch.as_tensor(inv_orders[i])]
        x = x[:, flip]
        x, d1 = coupling_reverse(sd, p + "coupling2.", x, train)
        x, d2 = actnorm_reverse(sd, p + "actnorm2.", x)
        x = x[:, flip]
        x, d3 = coupling_reverse(sd, p + "coupling.", x, train)
        x, d4 = actnorm_reverse(sd, p + "actnorm.", x)
        logdet = logdet + (d1 + d2 + d3 + d4)
    return x, logdet


def realnvp_forward(sd: State, x: Tensor, train: bool = True, permute: str = "random") -> Tuple[Tensor, Tensor]:
    """RealNVP.forward (:583-592) over Flow.forward (:439-455)."""
    ndim, n_flow = x.shape[1], n_flow_of(sd)
    orders, _ = make_orders(ndim, n_flow, permute)
    flip = torch.arange(ndim - 1, -1, -1)
    out, logdet = x, 0
    for i in range(n_flow):
        p = f"blocks.{i}."
        out, d1 = actnorm_forward(sd, p + "actnorm.", out)
        out, d2 = coupling_forward(sd, p + "coupling.", out, train)
        out = out[:, flip]
        out, d3 = actnorm_forward(sd, p + "actnorm2.", out)
        out, d4 = coupling_forward(sd, p + "coupling2.", out, train)
        out = out[:, flip]
        logdet = logdet + (d1 + d2 + d3 + d4)
        out = out[:, torch.as_tensor(orders[i])]
    return out, logdet


# ----------------------------------------------------------------------------------------------
# positivity + scale (DPItorch/DPI_interferometry.py:205-211, DPI_MRI.py:174-179)
# ----------------------------------------------------------------------------------------------
def positivity(x: Tensor, log_scale: Tensor, npix: int) -> Tuple[Tensor, Tensor]:
    x = x.reshape(-1, npix, npix)
    sp = F.softplus(x)
    img = sp * torch.exp(log_scale)
    det = (x - sp).sum((1, 2)) + log_scale * npix * npix
    return img, det


# ----------------------------------------------------------------------------------------------
# interferometry forward model + losses (DPItorch/interferometry_helpers.py)
# ----------------------------------------------------------------------------------------------
def eht_observe(img: Tensor, dft_mat: Tensor, cphase_ind: Sequence[Tensor], cphase_sign: Sequence[Tensor],
                camp_ind: Sequence[Tensor]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """eht_observation_pytorch.func, ttype='direct' (:257-291)."""
    eps = 1e-16
    npix2 = dft_mat.shape[0]
    x = img.reshape(-1, npix2).to(dft_mat.dtype)
    vis = torch.stack([x @ dft_mat[:, :, 0], x @ dft_mat[:, :, 1]], 1)  # [B,2,n_vis] (:36-39)
    amp = torch.sqrt(vis[:, 0] ** 2 + vis[:, 1] ** 2 + eps)
    ang = [torch.atan2(vis[:, 1, i], vis[:, 0, i]) for i in cphase_ind]
    cphase = (cphase_sign[0] * ang[0] + cphase_sign[1] * ang[1] + cphase_sign[2] * ang[2]) * 180 / np.pi

    def a(i):
        v = vis[:, :, i]
        return torch.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2 + eps)

    logcamp = torch.log(a(camp_ind[0])) + torch.log(a(camp_ind[1])) - torch.log(a(camp_ind[2])) \
        - torch.log(a(camp_ind[3]))
    return vis, amp, cphase, logcamp


def loss_angle_diff(sigma: Tensor, y_true: Tensor, y_pred: Tensor) -> Tensor:
    """Loss_angle_diff (:299-307)."""
    d = y_true * np.pi / 180 - y_pred * np.pi / 180
    return 2.0 * torch.mean((1 - torch.cos(d)) / (sigma * np.pi / 180) ** 2, 1)


def loss_logca_diff2(sigma, y_true, y_pred):
    """Loss_logca_diff2 (:318-323); Loss_visamp_diff (:342-347) has the same form."""
    return torch.mean((y_true - y_pred) ** 2 / sigma ** 2, 1)


def loss_vis_diff(sigma, y_true, y_pred):
    """Loss_vis_diff (:326-331)."""
    return torch.mean(((y_true[0] - y_pred[:, 0]) ** 2 + (y_true[1] - y_pred[:, 1]) ** 2) / sigma ** 2, 1)


def loss_logamp_diff(sigma, y_true, y_pred):
    """Loss_logamp_diff (:334-339)."""
    return torch.mean(y_true ** 2 / sigma ** 2 * (torch.log(y_true) - torch.log(y_pred)) ** 2, 1)


def loss_l1(img):
    return torch.mean(torch.abs(img), (-1, -2))  # :349


def loss_tsv(img):
    return torch.mean((img[:, 1:, :] - img[:, :-1, :]) ** 2, (-1, -2)) + \
        torch.mean((img[:, :, 1:] - img[:, :, :-1]) ** 2, (-1, -2))  # :354


def loss_tv(img):
    return torch.mean(torch.abs(img[:, 1:, :] - img[:, :-1, :]), (-1, -2)) + \
        torch.mean(torch.abs(img[:, :, 1:] - img[:, :, :-1]), (-1, -2))  # :358


def loss_flux(flux, img):
    return (torch.sum(img, (-1, -2)) - flux) ** 2  # :363


def loss_center(img, center, dim):
    """Loss_center (:369-382): X is the column index, Y the row index."""
    X = torch.arange(dim, dtype=img.dtype).reshape(1, dim).expand(dim, dim)
    Y = torch.arange(dim, dtype=img.dtype).reshape(dim, 1).expand(dim, dim)
    fl = torch.mean(img, (-1, -2))
    xc = torch.mean(img * X, (-1, -2)) / fl
    yc = torch.mean(img * Y, (-1, -2)) / fl
    return 0.5 * ((xc - center) ** 2 + (yc - center) ** 2)


def loss_cross_entropy(prior, img):
    return torch.mean(img * (torch.log(img + 1e-12) - torch.log(prior + 1e-12)), (-1, -2))  # :384


def interferometry_weights(npix, flux, n_cp, n_ca, l1=1.0, tsv=100.0, flux_w=1000.0, center=1.0,
                           mem=1024.0, logdet=1.0):
    """Loss weights, DPI_interferometry.py:151-159."""
    return dict(camp=1.0, cphase=n_cp / n_ca, visamp=0.0,
                l1=l1 * npix * npix / flux, tsv=tsv * npix * npix, flux=flux_w,
                center=center * 1e5 / (npix * npix), mem=mem, logdet=2.0 * logdet / n_ca)


def interferometry_loss(sd: State, log_scale: Tensor, z: Tensor, c: dict, train: bool = True):
    """Loop body DPI_interferometry.py:201-229 (realnvp branch). `c` holds the operator constants:
    npix, dft_mat, cphase_ind/sign, camp_ind, cphase_true, logcamp_true, sigma_cp, sigma_ca, prior,
    flux, w (weights dict). Returns (loss, dict of logged terms)."""
    npix, w = c["npix"], c["w"]
    x, logdet = realnvp_reverse(sd, z, train)
    img, det = positivity(x, log_scale, npix)
    logdet = logdet + det
    vis, amp, cphase, logcamp = eht_observe(img, c["dft_mat"], c["cphase_ind"], c["cphase_sign"], c["camp_ind"])
    l_center = loss_center(img, npix / 2 - 0.5, npix)
    l_l1, l_tsv = loss_l1(img), loss_tsv(img)
    l_mem = loss_cross_entropy(c["prior"], img)
    l_flux = loss_flux(c["flux"], img)
    l_cp = loss_angle_diff(c["sigma_cp"], c["cphase_true"], cphase)
    l_ca = loss_logca_diff2(c["sigma_ca"], c["logcamp_true"], logcamp)
    loss_data = w["camp"] * l_ca + w["cphase"] * l_cp
    loss_prior = w["mem"] * l_mem + w["flux"] * l_flux + w["tsv"] * l_tsv + w["center"] * l_center + w["l1"] * l_l1
    loss = torch.mean(loss_data) + torch.mean(loss_prior) - w["logdet"] * torch.mean(logdet)
    terms = dict(loss=loss, cphase=l_cp.mean(), logca=l_ca.mean(), prior=loss_prior.mean(),
                 logdet=logdet.mean(), flux=l_flux.mean(), tsv=l_tsv.mean(), center=l_center.mean(),
                 mem=l_mem.mean(), l1=l_l1.mean(), img=img, x=x, cphase_pred=cphase, logcamp_pred=logcamp,
                 vis=vis, visamp=amp, logdet_vec=logdet)
    return loss, terms


# ----------------------------------------------------------------------------------------------
# alpha-DPI: geometric crescent renderer + alpha-divergence objective
# (DPItorch/geometric_model.py:237-389, DPItorch/DPIx_interferometry.py:329-380)
# ----------------------------------------------------------------------------------------------
CRESCENT_RANGES = dict(r_range=(10.0, 40.0), asym_range=(1e-3, 0.99), width_range=(1.0, 40.0),
                       floor_range=(0.0, 1.0), crescent_flux_range=(1e-3, 2.0), shift_range=(-200.0, 200.0),
                       sigma_range=(1.0, 100.0), gaussian_scale_range=(1e-3, 2.0))   # geometric_model.py:238-240


def crescent_grids(npix: int, dtype=torch.float32):
    """geometric_model.py:265-272: xs = arange(-1 + 1/npix, 1, 2/npix); grid_y, grid_x = meshgrid(-xs, xs)."""
    gap = 1.0 / npix
    xs = torch.arange(-1 + gap, 1, 2 * gap, dtype=dtype)
    grid_y, grid_x = torch.meshgrid(-xs, xs, indexing="ij")
    return grid_x, grid_y, torch.sqrt(grid_x ** 2 + grid_y ** 2), torch.atan2(grid_y, grid_x)


def crescent_render(params: Tensor, npix: int, fov: float = 160.0, n_gaussian: int = 2, eps: float = 1e-4) -> Tensor:
    """SimpleCrescentNuisanceFloor_Param2Img.forward with flux_flag=False (geometric_model.py:275-381):
    params [B, 6 + 6 g] in (0, 1) -> unit-flux images [B, npix, npix]."""
    R = CRESCENT_RANGES
    hf = 0.5 * fov
    grid_x, grid_y, grid_r, grid_th = crescent_grids(npix, params.dtype)
    col = lambda i: params[:, i].unsqueeze(-1).unsqueeze(-1)
    r = R["r_range"][0] / hf + col(0) * (R["r_range"][1] - R["r_range"][0]) / hf
    sigma = R["width_range"][0] / hf + col(1) * (R["width_range"][1] - R["width_range"][0]) / hf
    s = R["asym_range"][0] + col(2) * (R["asym_range"][1] - R["asym_range"][0])
    eta = 181 / 180 * np.pi * (2.0 * col(3) - 1.0)
    floor = R["floor_range"][0] + (R["floor_range"][1] - R["floor_range"][0]) * col(4 + n_gaussian * 6)
    cflux = R["crescent_flux_range"][0] + (R["crescent_flux_range"][1] - R["crescent_flux_range"][0]) * col(5 + n_gaussian * 6)
    ring = torch.exp(-0.5 * (grid_r - r) ** 2 / sigma ** 2)
    S = 1 + s * torch.cos(grid_th - eta)
    crescent = S * ring
    disk = 0.5 * (1 + torch.erf((r - grid_r) / (np.sqrt(2) * sigma)))
    crescent = crescent / (torch.sum(crescent, (-1, -2)).unsqueeze(-1).unsqueeze(-1) + eps)
    disk = disk / (torch.sum(disk, (-1, -2)).unsqueeze(-1).unsqueeze(-1) + eps)
    crescent = cflux * ((1 - floor) * crescent + floor * disk)
    for k in range(n_gaussian):
        sh = R["shift_range"]
        x_shift = sh[0] / hf + col(4 + k * 6) * (sh[1] - sh[0]) / hf
        y_shift = sh[0] / hf + col(5 + k * 6) * (sh[1] - sh[0]) / hf
        scale = R["gaussian_scale_range"][0] + col(6 + k * 6) * (R["gaussian_scale_range"][1] - R["gaussian_scale_range"][0])
        sigma_x = R["sigma_range"][0] / hf + col(7 + k * 6) * (R["sigma_range"][1] - R["sigma_range"][0]) / hf
        sigma_y = R["sigma_range"][0] / hf + col(8 + k * 6) * (R["sigma_range"][1] - R["sigma_range"][0]) / hf
        theta = 181 / 180 * 0.5 * np.pi * col(9 + k * 6)
        x_c, y_c = grid_x - x_shift, grid_y - y_shift
        x_rot = x_c * torch.cos(theta) + y_c * torch.sin(theta)
        y_rot = -x_c * torch.sin(theta) + y_c * torch.cos(theta)
        delta = 0.5 * (x_rot ** 2 / sigma_x ** 2 + y_rot ** 2 / sigma_y ** 2)
        nuis = 1 / (2 * np.pi * sigma_x * sigma_y) * torch.exp(-delta)
        nuis = nuis / (torch.sum(nuis, (-1, -2)).unsqueeze(-1).unsqueeze(-1) + eps)
        crescent = crescent + scale * nuis
    return crescent / (torch.sum(crescent, (-1, -2)).unsqueeze(-1).unsqueeze(-1) + eps)


def dpix_weights(n_cp: int, n_ca: int, logdet: float = 1.0):
    """data_product 'cphase_logcamp' (DPIx_interferometry.py:268-274)."""
    return dict(camp=1.0, cphase=n_cp / n_ca, logdet=2.0 * logdet / n_ca, scale_factor=1.0 / n_ca)


def dpix_loss(sd: State, z: Tensor, c: dict, data_weight: float = 1.0, train: bool = True):
    """Loop body DPIx_interferometry.py:336-380 for data_product 'cphase_logcamp'. `c`: npix, fov, n_gaussian,
    dft_mat, cphase_ind/sign, camp_ind, cphase_true, logcamp_true, sigma_cp, sigma_ca, w, alpha ('KL' when 1)."""
    w, sf, alpha = c["w"], c["w"]["scale_factor"], c["alpha"]
    ps, logdet = realnvp_reverse(sd, z, train)
    params = torch.sigmoid(ps)
    img = crescent_render(params, c["npix"], c["fov"], c["n_gaussian"])
    det_sigmoid = torch.sum(-ps - 2 * torch.nn.functional.softplus(-ps), -1)
    logdet = logdet + det_sigmoid
    vis, amp, cphase, logcamp = eht_observe(img, c["dft_mat"], c["cphase_ind"], c["cphase_sign"], c["camp_ind"])
    l_cp = loss_angle_diff(c["sigma_cp"], c["cphase_true"], cphase)
    l_ca = loss_logca_diff2(c["sigma_ca"], c["logcamp_true"], logcamp)
    loss_data = 0.5 * (w["camp"] * l_ca + w["cphase"] * l_cp) / sf
    logprob = -logdet - 0.5 * torch.sum(z ** 2, 1)
    li = data_weight * loss_data + logprob
    lo = loss_data + logprob
    if alpha == 1:
        loss = torch.mean(sf * li)
        loss_orig = torch.mean(sf * lo)
    else:
        rej = torch.softmax(-(1 - alpha) * li, 0).detach()
        loss = torch.sum(rej * sf * li)
        loss_orig = sf * torch.log(torch.mean(torch.exp(-(1 - alpha) * lo))) / (alpha - 1)
    terms = dict(loss=loss, loss_orig=loss_orig, cphase=l_cp.mean(), logca=l_ca.mean(), logdet=logdet.mean(),
                 img=img, params=params, logdet_vec=logdet, loss_i=li)
    return loss, terms


# ----------------------------------------------------------------------------------------------
# MRI forward model + losses (DPItorch/DPI_MRI.py, MRI_helpers.py)
# ----------------------------------------------------------------------------------------------
def fft2c(img: Tensor) -> Tensor:
    """fft2c_torch (DPI_MRI.py:70-74): complex ortho FFT2 of a real image, [B,n,n,2].
    The reference calls the removed torch.fft(x, 2, normalized=True); this is the same transform
    (identical to its numpy twin fft2c at :63-67)."""
    return torch.view_as_real(torch.fft.fft2(img, norm="ortho"))


def loss_kspace_diff2(sigma, y_true, y_pred):
    return torch.mean((y_pred - y_true) ** 2, (1, 2, 3)) / sigma ** 2  # MRI_helpers.py:23-26


def loss_kspace_diff(sigma, y_true, y_pred):
    return torch.mean(torch.abs(y_pred - y_true), (1, 2, 3)) / sigma  # MRI_helpers.py:18-21


def mri_weights(npix, flux, mask, l1=0.0, tv=1e3, logdet=1.0):
    """DPI_MRI.py:137-142."""
    return dict(kspace=1.0, l1=l1 / flux, tv=tv * npix / flux, logdet=logdet / (0.5 * float(np.sum(mask))))


def mri_loss(sd: State, log_scale: Tensor, z: Tensor, c: dict, train: bool = True):
    """Loop body DPI_MRI.py:170-193. `c`: npix, mask [n,n,2], kspace_true [n,n,2] (already masked),
    sigma, w."""
    npix, w = c["npix"], c["w"]
    x, logdet = realnvp_reverse(sd, z, train)
    img, det = positivity(x, log_scale, npix)
    logdet = logdet + det
    kpred = fft2c(img)
    l_data = loss_kspace_diff2(c["sigma"], c["kspace_true"], kpred * c["mask"]) / float(c["mask"].mean())
    l_l1 = loss_l1(img) if w["l1"] > 0 else 0
    l_tv = loss_tv(img) if w["tv"] > 0 else 0
    l_prior = w["tv"] * l_tv + w["l1"] * l_l1
    loss = torch.mean(l_data) + torch.mean(l_prior) - w["logdet"] * torch.mean(logdet)
    terms = dict(loss=loss, kspace=l_data.mean(), prior=torch.as_tensor(l_prior).mean(), logdet=logdet.mean(),
                 img=img, x=x, logdet_vec=logdet, kspace_vec=l_data)
    return loss, terms


# ----------------------------------------------------------------------------------------------
# toy 2-D posteriors (DPItorch/DPI_2Dtoydist.py:15-37,64-72)
# ----------------------------------------------------------------------------------------------
def toy_logprob(kind: int, x: Tensor, y: Tensor) -> Tensor:
    if kind == 1:
        s = 0.4
        cs = [(-0.5, -0.5), (-0.5, 0.5), (0.5, -0.5), (0.5, 0.5)]
        return torch.log(sum(torch.exp(-1 / s ** 2 * ((x - a) ** 2 + (y - b) ** 2)) for a, b in cs))
    if kind == 2:
        return -0.5 * ((torch.sqrt((4 * x) ** 2 + (4 * y) ** 2) - 2) / 0.6) ** 2 + \
            torch.log(torch.exp(-0.5 * (4 * x - 2) ** 2 / 0.6 ** 2) + torch.exp(-0.5 * (4 * x + 2) ** 2 / 0.6 ** 2))
    return -0.5 * ((4 * y - torch.sin(2 * np.pi * x)) / 0.4) ** 2


def toy_loss(sd: State, z: Tensor, kind: int = 1, diversity: float = 1.0, train: bool = True):
    xt, logdet = realnvp_reverse(sd, z, train)
    xs = 2 * torch.sigmoid(xt) - 1
    det_sig = torch.sum(-xt - 2 * F.softplus(-xt), -1)
    logdet = logdet + det_sig
    loss = torch.mean(-toy_logprob(kind, xs[:, 0], xs[:, 1]) - diversity * logdet)
    return loss, dict(loss=loss, x=xt, logdet_vec=logdet)


# ----------------------------------------------------------------------------------------------
# optimiser (torch.nn.utils.clip_grad_norm_ + torch.optim.Adam defaults;
# DPI_interferometry.py:172,233-235)
# ----------------------------------------------------------------------------------------------
def clip_grad_norm(grads: Sequence[Tensor], max_norm: float) -> Tuple[Tensor, Tensor]:
    """Returns (total_norm, clip_coef) with torch's exact form coef=min(1, max_norm/(norm+1e-6))."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, coef


def adam_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """One torch.optim.Adam step (no weight decay, no amsgrad), in place on p, m, v.
    p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)  (SURVEY Appendix B)."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


class OracleTrainer:
    """The reference training loop assembled from the functions above, with autograd supplying
    the gradients, torch's own clip_grad_norm_ and optim.Adam doing the update. This is what the
    CPU-baseline leg of bench.py times."""

    def __init__(self, sd: State, log_scale: Optional[Tensor], loss_fn, consts: dict, lr: float, clip: float):
        self.sd = sd
        self.keys = param_keys(sd)
        for k in self.keys:
            sd[k].requires_grad_(True)
        self.log_scale = log_scale
        if log_scale is not None:
            log_scale.requires_grad_(True)
        self.params = [sd[k] for k in self.keys] + ([log_scale] if log_scale is not None else [])
        self.opt = torch.optim.Adam(self.params, lr=lr)
        self.loss_fn, self.consts, self.clip = loss_fn, consts, clip

    def loss_and_grads(self, z: Tensor):
        self.opt.zero_grad()
        if self.log_scale is not None:
            loss, terms = self.loss_fn(self.sd, self.log_scale, z, self.consts)
        elif self.loss_fn is dpix_loss:
            loss, terms = self.loss_fn(self.sd, z, self.consts, data_weight=getattr(self, "data_weight", 1.0))
        else:
            loss, terms = self.loss_fn(self.sd, z, **self.consts)
        loss.backward()
        return loss, terms

    def step(self, z: Tensor):
        loss, terms = self.loss_and_grads(z)
        norm = torch.nn.utils.clip_grad_norm_(self.params, self.clip)
        self.opt.step()
        terms["grad_norm"] = norm
        return loss.detach(), terms
