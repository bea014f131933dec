"""Manual backward of one Glow flow step in the sampling direction (TEST INFRASTRUCTURE ONLY).

The CUDA path of Glow (dpi_b200/generative_model/glow_model.py) runs both inference directions; its training backward
is not built yet. This file writes that backward out op by op - the formulas the kernels will implement - and
tests/test_oracle_golden.py checks every one of them against torch autograd through oracle/glow_oracle.py (which is
bit-identical to the reference). Reference lines: generative_model/glow_model.py Flow.reverse :349-358,
AffineCoupling.reverse :299-320, InvConv2dLU :183-234, ActNorm.reverse :132-148, ZeroConv2d :238-251.

Conventions: x_out = Flow.reverse(y); upstream gradients g_x [B, C, h, w] (of x_out) and g_ld [B] (of the flow's
logdet contribution); returns g_y and a dict of parameter gradients keyed like the state dict.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from oracle import glow_oracle as G

Tensor = torch.Tensor


def _bn2d_backward(g: Tensor, pre: Tensor, gamma: Tensor, eps: float = 1e-5):
    """Training-mode BatchNorm2d backward. pre: the BN input [B, C, h, w]. Returns (g_pre, g_gamma, g_beta)."""
    dims = (0, 2, 3)
    n = pre.numel() / pre.shape[1]
    mean = pre.mean(dims, keepdim=True)
    var = pre.var(dims, unbiased=False, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + eps)
    xh = (pre - mean) * rstd
    g_beta = g.sum(dims)
    g_gamma = (g * xh).sum(dims)
    g_pre = gamma.view(1, -1, 1, 1) * rstd * (g - g_beta.view(1, -1, 1, 1) / n - xh * g_gamma.view(1, -1, 1, 1) / n)
    return g_pre, g_gamma, g_beta


def _conv_backward(g: Tensor, x: Tensor, w: Tensor, pad: int, pad_value: float = 0.0):
    """Gradients of conv2d(pad(x, pad_value), w) + b. Returns (g_x, g_w, g_b)."""
    xp = F.pad(x, [pad] * 4, value=pad_value) if pad else x
    g_w = torch.nn.grad.conv2d_weight(xp, w.shape, g)
    g_xp = torch.nn.grad.conv2d_input(xp.shape, w, g)
    g_x = g_xp[:, :, pad:xp.shape[2] - pad, pad:xp.shape[3] - pad] if pad else g_xp
    return g_x, g_w, g.sum((0, 2, 3))


def coupling_reverse_backward(sd: Dict[str, Tensor], p: str, y: Tensor, g_x: Tensor, g_ld: Tensor):
    """AffineCoupling.reverse: x = cat[a, b / exp(tanh(ls0)) - t], logdet = -sum tanh(ls0), (ls0, t) = net(a), train mode."""
    C = y.shape[1]
    a, b = y[:, : C // 2], y[:, C // 2:]
    q = p + "net."
    # forward, keeping what the backward needs
    c1 = F.conv2d(a, sd[q + "0.weight"], sd[q + "0.bias"], padding=1)
    l1 = F.leaky_relu(c1, G.LRELU_SLOPE)
    n1 = F.batch_norm(l1, None, None, sd[q + "2.weight"], sd[q + "2.bias"], True, 0.0, 1e-5)
    c2 = F.conv2d(n1, sd[q + "3.weight"], sd[q + "3.bias"])
    l2 = F.leaky_relu(c2, G.LRELU_SLOPE)
    n2 = F.batch_norm(l2, None, None, sd[q + "5.weight"], sd[q + "5.bias"], True, 0.0, 1e-5)
    zc = F.conv2d(F.pad(n2, [1, 1, 1, 1], value=1), sd[q + "6.conv.weight"], sd[q + "6.conv.bias"])
    e3 = torch.exp(sd[q + "6.scale"] * 3)
    out = zc * e3
    ls0, t = out[:, : C // 2], out[:, C // 2:]
    ls = torch.tanh(ls0)
    ei = torch.exp(-ls)
    # backward of the tail
    g_a, g_inb = g_x[:, : C // 2], g_x[:, C // 2:]
    g_b = g_inb * ei
    g_t = -g_inb
    g_ls = -g_inb * b * ei - g_ld.view(-1, 1, 1, 1)
    g_out = torch.cat([g_ls * (1 - ls * ls), g_t], 1)
    grads = {}
    # ZeroConv2d: out = zc * exp(3 scale)
    grads[q + "6.scale"] = (3 * g_out * out).sum((0, 2, 3)).view(1, -1, 1, 1)
    g_zc = g_out * e3
    g_n2, grads[q + "6.conv.weight"], grads[q + "6.conv.bias"] = _conv_backward(g_zc, n2, sd[q + "6.conv.weight"], 1, 1.0)
    g_l2, grads[q + "5.weight"], grads[q + "5.bias"] = _bn2d_backward(g_n2, l2, sd[q + "5.weight"])
    g_c2 = g_l2 * torch.where(c2 > 0, torch.ones_like(c2), torch.full_like(c2, G.LRELU_SLOPE))
    g_n1, grads[q + "3.weight"], grads[q + "3.bias"] = _conv_backward(g_c2, n1, sd[q + "3.weight"], 0)
    g_l1, grads[q + "2.weight"], grads[q + "2.bias"] = _bn2d_backward(g_n1, l1, sd[q + "2.weight"])
    g_c1 = g_l1 * torch.where(c1 > 0, torch.ones_like(c1), torch.full_like(c1, G.LRELU_SLOPE))
    g_a_net, grads[q + "0.weight"], grads[q + "0.bias"] = _conv_backward(g_c1, a, sd[q + "0.weight"], 1, 0.0)
    x_mid = torch.cat([a, b * ei - t], 1)
    return torch.cat([g_a + g_a_net, g_b], 1), grads, x_mid


def invconv_reverse_backward(sd: Dict[str, Tensor], p: str, u: Tensor, g_x: Tensor, g_ld: Tensor):
    """InvConv2dLU.reverse: x = W^-1 u per position, logdet = -h w sum w_s; W = P (L*lm + I) (U*um + diag(s exp(w_s)))."""
    h, w = u.shape[2], u.shape[3]
    Lm = sd[p + "w_l"] * sd[p + "l_mask"] + sd[p + "l_eye"]
    d = sd[p + "s_sign"] * torch.exp(sd[p + "w_s"])
    Um = sd[p + "w_u"] * sd[p + "u_mask"] + torch.diag(d)
    P = sd[p + "w_p"]
    V = (P @ Lm @ Um).inverse()
    C = u.shape[1]
    uf = u.permute(1, 0, 2, 3).reshape(C, -1)
    gf = g_x.permute(1, 0, 2, 3).reshape(C, -1)
    x = (V @ uf).view(C, u.shape[0], h, w).permute(1, 0, 2, 3)
    g_u = (V.t() @ gf).view(C, u.shape[0], h, w).permute(1, 0, 2, 3)
    G_V = gf @ uf.t()                                  # d / d V
    G_W = -V.t() @ G_V @ V.t()                         # through the inverse
    G_Lm = P.t() @ G_W @ Um.t()
    G_Um = (P @ Lm).t() @ G_W
    grads = {p + "w_l": G_Lm * sd[p + "l_mask"], p + "w_u": G_Um * sd[p + "u_mask"],
             p + "w_s": torch.diagonal(G_Um) * d - h * w * g_ld.sum()}
    return g_u, grads, x


def actnorm_reverse_backward(sd: Dict[str, Tensor], p: str, v: Tensor, g_x: Tensor, g_ld: Tensor):
    """ActNorm.reverse: x = v * scale_inv - loc, logdet = +h w sum log|scale_inv|."""
    h, w = v.shape[2], v.shape[3]
    si = sd[p + "scale_inv"]
    grads = {p + "loc": -g_x.sum((0, 2, 3), keepdim=True),
             p + "scale_inv": (g_x * v).sum((0, 2, 3), keepdim=True) + h * w * g_ld.sum() / si}
    return g_x * si, grads


def flow_reverse_backward(sd: Dict[str, Tensor], p: str, y: Tensor, g_x: Tensor, g_ld: Tensor) -> Tuple[Tensor, Dict[str, Tensor]]:
    """Backward of Flow.reverse (coupling^-1 -> invconv^-1 -> actnorm^-1), train-mode BatchNorm."""
    # recompute the intermediates in forward order
    _, _, u = coupling_reverse_backward(sd, p + "coupling.", y, torch.zeros_like(y), torch.zeros_like(g_ld))
    _, _, v = invconv_reverse_backward(sd, p + "invconv.", u, torch.zeros_like(u), torch.zeros_like(g_ld))
    g_v, grads = actnorm_reverse_backward(sd, p + "actnorm.", v, g_x, g_ld)
    g_u, g2, _ = invconv_reverse_backward(sd, p + "invconv.", u, g_v, g_ld)
    g_y, g3, _ = coupling_reverse_backward(sd, p + "coupling.", y, g_u, g_ld)
    grads.update(g2)
    grads.update(g3)
    return g_y, grads


def prior_sample_backward(sd: Dict[str, Tensor], p: str, y: Tensor, eps: Tensor, g_full: Tensor, g_ld: Tensor):
    """Split prior of Block.reverse: full = cat[y, mean + exp(log_sd) * eps], logdet = sum log_sd, (mean, log_sd) = ZeroConv2d(y)."""
    c = y.shape[1]
    zc = F.conv2d(F.pad(y, [1, 1, 1, 1], value=1), sd[p + "conv.weight"], sd[p + "conv.bias"])
    e3 = torch.exp(sd[p + "scale"] * 3)
    out = zc * e3
    c2 = out.shape[1] // 2
    lsd = out[:, c2:]
    g_y_direct, g_z = g_full[:, :c], g_full[:, c:]
    g_mean = g_z
    g_lsd = g_z * torch.exp(lsd) * eps + g_ld.view(-1, 1, 1, 1)
    g_out = torch.cat([g_mean, g_lsd], 1)
    grads = {p + "scale": (3 * g_out * out).sum((0, 2, 3)).view(1, -1, 1, 1)}
    g_y_conv, grads[p + "conv.weight"], grads[p + "conv.bias"] = _conv_backward(g_out * e3, y, sd[p + "conv.weight"], 1, 1.0)
    g_eps = g_z * torch.exp(lsd)
    return g_y_direct + g_y_conv, g_eps, grads
