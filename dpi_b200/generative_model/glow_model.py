"""Glow flow of DPI (`--model_form glow`), sampling direction, on the B200 kernels.

Mirrors the module and state_dict surface of the reference's generative_model/glow_model.py (Glow, Block, Flow,
ActNorm, InvConv2dLU, AffineCoupling, ZeroConv2d, calc_z_shapes; same parameter / buffer names and shapes, so
checkpoints written by the reference load unchanged). What runs on the GPU is `Glow.reverse(z_list) -> (image,
logdet)` (:526-539) and `Glow.forward(image) -> (log_p, logdet, z_outs)` (:508-524), forward only: train mode
(BatchNorm batch statistics, running statistics updated) and eval mode. NOT built yet: gradients through them (the
calls run under no_grad and return tensors without history). SURVEY.md 8 row a20.

All activations are feature-major image stacks [channel][b*h*w]; every convolution is a tcgen05 GEMM (conv_tc.cu),
the rest are the small kernels of glow.cu. There is no CPU path.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
from torch import nn

from dpi_b200 import _lib
from dpi_b200.generative_model import glow_ops

LRELU_SLOPE = 0.1     # glow_model.py:262,265
BN_EPS = 1e-5         # nn.BatchNorm2d default (:263,266)
FILTER = 512          # AffineCoupling(filter_size=512)


def calc_z_shapes(n_channel: int, input_size: int, n_flow: int, n_block: int) -> List[Tuple[int, int, int]]:
    """glow_model.py:541-553."""
    shapes = []
    for _ in range(n_block - 1):
        input_size //= 2
        n_channel *= 2
        shapes.append((n_channel, input_size, input_size))
    input_size //= 2
    shapes.append((n_channel * 4, input_size, input_size))
    return shapes


class ActNorm(nn.Module):
    """Per-channel affine (:80-148); `scale_inv` is stored directly, not as a logarithm."""

    def __init__(self, in_channel: int):
        super().__init__()
        self.loc = nn.Parameter(torch.zeros(1, in_channel, 1, 1))
        self.scale_inv = nn.Parameter(torch.ones(1, in_channel, 1, 1))
        self.register_buffer("initialized", torch.tensor(0, dtype=torch.uint8))


class InvConv2dLU(nn.Module):
    """Invertible 1x1 convolution parametrised as P (L + I) (U + diag(sign exp(w_s))) (:183-234)."""

    def __init__(self, in_channel: int):
        super().__init__()
        from scipy import linalg as la
        q, _ = la.qr(np.random.randn(in_channel, in_channel))
        p_mat, l_mat, u_mat = la.lu(q.astype(np.float32))
        s_vec = np.diag(u_mat).copy()
        upper = np.triu(np.ones((in_channel, in_channel), dtype=np.float32), 1)
        self.register_buffer("w_p", torch.from_numpy(np.ascontiguousarray(p_mat)))
        self.register_buffer("u_mask", torch.from_numpy(upper.copy()))
        self.register_buffer("l_mask", torch.from_numpy(upper.T.copy()))
        self.register_buffer("s_sign", torch.sign(torch.from_numpy(s_vec)))
        self.register_buffer("l_eye", torch.eye(in_channel))
        self.w_l = nn.Parameter(torch.from_numpy(np.ascontiguousarray(l_mat)))
        self.w_s = nn.Parameter(torch.log(torch.abs(torch.from_numpy(s_vec))))
        self.w_u = nn.Parameter(torch.from_numpy(np.triu(u_mat, 1).copy()))


class ZeroConv2d(nn.Module):
    """3x3 convolution over an input padded with ONES, times exp(3 scale); zero-initialised (:238-251)."""

    def __init__(self, in_channel: int, out_channel: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channel, out_channel, 3, padding=0)
        nn.init.zeros_(self.conv.weight)
        nn.init.zeros_(self.conv.bias)
        self.scale = nn.Parameter(torch.zeros(1, out_channel, 1, 1))


class AffineCoupling(nn.Module):
    def __init__(self, in_channel: int, filter_size: int = FILTER):
        super().__init__()
        # containers with the reference's indices (:260-268); their own forward() is never called
        self.net = nn.Sequential(nn.Conv2d(in_channel // 2, filter_size, 3, padding=1), nn.LeakyReLU(LRELU_SLOPE),
                                 nn.BatchNorm2d(filter_size), nn.Conv2d(filter_size, filter_size, 1),
                                 nn.LeakyReLU(LRELU_SLOPE), nn.BatchNorm2d(filter_size), ZeroConv2d(filter_size, in_channel))
        for i in (0, 3):                                                   # :270-274
            nn.init.normal_(self.net[i].weight, 0, 0.05)
            nn.init.zeros_(self.net[i].bias)


class Flow(nn.Module):
    def __init__(self, in_channel: int):
        super().__init__()
        self.actnorm = ActNorm(in_channel)
        self.invconv = InvConv2dLU(in_channel)
        self.coupling = AffineCoupling(in_channel)


class Block(nn.Module):
    def __init__(self, in_channel: int, n_flow: int, split: bool = True):
        super().__init__()
        self.flows = nn.ModuleList([Flow(in_channel * 4) for _ in range(n_flow)])
        self.split = split
        self.prior = ZeroConv2d(in_channel * 2, in_channel * 4) if split else ZeroConv2d(in_channel * 4, in_channel * 8)


class Glow(nn.Module):
    """Glow(in_channel, n_flow, n_block, affine=True, conv_lu=True) (:495-506); only the affine / LU form is built."""

    def __init__(self, in_channel: int, n_flow: int, n_block: int, affine: bool = True, conv_lu: bool = True):
        super().__init__()
        if not affine or not conv_lu:
            raise NotImplementedError("dpi_b200 Glow: only affine=True, conv_lu=True is built")
        self.in_channel, self.n_flow, self.n_block = in_channel, n_flow, n_block
        self.blocks = nn.ModuleList()
        n_channel = in_channel
        for _ in range(n_block - 1):
            self.blocks.append(Block(n_channel, n_flow, split=True))
            n_channel *= 2
        self.blocks.append(Block(n_channel, n_flow, split=False))

    def _no_gradients_yet(self, inputs, what):
        """The backward of the Glow path is not built. A training loop (gradients enabled, trainable parameters) must
        not receive history-free tensors silently - loss.backward() would then reach only `logscale_factor`, Adam would
        never move the flow and the BatchNorm running statistics would still drift: raise. Sampling / density
        evaluation is opt-in through torch.no_grad() (or a model whose parameters do not require grad)."""
        if any(t.requires_grad for t in inputs):
            raise NotImplementedError(f"dpi_b200 Glow.{what}: gradients are not built yet (input requires grad); "
                                      "use the RealNVP flow for training or call under torch.no_grad() for sampling")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                f"dpi_b200 Glow.{what}: the training backward of Glow is not built yet, so this call cannot record "
                "autograd history for its parameters. Wrap inference in torch.no_grad() (or freeze the parameters); "
                "train with generative_model.realnvpfc_model.RealNVP.")

    def forward(self, input):
        """Density direction (:508-524): image [B, in_channel, npix, npix] -> (log_p_sum [B], logdet [B], z_outs)."""
        self._no_gradients_yet([input], "forward")
        with torch.no_grad():
            return self._forward(input)

    def _forward(self, input):
        x_in = input.detach().to(torch.float32)
        if not x_in.is_cuda:
            raise RuntimeError("dpi_b200 has no CPU path: Glow.forward needs CUDA tensors")
        lib = _lib.load()
        dev = x_in.device
        B, _, h, w = x_in.shape
        logdet = torch.zeros(B, dtype=torch.float32, device=dev)
        log_p = torch.zeros(B, dtype=torch.float32, device=dev)
        z_outs = []
        with torch.cuda.device(dev):
            x = glow_ops.to_feature_major(x_in)
            for blk in self.blocks:
                st = _lib.stream_ptr()
                C_in = x.shape[0]
                h, w = h // 2, w // 2
                M = B * h * w
                sq = torch.empty(4 * C_in, M, dtype=torch.float32, device=dev)
                _lib.check(lib.dpi_glow_squeeze(B, h, w, C_in, x.data_ptr(), sq.data_ptr(), st), "dpi_glow_squeeze")
                x = sq
                for fl in blk.flows:
                    x = self._flow_forward(fl, x, logdet, B, h, w)
                C = x.shape[0]
                if blk.split:
                    keep, z_new = x[: C // 2], x[C // 2:]
                    ml = self._zero_conv(blk.prior, keep, B, h, w)
                else:
                    keep, z_new = x, x
                    ml = self._zero_conv(blk.prior, torch.zeros(C, M, dtype=torch.float32, device=dev), B, h, w)
                c2 = z_new.shape[0]
                z_eps = torch.empty(c2, M, dtype=torch.float32, device=dev)
                _lib.check(lib.dpi_glow_prior_forward(B, h * w, c2, ml.data_ptr(), M, z_new.data_ptr(), M, z_eps.data_ptr(), M,
                                                      logdet.data_ptr(), log_p.data_ptr(), st), "dpi_glow_prior_forward")
                z_outs.append(glow_ops.from_feature_major(z_eps, B, h, w))
                x = keep.contiguous() if blk.split else keep
        return log_p, logdet, z_outs

    def _flow_forward(self, fl: "Flow", x_fm, logdet, B, h, w):
        """Flow.forward :338-347: actnorm -> invconv -> coupling."""
        lib, st = _lib.load(), _lib.stream_ptr()
        C, M = x_fm.shape
        ic, an = fl.invconv, fl.actnorm
        if not getattr(an, "_checked", False):
            if int(an.initialized) == 0:          # data-dependent init (:91-111): loc = -mean_c, scale_inv = std_c + 1e-6
                with torch.no_grad():
                    an.loc.copy_(-x_fm.mean(1).view(1, C, 1, 1))
                    an.scale_inv.copy_((x_fm.std(1) + 1e-6).view(1, C, 1, 1))
                an.initialized.fill_(1)
            an._checked = True
        y = torch.empty_like(x_fm)
        _lib.check(lib.dpi_glow_row_affine(C, M, x_fm.data_ptr(), an.loc.detach().data_ptr(), an.scale_inv.detach().data_ptr(),
                                           y.data_ptr(), st), "dpi_glow_row_affine")
        wmat = torch.empty(C, C, 1, 1, dtype=torch.float32, device=x_fm.device)
        _lib.check(lib.dpi_glow_lu_weight(C, ic.w_p.data_ptr(), ic.w_l.detach().data_ptr(), ic.w_s.detach().data_ptr(),
                                          ic.w_u.detach().data_ptr(), ic.l_mask.data_ptr(), ic.u_mask.data_ptr(),
                                          ic.s_sign.data_ptr(), ic.l_eye.data_ptr(), wmat.data_ptr(), st), "dpi_glow_lu_weight")
        x2 = glow_ops.conv2d_fm(y, wmat, None, B, h, w)
        # ActNorm.forward: -h w sum log|scale_inv| (:123-125); InvConv2dLU.forward: +h w sum w_s (:213)
        _lib.check(lib.dpi_glow_logdet_scalar(B, h * w, C, ic.w_s.detach().data_ptr(), 1.0, an.scale_inv.detach().data_ptr(), -1.0,
                                              logdet.data_ptr(), st), "dpi_glow_logdet_scalar")
        net = fl.coupling.net
        a = x2[: C // 2]
        h1 = self._conv_lrelu_bn(net[0], net[2], a, B, h, w)
        h2 = self._conv_lrelu_bn(net[3], net[5], h1, B, h, w)
        out = self._zero_conv(net[6], h2, B, h, w)
        _lib.check(lib.dpi_glow_affine_forward(B, h * w, C, out.data_ptr(), M, x2.data_ptr(), M, logdet.data_ptr(), st),
                   "dpi_glow_affine_forward")
        return x2

    # ------------------------------------------------------------------------------------------ kernels
    def _zero_conv(self, zc: ZeroConv2d, x_fm, B, h, w):
        return glow_ops.conv2d_fm(x_fm, zc.conv.weight.detach(), zc.conv.bias.detach(), B, h, w, pad_value=1.0,
                                  scale=zc.scale.detach())

    def _conv_lrelu_bn(self, conv: nn.Conv2d, bn: nn.BatchNorm2d, x_fm, B, h, w, rec=None):
        """rec (a list): training pass that records what the backward needs - the LeakyReLU output is kept and the
        BatchNorm writes into a second buffer instead of in place."""
        lib = _lib.load()
        wt = conv.weight.detach().contiguous()
        c_out, c_in, k, _ = wt.shape
        M = B * h * w
        y = torch.empty(c_out, M, dtype=torch.float32, device=x_fm.device)
        ws = torch.empty(lib.dpi_conv2d_fm_workspace_bytes(B, h, w, c_in, c_out, k), dtype=torch.uint8, device=x_fm.device)
        st = _lib.stream_ptr()
        _lib.check(lib.dpi_conv2d_fm_lrelu(B, h, w, c_in, c_out, k, 0.0, x_fm.data_ptr(), x_fm.shape[1], wt.data_ptr(),
                                           conv.bias.detach().data_ptr(), LRELU_SLOPE, y.data_ptr(), M, ws.data_ptr(),
                                           ws.numel(), st), "dpi_conv2d_fm_lrelu")
        n = torch.empty_like(y) if rec is not None else y
        _lib.check(lib.dpi_glow_bn_rows(c_out, M, M, y.data_ptr(), bn.weight.detach().data_ptr(), bn.bias.detach().data_ptr(),
                                        bn.running_mean.data_ptr(), bn.running_var.data_ptr(), int(self.training), BN_EPS,
                                        n.data_ptr(), st), "dpi_glow_bn_rows")
        if self.training:
            bn.num_batches_tracked += 1
        if rec is not None:
            rec.append(y)
        return n

    def _flow_reverse(self, fl: Flow, x_fm, logdet, B, h, w, tape=None):
        lib, st = _lib.load(), _lib.stream_ptr()
        C, M = x_fm.shape
        net = fl.coupling.net
        if tape is not None:
            x_fm = x_fm.clone()                                      # the producer's backward needs its output intact
        saved = [] if tape is not None else None
        a = x_fm[: C // 2]                                           # rows of the a-half: a contiguous block
        h1 = self._conv_lrelu_bn(net[0], net[2], a, B, h, w, saved)
        h2 = self._conv_lrelu_bn(net[3], net[5], h1, B, h, w, saved)
        out = self._zero_conv(net[6], h2, B, h, w)                   # rows [0, C/2): log_s0, rows [C/2, C): t
        _lib.check(lib.dpi_glow_affine_reverse(B, h * w, C, out.data_ptr(), M, x_fm.data_ptr(), M, logdet.data_ptr(), st),
                   "dpi_glow_affine_reverse")
        ic, an = fl.invconv, fl.actnorm
        winv = torch.empty(C, C, dtype=torch.float32, device=x_fm.device)
        _lib.check(lib.dpi_glow_lu_inverse(C, ic.w_p.data_ptr(), ic.w_l.detach().data_ptr(), ic.w_s.detach().data_ptr(),
                                           ic.w_u.detach().data_ptr(), ic.l_mask.data_ptr(), ic.u_mask.data_ptr(),
                                           ic.s_sign.data_ptr(), ic.l_eye.data_ptr(), winv.data_ptr(), st), "dpi_glow_lu_inverse")
        if not getattr(an, "_checked", False):                       # one host read per layer, not one per call
            if int(an.initialized) == 0:                             # first call in reverse: zeros / ones (:135-137)
                with torch.no_grad():
                    an.loc.zero_()
                    an.scale_inv.fill_(1.0)
                an.initialized.fill_(1)
            an._checked = True
        y = torch.empty_like(x_fm)
        ws = torch.empty(lib.dpi_conv2d_fm_workspace_bytes(B, h, w, C, C, 1), dtype=torch.uint8, device=x_fm.device)
        _lib.check(lib.dpi_conv2d_fm_affine(B, h, w, C, C, x_fm.data_ptr(), M, winv.data_ptr(), an.scale_inv.detach().data_ptr(),
                                            an.loc.detach().data_ptr(), y.data_ptr(), M, ws.data_ptr(), ws.numel(), st),
                   "dpi_conv2d_fm_affine")
        # InvConv2dLU.reverse: -h w sum w_s (:231); ActNorm.reverse: +h w sum log|scale_inv| (:141-143)
        _lib.check(lib.dpi_glow_logdet_scalar(B, h * w, C, ic.w_s.detach().data_ptr(), -1.0, an.scale_inv.detach().data_ptr(), 1.0,
                                              logdet.data_ptr(), st), "dpi_glow_logdet_scalar")
        if tape is not None:
            tape.append(dict(fl=fl, l1=saved[0], n1=h1, l2=saved[1], n2=h2, out=out, x_mid=x_fm, winv=winv, y=y))
        return y

    def reverse(self, z_list):
        """z_list[i]: [B, c_i, h_i, w_i] as calc_z_shapes gives them -> (image [B, in_channel, npix, npix], logdet [B]).
        With gradients enabled (the training loop, DPI_interferometry.py:129-136 / :201-204 with --model_form glow) the call
        is one autograd node whose backward runs the op-by-op backward kernels (training-mode BatchNorm only)."""
        z_list = list(z_list)
        params = [p for p in self.parameters() if p.requires_grad]
        if torch.is_grad_enabled() and (params or any(z.requires_grad for z in z_list)):
            if not self.training:
                raise NotImplementedError("dpi_b200 Glow.reverse: gradients through an eval-mode pass (running BatchNorm "
                                          "statistics) are not built; call .train() or wrap inference in torch.no_grad()")
            out = _GlowReverse.apply(self, len(z_list), *z_list, *params)
            return out[0], out[1]
        with torch.no_grad():
            return self._reverse(z_list)

    def _reverse(self, z_list, tape=None):
        z_list = [z.detach().to(torch.float32) for z in z_list]
        if not all(z.is_cuda for z in z_list):
            raise RuntimeError("dpi_b200 has no CPU path: Glow.reverse needs CUDA tensors")
        lib = _lib.load()
        dev = z_list[0].device
        B = z_list[-1].shape[0]
        logdet = torch.zeros(B, dtype=torch.float32, device=dev)
        x = None
        with torch.cuda.device(dev):
            for i, blk in enumerate(reversed(self.blocks)):
                eps = z_list[-(i + 1)]
                _, c_eps, h, w = eps.shape
                M = B * h * w
                st = _lib.stream_ptr()
                eps_fm = glow_ops.to_feature_major(eps)
                if blk.split:
                    prior_in = x
                    ml = self._zero_conv(blk.prior, x, B, h, w)                     # [mean ; log_sd] of the dropped half
                    full = torch.empty(x.shape[0] + c_eps, M, dtype=torch.float32, device=dev)
                    full[: x.shape[0]].copy_(x)
                    z_rows = full[x.shape[0]:]
                else:
                    prior_in = torch.zeros(c_eps, M, dtype=torch.float32, device=dev)
                    ml = self._zero_conv(blk.prior, prior_in, B, h, w)
                    full = torch.empty(c_eps, M, dtype=torch.float32, device=dev)
                    z_rows = full
                flows_tape = [] if tape is not None else None
                _lib.check(lib.dpi_glow_prior_sample(B, h * w, c_eps, ml.data_ptr(), M, eps_fm.data_ptr(), M, z_rows.data_ptr(), M,
                                                     logdet.data_ptr(), st), "dpi_glow_prior_sample")
                x = full
                for fl in reversed(blk.flows):
                    x = self._flow_reverse(fl, x, logdet, B, h, w, flows_tape)
                if tape is not None:
                    tape.append(dict(blk=blk, B=B, h=h, w=w, prior_in=prior_in, ml=ml, eps=eps_fm, flows=flows_tape))
                C = x.shape[0]
                up = torch.empty(C // 4, 4 * M, dtype=torch.float32, device=dev)
                _lib.check(lib.dpi_glow_unsqueeze(B, h, w, C, x.data_ptr(), up.data_ptr(), st), "dpi_glow_unsqueeze")
                x = up
                h, w = 2 * h, 2 * w
        return glow_ops.from_feature_major(x, B, h, w), logdet


# ------------------------------------------------------------------------------------------------
# training backward of the sampling direction (formulas: oracle/glow_backward_spec.py)
# ------------------------------------------------------------------------------------------------
def _acc(grads, p, g):
    g = g.reshape(p.shape)
    grads[p] = g if p not in grads else grads[p] + g


def _zero_conv_backward(zc: "ZeroConv2d", x_in, out, g_out, B, h, w, grads, need_input_grad=True):
    """ZeroConv2d (:246-251). g_out is the gradient of the scaled output and is consumed (scaled in place)."""
    lib, st = _lib.load(), _lib.stream_ptr()
    C, M = out.shape
    g_scale = torch.empty(C, dtype=torch.float32, device=out.device)
    g_bias = torch.empty(C, dtype=torch.float32, device=out.device)
    _lib.check(lib.dpi_glow_zeroconv_bwd(C, M, g_out.data_ptr(), out.data_ptr(), zc.scale.detach().data_ptr(), g_scale.data_ptr(),
                                         g_bias.data_ptr(), st), "dpi_glow_zeroconv_bwd")
    _acc(grads, zc.scale, g_scale)
    _acc(grads, zc.conv.bias, g_bias)
    _acc(grads, zc.conv.weight, glow_ops.conv2d_fm_wgrad(x_in, g_out, 3, B, h, w, pad_value=1.0))
    return glow_ops.conv2d_fm_dgrad(g_out, zc.conv.weight.detach(), B, h, w) if need_input_grad else None


def _conv_bn_backward(conv: nn.Conv2d, bn: nn.BatchNorm2d, x_in, l, g, B, h, w, grads):
    """Conv -> LeakyReLU -> BatchNorm2d (:260-268), training mode. g: gradient of the BatchNorm output (consumed)."""
    lib, st = _lib.load(), _lib.stream_ptr()
    C, M = l.shape
    gg, gb, gbias = (torch.empty(C, dtype=torch.float32, device=l.device) for _ in range(3))
    _lib.check(lib.dpi_glow_bn_lrelu_bwd(C, M, g.data_ptr(), l.data_ptr(), bn.weight.detach().data_ptr(), BN_EPS, LRELU_SLOPE,
                                         gg.data_ptr(), gb.data_ptr(), gbias.data_ptr(), st), "dpi_glow_bn_lrelu_bwd")
    _acc(grads, bn.weight, gg)
    _acc(grads, bn.bias, gb)
    _acc(grads, conv.bias, gbias)
    k = conv.weight.shape[-1]
    _acc(grads, conv.weight, glow_ops.conv2d_fm_wgrad(x_in, g, k, B, h, w, pad_value=0.0))
    return glow_ops.conv2d_fm_dgrad(g, conv.weight.detach(), B, h, w)


def _flow_reverse_backward(rec, g, g_ld, hw_gld, B, h, w, grads):
    """Flow.reverse = coupling^-1 -> invconv^-1 -> actnorm^-1 (:349-358); g: gradient of the flow's output [C][M]."""
    lib, st = _lib.load(), _lib.stream_ptr()
    fl = rec["fl"]
    C, M = g.shape
    dev = g.device
    an, ic, net = fl.actnorm, fl.invconv, fl.coupling.net
    f32 = torch.float32
    # ActNorm.reverse
    g_v = torch.empty_like(g)
    g_loc, g_si = torch.empty(C, dtype=f32, device=dev), torch.empty(C, dtype=f32, device=dev)
    _lib.check(lib.dpi_glow_actnorm_reverse_bwd(C, M, rec["y"].data_ptr(), an.loc.detach().data_ptr(), an.scale_inv.detach().data_ptr(),
                                                g.data_ptr(), hw_gld.data_ptr(), g_v.data_ptr(), g_loc.data_ptr(), g_si.data_ptr(),
                                                st), "dpi_glow_actnorm_reverse_bwd")
    _acc(grads, an.loc, g_loc)
    _acc(grads, an.scale_inv, g_si)
    # InvConv2dLU.reverse: x = V u
    winv4 = rec["winv"].view(C, C, 1, 1)
    g_u = glow_ops.conv2d_fm_dgrad(g_v, winv4, B, h, w)
    G_V = glow_ops.conv2d_fm_wgrad(rec["x_mid"], g_v, 1, B, h, w).view(C, C).contiguous()
    g_wl, g_wu = torch.empty(C, C, dtype=f32, device=dev), torch.empty(C, C, dtype=f32, device=dev)
    g_ws = torch.empty(C, dtype=f32, device=dev)
    scratch = torch.empty(3 * C * C, dtype=f32, device=dev)
    _lib.check(lib.dpi_glow_lu_bwd(C, ic.w_p.data_ptr(), ic.w_l.detach().data_ptr(), ic.w_s.detach().data_ptr(),
                                   ic.w_u.detach().data_ptr(), ic.l_mask.data_ptr(), ic.u_mask.data_ptr(), ic.s_sign.data_ptr(),
                                   ic.l_eye.data_ptr(), rec["winv"].data_ptr(), G_V.data_ptr(), hw_gld.data_ptr(),
                                   scratch.data_ptr(), g_wl.data_ptr(), g_wu.data_ptr(), g_ws.data_ptr(), st), "dpi_glow_lu_bwd")
    _acc(grads, ic.w_l, g_wl)
    _acc(grads, ic.w_u, g_wu)
    _acc(grads, ic.w_s, g_ws)
    # AffineCoupling.reverse
    g_out = torch.empty(C, M, dtype=f32, device=dev)
    _lib.check(lib.dpi_glow_affine_reverse_bwd(B, h * w, C, rec["out"].data_ptr(), M, rec["x_mid"].data_ptr(), M, g_u.data_ptr(), M,
                                               g_ld.data_ptr(), g_out.data_ptr(), M, st), "dpi_glow_affine_reverse_bwd")
    g_n2 = _zero_conv_backward(net[6], rec["n2"], rec["out"], g_out, B, h, w, grads)
    g_n1 = _conv_bn_backward(net[3], net[5], rec["n1"], rec["l2"], g_n2, B, h, w, grads)
    g_a = _conv_bn_backward(net[0], net[2], rec["x_mid"][: C // 2], rec["l1"], g_n1, B, h, w, grads)
    g_u[: C // 2].add_(g_a)
    return g_u


def _reverse_backward(tape, g_img, g_ld):
    """-> (gradients of z_list in its order, {parameter: gradient})."""
    lib = _lib.load()
    grads, g_z = {}, []
    dev = g_img.device
    f32 = torch.float32
    with torch.cuda.device(dev):
        st = _lib.stream_ptr()
        g = glow_ops.to_feature_major(g_img.contiguous().float())
        g_ld = g_ld.contiguous().float()
        hw_gld = torch.empty(1, dtype=f32, device=dev)
        for rec in reversed(tape):
            B, h, w = rec["B"], rec["h"], rec["w"]
            M = B * h * w
            C = g.shape[0] * 4
            gx = torch.empty(C, M, dtype=f32, device=dev)               # unsqueeze^T = squeeze
            _lib.check(lib.dpi_glow_squeeze(B, h, w, C // 4, g.data_ptr(), gx.data_ptr(), st), "dpi_glow_squeeze")
            g = gx
            _lib.check(lib.dpi_glow_scaled_sum(B, g_ld.data_ptr(), float(h * w), hw_gld.data_ptr(), st), "dpi_glow_scaled_sum")
            for frec in reversed(rec["flows"]):
                g = _flow_reverse_backward(frec, g, g_ld, hw_gld, B, h, w, grads)
            blk = rec["blk"]
            c_eps = rec["eps"].shape[0]
            c_x = C - c_eps
            gz_rows = g[c_x:]
            g_ml = torch.empty(2 * c_eps, M, dtype=f32, device=dev)
            g_eps = torch.empty(c_eps, M, dtype=f32, device=dev)
            _lib.check(lib.dpi_glow_prior_sample_bwd(B, h * w, c_eps, rec["ml"].data_ptr(), M, rec["eps"].data_ptr(), M,
                                                     gz_rows.data_ptr(), M, g_ld.data_ptr(), g_ml.data_ptr(), M,
                                                     g_eps.data_ptr(), M, st), "dpi_glow_prior_sample_bwd")
            g_prior_in = _zero_conv_backward(blk.prior, rec["prior_in"], rec["ml"], g_ml, B, h, w, grads, need_input_grad=blk.split)
            g_z.append(glow_ops.from_feature_major(g_eps, B, h, w))
            if blk.split:
                g = g[:c_x] + g_prior_in
    return g_z, grads


class _GlowReverse(torch.autograd.Function):
    """Glow.reverse as ONE autograd node: inputs (z_0 .. z_{n-1}, trainable parameters), outputs (image, logdet)."""

    @staticmethod
    def forward(ctx, module, n_z, *args):
        z_list, params = list(args[:n_z]), list(args[n_z:])
        tape = []
        with torch.no_grad():
            x, ld = module._reverse(z_list, tape)
        ctx.tape, ctx.params, ctx.n_z = tape, params, n_z
        return x, ld

    @staticmethod
    def backward(ctx, g_x, g_ld):
        tape = ctx.tape
        if tape is None:
            raise RuntimeError("dpi_b200 Glow.reverse: backward called twice (the saved activations are released after one)")
        ctx.tape = None
        if g_ld is None:
            g_ld = torch.zeros(tape[0]["B"], dtype=torch.float32, device=g_x.device)
        g_z, grads = _reverse_backward(tape, g_x, g_ld)
        out = [None, None]
        out += [g_z[i] if ctx.needs_input_grad[2 + i] else None for i in range(ctx.n_z)]
        out += [grads.get(p) for p in ctx.params]
        return tuple(out)
