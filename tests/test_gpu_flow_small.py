"""Small-batch cluster kernel (dpi_b200/csrc/flow_small.cu) against the CPU oracle and against the tiled
phase program on the same inputs: every decomposition it can choose (split-K over a cluster through
DSMEM, single-CTA full-K, chunked K, no-cluster fallback), persistent and one-launch-per-phase.
Tolerance: 1e-4 relative (north_star, fp32)."""
import os

import pytest
import torch

from conftest import RTOL, rel_err
from oracle import dpi_oracle as O

pytestmark = pytest.mark.gpu


def _run(D, NF, B, seqfrac, env, persistent=True, bn=True, seed=11):
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    env = {"DPI_B200_SR": "0", **env}           # this file tests flow_small.cu; the sample-resident kernels have their own
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        sd = O.init_state(D, NF, seqfrac=seqfrac, seed=seed, randomize=True, batch_norm=bn)
        m = RealNVP(D, NF, seqfrac=seqfrac, batch_norm=bn)
        m.load_state_dict(sd)
        m = m.cuda()
        m.set_flow_mode(persistent)
        torch.manual_seed(seed)
        z = torch.randn(B, D)
        cx, cl = torch.randn(B, D), torch.randn(B)
        zg = z.cuda().requires_grad_(True)
        x, ld = m.reverse(zg)                       # the handle (and its plan) is created here, under `env`
        ((x * cx.cuda()).sum() + (ld * cl.cuda()).sum()).backward()
        grads = {k: v.detach().cpu().clone() for k, v in m.named_grad_views()}
        with torch.no_grad():
            m.eval()
            xe, lde = m.reverse(z.cuda())
            m.train()
            zb, ldb = m.forward(x.detach())
        out = dict(x=x.detach().cpu(), ld=ld.detach().cpu(), gz=zg.grad.cpu(), grads=grads, zb=zb.cpu(), ldb=ldb.cpu(),
                   xe=xe.cpu(), lde=lde.cpu(), z=z, cx=cx, cl=cl, sd=sd, bn_after={k: v.cpu() for k, v in m.state_dict().items()})
        del m
        return out
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _oracle(r, bn=True):
    sd = {k: v.clone() for k, v in r["sd"].items()}
    for k in O.param_keys(sd):
        sd[k].requires_grad_(True)
    zo = r["z"].clone().requires_grad_(True)
    xo, ldo = O.realnvp_reverse(sd, zo, train=True)
    ((xo * r["cx"]).sum() + (ldo * r["cl"]).sum()).backward()
    with torch.no_grad():
        sde = {k: v.detach() for k, v in sd.items()}
        xe, lde = O.realnvp_reverse(sde, r["z"], train=False)
    return sd, zo, xo.detach(), ldo.detach(), xe, lde


def _check(r, bn=True, tol=RTOL):
    sd, zo, xo, ldo, xe, lde = _oracle(r, bn)
    assert rel_err(r["x"], xo) < tol, rel_err(r["x"], xo)
    assert rel_err(r["ld"], ldo) < tol, rel_err(r["ld"], ldo)
    assert rel_err(r["gz"], zo.grad) < tol, rel_err(r["gz"], zo.grad)
    worst = max(((rel_err(gv, sd[k].grad), k) for k, gv in r["grads"].items() if "actnorm" not in k))
    assert worst[0] < 3 * tol, worst
    for kind in ("loc", "log_scale_inv"):
        ks = [k for k in r["grads"] if k.endswith("." + kind)]
        mine = torch.cat([r["grads"][k].reshape(-1) for k in ks])
        ref = torch.cat([sd[k].grad.reshape(-1) for k in ks])
        assert rel_err(mine, ref) < 3 * tol, (kind, rel_err(mine, ref))
    # eval-mode reverse uses the running statistics the train-mode pass just updated
    assert rel_err(r["xe"], xe) < tol and rel_err(r["lde"], lde) < tol
    # invertibility / logdet antisymmetry
    assert float((r["zb"] - r["z"]).abs().max()) < 5e-4
    assert float((r["ldb"] + r["ld"]).abs().max()) < 1e-3 * max(1.0, float(r["ld"].abs().max()))


SHAPES = [
    # D, NF, B, seqfrac
    (64, 2, 8, 1),        # H = 32
    (64, 2, 6, 1),        # batch not a multiple of 4 -> 4-byte copies
    (50, 2, 7, 1),        # H = 25: nothing aligned
    (9, 2, 5, 0.25),      # odd D
    (256, 3, 64, 4),      # full 64-sample rows
    (1024, 2, 33, 4),     # C1 layer shapes: K = 512 / 1024 split over the cluster by default
]


@pytest.mark.parametrize("persistent", [True, False])
@pytest.mark.parametrize("D,NF,B,seqfrac", SHAPES)
def test_small_kernel_vs_oracle(D, NF, B, seqfrac, persistent):
    _check(_run(D, NF, B, seqfrac, {"DPI_B200_SMALL": "1"}, persistent))


@pytest.mark.parametrize("D,NF,B,seqfrac", SHAPES[:5])
def test_small_kernel_forced_cluster_split(D, NF, B, seqfrac):
    """Every GEMM with K >= 8 goes through the DSMEM split-K reduction."""
    _check(_run(D, NF, B, seqfrac, {"DPI_B200_SMALL": "1", "DPI_B200_SMALL_KSPLIT_MIN": "8"}))


@pytest.mark.parametrize("D,NF,B,seqfrac", [SHAPES[0], SHAPES[5]])
def test_small_kernel_without_clusters(D, NF, B, seqfrac):
    _check(_run(D, NF, B, seqfrac, {"DPI_B200_SMALL": "1", "DPI_B200_SMALL_NOCLUSTER": "1"}))


@pytest.mark.parametrize("cl", ["8", "4", "2"])
@pytest.mark.parametrize("D,NF,B,seqfrac", [SHAPES[0], SHAPES[5]])
def test_small_kernel_every_cluster_size(D, NF, B, seqfrac, cl):
    """Cluster sizes 8 / 4 / 2 (the planner normally picks the cheapest per shape), K split from 8 up."""
    _check(_run(D, NF, B, seqfrac, {"DPI_B200_SMALL": "1", "DPI_B200_SMALL_CL": cl, "DPI_B200_SMALL_KSPLIT_MIN": "8"}))


def test_small_kernel_no_batchnorm():
    _check(_run(64, 2, 8, 1, {"DPI_B200_SMALL": "1"}, bn=False), bn=False)


def test_small_kernel_mri_layer_shapes():
    """C2 layer shapes (D = 4096, H = 512): chunked K, 32-feature tiles, 32 x 128 weight-gradient tiles."""
    _check(_run(4096, 1, 16, 4, {"DPI_B200_SMALL": "1"}))


def test_small_kernel_matches_tiled_program_and_is_deterministic():
    a = _run(256, 3, 64, 4, {"DPI_B200_SMALL": "1"})
    b = _run(256, 3, 64, 4, {"DPI_B200_SMALL": "1"})
    c = _run(256, 3, 64, 4, {"DPI_B200_SMALL": "0"})
    assert torch.equal(a["x"], b["x"]) and torch.equal(a["ld"], b["ld"]) and torch.equal(a["gz"], b["gz"])
    for k in a["grads"]:
        assert torch.equal(a["grads"][k], b["grads"][k]), k      # fixed-order reductions: bit-reproducible
    assert rel_err(a["x"], c["x"]) < 1e-5 and rel_err(a["ld"], c["ld"]) < 1e-5
    for k in a["bn_after"]:
        assert rel_err(a["bn_after"][k].float(), c["bn_after"][k].float()) < 1e-5, k
