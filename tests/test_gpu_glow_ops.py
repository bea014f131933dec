"""GPU: the first building block of the Glow path (not built yet as a whole): the coupling-net
convolutions as tcgen05 GEMMs with the 3x3 window gather in the operand pack, against torch's CPU convolution and
against the ZeroConv2d of the Glow oracle. 3xTF32: the fp32 bar (1e-4 relative)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import RTOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,cin,cout,hw,k,pad", [(3, 6, 20, 8, 3, 1.0), (4, 2, 512, 4, 3, 0.0), (4, 512, 512, 4, 1, 0.0),
                                                   (5, 512, 8, 6, 3, 1.0), (2, 3, 7, 5, 3, 0.0)])
def test_conv2d_feature_major_matches_torch(B, cin, cout, hw, k, pad):
    from dpi_b200.generative_model import glow_ops as G
    torch.manual_seed(B * 100 + cin)
    x = torch.randn(B, cin, hw, hw)
    wt = torch.randn(cout, cin, k, k) * 0.1
    bias = torch.randn(cout) * 0.3
    scale = torch.randn(1, cout, 1, 1) * 0.2 if pad == 1.0 else None
    if k == 3:
        ref = F.conv2d(F.pad(x, [1, 1, 1, 1], value=pad), wt, bias)
    else:
        ref = F.conv2d(x, wt, bias)
    if scale is not None:
        ref = ref * torch.exp(scale * 3)
    y = G.conv2d_fm(G.to_feature_major(x.cuda()), wt.cuda(), bias.cuda(), B, hw, hw, pad_value=pad,
                    scale=None if scale is None else scale.cuda())
    out = G.from_feature_major(y, B, hw, hw).cpu()
    assert out.shape == ref.shape
    assert rel_err(out, ref) < RTOL, rel_err(out, ref)


def test_zero_conv_matches_glow_oracle():
    """ZeroConv2d exactly as the Glow oracle (and therefore the reference, glow_model.py:246-251) computes it."""
    from dpi_b200.generative_model import glow_ops as G
    from oracle import glow_oracle as GO
    torch.manual_seed(4)
    B, cin, cout, hw = 6, 512, 4, 8
    sd = {"conv.weight": torch.randn(cout, cin, 3, 3) * 0.05, "conv.bias": torch.randn(cout) * 0.1,
          "scale": torch.randn(1, cout, 1, 1) * 0.3}
    x = torch.randn(B, cin, hw, hw)
    ref = GO.zero_conv(sd, "", x)
    y = G.conv2d_fm(G.to_feature_major(x.cuda()), sd["conv.weight"].cuda(), sd["conv.bias"].cuda(), B, hw, hw, pad_value=1.0,
                    scale=sd["scale"].cuda())
    assert rel_err(G.from_feature_major(y, B, hw, hw).cpu(), ref) < RTOL


def _glow_from_fixture():
    from conftest import load_golden, split_sd
    from dpi_b200.generative_model.glow_model import Glow
    from oracle import glow_oracle as GO
    g = load_golden("glow_small")
    sd = split_sd(g)
    flows = sorted({k[: k.index("coupling.")] for k in sd if "coupling.net.0.weight" in k})
    for tag, k in enumerate(GO.big_weight_keys([f + "coupling.net.3.weight" for f in flows])):
        sd[k] = GO.formula_weight((GO.FILTER, GO.FILTER, 1, 1), tag)
    m = Glow(1, int(g["n_flow"]), int(g["n_block"]))
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return g, sd, m.cuda()


def test_glow_reverse_matches_reference_fixture():
    """SURVEY 8 a20, sampling direction: a state dict WRITTEN BY THE REFERENCE loads into the drop-in Glow and
    Glow.reverse reproduces the reference's image and log-determinant (train mode incl. the BatchNorm running
    statistics, then eval mode)."""
    g, sd, m = _glow_from_fixture()
    zs = [torch.from_numpy(g[f"z{i}"]).cuda() for i in range(int(g["n_block"]))]
    m.train()
    with pytest.raises(NotImplementedError):        # the density direction has no backward: never history-free silently
        m.forward(torch.from_numpy(g["x_rev"]).cuda())
    torch.set_grad_enabled(False)                   # inference is opt-in (the whole test runs under no_grad)
    x, ld = m.reverse(zs)
    assert rel_err(x, g["x_rev"]) < RTOL, rel_err(x, g["x_rev"])
    assert rel_err(ld, g["ld_rev"]) < RTOL, rel_err(ld, g["ld_rev"])
    mine = m.state_dict()
    for k, v in g.items():
        if k.startswith("after."):
            assert rel_err(mine[k[len("after."):]].float(), v) < RTOL, k
    # the fixture script then ran the reference's forward() in train mode (which moves the running statistics again)
    lp, ldf, zf = m.forward(torch.from_numpy(g["x_rev"]).cuda())
    assert rel_err(lp, g["logp_fwd"]) < RTOL, rel_err(lp, g["logp_fwd"])
    assert rel_err(ldf, g["ld_fwd"]) < RTOL, rel_err(ldf, g["ld_fwd"])
    for i, z in enumerate(zf):
        assert rel_err(z, g[f"zf{i}"]) < RTOL, (i, rel_err(z, g[f"zf{i}"]))
    m.eval()
    xe, lde = m.reverse(zs)
    torch.set_grad_enabled(True)
    assert rel_err(xe, g["x_rev_eval"]) < RTOL and rel_err(lde, g["ld_rev_eval"]) < RTOL


@pytest.mark.parametrize("npix,n_flow,n_block,B", [(16, 2, 3, 4), (8, 1, 2, 20)])
def test_glow_reverse_matches_oracle(npix, n_flow, n_block, B):
    """A fresh drop-in Glow with its zero-initialised pieces randomised, against the Glow oracle on its state dict."""
    from dpi_b200.generative_model.glow_model import Glow, calc_z_shapes
    from oracle import glow_oracle as GO
    torch.manual_seed(npix + B)
    import numpy as np
    np.random.seed(npix + B)
    m = Glow(1, n_flow, n_block)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if ".net.6." in name or ".prior." in name:
                p.normal_(0, 0.05)
            elif name.endswith("actnorm.loc"):
                p.normal_(0, 0.2)
            elif name.endswith("actnorm.scale_inv"):
                p.copy_(0.7 + 0.6 * torch.rand_like(p))
        for name, b in m.named_buffers():
            if name.endswith("initialized"):
                b.fill_(1)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    shapes = calc_z_shapes(1, npix, n_flow, n_block)
    assert shapes == GO.calc_z_shapes(1, npix, n_flow, n_block)
    zs = [torch.randn(B, *s) for s in shapes]
    xo, lo = GO.glow_reverse(sd, zs, train=True)
    m = m.cuda().train()
    for p_ in m.parameters():                                # frozen parameters: the other way to opt in to inference
        p_.requires_grad_(False)
    x, ld = m.reverse([z.cuda() for z in zs])
    assert x.shape == (B, 1, npix, npix)
    assert rel_err(x, xo) < RTOL, rel_err(x, xo)
    assert rel_err(ld, lo) < RTOL, rel_err(ld, lo)
    mine = m.state_dict()
    for k, v in sd.items():
        if "running" in k:
            assert rel_err(mine[k].float(), v.float()) < RTOL, k
    xoe, loe = GO.glow_reverse(sd, zs, train=False)
    m.eval()
    xe, lde = m.reverse([z.cuda() for z in zs])
    assert rel_err(xe, xoe) < RTOL and rel_err(lde, loe) < RTOL
    # density direction in eval mode: forward(reverse(z)) returns z and the log-determinants cancel
    lpo, ldo, zo = GO.glow_forward(sd, xoe, train=False)
    lp, ldf, zf = m.forward(xoe.cuda())                      # the same input on both sides
    assert rel_err(lp, lpo) < RTOL and rel_err(ldf, ldo) < RTOL
    if n_flow * n_block <= 2:
        # latents and invertibility only for the shallow net: through the 6 randomised flows of the other case the
        # density direction is expansive (measured 4e-3 on z there while log_p and logdet hold 1e-4)
        for a, b, z in zip(zf, zo, zs):
            assert rel_err(a, b) < 10 * RTOL, rel_err(a, b)
            assert float((a.cpu() - z).abs().max()) < 2e-2
        assert float((ldf + lde).abs().max()) < 1e-2 * max(1.0, float(lde.abs().max()))


@pytest.mark.parametrize("B,cin,cout,hw,k,pad", [(3, 6, 20, 8, 3, 1.0), (4, 2, 512, 4, 3, 0.0), (4, 512, 512, 4, 1, 0.0),
                                                   (5, 512, 8, 6, 3, 1.0)])
def test_conv2d_gradients_match_torch_autograd(B, cin, cout, hw, k, pad):
    """Building blocks of the Glow backward (not assembled yet): data and weight gradient of the coupling-net
    convolutions on tcgen05, against torch autograd on the CPU."""
    from dpi_b200.generative_model import glow_ops as G
    torch.manual_seed(B + cout)
    x = torch.randn(B, cin, hw, hw, requires_grad=True)
    wt = (torch.randn(cout, cin, k, k) * 0.1).requires_grad_(True)
    c = torch.randn(B, cout, hw, hw)
    y = F.conv2d(F.pad(x, [1, 1, 1, 1], value=pad), wt) if k == 3 else F.conv2d(x, wt)
    (y * c).sum().backward()
    g_fm = G.to_feature_major(c.cuda())
    gx = G.from_feature_major(G.conv2d_fm_dgrad(g_fm, wt.detach().cuda(), B, hw, hw), B, hw, hw).cpu()
    gw = G.conv2d_fm_wgrad(G.to_feature_major(x.detach().cuda()), g_fm, k, B, hw, hw, pad_value=pad).cpu()
    assert rel_err(gx, x.grad) < RTOL, rel_err(gx, x.grad)
    assert rel_err(gw, wt.grad) < RTOL, rel_err(gw, wt.grad)


def test_glow_training_backward_matches_reference_fixture():
    """SURVEY 8 a20, the training direction: Glow.reverse with gradients enabled is one autograd node; the gradients of the
    reference's objective sum(x * c) + sum(logdet) with respect to the latents and to EVERY parameter (ActNorm, the P L U
    factors of the invertible 1x1 convolutions, coupling-net convolutions, BatchNorm affine, ZeroConv2d weight / bias /
    scale, split priors) match the ones the unmodified reference produced (tests/golden/make_golden.py glow_case). The
    convolutions run 3xTF32 on the tensor cores: gradients are held to the suite's gradient bar, 3e-4 normwise."""
    from oracle import glow_oracle as GO
    g, sd, m = _glow_from_fixture()
    m.train()
    nb = int(g["n_block"])
    zs = [torch.from_numpy(g[f"z{i}"]).cuda().requires_grad_(True) for i in range(nb)]
    x, ld = m.reverse(zs)
    assert x.requires_grad and ld.requires_grad
    assert rel_err(x, g["x_rev"]) < RTOL and rel_err(ld, g["ld_rev"]) < RTOL
    c = torch.from_numpy(g["c"]).cuda()
    ((x * c).sum() + ld.sum()).backward()
    for i in range(nb):
        assert rel_err(zs[i].grad, g[f"g_z{i}"]) < 3 * RTOL, (i, rel_err(zs[i].grad, g[f"g_z{i}"]))
    big = set(GO.big_weight_keys(m.state_dict().keys()))
    worst = []
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        mine = p.grad.detach().reshape(-1)[::1021] if k in big else p.grad.detach()
        worst.append((rel_err(mine.cpu().reshape(-1), torch.from_numpy(g["grad." + k]).reshape(-1)), k))
    bad = [w for w in worst if not w[0] < 3 * RTOL]
    assert not bad, sorted(bad, reverse=True)[:8]
    assert max(w[0] for w in worst) < RTOL          # measured 1.1e-5: the fp32 bar holds for every Glow gradient on this fixture


def test_glow_training_step_moves_every_parameter():
    """A torch.optim.Adam step on the drop-in Glow (what DPI_interferometry.py --model_form glow does) changes every
    parameter tensor and keeps everything finite."""
    g, sd, m = _glow_from_fixture()
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    torch.manual_seed(0)
    zs = [torch.randn(4, *s, device="cuda") for s in [tuple(g[f"z{i}"].shape[1:]) for i in range(int(g["n_block"]))]]
    x, ld = m.reverse(zs)
    loss = (x ** 2).mean() - ld.mean()
    loss.backward()
    opt.step()
    for k, p in m.named_parameters():
        assert torch.isfinite(p).all(), k
        assert not torch.equal(p.detach(), before[k]), k


def test_glow_training_trajectory_matches_reference_run():
    """three optimisation steps of the DPI_interferometry.py loop body with --model_form glow (:129-136,
    :193-235: Glow.reverse -> Softplus * exp(log_scale) -> closure quantities -> chi^2 + priors - logdet -> backward ->
    clip_grad_norm_ -> Adam), once over the UNMODIFIED reference modules on the CPU (baseline/_ref, installed by
    __graft_entry__.build()) and once over this package's drop-in modules on the GPU, same initial state, same latents."""
    from baseline.ref_step import ReferenceStep, reference_root
    if reference_root() is None:
        pytest.skip("reference modules not installed (baseline/_ref)")
    import dpi_b200.MRI_helpers as MH
    import dpi_b200.interferometry_helpers as IH
    import dpi_b200.generative_model.glow_model as GM
    import dpi_b200.generative_model.realnvpfc_model as RN
    from dpi_b200.workloads import interferometry_workload
    npix, NF, NB, B, lr, clip = 8, 2, 2, 6, 1e-3, 1e-2
    wl = interferometry_workload(npix=npix, seed=3, n_stations=5, n_scans=8)
    torch.manual_seed(7)
    import numpy as np
    np.random.seed(7)
    ref = ReferenceStep("interferometry", wl, NF, lr, clip, device="cpu", model_form="glow", n_block=NB)
    with torch.no_grad():
        for name, p in ref.gen.named_parameters():          # ZeroConv2d layers start at zero: make every term live
            if ".net.6." in name or ".prior." in name:
                p.normal_(0, 0.05)
    mine = ReferenceStep("interferometry", wl, NF, lr, clip, device="cuda",
                         modules=dict(realnvp=RN, interferometry=IH, mri=MH, glow=GM), model_form="glow", n_block=NB)
    mine.gen.load_state_dict(ref.gen.state_dict(), strict=True)
    p0 = {k: v.detach().clone() for k, v in ref.gen.named_parameters()}
    losses = []
    for k in range(3):
        z = [torch.randn(B, *s) for s in ref.z_shapes]
        losses.append((ref.step(z), mine.step(z)))
    for k, (a, b) in enumerate(losses):
        assert abs(a - b) < (RTOL if k == 0 else 10 * RTOL) * max(1.0, abs(a)), (k, a, b)
    mine_p = dict(mine.gen.named_parameters())
    for k, p in ref.gen.named_parameters():
        up_ref = (p.detach() - p0[k]).double()
        up_mine = (mine_p[k].detach().cpu() - p0[k]).double()
        assert float(up_ref.norm()) > 0, k
        # Adam normalises every element by its own |g| (an element whose gradient is at round-off level moves by +lr on one
        # side and -lr on the other): judge the update vector per tensor in L2, as the full-size RealNVP trajectory test
        # does, and bound every element by one such flip
        assert float((up_mine - up_ref).norm() / up_ref.norm()) < 0.05, k
        assert float((up_mine - up_ref).abs().max()) < 2.1 * lr, k
    assert abs(float(mine.log_scale) - float(ref.log_scale)) < 1e-5
