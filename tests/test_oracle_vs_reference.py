"""CPU, build container only: the oracle restatement against the LIVE reference modules imported from
/root/reference (skipped where that tree is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.needs_reference


@pytest.fixture(scope="module")
def R():
    import ref_shim
    return ref_shim.import_reference()


@pytest.mark.parametrize("D,NF,B,seqfrac,bn", [(64, 3, 16, 4, True), (1024, 2, 8, 4, True), (18, 2, 40, 1 / 16, True),
                                              (2, 2, 30, 1 / 128, True), (33, 2, 9, 1, False)])
def test_flow_bit_identical_to_reference(R, D, NF, B, seqfrac, bn):
    from oracle import dpi_oracle as O
    ref = R["realnvp"].RealNVP(D, NF, seqfrac=seqfrac, batch_norm=bn)
    sd = O.init_state(D, NF, seqfrac=seqfrac, seed=1, randomize=True, batch_norm=bn)
    ref.load_state_dict(sd)
    assert list(ref.state_dict().keys()) == list(sd.keys())
    torch.manual_seed(0)
    z = torch.randn(B, D)
    orders, inv = O.make_orders(D, NF)
    for i in range(NF):
        assert np.array_equal(orders[i], ref.orders[i]) and np.array_equal(inv[i], ref.inverse_orders[i])
    sd1 = {k: v.clone() for k, v in sd.items()}
    xr, lr = ref.reverse(z)
    xo, lo = O.realnvp_reverse(sd1, z, train=True)
    assert torch.equal(xr, xo) and torch.equal(lr, lo)
    zr, l2 = ref.forward(xr.detach())
    zo, l2o = O.realnvp_forward(sd1, xr.detach(), train=True)
    assert torch.equal(zr, zo) and torch.equal(l2, l2o)
    for k, v in ref.state_dict().items():
        assert torch.equal(v, sd1[k]), k
    ref.eval()
    xe, le = ref.reverse(z)
    xeo, leo = O.realnvp_reverse(sd1, z, train=False)
    assert torch.equal(xe, xeo) and torch.equal(le, leo)


def test_forward_model_and_losses_match_reference(R):
    from oracle import dpi_oracle as O
    from dpi_b200.workloads import interferometry_workload
    import types
    IH = R["interferometry"]
    wl = interferometry_workload(npix=16, seed=4, n_stations=6, n_scans=12)
    dev = torch.device("cpu")
    f = IH.eht_observation_pytorch(16, types.SimpleNamespace(to=lambda **k: None), wl["dft_mat"], wl["ktraj"],
                                   wl["pulsefac"], wl["cphase_ind"], wl["cphase_sign"], wl["camp_ind"], dev, ttype="direct")
    torch.manual_seed(0)
    img = torch.rand(5, 16, 16) * 0.02
    a = f(img)
    b = O.eht_observe(img, wl["dft_mat"], wl["cphase_ind"], wl["cphase_sign"], wl["camp_ind"])
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and rel_err(y, x) < 1e-6
    pairs = [(IH.Loss_l1(img), O.loss_l1(img)), (IH.Loss_TSV(img), O.loss_tsv(img)), (IH.Loss_TV(img), O.loss_tv(img)),
             (IH.Loss_flux(2.0)(img), O.loss_flux(2.0, img)),
             (IH.Loss_center(dev, center=7.5, dim=16)(img), O.loss_center(img, 7.5, 16)),
             (IH.Loss_cross_entropy(wl["prior"], img), O.loss_cross_entropy(wl["prior"], img)),
             (IH.Loss_angle_diff(wl["sigma_cp"].numpy(), dev)(wl["cphase_true"], a[2]),
              O.loss_angle_diff(wl["sigma_cp"], wl["cphase_true"], b[2])),
             (IH.Loss_logca_diff2(wl["sigma_ca"].numpy(), dev)(wl["logcamp_true"], a[3]),
              O.loss_logca_diff2(wl["sigma_ca"], wl["logcamp_true"], b[3])),
             (IH.Loss_visamp_diff(np.ones(wl["n_vis"]), dev)(a[1] * 1.1, a[1]),
              O.loss_logca_diff2(torch.ones(wl["n_vis"]), a[1] * 1.1, b[1])),
             (IH.Loss_logamp_diff(np.ones(wl["n_vis"]), dev)(a[1] * 1.1, a[1]),
              O.loss_logamp_diff(torch.ones(wl["n_vis"]), a[1] * 1.1, b[1])),
             (IH.Loss_vis_diff(np.ones(wl["n_vis"]), dev)(a[0][0] * 1.1, a[0]),
              O.loss_vis_diff(torch.ones(wl["n_vis"]), a[0][0] * 1.1, b[0]))]
    for i, (x, y) in enumerate(pairs):
        assert x.dtype == y.dtype and x.shape == y.shape and rel_err(y, x) < 1e-6, i
    MH = R["mri"]
    k = torch.randn(5, 8, 8, 2)
    t = torch.randn(8, 8, 2)
    assert rel_err(O.loss_kspace_diff2(0.3, t, k), MH.Loss_kspace_diff2(0.3)(t, k)) < 1e-6
    assert rel_err(O.loss_kspace_diff(0.3, t, k), MH.Loss_kspace_diff(0.3)(t, k)) < 1e-6
    ref_fft = np.stack((np.fft.fft2(img.numpy(), norm="ortho").real, np.fft.fft2(img.numpy(), norm="ortho").imag), -1)
    assert rel_err(O.fft2c(img), ref_fft) < 1e-5  # numpy twin fft2c (DPI_MRI.py:63-67)


def test_vectorised_closure_maps_equal_reference_loops(R):
    from dpi_b200 import synthetic as syn
    from dpi_b200.interferometry_helpers import Obs_params_torch
    for seed, ns, nsc, flip in [(0, 6, 15, 0.5), (1, 9, 30, 0.25), (2, 4, 10, 0.0), (3, 5, 12, 1.0)]:
        obs = syn.SyntheticObs(n_stations=ns, n_scans=nsc, seed=seed, min_vis=2, flip_frac=flip)
        obs2 = syn.SyntheticObs(n_stations=ns, n_scans=nsc, seed=seed, min_vis=2, flip_frac=flip)
        sim = syn.SimImage(8, 160.0)
        a = R["interferometry"].Obs_params_torch(obs, sim, ttype="direct")
        b = Obs_params_torch(obs2, sim, ttype="direct")
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
        for la, lb in zip(a[3:], b[3:]):
            for x, y in zip(la, lb):
                assert x.dtype == y.dtype and torch.equal(x, y)


@pytest.mark.parametrize("npix,B,ng", [(16, 6, 2), (64, 4, 2), (32, 5, 1)])
def test_crescent_renderer_matches_reference_class(R, npix, B, ng):
    """geometric_model.SimpleCrescentNuisanceFloor_Param2Img (flux_flag=False) vs the oracle restatement."""
    from oracle import dpi_oracle as O
    conv = R["geometric"].SimpleCrescentNuisanceFloor_Param2Img(npix, r_range=[10.0, 40.0], fov=160.0, n_gaussian=ng,
                                                                flux_flag=False)
    torch.manual_seed(npix)
    p = torch.rand(B, conv.nparams)
    c = torch.randn(B, npix, npix)
    pr = p.clone().requires_grad_(True)
    ir = conv.forward(pr)
    (ir * c).sum().backward()
    po = p.clone().requires_grad_(True)
    io = O.crescent_render(po, npix, 160.0, ng)
    (io * c).sum().backward()
    assert rel_err(io, ir) < 1e-6 and rel_err(po.grad, pr.grad) < 1e-5


def _randomised_reference_glow(R, n_flow, n_block, seed):
    """A reference Glow whose zero-initialised pieces (ZeroConv2d, ActNorm) are made non-trivial, marked initialised."""
    torch.manual_seed(seed)
    np.random.seed(seed)                       # InvConv2dLU draws its matrix with numpy (:185)
    ref = R["glow"].Glow(1, n_flow, n_block, affine=True, conv_lu=True)
    with torch.no_grad():
        for name, p in ref.named_parameters():
            if ".net.6." in name or ".prior." in name:
                p.normal_(0, 0.05)
            elif name.endswith("actnorm.loc"):
                p.normal_(0, 0.2)
            elif name.endswith("actnorm.scale_inv"):
                p.copy_(0.7 + 0.6 * torch.rand_like(p))
        for name, b in ref.named_buffers():
            if name.endswith("initialized"):
                b.fill_(1)
    return ref


@pytest.mark.parametrize("npix,n_flow,n_block,B", [(8, 2, 2, 5), (16, 1, 3, 3)])
def test_glow_oracle_bit_identical_to_reference(R, npix, n_flow, n_block, B):
    """the restatement in oracle/glow_oracle.py against the unmodified reference class -
    reverse (the training direction) and forward, train and eval mode, running statistics, input gradients."""
    from oracle import glow_oracle as G
    ref = _randomised_reference_glow(R, n_flow, n_block, seed=3)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    shapes = G.calc_z_shapes(1, npix, n_flow, n_block)
    assert shapes == R["glow"].calc_z_shapes(1, npix, n_flow, n_block)
    torch.manual_seed(1)
    zs = [torch.randn(B, *s) for s in shapes]
    zr = [z.clone().requires_grad_(True) for z in zs]
    zo = [z.clone().requires_grad_(True) for z in zs]
    xr, lr = ref.reverse(zr)
    xo, lo = G.glow_reverse(sd, zo, train=True)
    assert xr.shape == (B, 1, npix, npix)
    assert torch.equal(xr, xo) and torch.equal(lr, lo)
    c = torch.randn_like(xr)
    ((xr * c).sum() + lr.sum()).backward()
    ((xo * c).sum() + lo.sum()).backward()
    for a, b in zip(zr, zo):
        assert torch.equal(a.grad, b.grad)
    for k, v in ref.state_dict().items():           # BatchNorm running statistics moved identically
        assert torch.equal(v, sd[k]), k
    with torch.no_grad():
        pr, dr, zl = ref.forward(xr.detach())
        po, do, zlo = G.glow_forward(sd, xr.detach(), train=True)
        assert torch.equal(pr, po) and torch.equal(dr, do)
        for a, b in zip(zl, zlo):
            assert torch.equal(a, b)
        ref.eval()
        xe, le = ref.reverse(zs)
        xeo, leo = G.glow_reverse(sd, zs, train=False)
        assert torch.equal(xe, xeo) and torch.equal(le, leo)
        # invertibility in eval mode (the BatchNorm statistics are then fixed): forward(reverse(z)) returns z
        _, ldf, zb = ref.forward(xe)
        for a, b in zip(zs, zb):
            assert float((a - b).abs().max()) < 2e-3
        assert float((ldf + le).abs().max()) < 1e-2 * max(1.0, float(le.abs().max()))


def test_glow_oracle_data_dependent_actnorm_init(R):
    """First forward call of a fresh ActNorm: loc = -mean_c, scale_inv = std_c + 1e-6 (glow_model.py:91-111)."""
    from oracle import glow_oracle as G
    an = R["glow"].ActNorm(6)
    sd = {k: v.detach().clone() for k, v in an.state_dict().items()}
    torch.manual_seed(2)
    x = torch.randn(7, 6, 4, 4) * 1.3 + 0.4
    yr, lr = an(x)
    G.actnorm_initialize(sd, "", x, inv_init=False)
    yo, lo = G.actnorm_forward(sd, "", x)
    assert torch.equal(yr, yo) and torch.equal(lr, lo)
    for k, v in an.state_dict().items():
        assert torch.equal(v, sd[k]), k
