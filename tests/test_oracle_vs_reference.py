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
