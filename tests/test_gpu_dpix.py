"""GPU parity of the alpha-DPI path: crescent renderer (geometric_model.py:237-389) and the
alpha-divergence training step (DPIx_interferometry.py:329-386) against the fixtures generated from the unmodified
reference modules and against the CPU oracle. Tolerance 1e-4 relative (north_star, fp32)."""
import numpy as np
import pytest
import torch

from conftest import RTOL, load_golden, rel_err, split_sd
from oracle import dpi_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["dpix_small", "dpix_kl"])
def test_renderer_matches_reference_fixture(name):
    from dpi_b200.geometric_model import SimpleCrescentNuisanceFloor_Param2Img
    g = load_golden(name)
    conv = SimpleCrescentNuisanceFloor_Param2Img(int(g["npix"]), r_range=[10.0, 40.0], fov=160.0, n_gaussian=2,
                                                 flux_flag=False).to(device="cuda")
    assert conv.nparams == 18
    p = torch.from_numpy(g["render_params"]).cuda().requires_grad_(True)
    img = conv.forward(p)
    (img * torch.from_numpy(g["render_cimg"]).cuda()).sum().backward()
    assert rel_err(img, g["render_img"]) < RTOL, rel_err(img, g["render_img"])
    assert rel_err(p.grad, g["render_gparams"]) < RTOL, rel_err(p.grad, g["render_gparams"])


def test_renderer_with_total_flux_matches_reference_fixture():
    """flux_flag=True (geometric_model.py:253-256, :305-309, :373-376): 7 + 6 g parameters, image = total flux x unit-flux
    image; image, parameter gradient and the feature tuple against the unmodified reference (tests/golden/make_golden.py
    crescent_flux_case)."""
    from dpi_b200.geometric_model import SimpleCrescentNuisanceFloor_Param2Img
    g = load_golden("crescent_flux")
    conv = SimpleCrescentNuisanceFloor_Param2Img(int(g["npix"]), r_range=[10.0, 40.0], fov=160.0, n_gaussian=int(g["ng"]),
                                                 flux_flag=True, flux_range=[float(g["flux_lo"]), float(g["flux_hi"])])
    assert conv.nparams == int(g["nparams"]) == 19
    p = torch.from_numpy(g["render_params"]).cuda().requires_grad_(True)
    img = conv.forward(p)
    (img * torch.from_numpy(g["render_cimg"]).cuda()).sum().backward()
    assert rel_err(img, g["render_img"]) < RTOL, rel_err(img, g["render_img"])
    assert rel_err(p.grad, g["render_gparams"]) < RTOL, rel_err(p.grad, g["render_gparams"])
    feats = conv.compute_features(p.detach())
    assert len(feats) == int(g["n_feats"]) == 13
    assert rel_err(feats[4], g["feat_flux"]) < 1e-6 and rel_err(feats[5], g["feat_crescent_flux"]) < 1e-6
    assert rel_err(feats[6], g["feat_floor"]) < 1e-6


@pytest.mark.parametrize("npix,B,ng", [(64, 32, 2), (32, 7, 1), (24, 5, 0)])
def test_renderer_vs_oracle(npix, B, ng):
    from dpi_b200.geometric_model import SimpleCrescentNuisanceFloor_Param2Img
    torch.manual_seed(npix + B)
    p0 = torch.rand(B, 6 + 6 * ng) * 0.9 + 0.05
    c = torch.randn(B, npix, npix)
    po = p0.clone().double().requires_grad_(True)          # fp64 oracle: the 1e-4 bar is against exact math here
    io = O.crescent_render(po, npix, 160.0, ng)
    (io * c.double()).sum().backward()
    conv = SimpleCrescentNuisanceFloor_Param2Img(npix, n_gaussian=ng, fov=160)
    pg = p0.cuda().requires_grad_(True)
    ig = conv(pg)
    (ig * c.cuda()).sum().backward()
    assert rel_err(ig, io) < RTOL, rel_err(ig, io)
    assert rel_err(pg.grad, po.grad) < RTOL, rel_err(pg.grad, po.grad)
    assert float((ig.sum((1, 2)) - 1).abs().max()) < 1e-3        # unit flux (up to the reference's eps)


@pytest.mark.parametrize("name", ["dpix_small", "dpix_kl"])
def test_dpix_training_trajectory_matches_reference_run(name):
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    from dpi_b200.trainer import DPIxTrainer
    from dpi_b200.workloads import dpix_workload
    g = load_golden(name)
    npix, NF, B = int(g["npix"]), int(g["NF"]), int(g["B"])
    wl = dpix_workload(npix=npix, seed=int(g["seed"]), n_stations=int(g["n_stations"]), n_scans=int(g["n_scans"]),
                       alpha=float(g["alpha"]))
    flow = RealNVP(18, NF, seqfrac=1 / 16)
    flow.load_state_dict(split_sd(g))
    tr = DPIxTrainer(flow, wl, B, lr=float(g["lr"]), clip=float(g["clip"]), start_order=int(g["start_order"]),
                     decay_rate=int(g["decay_rate"]))
    for k in range(int(g["steps"])):
        z = torch.from_numpy(g["z"][k]).cuda()
        if k == 0:
            bn = tr.flow.bn_running.clone()
            tr.loss_and_grad(z)
            torch.cuda.synchronize()
            assert rel_err(tr.img.view(B, npix, npix), g["img0"]) < RTOL
            assert rel_err(tr.prm, g["params0"]) < RTOL
            assert rel_err(tr.logdet_flow + tr.det, g["logdet0"]) < RTOL
            assert rel_err(tr.loss_cp, g["l_cp0"]) < RTOL and rel_err(tr.loss_ca, g["l_ca0"]) < RTOL
            views = dict(tr.flow.named_grad_views(tr.grads))
            worst = max((rel_err(v, g["grad0." + kk]), kk) for kk, v in views.items() if "actnorm" not in kk)
            assert worst[0] < 5 * RTOL, worst
            tr.flow.bn_running.copy_(bn)
            tr.flow._nbt -= 1
        t = tr.step(z).cpu()
        got = [float(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[4]), float(tr.norm_out[0])]
        assert np.allclose(got, g["hist"][k], rtol=5e-4, atol=1e-5), (k, got, list(g["hist"][k]))
    mine = tr.flow.state_dict()
    for kk, v in split_sd(g, "sd_final.").items():
        assert rel_err(mine[kk].float().cpu(), v.float()) < 2 * RTOL, kk


def test_dpix_config_size_step_vs_oracle():
    """C3 shapes (npix = 64, 18 parameters, 7 stations / 25 scans) at a reduced batch: one step against the oracle."""
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    from dpi_b200.trainer import DPIxTrainer
    from dpi_b200.workloads import dpix_workload
    wl = dpix_workload(npix=64)
    B, NF = 96, 4
    sd = O.init_state(18, NF, seqfrac=1 / 16, seed=5, randomize=True)
    flow = RealNVP(18, NF, seqfrac=1 / 16)
    flow.load_state_dict({k: v.clone() for k, v in sd.items()})
    tr = DPIxTrainer(flow, wl, B, lr=1e-4, clip=1e-4)
    c = dict(npix=64, fov=160.0, n_gaussian=2, dft_mat=wl["dft_mat"], cphase_ind=wl["cphase_ind"],
             cphase_sign=wl["cphase_sign"], camp_ind=wl["camp_ind"], cphase_true=wl["cphase_true"],
             logcamp_true=wl["logcamp_true"], sigma_cp=wl["sigma_cp"], sigma_ca=wl["sigma_ca"], w=wl["w"], alpha=0.95)
    ot = O.OracleTrainer(sd, None, O.dpix_loss, c, lr=1e-4, clip=1e-4)
    ot.data_weight = 1e-4
    torch.manual_seed(2)
    z = torch.randn(B, 18)
    lo, to = ot.step(z)
    t = tr.step(z.cuda(), data_weight=1e-4).cpu()
    assert abs(float(t[0]) - float(lo)) < 5e-4 * abs(float(lo)), (float(t[0]), float(lo))
    assert abs(float(tr.norm_out[0]) - float(to["grad_norm"])) < 5e-4 * float(to["grad_norm"])
    mine = tr.flow.state_dict()
    worst = max(rel_err(mine[k].cpu(), sd[k].detach()) for k in ot.keys)
    assert worst < 2 * RTOL, worst


def test_posterior_sampling_weights_vs_oracle():
    """SURVEY 8(f) n1, `Gen_samples` (DPIx_interferometry.py:433-474): eval-mode flow, importance weights
    softmax(-(0.5 loss_data + logprob)) and the feature extraction used for the posterior summaries."""
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    from dpi_b200.geometric_model import SimpleCrescentNuisanceFloor_Param2Img
    from dpi_b200.trainer import DPIxTrainer
    from dpi_b200.workloads import dpix_workload
    wl = dpix_workload(npix=32, n_scans=12)
    B, NF = 128, 3
    sd = O.init_state(18, NF, seqfrac=1 / 16, seed=8, randomize=True)
    for k in sd:                                    # non-trivial running statistics for the eval-mode pass
        if k.endswith("running_mean"):
            sd[k] = torch.randn_like(sd[k]) * 0.05
        if k.endswith("running_var"):
            sd[k] = 1 + 0.2 * torch.rand_like(sd[k])
    flow = RealNVP(18, NF, seqfrac=1 / 16)
    flow.load_state_dict({k: v.clone() for k, v in sd.items()})
    tr = DPIxTrainer(flow, wl, B)
    torch.manual_seed(4)
    z = torch.randn(B, 18)
    res = tr.sample_posterior(z=z.cuda(), rejsamp=True)
    with torch.no_grad():
        ps, logdet = O.realnvp_reverse({k: v.clone() for k, v in sd.items()}, z, train=False)
        prm = torch.sigmoid(ps)
        img = O.crescent_render(prm, 32, 160.0, 2)
        logdet = logdet + torch.sum(-ps - 2 * torch.nn.functional.softplus(-ps), -1)
        _, _, cph, lca = O.eht_observe(img, wl["dft_mat"], wl["cphase_ind"], wl["cphase_sign"], wl["camp_ind"])
        ld = wl["w"]["camp"] * O.loss_logca_diff2(wl["sigma_ca"], wl["logcamp_true"], lca) \
            + wl["w"]["cphase"] * O.loss_angle_diff(wl["sigma_cp"], wl["cphase_true"], cph)
        lo = 0.5 * ld + (-logdet - 0.5 * (z ** 2).sum(1))
        w = torch.softmax(-lo.double(), 0)
    assert rel_err(res["params"], prm) < RTOL and rel_err(res["img"], img) < RTOL
    assert rel_err(res["weights"], w) < 5e-3          # exp() of O(1e2) exponents amplifies 1e-5 input differences
    assert bool(res["accepted"][torch.argmax(res["weights"])])
    assert res["img_mean"].shape == (32, 32) and res["img_std"].shape == (32, 32)
    # Gen_samples2 (:665-709): no rejection, importance-weighted summaries
    res2 = tr.sample_posterior2(z=z.cuda())
    wm = (w.float().view(-1, 1, 1) * img).sum(0)
    ws = torch.sqrt((w.float().view(-1, 1, 1) * (img - wm) ** 2).sum(0))
    assert "accepted" not in res2 and res2["img_mean"].shape == (1, 32, 32)
    assert rel_err(res2["img_mean"][0], wm) < 5e-3 and rel_err(res2["img_std"][0], ws) < 5e-3
    assert abs(float(res2["weights"].sum()) - 1.0) < 1e-4
    conv = SimpleCrescentNuisanceFloor_Param2Img(32, n_gaussian=2, fov=160)
    feats = conv.compute_features(res["params"])
    assert len(feats) == 12 and feats[0].shape == (B, 1, 1)
    assert float(feats[0].min()) >= 10.0 / 80.0 - 1e-6 and float(feats[0].max()) <= 40.0 / 80.0 + 1e-6
