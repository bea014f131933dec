"""Generate tests/golden/*.npz from the UNMODIFIED reference modules (build container only).

    python tests/golden/make_golden.py

Every fixture is produced by importing /root/reference/DPItorch (see ref_shim.py) and running the
reference classes / functions on seeded inputs; the loop bodies of the reference scripts
(DPI_interferometry.py:201-235, DPI_MRI.py:170-207 - scripts that cannot be imported because they
load ehtim data at module scope) are assembled here around those imported pieces, with
torch.optim.Adam and nn.utils.clip_grad_norm_ as the scripts use them.
The oracle (oracle/dpi_oracle.py) is NOT used here: these files pin the oracle, not vice versa.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402
from dpi_b200 import synthetic as syn  # noqa: E402
from oracle.dpi_oracle import init_state  # noqa: E402  (weight generator only: a seeded tensor factory)

R = ref_shim.import_reference()
RealNVP = R["realnvp"].RealNVP
IH = R["interferometry"]
MH = R["mri"]


def sd_to_np(sd, prefix="sd."):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def flow_case(name, D, NF, B, seqfrac, seed, batch_norm=True):
    torch.manual_seed(seed)
    ref = RealNVP(D, NF, seqfrac=seqfrac, batch_norm=batch_norm)
    sd = init_state(D, NF, seqfrac=seqfrac, seed=seed, randomize=True, batch_norm=batch_norm)
    ref.load_state_dict(sd)
    out = dict(D=D, NF=NF, B=B, seqfrac=seqfrac, batch_norm=int(batch_norm))
    out.update(sd_to_np(sd))
    z = torch.randn(B, D)
    out["z"] = z.numpy()
    # --- reverse, train mode, with gradients of a generic scalar objective
    zz = z.clone().requires_grad_(True)
    x, ld = ref.reverse(zz)
    cx = torch.randn(B, D)
    cl = torch.randn(B)
    loss = (x * cx).sum() + (ld * cl).sum()
    loss.backward()
    out.update(x_rev=x.detach().numpy(), ld_rev=ld.detach().numpy(), cx=cx.numpy(), cl=cl.numpy(),
               g_z=zz.grad.numpy())
    for k, p in ref.named_parameters():
        out["grad." + k] = p.grad.detach().numpy()
    out.update(sd_to_np(ref.state_dict(), "sd_after_rev."))
    # --- forward (density evaluation) from the produced x, still train mode
    with torch.no_grad():
        zf, ldf = ref.forward(x.detach())
        out.update(z_fwd=zf.numpy(), ld_fwd=ldf.numpy())
        # --- eval mode (running statistics)
        ref.eval()
        xe, lde = ref.reverse(z)
        ze, ldze = ref.forward(xe)
        out.update(x_rev_eval=xe.numpy(), ld_rev_eval=lde.numpy(), z_fwd_eval=ze.numpy(), ld_fwd_eval=ldze.numpy())
    # --- data-dependent ActNorm init on a fresh model's first forward
    torch.manual_seed(seed + 100)
    ref2 = RealNVP(D, NF, seqfrac=seqfrac, batch_norm=batch_norm)
    out.update(sd_to_np(ref2.state_dict(), "sd_fresh."))
    with torch.no_grad():
        x0 = torch.randn(B, D) * 1.7 + 0.3
        z0, l0 = ref2.forward(x0)
    out.update(x_fresh=x0.numpy(), z_fresh=z0.numpy(), ld_fresh=l0.numpy())
    out.update(sd_to_np(ref2.state_dict(), "sd_fresh_after."))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", sum(v.nbytes for v in out.values() if hasattr(v, "nbytes")) // 1024, "KiB")


def obs_arrays(obs):
    """Station names -> integer codes so the fixture is plain numeric."""
    names = sorted(set(obs.data["t1"]) | set(obs.data["t2"]))
    code = {n: i for i, n in enumerate(names)}
    enc = lambda a: np.array([code[x] for x in a], dtype=np.int64)  # noqa: E731
    return dict(time=obs.data["time"].copy(), t1=enc(obs.data["t1"]), t2=enc(obs.data["t2"]), u=obs.data["u"].copy(),
                v=obs.data["v"].copy(),
                cp_time=obs.cphase["time"].copy(), cp_t1=enc(obs.cphase["t1"]), cp_t2=enc(obs.cphase["t2"]),
                cp_t3=enc(obs.cphase["t3"]),
                ca_time=obs.camp["time"].copy(), ca_t1=enc(obs.camp["t1"]), ca_t2=enc(obs.camp["t2"]),
                ca_t3=enc(obs.camp["t3"]), ca_t4=enc(obs.camp["t4"]))


def closure_map_case(name, n_stations, n_scans, seed, npix, min_vis=2):
    obs = syn.SyntheticObs(n_stations=n_stations, n_scans=n_scans, seed=seed, min_vis=min_vis, flip_frac=0.4)
    sim = syn.SimImage(npix, 160.0)
    dft, ktraj, pf, cpi, cps, cai = IH.Obs_params_torch(obs, sim, snrcut=0.0, ttype="direct")
    out = obs_arrays(obs)
    out.update(npix=npix, n_stations=n_stations, n_scans=n_scans, seed=seed, min_vis=min_vis,
               dft_mat=dft.numpy() if npix <= 8 else dft.numpy()[:, :16], ktraj=ktraj.numpy(), pulsefac=pf.numpy(),
               cp_ind=np.stack([t.numpy() for t in cpi]), cp_sign=np.stack([t.numpy() for t in cps]),
               ca_ind=np.stack([t.numpy() for t in cai]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok n_vis", len(obs.data), "n_cp", len(obs.cphase), "n_ca", len(obs.camp))


def interferometry_case(name, npix, NF, B, seed, n_stations=5, n_scans=8, steps=3):
    """DPI_interferometry.py:201-235 assembled from the reference modules."""
    from dpi_b200.workloads import interferometry_workload
    wl = interferometry_workload(npix=npix, seed=seed, n_stations=n_stations, n_scans=n_scans)
    # operator tables from the reference's own loops (checked equal to ours in test_host_logic)
    obs = syn.SyntheticObs(n_stations=n_stations, n_scans=n_scans, seed=seed, min_vis=2)
    dft, ktraj, pf, cpi, cps, cai = IH.Obs_params_torch(obs, syn.SimImage(npix, 160.0), ttype="direct")
    dev = torch.device("cpu")
    eht = IH.eht_observation_pytorch(npix, ref_shim.types.SimpleNamespace(to=lambda **k: None), dft, ktraj, pf, cpi, cps,
                                     cai, dev, ttype="direct")
    w = wl["w"]
    torch.manual_seed(seed)
    gen = RealNVP(npix * npix, NF)
    sd = init_state(npix * npix, NF, seed=seed, randomize=True)
    gen.load_state_dict(sd)
    log_scale = torch.nn.Parameter(torch.tensor([wl["log_scale_init"]], dtype=torch.float32))
    L_center = IH.Loss_center(dev, center=npix / 2 - 0.5, dim=npix)
    L_flux = IH.Loss_flux(wl["flux"])
    L_cp = IH.Loss_angle_diff(wl["sigma_cp"].numpy(), dev)
    L_ca = IH.Loss_logca_diff2(wl["sigma_ca"].numpy(), dev)
    params = list(gen.parameters()) + [log_scale]
    opt = torch.optim.Adam(params, lr=1e-4)
    out = dict(npix=npix, NF=NF, B=B, seed=seed, n_stations=n_stations, n_scans=n_scans, steps=steps, lr=1e-4,
               clip=1e-3)
    out.update(sd_to_np(sd))
    zs, hist = [], []
    for k in range(steps):
        z = torch.randn(B, npix * npix)
        zs.append(z.numpy())
        img_samp, logdet = gen.reverse(z)
        img_samp = img_samp.reshape((-1, npix, npix))
        scale_factor = torch.exp(log_scale)
        img = torch.nn.Softplus()(img_samp) * scale_factor
        det_softplus = torch.sum(img_samp - torch.nn.Softplus()(img_samp), (1, 2))
        logdet = logdet + det_softplus + log_scale * npix * npix
        vis, visamp, cphase, logcamp = eht(img)
        l_center, l_l1, l_tsv = L_center(img), IH.Loss_l1(img), IH.Loss_TSV(img)
        l_mem = IH.Loss_cross_entropy(wl["prior"], img)
        l_flux = L_flux(img)
        l_cp = L_cp(wl["cphase_true"], cphase)
        l_ca = L_ca(wl["logcamp_true"], logcamp)
        loss_data = w["camp"] * l_ca + w["cphase"] * l_cp
        loss_prior = w["mem"] * l_mem + w["flux"] * l_flux + w["tsv"] * l_tsv + w["center"] * l_center + w["l1"] * l_l1
        loss = torch.mean(loss_data) + torch.mean(loss_prior) - w["logdet"] * torch.mean(logdet)
        opt.zero_grad()
        loss.backward()
        if k == 0:
            out.update(img0=img.detach().numpy(), vis0=vis.detach().numpy(), visamp0=visamp.detach().numpy(),
                       cphase0=cphase.detach().numpy(), logcamp0=logcamp.detach().numpy(),
                       l_cp0=l_cp.detach().numpy(), l_ca0=l_ca.detach().numpy(), l_center0=l_center.detach().numpy(),
                       l_l10=l_l1.detach().numpy(), l_tsv0=l_tsv.detach().numpy(), l_mem0=l_mem.detach().numpy(),
                       l_flux0=l_flux.detach().numpy(), l_tv0=IH.Loss_TV(img).detach().numpy(),
                       logdet0=logdet.detach().numpy(), g_log_scale0=log_scale.grad.numpy().copy())
            for kk, p in gen.named_parameters():
                out["grad0." + kk] = p.grad.detach().numpy().copy()
        norm = torch.nn.utils.clip_grad_norm_(params, 1e-3)
        opt.step()
        hist.append([float(loss), float(l_cp.mean()), float(l_ca.mean()), float(loss_prior.mean()),
                     float(logdet.mean()), float(norm), float(log_scale)])
    out["z"] = np.stack(zs)
    out["hist"] = np.array(hist, dtype=np.float64)
    out.update(sd_to_np(gen.state_dict(), "sd_final."))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", hist[-1])


def mri_case(name, npix, NF, B, seed, steps=3):
    """DPI_MRI.py:170-207 assembled from the reference modules (fft2c_torch restated with torch.fft.fft2,
    the removed torch.fft(x, 2, normalized=True) it calls is the same transform as its numpy twin :63-67)."""
    from dpi_b200.workloads import mri_workload
    wl = mri_workload(npix=npix, seed=seed)
    w = wl["w"]
    torch.manual_seed(seed)
    gen = RealNVP(npix * npix, NF)
    sd = init_state(npix * npix, NF, seed=seed, randomize=True)
    gen.load_state_dict(sd)
    log_scale = torch.nn.Parameter(torch.tensor([wl["log_scale_init"]], dtype=torch.float32))
    L_k = MH.Loss_kspace_diff2(wl["sigma"])
    params = list(gen.parameters()) + [log_scale]
    opt = torch.optim.Adam(params, lr=1e-5)
    out = dict(npix=npix, NF=NF, B=B, seed=seed, steps=steps, lr=1e-5, clip=1e-2)
    out.update(sd_to_np(sd))
    zs, hist = [], []
    mask, ktrue = wl["mask"], wl["kspace_true"]
    for k in range(steps):
        z = torch.randn(B, npix * npix)
        zs.append(z.numpy())
        img_samp, logdet = gen.reverse(z)
        img_samp = img_samp.reshape((-1, npix, npix))
        img = torch.nn.Softplus()(img_samp) * torch.exp(log_scale)
        det_softplus = torch.sum(img_samp - torch.nn.Softplus()(img_samp), (1, 2))
        logdet = logdet + det_softplus + log_scale * npix * npix
        kpred = torch.view_as_real(torch.fft.fft2(img, norm="ortho"))
        loss_data = L_k(ktrue, kpred * mask) / float(mask.mean())
        l_tv = MH.Loss_TV(img)
        loss_prior = w["tv"] * l_tv
        loss = torch.mean(loss_data) + torch.mean(loss_prior) - w["logdet"] * torch.mean(logdet)
        opt.zero_grad()
        loss.backward()
        if k == 0:
            out.update(img0=img.detach().numpy(), kpred0=kpred.detach().numpy(), l_data0=loss_data.detach().numpy(),
                       l_tv0=l_tv.detach().numpy(), logdet0=logdet.detach().numpy(),
                       g_log_scale0=log_scale.grad.numpy().copy())
            for kk, p in gen.named_parameters():
                out["grad0." + kk] = p.grad.detach().numpy().copy()
        norm = torch.nn.utils.clip_grad_norm_(params, 1e-2)
        opt.step()
        hist.append([float(loss), float(loss_data.mean()), float(loss_prior.mean()), float(logdet.mean()),
                     float(norm), float(log_scale)])
    out["z"] = np.stack(zs)
    out["hist"] = np.array(hist, dtype=np.float64)
    out.update(sd_to_np(gen.state_dict(), "sd_final."))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", hist[-1])


def dpix_case(name, npix, NF, B, seed, alpha, n_stations=5, n_scans=8, steps=3):
    """DPIx_interferometry.py:329-386 (cphase_logcamp, simple_crescent_floor_nuisance) assembled from the
    reference modules: RealNVP(18, NF, seqfrac=1/16), SimpleCrescentNuisanceFloor_Param2Img, eht closure operator."""
    from dpi_b200.workloads import dpix_workload
    wl = dpix_workload(npix=npix, seed=seed, n_stations=n_stations, n_scans=n_scans, alpha=alpha)
    obs = syn.SyntheticObs(n_stations=n_stations, n_scans=n_scans, seed=seed, min_vis=2)
    dft, ktraj, pf, cpi, cps, cai = IH.Obs_params_torch(obs, syn.SimImage(npix, 160.0), ttype="direct")
    dev = torch.device("cpu")
    eht = IH.eht_observation_pytorch(npix, ref_shim.types.SimpleNamespace(to=lambda **k: None), dft, ktraj, pf, cpi, cps,
                                     cai, dev, ttype="direct")
    GM = R["geometric"]
    conv = GM.SimpleCrescentNuisanceFloor_Param2Img(npix, r_range=[10.0, 40.0], fov=160.0, n_gaussian=2, flux_flag=False)
    nparams = conv.nparams
    w, sf = wl["w"], wl["w"]["scale_factor"]
    torch.manual_seed(seed)
    gen = RealNVP(nparams, NF, seqfrac=1 / 16)
    sd = init_state(nparams, NF, seqfrac=1 / 16, seed=seed, randomize=True)
    gen.load_state_dict(sd)
    L_cp = IH.Loss_angle_diff(wl["sigma_cp"].numpy(), dev)
    L_ca = IH.Loss_logca_diff2(wl["sigma_ca"].numpy(), dev)
    params_l = list(gen.parameters())
    opt = torch.optim.Adam(params_l, lr=1e-4)
    out = dict(npix=npix, NF=NF, B=B, seed=seed, n_stations=n_stations, n_scans=n_scans, steps=steps, lr=1e-4,
               clip=1e-4, alpha=alpha, start_order=4, decay_rate=2000)
    out.update(sd_to_np(sd))
    # renderer alone: image and parameter gradient of a random linear functional
    p0 = torch.rand(B, nparams, generator=torch.Generator().manual_seed(seed)).requires_grad_(True)
    cimg = torch.randn(B, npix, npix, generator=torch.Generator().manual_seed(seed + 1))
    im0 = conv.forward(p0)
    (im0 * cimg).sum().backward()
    out.update(render_params=p0.detach().numpy(), render_cimg=cimg.numpy(), render_img=im0.detach().numpy(),
               render_gparams=p0.grad.numpy().copy())
    zs, hist = [], []
    for k in range(steps):
        data_weight = min(10 ** (-4 + k / 2000), 1.0)
        z = torch.randn(B, nparams)
        zs.append(z.numpy())
        ps, logdet = gen.reverse(z)
        prm = torch.sigmoid(ps)
        img = conv.forward(prm)
        det_sigmoid = torch.sum(-ps - 2 * torch.nn.Softplus()(-ps), -1)
        logdet = logdet + det_sigmoid
        vis, visamp, cphase, logcamp = eht(img)
        l_cp = L_cp(wl["cphase_true"], cphase)
        l_ca = L_ca(wl["logcamp_true"], logcamp)
        loss_data = 0.5 * (w["camp"] * l_ca + w["cphase"] * l_cp) / sf
        logprob = -logdet - 0.5 * torch.sum(z ** 2, 1)
        loss = data_weight * loss_data + logprob
        if alpha == 1:
            loss = torch.mean(sf * loss)
        else:
            rej = torch.nn.Softmax(dim=0)(-(1 - alpha) * loss).detach()
            loss = torch.sum(rej * sf * loss)
        loss_orig = loss_data + logprob
        if alpha == 1:
            loss_orig = torch.mean(sf * loss_orig)
        else:
            loss_orig = sf * torch.log(torch.mean(torch.exp(-(1 - alpha) * loss_orig))) / (alpha - 1)
        opt.zero_grad()
        loss.backward()
        if k == 0:
            out.update(img0=img.detach().numpy(), params0=prm.detach().numpy(), logdet0=logdet.detach().numpy(),
                       l_cp0=l_cp.detach().numpy(), l_ca0=l_ca.detach().numpy())
            for kk, p in gen.named_parameters():
                out["grad0." + kk] = p.grad.detach().numpy().copy()
        norm = torch.nn.utils.clip_grad_norm_(params_l, 1e-4)
        opt.step()
        hist.append([float(loss), float(loss_orig), float(l_cp.mean()), float(l_ca.mean()), float(logdet.mean()), float(norm)])
    out["z"] = np.stack(zs)
    out["hist"] = np.array(hist, dtype=np.float64)
    out.update(sd_to_np(gen.state_dict(), "sd_final."))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", hist[-1])


def glow_case(name, npix, n_flow, n_block, B, seed):
    """the reference Glow (generative_model/glow_model.py) in reverse (training direction) with
    gradients, forward, and eval mode. The 512 x 512 1x1-conv weights are set from a closed form shared with the
    test (oracle.glow_oracle.formula_weight) instead of being stored."""
    from oracle import glow_oracle as G
    torch.manual_seed(seed)
    np.random.seed(seed)
    ref = R["glow"].Glow(1, n_flow, n_block, affine=True, conv_lu=True)
    with torch.no_grad():
        for pname, p in ref.named_parameters():
            if ".net.6." in pname or ".prior." in pname:
                p.normal_(0, 0.05)
            elif pname.endswith("actnorm.loc"):
                p.normal_(0, 0.2)
            elif pname.endswith("actnorm.scale_inv"):
                p.copy_(0.7 + 0.6 * torch.rand_like(p))
        for bname, b in ref.named_buffers():
            if bname.endswith("initialized"):
                b.fill_(1)
        big = G.big_weight_keys(ref.state_dict().keys())
        params = dict(ref.named_parameters())
        for tag, k in enumerate(big):
            params[k].copy_(G.formula_weight(params[k].shape, tag))
    sd0 = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    out = dict(npix=npix, n_flow=n_flow, n_block=n_block, B=B)
    out.update({("sd." + k): v.numpy().copy() for k, v in sd0.items() if k not in big})
    shapes = R["glow"].calc_z_shapes(1, npix, n_flow, n_block)
    zs = [torch.randn(B, *s) for s in shapes]
    zr = [z.clone().requires_grad_(True) for z in zs]
    x, ld = ref.reverse(zr)
    c = torch.randn_like(x)
    ((x * c).sum() + ld.sum()).backward()
    out.update(x_rev=x.detach().numpy(), ld_rev=ld.detach().numpy(), c=c.numpy())
    for i, (z, zz) in enumerate(zip(zs, zr)):
        out[f"z{i}"] = z.numpy()
        out[f"g_z{i}"] = zz.grad.numpy()
    # parameter gradients of the same objective (for the backward pass of the CUDA path): small tensors verbatim, the
    # 512 x 512 ones as every 1021st entry of the flattened gradient
    for k, p in ref.named_parameters():
        gflat = p.grad.detach().reshape(-1)
        out["grad." + k] = gflat[::1021].numpy().copy() if k in big else p.grad.detach().numpy().copy()
    for k, v in ref.state_dict().items():
        if "running" in k:
            out["after." + k] = v.numpy().copy()
    with torch.no_grad():
        lp, ldf, zf = ref.forward(x.detach())
        out.update(logp_fwd=lp.numpy(), ld_fwd=ldf.numpy())
        for i, z in enumerate(zf):
            out[f"zf{i}"] = z.numpy()
        ref.eval()
        xe, lde = ref.reverse(zs)
        out.update(x_rev_eval=xe.numpy(), ld_rev_eval=lde.numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", sum(v.nbytes for v in out.values() if hasattr(v, "nbytes")) // 1024, "KiB")


def crescent_flux_case(name, npix, B, ng, seed, flux_range):
    """SURVEY 8 a17 with flux_flag=True (geometric_model.py:253-256, :305-309, :373-376; the renderer DPIx_interferometry.py
    builds for every data product except cphase_logcamp, :214-229): image, parameter gradient and the feature tuple."""
    conv = R["geometric"].SimpleCrescentNuisanceFloor_Param2Img(npix, r_range=[10.0, 40.0], fov=160.0, n_gaussian=ng,
                                                                flux_flag=True, flux_range=flux_range)
    p0 = torch.rand(B, conv.nparams, generator=torch.Generator().manual_seed(seed)).requires_grad_(True)
    cimg = torch.randn(B, npix, npix, generator=torch.Generator().manual_seed(seed + 1))
    im0 = conv.forward(p0)
    (im0 * cimg).sum().backward()
    feats = conv.compute_features(p0.detach())
    out = dict(npix=npix, B=B, ng=ng, nparams=conv.nparams, flux_lo=flux_range[0], flux_hi=flux_range[1],
               render_params=p0.detach().numpy(), render_cimg=cimg.numpy(), render_img=im0.detach().numpy(),
               render_gparams=p0.grad.numpy().copy(), feat_flux=feats[4].numpy(), feat_crescent_flux=feats[5].numpy(),
               feat_floor=feats[6].numpy(), n_feats=len(feats))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", sum(v.nbytes for v in out.values() if hasattr(v, "nbytes")) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(4)
    if len(sys.argv) > 1 and sys.argv[1] == "crescent_flux":      # add this fixture without regenerating the others
        crescent_flux_case("crescent_flux", npix=16, B=9, ng=2, seed=33, flux_range=[0.8 * 0.6, 1.2 * 0.6])
        sys.exit(0)
    flow_case("flow_even", D=20, NF=2, B=6, seqfrac=1, seed=11)          # H=10, one row tile
    flow_case("flow_odd", D=9, NF=2, B=5, seqfrac=0.25, seed=12)         # odd ndim: chunk gives 5 + 4, H=18
    flow_case("flow_bigbatch", D=12, NF=2, B=80, seqfrac=0.5, seed=13)   # batch spans two 64-row tiles, H=12
    flow_case("flow_nobn", D=16, NF=2, B=7, seqfrac=1, seed=14, batch_norm=False)
    flow_case("flow_toy", D=2, NF=3, B=33, seqfrac=1 / 16, seed=15)      # DPI_2Dtoydist shape class (H=16)
    closure_map_case("closure_small", n_stations=5, n_scans=8, seed=3, npix=8)
    closure_map_case("closure_eht", n_stations=7, n_scans=100, seed=190, npix=32)
    interferometry_case("interferometry_small", npix=8, NF=2, B=6, seed=3)
    mri_case("mri_small", npix=8, NF=2, B=6, seed=5)
    dpix_case("dpix_small", npix=16, NF=2, B=12, seed=9, alpha=0.95)
    dpix_case("dpix_kl", npix=16, NF=2, B=12, seed=10, alpha=1.0)
    glow_case("glow_small", npix=8, n_flow=1, n_block=2, B=3, seed=21)
