"""The reference's own training step, assembled from the UNMODIFIED reference modules (bench.py --impl reference).

HeSunPU/DPI is pure Python/PyTorch with no entry point that runs without ehtim / uvfits files, so the loop bodies of
DPI_interferometry.py:193-235 and DPI_MRI.py:162-207 are restated here around the reference's own classes and
functions (generative_model.realnvpfc_model.RealNVP, interferometry_helpers.eht_observation_pytorch / Loss_*,
MRI_helpers.Loss_*; torch.optim.Adam, nn.utils.clip_grad_norm_) exactly as tests/golden/make_golden.py does for the
fixtures. The modules are imported from the reference install `baseline/_ref/DPItorch` (git-ignored; written by
`__graft_entry__.build()` from /root/reference where that exists, shipped to the GPU box with the snapshot), or from
/root/reference itself. None of dpi_b200's kernels, models or trainers is on this path: only the synthetic workload
tables (the same inputs both arms see) come from dpi_b200.workloads.
"""
from __future__ import annotations

import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CANDIDATES = [os.path.join(HERE, "_ref", "DPItorch"), "/root/reference/DPItorch"]


def reference_root():
    for p in CANDIDATES:
        if os.path.isfile(os.path.join(p, "generative_model", "realnvpfc_model.py")):
            return p
    return None


def _modules():
    root = reference_root()
    if root is None:
        raise RuntimeError("reference modules not installed (baseline/_ref is written by __graft_entry__.build())")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_shim
    ref_shim.REF_ROOT = root
    return ref_shim.import_reference()


class ReferenceStep:
    """One optimisation step of the reference loop body per call: step(z [B, D]) -> loss (python float)."""

    def __init__(self, kind, wl, n_flow, lr, clip, device="cpu", modules=None, model_form="realnvp", n_block=4):
        """device="cuda": the same stock-PyTorch eager path on the GPU (`.to(device)`, DPI_interferometry.py:91-92,162-168)
        - the "reference on the same B200" context number; device="cpu": the CPU baseline.
        modules: dict(realnvp=, interferometry=, mri=) to run the SAME loop body over another implementation of the
        reference's module surface (bench.py passes the dpi_b200 drop-in modules to time the autograd surface a
        reference script would use); default: the unmodified reference."""
        R = modules if modules is not None else _modules()
        npix = wl["npix"]
        dev = torch.device(device)
        wl = {k: (v.to(dev) if torch.is_tensor(v) and k in ("prior", "cphase_true", "logcamp_true", "mask", "kspace_true")
                  else v) for k, v in wl.items()}
        self.kind, self.wl, self.npix, self.clip, self.device = kind, wl, npix, clip, dev
        self.model_form = model_form
        if model_form == "glow":
            self.gen = R["glow"].Glow(1, n_flow, n_block, affine=True, conv_lu=True).to(dev)      # :129-136
            self.z_shapes = R["glow"].calc_z_shapes(1, npix, n_flow, n_block)
        else:
            self.gen = R["realnvp"].RealNVP(npix * npix, n_flow).to(dev)  # DPI_interferometry.py:127,164 / DPI_MRI.py:118
        self.log_scale = torch.nn.Parameter(torch.tensor([wl["log_scale_init"]], dtype=torch.float32, device=dev))
        self.params = list(self.gen.parameters()) + [self.log_scale]
        self.opt = torch.optim.Adam(self.params, lr=lr)                   # :172 / :146
        if kind == "interferometry":
            IH = R["interferometry"]
            self.IH = IH
            self.eht = IH.eht_observation_pytorch(npix, types.SimpleNamespace(to=lambda **k: None), wl["dft_mat"],
                                                  wl["ktraj"], wl["pulsefac"], wl["cphase_ind"], wl["cphase_sign"],
                                                  wl["camp_ind"], dev, ttype="direct")
            self.L_center = IH.Loss_center(dev, center=npix / 2 - 0.5, dim=npix)
            self.L_flux = IH.Loss_flux(wl["flux"])
            self.L_cp = IH.Loss_angle_diff(wl["sigma_cp"].cpu().numpy(), dev)
            self.L_ca = IH.Loss_logca_diff2(wl["sigma_ca"].cpu().numpy(), dev)
        else:
            self.MH = R["mri"]
            self.L_k = self.MH.Loss_kspace_diff2(wl["sigma"])

    def step(self, z):
        """z: host tensor, as DPI_interferometry.py:193 draws it (`torch.randn(B, D).to(device)`)."""
        npix, w, wl = self.npix, self.wl["w"], self.wl
        z = [zz.to(self.device) for zz in z] if self.model_form == "glow" else z.to(self.device)
        img_samp, logdet = self.gen.reverse(z)                                     # :201 / :197-200
        img_samp = img_samp.reshape((-1, npix, npix))
        img = torch.nn.Softplus()(img_samp) * torch.exp(self.log_scale)            # :205-208
        det_softplus = torch.sum(img_samp - torch.nn.Softplus()(img_samp), (1, 2))
        logdet = logdet + det_softplus + self.log_scale * npix * npix              # :209-211
        if self.kind == "interferometry":
            IH = self.IH
            vis, visamp, cphase, logcamp = self.eht(img)                           # :213
            loss_data = w["camp"] * self.L_ca(wl["logcamp_true"], logcamp) + w["cphase"] * self.L_cp(wl["cphase_true"], cphase)
            loss_prior = w["mem"] * IH.Loss_cross_entropy(wl["prior"], img) + w["flux"] * self.L_flux(img) + \
                w["tsv"] * IH.Loss_TSV(img) + w["center"] * self.L_center(img) + w["l1"] * IH.Loss_l1(img)   # :214-227
        else:
            kpred = torch.view_as_real(torch.fft.fft2(img, norm="ortho"))         # DPI_MRI.py:70-74 (removed torch.fft API)
            loss_data = self.L_k(wl["kspace_true"], kpred * wl["mask"]) / float(wl["mask"].mean())   # :181-182
            loss_prior = w["tv"] * self.MH.Loss_TV(img)                            # :185
        loss = torch.mean(loss_data) + torch.mean(loss_prior) - w["logdet"] * torch.mean(logdet)     # :229 / :193
        self.opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, self.clip)                     # :233
        self.opt.step()
        return float(loss.detach())
