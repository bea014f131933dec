"""Where does the C4-shape gradient error come from? Compares, on the same weights and z (D=4096, H=512, B=1024,
4 flows): CPU oracle fp32, CPU oracle fp64, the tcgen05 (3xTF32) path and the fp32 phase-program path.
Run on the GPU box: python tools/c4_parity_probe.py"""
import os
import sys
import subprocess
import json

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def gpu_grads(out):
    from oracle import dpi_oracle as O
    from dpi_b200.trainer import InterferometryTrainer
    from dpi_b200.workloads import interferometry_workload
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    wl = interferometry_workload(npix=64)
    D, NF, B = 4096, 4, 1024
    sd = O.init_state(D, NF, seed=21, randomize=True)
    m = RealNVP(D, NF).cuda()
    m.load_state_dict(sd)
    tr = InterferometryTrainer(m, wl, B, lr=1e-4, clip=1e-3, use_graph=False)
    torch.manual_seed(9)
    z = torch.randn(B, D)
    tr.loss_and_grad(z.cuda())
    torch.cuda.synchronize()
    g = {k: v.detach().cpu().clone() for k, v in tr.flow.named_grad_views(tr.grads)}
    g["_x"] = tr.x.cpu().clone(); g["_img"] = tr.img.cpu().clone().view(B, -1); g["_gx"] = tr.g_x.cpu().clone()
    g["_gimg"] = (tr.g_img_data + tr.g_img_prior).cpu().clone(); g["_gimg_data"] = tr.g_img_data.cpu().clone()
    torch.save(g, out)


def cpu_grads(dtype):
    from oracle import dpi_oracle as O
    from dpi_b200.workloads import interferometry_workload
    from test_gpu_trainer import _consts
    wl = interferometry_workload(npix=64)
    D, NF, B = 4096, 4, 1024
    sd = O.init_state(D, NF, seed=21, randomize=True)
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    c = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else
             (v.to(torch.complex128) if torch.is_tensor(v) and v.is_complex() and dtype == torch.float64 else v))
         for k, v in _consts(wl).items()}
    ls = torch.tensor([wl["log_scale_init"]], dtype=dtype)
    ot = O.OracleTrainer(sd, ls, O.interferometry_loss, c, lr=1e-4, clip=1e-3)
    torch.manual_seed(9)
    z = torch.randn(B, D).to(dtype)
    loss, terms = O.interferometry_loss(sd, ls, z, c)
    terms["x"].retain_grad(); terms["img"].retain_grad()
    loss.backward()
    g = {k: sd[k].grad.detach().clone() for k in ot.keys}
    g["_x"] = terms["x"].detach(); g["_img"] = terms["img"].detach().reshape(B, -1); g["_gx"] = terms["x"].grad
    g["_gimg"] = terms["img"].grad.reshape(B, -1)
    return g


def main():
    if len(sys.argv) > 1:
        gpu_grads(sys.argv[1])
        return
    res = {}
    for name, env in (("tc", {}), ("fp32", {"DPI_B200_FLOW_TC": "0"})):
        e = dict(os.environ)
        e.update(env)
        subprocess.check_call([sys.executable, __file__, f"/tmp/g_{name}.pt"], env=e)
        res[name] = torch.load(f"/tmp/g_{name}.pt")
    res["o32"] = cpu_grads(torch.float32)
    res["o64"] = cpu_grads(torch.float64)

    outp0 = {}
    for k in ("_x", "_img", "_gx", "_gimg"):
        r = res["o64"][k].double()
        outp0[k] = {n: [float((res[n][k].double() - r).abs().max() / r.abs().max()), float((res[n][k].double() - r).norm() / r.norm())]
                    for n in ("tc", "fp32", "o32")}
    print(json.dumps(outp0, indent=1))
    for d in res.values():
        for k in [k for k in d if k.startswith("_")]:
            d.pop(k)

    def vec(d):
        return torch.cat([d[k].double().reshape(-1) for k in sorted(res["o64"])])
    v = {k: vec(d) for k, d in res.items()}
    outp = {"norm": {k: float(x.norm()) for k, x in v.items()}}
    for a in ("tc", "fp32", "o32"):
        for b in ("o64", "o32"):
            if a == b:
                continue
            outp[f"{a}_vs_{b}"] = {"l2": float((v[a] - v[b]).norm() / v[b].norm()),
                                   "max": float((v[a] - v[b]).abs().max() / v[b].abs().max()),
                                   "norm_rel": float(abs(v[a].norm() - v[b].norm()) / v[b].norm())}
    worst = {}
    for k in sorted(res["o64"]):
        r = res["o64"][k].double()
        worst[k] = {n: float((res[n][k].double() - r).abs().max() / r.abs().max().clamp_min(1e-300)) for n in ("tc", "fp32", "o32")}
    top = sorted(worst.items(), key=lambda kv: -kv[1]["tc"])[:8]
    outp["worst_tensors_vs_o64"] = top
    print(json.dumps(outp, indent=1))


if __name__ == "__main__":
    main()
