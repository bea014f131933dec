import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from conftest import rel_err
from oracle import dpi_oracle as O
from dpi_b200.generative_model.realnvpfc_model import RealNVP
D, NF, B, seqfrac = int(os.environ.get("TD", 2)), int(os.environ.get("TNF", 16)), int(os.environ.get("TB", 512)), 1 / 128
if D != 2: seqfrac = float(os.environ.get("TSEQ", 1))
sd = O.init_state(D, NF, seqfrac=seqfrac, seed=31, randomize=True)
torch.manual_seed(8)
z = torch.randn(B, D)
cx, cl = torch.randn(B, D), torch.randn(B)
sdo = {k: v.detach().clone() for k, v in sd.items()}
for k in O.param_keys(sdo):
    sdo[k].requires_grad_(True)
zo = z.clone().requires_grad_(True)
xo, ldo = O.realnvp_reverse(sdo, zo, train=True)
mode = os.environ.get("TMODE", "toy")
def objective(x, ld, cx, cl):
    if mode == "toy":
        xs = 2 * torch.sigmoid(x) - 1
        ds = torch.sum(-x - 2 * torch.nn.Softplus()(-x), -1)
        return torch.mean(-O.toy_logprob(1, xs[:, 0], xs[:, 1]) - (ld + ds))
    return (x * cx).sum() + (ld * cl).sum()
objective(xo, ldo, cx, cl).backward()
m = RealNVP(D, NF, seqfrac=seqfrac)
m.load_state_dict({k: v.detach().clone() for k, v in sd.items()})
m = m.cuda()
zg = z.cuda().requires_grad_(True)
x, ld = m.reverse(zg)
objective(x, ld, cx.cuda(), cl.cuda()).backward()
views = dict(m.named_grad_views())
print("x", rel_err(x, xo), "ld", rel_err(ld, ldo), "g_z", rel_err(zg.grad, zo.grad))
for blk in (NF - 1, 0):
    for k, gv in views.items():
        if k.startswith(f"blocks.{blk}.") and ("coupling2" in k or "actnorm2" in k):
            print("   %.3e  %-44s |ref|max %.3e" % (rel_err(gv, sdo[k].grad), k, float(sdo[k].grad.abs().max())))
worst = max((rel_err(gv, sdo[k].grad), k) for k, gv in views.items() if "actnorm" not in k)
print("worst non-actnorm", worst)
k = f"blocks.{NF-1}.coupling2.net.0.bias"
print(k, views[k][:6].cpu().tolist(), sdo[k].grad[:6].tolist())
k = f"blocks.{NF-1}.actnorm2.loc"
print(k, views[k].cpu().tolist(), sdo[k].grad.tolist())
