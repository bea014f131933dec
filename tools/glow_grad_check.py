"""Which Glow parameter gradients are non-finite / large at the reference's shape (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dpi_b200.generative_model.glow_model import Glow, calc_z_shapes
npix, n_flow, n_block, B = 32, int(os.environ.get("GLOW_NFLOW", 16)), 4, int(os.environ.get("B", 64))
torch.manual_seed(0); np.random.seed(0)
m = Glow(1, n_flow, n_block).cuda().train()
zs = [torch.randn(B, *s, device="cuda") for s in calc_z_shapes(1, npix, n_flow, n_block)]
c = torch.randn(B, 1, npix, npix, device="cuda")
for it in range(3):
    x, ld = m.reverse(zs)
    loss = (x * c).mean() - 1e-3 * ld.mean()
    m.zero_grad()
    loss.backward()
    bad = [(k, float(p.grad.abs().max())) for k, p in m.named_parameters() if not torch.isfinite(p.grad).all()]
    top = sorted(((float(p.grad.abs().max()), k) for k, p in m.named_parameters() if torch.isfinite(p.grad).all()), reverse=True)[:4]
    print(it, "loss", float(loss), "x finite", bool(torch.isfinite(x).all()), "non-finite grads:", bad[:6], "largest:", top)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(p.grad.sign(), alpha=-1e-5)
