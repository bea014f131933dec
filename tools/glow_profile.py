"""Kernel-time table of one Glow training step at the reference's shape (torch.profiler / CUPTI; diagnosis only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from dpi_b200.generative_model.glow_model import Glow, calc_z_shapes
npix, n_flow, n_block, B = 32, 16, 4, 64
torch.manual_seed(0); np.random.seed(0)
m = Glow(1, n_flow, n_block).cuda().train()
zs = [torch.randn(B, *s, device="cuda") for s in calc_z_shapes(1, npix, n_flow, n_block)]
c = torch.randn(B, 1, npix, npix, device="cuda")
opt = torch.optim.Adam(m.parameters(), lr=1e-7)
def step():
    x, ld = m.reverse(zs)
    loss = ((x - c) ** 2).mean() - 1e-3 * ld.mean()
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot/1e3:.2f} ms over {sum(e.count for e in rows)} kernels")
for e in rows[:22]:
    print(f"{e.key[:70]:70s} n={e.count:5d} total {e.device_time_total/1e3:8.2f} ms avg {e.device_time_total/max(1,e.count):7.1f} us")
