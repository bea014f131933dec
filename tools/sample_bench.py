"""Throughput of the posterior sampling path (Gen_samples, DPIx_interferometry.py:433-474) on the C3 workload:
eval-mode flow -> sigmoid -> renderer -> closure operator -> importance weights, forward only."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dpi_b200.generative_model.realnvpfc_model import RealNVP  # noqa: E402
from dpi_b200.trainer import DPIxTrainer  # noqa: E402
from dpi_b200.workloads import dpix_workload  # noqa: E402

B = int(os.environ.get("SAMPLE_B", 8192))
wl = dpix_workload(npix=64)
torch.manual_seed(0)
flow = RealNVP(wl["nparams"], 16, seqfrac=1 / 16)
tr = DPIxTrainer(flow, wl, B, device="cuda")
for _ in range(3):
    tr.step(torch.randn(B, wl["nparams"]).cuda())       # a few training steps so the running statistics are non-trivial
for n in (2, 10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = tr.sample_posterior(n_concat=n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"sample_posterior: {n} x {B} samples in {1e3 * dt:.1f} ms = {n * B / dt / 1e6:.2f} M samples/s; "
          f"accepted {int(res['accepted'].sum())}; weights sum {float(res['weights'].sum()):.4f}")
