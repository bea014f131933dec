"""Throughput of Glow.reverse (sampling direction) and of a training step (reverse + backward + Adam) at the reference's
`--model_form glow` shape: Glow(1, 16, 4), npix 32."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dpi_b200 import _lib  # noqa: E402
from dpi_b200.generative_model.glow_model import Glow, calc_z_shapes  # noqa: E402

npix, n_flow, n_block = 32, int(os.environ.get("GLOW_NFLOW", 16)), 4
torch.manual_seed(0)
np.random.seed(0)
m = Glow(1, n_flow, n_block).cuda()
print("parameters:", sum(p.numel() for p in m.parameters()))
for B in (32, 64):
    zs = [torch.randn(B, *s, device="cuda") for s in calc_z_shapes(1, npix, n_flow, n_block)]
    for mode in ("train", "eval"):
        m.train(mode == "train")
        torch.set_grad_enabled(False)
        for _ in range(2):
            x, ld = m.reverse(zs)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            x, ld = m.reverse(zs)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"B={B} {mode}: {1e3 * dt:.2f} ms per reverse pass, {B / dt:.0f} samples/s, "
              f"{(_lib.launch_count() - n0) // reps} library launches, finite={bool(torch.isfinite(x).all())}")
    # training step: the objective of the fixture, gradients of all 23.7 M parameters, torch.optim.Adam
    torch.set_grad_enabled(True)
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-7)
    c = torch.randn(B, 1, npix, npix, device="cuda")

    def step():
        x, ld = m.reverse(zs)
        loss = ((x - c) ** 2).mean() - 1e-3 * ld.mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(5):
        loss = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"B={B} training step: {1e3 * dt:.2f} ms, {B / dt:.0f} samples/s, {(_lib.launch_count() - n0) // 5} library launches, "
          f"loss finite={bool(torch.isfinite(loss))}")

# the same step captured into ONE CUDA graph (static latents, Adam(capturable=True)): what is left is kernel time
try:
    B = 64
    m2 = Glow(1, n_flow, n_block).cuda().train()
    zs2 = [torch.randn(B, *s, device="cuda") for s in calc_z_shapes(1, npix, n_flow, n_block)]
    c2 = torch.randn(B, 1, npix, npix, device="cuda")
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-7, capturable=True)

    def step2():
        x, ld = m2.reverse(zs2)
        loss = ((x - c2) ** 2).mean() - 1e-3 * ld.mean()
        opt2.zero_grad(set_to_none=False)
        loss.backward()
        opt2.step()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step2()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_static = step2()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        graph.replay()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"B={B} training step as one CUDA graph: {1e3 * dt:.2f} ms, {B / dt:.0f} samples/s, loss finite={bool(torch.isfinite(loss_static))}")
except Exception as e:  # noqa: BLE001
    print("graph capture of the training step failed:", type(e).__name__, str(e)[:300])
