"""One training-mode reverse pass + backward of a C4-shaped flow (D=4096, H=512, batch 1024) with few flows, for
`ncu --metrics gpu__time_duration.sum` launch lists of the tensor-core path (tools/ only; not part of the product)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dpi_b200.generative_model.realnvpfc_model import RealNVP  # noqa: E402

D = int(os.environ.get("PROF_D", 4096))
NF = int(os.environ.get("PROF_NF", 2))
B = int(os.environ.get("PROF_B", 1024))
reps = int(os.environ.get("PROF_REPS", 1))
torch.manual_seed(0)
m = RealNVP(D, NF).cuda()
with torch.no_grad():
    m.flat.add_(0.01 * torch.randn_like(m.flat))
z = torch.randn(B, D, device="cuda", requires_grad=True)
for it in range(reps + 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, ld = m.reverse(z)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    (x.sum() + ld.sum()).backward()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"iter {it}: fwd {1e3 * (t1 - t0):.3f} ms  bwd {1e3 * (t2 - t1):.3f} ms (wall, eager, {2 * NF} stages)")
