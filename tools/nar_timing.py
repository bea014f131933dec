"""Narrow-flow pass (flow_nar.cu) vs the tcgen05 GEMM path on the alpha-DPI flow RealNVP(18, 16, seqfrac 1/16), batch 2048,
and on the 2-D toy flow (batch 512): CUDA-event time of reverse and of reverse + backward (autograd surface)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpi_b200 import _lib  # noqa: E402
from dpi_b200.generative_model.realnvpfc_model import RealNVP  # noqa: E402
from oracle import dpi_oracle as O  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    lib = _lib.load()
    for D, NF, sf, B in ((18, 16, 1 / 16, 2048), (2, 16, 1 / 128, 512), (18, 16, 1 / 16, 8192)):
        sd = O.init_state(D, NF, seqfrac=sf, seed=3, randomize=True)
        res = {}
        for mode in ("1", "0"):
            os.environ["DPI_B200_NAR"] = mode
            m = RealNVP(D, NF, seqfrac=sf)
            m.load_state_dict(sd)
            m = m.cuda()
            torch.manual_seed(0)
            z = torch.randn(B, D, device="cuda", requires_grad=True)
            cx = torch.randn(B, D, device="cuda")
            act = lib.dpi_flow_nar_active(m._handle(z.device), B)

            def fwd():
                with torch.no_grad():
                    return m.reverse(z)

            def fwdbwd():
                x, ld = m.reverse(z)
                ((x * cx).sum() + ld.sum()).backward()

            def evalfwd():
                with torch.no_grad():
                    return m.reverse(z)

            n0 = _lib.launch_count()
            fwd()
            lf = _lib.launch_count() - n0
            t_f, t_fb = timed(fwd), timed(fwdbwd)
            m.eval()
            t_e = timed(evalfwd)
            m.train()
            x, ld = fwd()
            res[mode] = (x.cpu(), ld.cpu())
            if act and B == 2048:
                import numpy as np
                hnd = m._handle(z.device)
                for label, fn_ in (("train", fwd), ("eval", None)):
                    if fn_ is None:
                        m.eval()
                    _lib.check(lib.dpi_flow_debug_timing(hnd, 1))
                    fwd(); torch.cuda.synchronize()
                    stt = np.zeros(2 * 4096, dtype=np.uint64); ops = np.zeros(4096, dtype=np.int32)
                    lib.dpi_flow_debug_read(hnd, stt.ctypes.data, ops.ctypes.data, 4096)
                    _lib.check(lib.dpi_flow_debug_timing(hnd, 0))
                    S_ = 2 * NF
                    s_ = stt[:16 * S_].astype(np.int64).reshape(S_, 16)[1:-1]
                    seq = [0, 1, 2, 3, 5, 6, 7, 8, 9, 10]
                    names = ["wait operands", "issue small prefetch", "net.0", "save+BN1", "net.3", "issue W2", "save+BN2", "net.6", "tail"]
                    print(f"   [{label}] cycles per stage (CTA 0):", ", ".join(f"{n} {np.median(s_[:, b_] - s_[:, a_]):.0f}" for n, a_, b_ in zip(names, seq[:-1], seq[1:])),
                          f"| stage {np.median(np.diff(stt[:16 * S_].astype(np.int64).reshape(S_, 16)[:, 0])):.0f}")
                m.train()
            print(f"D={D} H={m.hidden} B={B} DPI_B200_NAR={mode} active={act}: reverse {t_f * 1e3:.0f} us ({lf} launches), "
                  f"eval reverse {t_e * 1e3:.0f} us, reverse+backward {t_fb * 1e3:.0f} us", flush=True)
        x1, l1 = res["1"]; x0, l0 = res["0"]
        print("   max rel diff x", float((x1 - x0).abs().max() / x0.abs().max()), "logdet", float((l1 - l0).abs().max() / l0.abs().max()))


if __name__ == "__main__":
    main()
