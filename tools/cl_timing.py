"""Cluster-resident tcgen05 flow pass vs the fp32 small-batch kernel on C1's flow (RealNVP(1024, 16), batch 64):
CUDA-event time per pass and the per-stage phase stamps of CTA 0 (debug aid): python tools/cl_timing.py"""
import os
import sys

import numpy as np
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
    D, NF, B = 1024, int(os.environ.get("NF", "16")), int(os.environ.get("B", "64"))
    lib = _lib.load()
    sd = O.init_state(D, NF, seqfrac=4, seed=3, randomize=True)
    res = {}
    for mode in ("1", "0"):
        os.environ["DPI_B200_CL"] = mode
        m = RealNVP(D, NF)
        m.load_state_dict(sd)
        m = m.cuda()
        torch.manual_seed(0)
        z = torch.randn(B, D, device="cuda", requires_grad=True)
        cx = torch.randn(B, D, device="cuda")
        h = m._handle(z.device)
        act = lib.dpi_flow_cl_active(h, B)

        def fwd():
            with torch.no_grad():
                return m.reverse(z)

        def fwdbwd():
            x, ld = m.reverse(z)
            ((x * cx).sum() + ld.sum()).backward()

        t_f = timed(fwd)
        t_fb = timed(fwdbwd)
        x, ld = fwd()
        res[mode] = (x.cpu(), ld.cpu())
        print(f"DPI_B200_CL={mode} active={act}: reverse {t_f * 1e3:.1f} us, reverse+backward (autograd) {t_fb * 1e3:.1f} us")
        if mode == "1" and act:
            _lib.check(lib.dpi_flow_debug_timing(h, 1))
            fwd(); torch.cuda.synchronize()
            st = np.zeros(2 * 4096, dtype=np.uint64)
            ops = np.zeros(4096, dtype=np.int32)
            lib.dpi_flow_debug_read(h, st.ctypes.data, ops.ctypes.data, 4096)
            S = 2 * NF
            T = st[:16 * 32 * 16].astype(np.int64).reshape(16, 32, 16)[:, :S, :]      # [cta][stage][slot], ns
            names = ["rows landed -> net.0 operand", "net.0 MMA", "net.0 partial push + wait", "BatchNorm 1 epilogue",
                     "net.3 MMA + partial push + wait", "BatchNorm 2 epilogue + slice push", "net.6 (wait slices + MMA)",
                     "TMEM -> tile", "coupling tail + row push", "(to next stage)"]
            nxt = np.roll(T[:, :, 0], -1, axis=1)
            seq = np.concatenate([T[:, :, :10], nxt[:, :, None]], axis=2)[:, 1:-1, :]
            d = np.diff(seq, axis=2)                                                    # [cta][stage][phase]
            print("per-stage ns (median over stages), CTA 0 / min / max over CTAs; and skew = spread over CTAs of the phase END time:")
            for i, nme in enumerate(names):
                per_cta = np.median(d[:, :, i], axis=1)
                skew = np.median(seq[:, :, i + 1].max(0) - seq[:, :, i + 1].min(0))
                print(f"   {nme:36s} {per_cta[0]:7.0f} {per_cta.min():7.0f} {per_cta.max():7.0f}   skew {skew:6.0f}")
            print(f"   stage total {np.median(d.sum(2), axis=1)[0]:.0f} ns")
            f = T[0, 1:-1, :]
            md = lambda a_, b_: np.median(f[:, a_] - f[:, b_])
            print("   detail CTA 0 (ns): stage top -> rows landed %.0f, -> operand written + signalled %.0f | BN1: reduce+stats+stores %.0f, operand+signal %.0f | "
                  "BN2: reduce+stats+stores %.0f, slice stores %.0f, fence+barrier %.0f, multicast issue %.0f" % (
                      md(10, 0), md(1, 10), md(15, 3), md(4, 15), md(11, 5), md(12, 11), md(13, 12), md(6, 13)))
            _lib.check(lib.dpi_flow_debug_timing(h, 0))
    if len(res) == 2:
        x1, l1 = res["1"]; x0, l0 = res["0"]
        print("max rel diff x", float((x1 - x0).abs().max() / x0.abs().max()), "logdet", float((l1 - l0).abs().max() / l0.abs().max()))


if __name__ == "__main__":
    main()
