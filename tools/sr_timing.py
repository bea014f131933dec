"""Sample-resident small-batch pass (flow_sr.cu) vs the feature-split fp32 kernel (flow_small.cu) on C1's flow
(RealNVP(1024, 16), batch 64): CUDA-event time of reverse, and reverse + backward through the C ABI (no autograd
overhead): python tools/sr_timing.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpi_b200 import _lib  # noqa: E402
from dpi_b200.generative_model.realnvpfc_model import RealNVP  # noqa: E402
from oracle import dpi_oracle as O  # noqa: E402


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    D, NF, B = int(os.environ.get("D", "1024")), int(os.environ.get("NF", "16")), int(os.environ.get("B", "64"))
    lib = _lib.load()
    sd = O.init_state(D, NF, seqfrac=4, seed=3, randomize=True)
    for mode in ("1", "0"):
        os.environ["DPI_B200_SR"] = mode
        m = RealNVP(D, NF)
        m.load_state_dict(sd)
        m = m.cuda()
        dev = torch.device("cuda", 0)
        h = m._handle(dev)
        torch.manual_seed(0)
        z = torch.randn(B, D, device=dev)
        x = torch.empty(B, D, device=dev)
        ld = torch.empty(B, device=dev)
        gx = torch.randn(B, D, device=dev)
        gld = torch.randn(B, device=dev)
        gz = torch.empty(B, D, device=dev)
        grads = torch.zeros_like(m.flat.data)
        ws = torch.empty(lib.dpi_flow_workspace_bytes(h, B), dtype=torch.uint8, device=dev)
        st = _lib.stream_ptr()
        P, R = m.flat.data, m.bn_running

        def fwd(train=1):
            _lib.check(lib.dpi_flow_run(h, 0, train, B, P.data_ptr(), R.data_ptr(), z.data_ptr(), x.data_ptr(), ld.data_ptr(),
                                        ws.data_ptr(), ws.numel(), st), "run")

        def bwd():
            _lib.check(lib.dpi_flow_backward(h, B, P.data_ptr(), grads.data_ptr(), z.data_ptr(), gx.data_ptr(), gld.data_ptr(),
                                             gz.data_ptr(), ws.data_ptr(), ws.numel(), st), "bwd")

        def both():
            fwd(); bwd()

        print(f"DPI_B200_SR={mode} sr_active={lib.dpi_flow_sr_active(h, B)}: reverse train {timed(fwd) * 1e3:7.1f} us, "
              f"eval {timed(lambda: fwd(0)) * 1e3:7.1f} us, backward {timed(bwd) * 1e3:7.1f} us, both {timed(both) * 1e3:7.1f} us")
        if mode == "1" and lib.dpi_flow_sr_active(h, B):
            import numpy as np
            S_ = 2 * NF

            def timeline(fn, names, slots):
                _lib.check(lib.dpi_flow_debug_timing(h, 1))
                fn(); torch.cuda.synchronize()
                stt = np.zeros(2 * 4096, dtype=np.uint64); ops = np.zeros(4096, dtype=np.int32)
                lib.dpi_flow_debug_read(h, stt.ctypes.data, ops.ctypes.data, 4096)
                _lib.check(lib.dpi_flow_debug_timing(h, 0))
                s_ = stt[:16 * S_].astype(np.int64).reshape(S_, 16)
                mid = s_[1:-1]
                return ", ".join(f"{n} {np.median(mid[:, b_] - mid[:, a_]):.0f}" for n, a_, b_ in zip(names, slots[:-1], slots[1:])) + \
                    f" | stage {np.median(np.diff(s_[:, 0])):.0f}"
            print("   fwd train cycles/stage:", timeline(fwd, ["wait", "prefetch+gather", "net.0", "reduce+save", "cluster sync", "BN1", "net.3", "cl sync", "BN2", "net.6", "tail"],
                                                         [0, 1, 2, 3, 4, 4, 5, 6, 7, 8, 9, 10]))
            print("   bwd cycles/stage:", timeline(bwd, ["wait", "issue next", "tail bwd", "dgrad W3", "BN2 bwd", "dgrad W2", "BN1 bwd", "dgrad W1", "scatter"],
                                                   [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]))
        del m


if __name__ == "__main__":
    main()
