"""Per-phase timeline of the flow program on the GPU (debug aid): python tools/phase_timing.py [c1|c2]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload  # noqa: E402
from dpi_b200 import _lib  # noqa: E402
from dpi_b200.generative_model.realnvpfc_model import RealNVP  # noqa: E402
from dpi_b200.trainer import InterferometryTrainer, MRITrainer  # noqa: E402

OPS = {1: "FWD_HID", 2: "FWD_OUT", 3: "BN_STATS", 4: "LOGDET", 5: "TRANSPOSE", 6: "B0", 7: "DGRAD_HID", 8: "WGRAD",
       9: "DGRAD_IN", 10: "BNBWD", 11: "BWD_FINAL",
       20: "s.T_IN", 21: "s.F1", 22: "s.F2", 23: "s.F3", 24: "s.F_FIN", 25: "s.B0", 26: "s.DG2+WG3", 27: "s.DG1+WG2",
       28: "s.DGIN+WG1", 29: "s.B_FIN"}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c1"
    flush = len(sys.argv) > 2 and sys.argv[2] == "flush"
    wl, kind, npix, nf, B, lr, clip, desc = build_workload(name)
    lib = _lib.load()
    torch.manual_seed(0)
    flow = RealNVP(npix * npix, nf)
    T = InterferometryTrainer if kind == "interferometry" else MRITrainer
    tr = T(flow, wl, B, lr=lr, clip=clip, device="cuda", use_graph=False)
    for _ in range(3):
        tr.sample_z(); tr.step()
    _lib.check(lib.dpi_flow_debug_timing(tr.handle, 1))
    big = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    for which in ("fwd", "bwd"):
        tr.sample_z()
        if flush:
            big.fill_(1)
        torch.cuda.synchronize()
        if which == "fwd":
            _lib.check(lib.dpi_flow_run(tr.handle, 0, 1, B, tr.params.data_ptr(), tr.flow.bn_running.data_ptr(),
                                        tr.z.data_ptr(), tr.x.data_ptr(), tr.logdet_flow.data_ptr(), tr.ws_flow.data_ptr(),
                                        tr.ws_flow.numel(), _lib.stream_ptr()))
        else:
            tr.g_x.normal_()
            _lib.check(lib.dpi_flow_backward(tr.handle, B, tr.params.data_ptr(), tr.grads.data_ptr(), tr.z.data_ptr(),
                                             tr.g_x.data_ptr(), tr.g_det.data_ptr(), None, tr.ws_flow.data_ptr(),
                                             tr.ws_flow.numel(), _lib.stream_ptr()))
        st = np.zeros(2 * 4096, dtype=np.uint64)
        ops = np.zeros(4096, dtype=np.int32)
        n = lib.dpi_flow_debug_read(tr.handle, st.ctypes.data, ops.ctypes.data, 4096)
        fine = st[1024:1024 + 8 * n].astype(np.int64).reshape(n, 8)
        st = st[:2 * n].astype(np.int64).reshape(n, 2)
        start, done = st[:, 0], st[:, 1]
        work = done - start
        wait = np.append(start[1:] - done[:-1], 0)
        print(f"== {name} {which}: {n} phases, total {(done[-1] - start[0]) / 1e3:.1f} us, "
              f"work(cta0) {work.sum() / 1e3:.1f} us, barrier/wait {wait.sum() / 1e3:.1f} us")
        by = {}
        for i in range(n):
            key = (OPS.get(ops[i] // 100, "?"), (ops[i] // 10) % 10, ops[i] % 10)
            by.setdefault(key, []).append((work[i], wait[i]))
        for key, v in by.items():
            v = np.array(v)
            print(f"   {key[0]:<12} layer={key[1]} two_subops={key[2]}  n={len(v):3d}  work {v[:, 0].mean() / 1e3:7.2f} us  "
                  f"wait-after {v[:, 1].mean() / 1e3:7.2f} us")
        if fine.any():
            ok = fine[:, 0] != 0
            if ok.sum() > 2:
                i0, i1 = np.flatnonzero(ok)[0], np.flatnonzero(ok)[-1]
                mhz = (fine[i1, 0] - fine[i0, 0]) / max(1, (start[i1] - start[i0])) * 1e3
                print(f"   SM clock during the pass (clock64 / globaltimer): {mhz:.0f} MHz")
            # small-batch kernel: stamps 0 phase start, 1 prefetched weights landed, 2 activation rows landed,
            # 3 micro-GEMM done, 4 k-part (+cluster) reduction done, 5 epilogue done, 6 arrive + next prefetch issued,
            # 7 shadow (weight gradients) done; the next phase's stamp 0 closes the barrier wait
            names = ["pf-wait", "act-load", "gemm", "reduce", "epilogue", "arrive+pf", "shadow", "barrier"]
            byk = {}
            for i in range(n - 1):
                f = fine[i]
                if f[0] == 0:
                    continue
                d = [f[1] - f[0], f[2] - f[1], f[3] - f[2], f[4] - f[3], f[5] - f[4], f[6] - f[5], f[7] - f[6],
                     fine[i + 1][0] - f[7]]
                if f[1] == 0 or f[4] == 0:      # phases without a GEMM tile on CTA 0
                    d = [0, 0, 0, 0, f[5] - f[0], f[6] - f[5], f[7] - f[6], fine[i + 1][0] - f[7]]
                byk.setdefault(OPS.get(ops[i] // 100, "?"), []).append(d)
            for k, v in byk.items():
                v = np.array(v[2:] if len(v) > 4 else v, dtype=np.float64) / 1.965e3   # SM cycles at 1965 MHz -> us
                print(f"   {k:<12} " + "  ".join(f"{nm} {x:5.2f}" for nm, x in zip(names, v.mean(0))) + f"  | total {v.mean(0).sum():5.2f} us")
        print("   first 12 phases (work, wait) us:", [(round(work[i] / 1e3, 1), round(wait[i] / 1e3, 1)) for i in range(min(12, n))])


if __name__ == "__main__":
    main()
