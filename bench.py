#!/usr/bin/env python
"""bench.py - DPI training-step throughput (samples/s) on N B200s, next to the CPU reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4] [--impl reference]

One "step" = one pass of the hot path over one synthetic batch: z ~ N(0,I) drawn on the device ->
RealNVP reverse with log|det J| -> positivity -> measurement operator -> data/prior/entropy losses
-> backward -> (gradient reduce-scatter / all-reduce) -> global-norm clip -> Adam. Default workload =
BASELINE.json configs[1]: DPI_interferometry npix=32, n_flow=16, batch=64 per GPU, synthetic EHT
uv-coverage. Prints ONE JSON line (rank 0). The default run also measures the other BASELINE
configurations in the same process (`configs`: c2, c3, c4; under --gpus N > 1 c4 strong + weak scaling and
a `dp_check` of the collective path), a stock-PyTorch-eager context number on the same GPU
(`gpu_eager_baseline`) and the drop-in autograd surface (`e2e_dropin`).

`--impl reference` times the reference's own CPU implementation of the step on the host cores: the
UNMODIFIED reference modules from the install `baseline/_ref` (baseline/ref_step.py) when present
(kind "reference"), else the oracle port (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (kind, npix, n_flow, per-GPU batch, lr, clip, description)
    "c1": ("interferometry", 32, 16, 64, 1e-4, 1e-3,
           "DPI_interferometry npix=32 n_flow=16 batch=64/GPU, synthetic EHT uv-coverage (938 vis, 582 cphase, "
           "493 logcamp), closure phase + log closure amp + MEM/TSV/L1/flux/center"),
    "c2": ("mri", 64, 16, 64, 1e-5, 1e-2,
           "DPI_MRI npix=64 n_flow=16 ratio=4 batch=64/GPU, knee-shaped phantom + 4x mask, masked FFT + TV, "
           "sigma=5e-7"),
    "c4": ("interferometry", 64, 32, 1024, 1e-4, 1e-3,
           "DPI_interferometry scale-up npix=64 n_flow=32 batch=1024/GPU (weak scaling), synthetic EHT uv-coverage"),
    "c3": ("dpix", 64, 16, 2048, 1e-4, 1e-4,
           "DPIx_interferometry crescent + 2 Gaussians (18 parameters, npix=64, fov=160), alpha-divergence 0.95, "
           "batch=2048/GPU, synthetic M87-like uv-coverage (7 stations, 25 scans), closure phase + log closure amp"),
}
METRIC = "DPI train samples/sec (flow+logdet+fwd-model, fwd+bwd)"
L2_NOTE = "flushed between timed steps (192 MiB fill)"


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, from the committed `ncu --set full`
    capture (bench.py cannot run ncu on itself): profiles/ncu_traffic.json, written by tools/ncu_traffic.py, stamped
    with the commit it was captured at."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return {}, None
    d = json.load(open(path))
    return d.get(workload, {}), d.get("commit")


def build_workload(name):
    from dpi_b200.workloads import dpix_workload, interferometry_workload, mri_workload
    kind, npix, nf, B, lr, clip, desc = WORKLOADS[name]
    wl = {"interferometry": interferometry_workload, "mri": mri_workload, "dpix": dpix_workload}[kind](npix=npix)
    return wl, kind, npix, nf, B, lr, clip, desc


def config_block(name, world, B, params, args, extra=None):
    """The SAME keys on both arms (the driver compares them)."""
    desc = WORKLOADS[name][6]
    c = {"workload": desc, "name": name, "global_batch": world * B, "batch_per_gpu": B, "params": params,
         "cuda_graph": not args.no_graph, "l2": L2_NOTE if not args.no_flush else "not flushed",
         "bn": "per-replica batch statistics", "host_sync_per_step": not args.free_run}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules from baseline/_ref (or the oracle port of the reference step; oracle/ is test
# infrastructure - this leg and the tests are the only places allowed to run it)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(name, steps, warmup, max_seconds=25.0, device="cpu"):
    wl, kind, npix, nf, B, lr, clip, desc = build_workload(name)
    torch.set_num_threads(os.cpu_count() or 1)
    D = npix * npix
    stepper, ref_kind = None, "port"
    if kind != "dpix":
        try:
            from baseline.ref_step import ReferenceStep, reference_root
            if reference_root() is not None:
                torch.manual_seed(0)
                stepper = ReferenceStep(kind, wl, nf, lr, clip, device=device)
                ref_kind = "reference"
        except Exception as e:      # a broken install must not take the bench down: fall back to the port, say so
            print(f"bench.py: reference install unusable ({e!r}); timing the oracle port", file=sys.stderr)
            stepper = None
    if stepper is None:
        if device != "cpu":
            return None
        from oracle import dpi_oracle as O
        if kind == "dpix":
            D = wl["nparams"]
            sd = O.init_state(D, nf, seqfrac=1 / 16, seed=0, randomize=False)
            c = dict(npix=npix, fov=wl["fov"], n_gaussian=wl["n_gaussian"], dft_mat=wl["dft_mat"], cphase_ind=wl["cphase_ind"],
                     cphase_sign=wl["cphase_sign"], camp_ind=wl["camp_ind"], cphase_true=wl["cphase_true"],
                     logcamp_true=wl["logcamp_true"], sigma_cp=wl["sigma_cp"], sigma_ca=wl["sigma_ca"], w=wl["w"],
                     alpha=wl["alpha"])
            stepper = O.OracleTrainer(sd, None, O.dpix_loss, c, lr=lr, clip=clip)
            stepper.data_weight = 1e-4
        else:
            sd = O.init_state(D, nf, seed=0, randomize=False)
            ls = torch.tensor([wl["log_scale_init"]], dtype=torch.float32)
            if kind == "interferometry":
                c = dict(npix=npix, dft_mat=wl["dft_mat"], cphase_ind=wl["cphase_ind"], cphase_sign=wl["cphase_sign"],
                         camp_ind=wl["camp_ind"], cphase_true=wl["cphase_true"], logcamp_true=wl["logcamp_true"],
                         sigma_cp=wl["sigma_cp"], sigma_ca=wl["sigma_ca"], prior=wl["prior"], flux=wl["flux"], w=wl["w"])
                stepper = O.OracleTrainer(sd, ls, O.interferometry_loss, c, lr=lr, clip=clip)
            else:
                c = dict(npix=npix, mask=wl["mask"], kspace_true=wl["kspace_true"], sigma=wl["sigma"], w=wl["w"])
                stepper = O.OracleTrainer(sd, ls, O.mri_loss, c, lr=lr, clip=clip)
    if kind == "dpix":
        D = wl["nparams"]
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    torch.manual_seed(1234)
    for _ in range(warmup):
        stepper.step(torch.randn(B, D))
    sync()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        stepper.step(torch.randn(B, D))      # DPI_interferometry.py:193 draws z on the host every step
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    sync()
    dt = time.perf_counter() - t0
    return dict(value=B * done / dt, ms_per_step=1e3 * dt / done, steps=done, cores=torch.get_num_threads(),
                batch=B, desc=desc, kind=ref_kind)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide bench state (device, ranks, the L2 flush buffer)."""

    def __init__(self, args):
        import torch.distributed as dist
        self.args, self.dist = args, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - the dpi_b200 product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.l2_flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)


def make_trainer(ctx, name, batch=None):
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    from dpi_b200.trainer import DPIxTrainer, InterferometryTrainer, MRITrainer
    wl, kind, npix, nf, B, lr, clip, desc = build_workload(name)
    B = int(batch) if batch else B
    D = npix * npix
    torch.manual_seed(0)                      # identical initial weights on every rank (and broadcast by the trainer)
    g = not ctx.args.no_graph
    if kind == "dpix":
        D = wl["nparams"]
        flow = RealNVP(D, nf, seqfrac=1 / 16)
        tr = DPIxTrainer(flow, wl, B, lr=lr, clip=clip, device=ctx.dev, use_graph=g)
    else:
        flow = RealNVP(D, nf)
        T = InterferometryTrainer if kind == "interferometry" else MRITrainer
        tr = T(flow, wl, B, lr=lr, clip=clip, device=ctx.dev, use_graph=g)
    return dict(tr=tr, flow=flow, wl=wl, kind=kind, npix=npix, nf=nf, B=B, D=D, lr=lr, clip=clip, desc=desc, name=name)


def measure(ctx, m, steps, warmup, clocks_index=None):
    """Device-timed throughput (`value`) and the end-to-end number through the public step API with host buffers."""
    args, tr, B, D, dev = ctx.args, m["tr"], m["B"], m["D"], ctx.dev
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + ctx.rank)              # per-rank z stream
    for _ in range(max(warmup, 3)):
        tr.sample_z(gen)
        tr.step()
    ctx.barrier()
    # ---- timed region: K steps, each between its own CUDA events; L2 flushed between steps
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sampler = ClockSampler(clocks_index) if clocks_index is not None else None
    if sampler:
        sampler.__enter__()
    ctx.barrier()
    t_wall0 = time.perf_counter()
    for s, e in evs:
        if not args.no_flush:
            ctx.l2_flush.fill_(1)
        s.record()
        tr.sample_z(gen)                  # z ~ N(0,I) drawn on the device, inside the timed step
        tr.step()
        e.record()
        if not args.free_run:
            e.synchronize()               # host waits for the step (as a training loop that logs its loss does)
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall0
    if sampler:
        sampler.__exit__()
    total_ms = ctx.max_over_ranks(sum(s.elapsed_time(e) for s, e in evs))
    # ---- end-to-end through the public step API with host buffers: pinned z -> H2D, loss -> D2H, every step; the
    #      L2 is flushed between steps here too (the flush is inside the wall-clock region: measured and subtracted)
    zs = [torch.randn(B, D).pin_memory() for _ in range(4)]
    host_terms = torch.empty(tr.terms.numel(), dtype=torch.float32).pin_memory()
    flush_s = 0.0
    if not args.no_flush:
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.l2_flush.fill_(1)
            torch.cuda.synchronize(dev)
        flush_s = time.perf_counter() - t0
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        if not args.no_flush:
            ctx.l2_flush.fill_(1)
            torch.cuda.synchronize(dev)
        tr.step(zs[k % 4])
        host_terms.copy_(tr.terms, non_blocking=False)     # the step's loss read back every step
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0 - flush_s)
    return dict(total_ms=total_ms, wall=t_wall, e2e_s=e2e_s, loss=float(host_terms[0]),
                clocks=sampler.summary() if sampler else None,
                launches=(tr.launches_per_step or 0) * steps)


def step_bytes_of(m):
    tr = m["tr"]
    if hasattr(tr, "algorithmic_bytes_per_step"):
        return tr.algorithmic_bytes_per_step()
    P_ = int(m["flow"].flat.numel())    # DPIx: parameter state + per-sample images + DFT matrix
    return 40 * P_ + 4 * m["B"] * (m["npix"] ** 2) * 4 + 2 * 4 * m["npix"] ** 2 * 2 * m["wl"]["n_vis"]


def algorithmic_flops(m):
    """B [3 * 2 * C (Da H + H^2 + H Dout) + DFT fwd + dgrad]."""
    flow, B, nf, npix = m["flow"], m["B"], m["nf"], m["npix"]
    H_, Dn = flow.hidden, flow.ndim
    flops = B * (3 * 2 * 2 * nf * ((Dn - Dn // 2) * H_ + H_ * H_ + H_ * 2 * (Dn // 2)))
    if m["kind"] in ("interferometry", "dpix"):
        flops += B * 2 * (2 * 2 * npix * npix * m["wl"]["n_vis"])
    return flops


def small_block(ctx, name, steps, batch=None, scaling=None):
    """One of the other BASELINE configurations, measured in this process: the same two numbers as the headline."""
    m = make_trainer(ctx, name, batch)
    r = measure(ctx, m, steps, 3)
    hbm_peak = PEAKS.get("hbm_gbs") or 6650.0
    sb = step_bytes_of(m)
    ms = r["total_ms"] / steps
    out = {"workload": m["desc"], "batch_per_gpu": m["B"], "global_batch": ctx.world * m["B"], "steps": steps,
           "ms_per_step": ms, "value": ctx.world * m["B"] * steps / (r["total_ms"] * 1e-3), "unit": "samples/s",
           "e2e": {"value": ctx.world * m["B"] * steps / r["e2e_s"], "ms_per_step": 1e3 * r["e2e_s"] / steps},
           "roofline": {"bound": "hbm", "step_algorithmic_bytes": sb, "achieved": sb / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                        "unit": "GB/s", "frac": sb / (ms * 1e-3) / 1e9 / hbm_peak,
                        "tensor_tflops": algorithmic_flops(m) / (ms * 1e-3) / 1e12},
           "gpu_launches": int(r["launches"]), "params": int(m["flow"].flat.numel())}
    if scaling:
        out["scaling"] = scaling
    if hasattr(m["tr"], "close"):
        m["tr"].close()
    del m
    torch.cuda.empty_cache()
    return out


def dp_check(ctx, m):
    """Hardware check of the data-parallel path (N > 1): (1) the collective's gradient equals the sum of the per-rank
    gradients gathered and added in fp64; (2) the replicas' parameters are identical after the timed steps."""
    from dpi_b200.dist import replicas_max_divergence
    tr, dist = m["tr"], ctx.dist
    div = replicas_max_divergence(tr.params, tr.pg)
    gen = torch.Generator(device=ctx.dev)
    gen.manual_seed(4321 + ctx.rank)
    tr.sample_z(gen)
    tr.loss_and_grad()
    torch.cuda.synchronize(ctx.dev)
    local = tr.grads_all.clone()
    parts = [torch.empty_like(local) for _ in range(ctx.world)]
    dist.all_gather(parts, local)
    ref = torch.zeros_like(local, dtype=torch.float64)
    for p_ in parts:
        ref += p_.double()
    tr._reduced = False
    tr._allreduce_impl()
    torch.cuda.synchronize(ctx.dev)
    if getattr(tr, "zero", False):
        got, want = tr.g_shard.double(), ref[tr.rank * tr.shard_n:(tr.rank + 1) * tr.shard_n]
    else:
        got, want = tr.grads_all.double(), ref
    err = float((got - want).abs().max() / ref.abs().max().clamp_min(1e-300))
    err = ctx.max_over_ranks(err)
    return {"replicas_max_divergence": div, "grad_sum_rel_err": err, "ranks": ctx.world,
            "collective": "reduce-scatter -> shard clip+Adam -> all-gather" if getattr(tr, "zero", False) else "all-reduce"}


PEAKS = {}


def gpu_run(args):
    from dpi_b200 import _lib
    global PEAKS
    ctx = Ctx(args)
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    _lib.load()
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        PEAKS = json.load(open(pk_path))
    m = make_trainer(ctx, args.workload)
    tr, flow, wl, kind, npix, nf, B, D = m["tr"], m["flow"], m["wl"], m["kind"], m["npix"], m["nf"], m["B"], m["D"]
    r = measure(ctx, m, args.steps, args.warmup, clocks_index=ctx.local)
    total_ms, e2e_s = r["total_ms"], r["e2e_s"]

    # ---- per-kernel timing of the dominant kernels (eager launches, CUDA events on the launch stream)
    step_bytes = step_bytes_of(m)
    if hasattr(tr, "profile_kernels"):
        prof = tr.profile_kernels(steps=min(args.steps, 10), flush=None if args.no_flush else ctx.l2_flush)
    else:   # DPIx: the step is reported as one unit
        prof = {"dpix_step": dict(ms=total_ms / args.steps, bytes=step_bytes)}
    check = dp_check(ctx, m) if world > 1 and hasattr(tr, "grads_all") else None
    ctx.barrier()

    # ---- the other BASELINE configurations, same process (default c1 run only)
    extra = {}
    if args.workload == "c1" and not args.no_extra:
        if hasattr(tr, "close"):
            tr.close()
        ks = max(5, min(args.steps, 20))
        if world == 1:
            for name in ("c2", "c3", "c4"):
                extra[name] = small_block(ctx, name, ks)
        else:
            extra["c4_weak"] = small_block(ctx, "c4", ks, scaling="weak: 1024 samples per GPU")
            extra["c4_strong"] = small_block(ctx, "c4", ks, batch=1024 // world,
                                             scaling=f"strong: global batch 1024 = {world} x {1024 // world}")
    ctx.barrier()
    if rank != 0:
        _teardown(tr, world)
        return
    hbm_peak, peak_src = (PEAKS.get("hbm_gbs"), "measured") if PEAKS.get("hbm_gbs") else (6650.0, "fallback")
    dom = max(prof, key=lambda k: prof[k]["ms"])
    traffic, traffic_commit = ncu_traffic(args.workload)
    roof = dict(bound="hbm", kernel=dom, achieved=prof[dom]["bytes"] / (prof[dom]["ms"] * 1e-3) / 1e9, peak=hbm_peak,
                peak_source=peak_src, unit="GB/s", traffic=traffic.get(dom), traffic_source=(
                    f"profiles/ncu_traffic.json (ncu --set full, commit {traffic_commit})" if traffic.get(dom) else None),
                algorithmic_bytes=prof[dom]["bytes"],
                ms_per_launch=prof[dom]["ms"],
                step=dict(algorithmic_bytes=step_bytes, achieved=step_bytes / (total_ms / args.steps * 1e-3) / 1e9),
                kernels={k: dict(ms=round(v["ms"], 5), gbs=round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)) for k, v in prof.items()})
    roof["frac"] = roof["achieved"] / hbm_peak
    roof["step"]["frac"] = roof["step"]["achieved"] / hbm_peak
    # tensor-pipe view of the same step against the measured dense bf16 peak (3xTF32 needs 6 bf16-equivalents per useful
    # flop, so 1/6 is the ceiling)
    flops = algorithmic_flops(m)
    tf_peak = PEAKS.get("bf16_tflops_sustained") or 1415.5
    tf = flops / (total_ms / args.steps * 1e-3) / 1e12
    roof["tensor"] = dict(algorithmic_gflop_per_step=flops / 1e9, achieved=tf, peak=tf_peak, unit="TFLOP/s", frac=tf / tf_peak,
                          path="tcgen05 3xTF32 (flow_tc.cu)" if B >= 96 else "fp32 FFMA hops (flow_small.cu), tcgen05 DFT")
    cpu = cpu_reference_run(args.workload, steps=args.cpu_steps, warmup=2)
    out = {
        "metric": METRIC, "value": world * B * args.steps / (total_ms * 1e-3), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args.workload, world, B, int(flow.flat.numel()), args),
        "wall_s_timed_region": round(r["wall"], 4),
        "clocks": r["clocks"],
        "e2e": {"value": world * B * args.steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": B * D * 4,
                "d2h_bytes_per_step": 64, "ms_per_step": 1e3 * e2e_s / args.steps,
                "l2": "flushed before every step (the fill is timed separately and subtracted)" if not args.no_flush else "warm"},
        "gpu_launches": int(r["launches"]),
        "roofline": roof,
        "cpu_baseline": {"value": cpu["value"], "unit": "samples/s", "cores": cpu["cores"], "kind": cpu["kind"],
                         "sample": f"{cpu['steps']} steps of the same workload (batch {cpu['batch']}) after 2 warm-up, "
                                   f"{cpu['ms_per_step']:.1f} ms/step, torch {torch.__version__} CPU"},
        "loss": r["loss"],
    }
    if check:
        out["dp_check"] = check
    if extra:
        out["configs"] = extra
    if world == 1 and not args.no_extra and kind != "dpix":
        # context numbers on the same GPU: (1) the unmodified reference, stock PyTorch eager, moved .to('cuda');
        # (2) dpi_b200's drop-in autograd surface driven by the reference's loop body + torch.optim.Adam
        try:
            eg = cpu_reference_run(args.workload, steps=20, warmup=3, max_seconds=20.0, device="cuda")
            if eg:
                out["gpu_eager_baseline"] = {"value": eg["value"], "unit": "samples/s", "ms_per_step": eg["ms_per_step"],
                                             "what": "unmodified reference modules .to('cuda'), stock PyTorch eager, same step "
                                                     "(host z + H2D every step as DPI_interferometry.py:193)"}
        except Exception as e:
            out["gpu_eager_baseline"] = {"unavailable": repr(e)[:200]}
        try:
            out["e2e_dropin"] = dropin_e2e(args.workload, 20)
        except Exception as e:
            out["e2e_dropin"] = {"unavailable": repr(e)[:200]}
    print(json.dumps(out), flush=True)
    _teardown(tr, world)


def dropin_e2e(name, steps):
    """The drop-in module surface exactly as a reference script would drive it (RealNVP.reverse -> Softplus -> the
    eht_observation_pytorch closure -> Loss_* closures -> loss.backward() -> clip_grad_norm_ -> torch.optim.Adam),
    host z every step, loss read back every step."""
    from baseline.ref_step import ReferenceStep
    import dpi_b200.MRI_helpers as MH
    import dpi_b200.interferometry_helpers as IH
    import dpi_b200.generative_model.realnvpfc_model as RN
    wl, kind, npix, nf, B, lr, clip, desc = build_workload(name)
    torch.manual_seed(0)
    st = ReferenceStep(kind, wl, nf, lr, clip, device="cuda", modules=dict(realnvp=RN, interferometry=IH, mri=MH))
    D = npix * npix
    for _ in range(3):
        st.step(torch.randn(B, D))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step(torch.randn(B, D))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": B * steps / dt, "unit": "samples/s", "ms_per_step": 1e3 * dt / steps,
            "what": "dpi_b200 drop-in modules behind the reference's loop body, autograd + torch.optim.Adam + clip_grad_norm_"}


def _teardown(tr, world):
    """Multi-rank exit. The step graph holds captured NCCL collectives, and NCCL will not destroy a communicator a
    live graph still references: release the graphs first, then destroy the group under a watchdog, then leave
    without running interpreter finalisers (rank 0 has already printed and flushed its line)."""
    if world <= 1:
        return
    import torch.distributed as dist
    if hasattr(tr, "close"):
        tr.close()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Timer(10.0, lambda: os._exit(0)).start()
    try:
        dist.destroy_process_group()
    except Exception:
        pass
    os._exit(0)


def reference_run(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = max(args.warmup, 3)                      # the same warm-up rule as the GPU arm
    r = cpu_reference_run(args.workload, steps=args.steps, warmup=warm, max_seconds=150.0)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    B = r["batch"]
    if r["kind"] == "reference":
        what = (f"{r['steps']} steps (batch {B}) of the UNMODIFIED reference modules (baseline/_ref) assembled into the "
                "loop body of DPI_interferometry.py:193-235 / DPI_MRI.py:162-207 (baseline/ref_step.py), all host cores")
    else:
        what = f"{r['steps']} steps (batch {B}) of the oracle port of the reference step on the host cores"
    nparams = None
    try:
        from dpi_b200 import _lib
        import ctypes as C
        kind, npix, nf = WORKLOADS[args.workload][0], WORKLOADS[args.workload][1], WORKLOADS[args.workload][2]
        if kind != "dpix":
            n_p, n_b = C.c_int64(), C.c_int64()
            D = npix * npix
            _lib.load().dpi_flow_layout(D, nf, int(D / 8), 1, None, None, C.byref(n_p), C.byref(n_b))
            nparams = int(n_p.value)
    except Exception:
        pass
    out = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "samples/s", "n_gpus": world,
           "steps": args.steps, "steps_timed": r["steps"], "warmup": warm, "ms_per_step": r["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_block(args.workload, world, B, nparams, args),
           "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": what},
           "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="dpi_b200", choices=["dpi_b200", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (skip configs / context numbers)")
    ap.add_argument("--free-run", action="store_true",
                    help="enqueue all timed steps without waiting for each one (default: the host waits for every step)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_run(args)
    else:
        gpu_run(args)


if __name__ == "__main__":
    main()
