"""Fused DPI training steps: the benchmarked hot path.

One `step(z)` = the loop body of DPI_interferometry.py:193-235 / DPI_MRI.py:162-207 (realnvp
branch): flow reverse with log|det J| -> softplus positivity and learned scale -> measurement
operator -> data + prior losses - entropy -> analytic backward through all of it -> global-norm
clip -> Adam, as ~12 launches of libdpi_b200 kernels on the current stream (no autograd, no host
synchronisation, no host reads). After warm-up the launch sequence is captured into a CUDA graph
and replayed.

Data parallel (one process per GPU): every rank runs the same step on its own z; the flat
gradient buffer (flow parameters + log_scale) is averaged with one NCCL all-reduce between the
backward pass and the clip (BatchNorm statistics stay per-replica = DDP semantics).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from dpi_b200 import _lib
from dpi_b200.generative_model.realnvpfc_model import RealNVP

_PRIOR_ORDER = ("l1", "tsv", "tv", "flux", "center", "mem")


class _FusedTrainer:
    def __init__(self, flow: RealNVP, log_scale_init: float, batch: int, npix: int, lr: float, clip: float,
                 device, betas=(0.9, 0.999), eps=1e-8, process_group=None, use_graph=True):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("dpi_b200 trainers need a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.flow = flow.to(self.device)
        self.flow.train()
        self.flow._init_reverse()
        self.B, self.npix, self.D = int(batch), int(npix), int(npix) * int(npix)
        assert self.D == flow.ndim
        self.lr, self.clip, self.betas, self.eps = float(lr), float(clip), betas, float(eps)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        dev, f32 = self.device, torch.float32
        P = self.flow.flat.numel()
        self.P = P
        # Data parallel, SURVEY 8(e): reduce-scatter the flat gradient -> per-shard ||g||^2 -> scalar all-reduce ->
        # clip + Adam on this rank's 1/world shard (optimizer state and traffic / world) -> all-gather the parameters.
        # Three collectives instead of one and no overlap with the backward. Measured on C1 (6.9 M parameters), sharded vs
        # all-reduce + replicated Adam: 2 GPUs 1.428 vs 1.407 ms, 8 GPUs 1.459 vs 1.517 ms - but the replicated optimiser
        # with the gradient exchange overlapped in two stage ranges beats both (2 / 4 / 8 GPUs: 1.387 / 1.413 / 1.442 ms), and
        # on C4 (218 M parameters, 8 GPUs) the bucketed all-reduce gives 21.3 ms against 21.6 ms sharded. So the sharded
        # optimiser is opt-in (DPI_B200_ZERO=1): it is the memory-saving mode (Adam state / world), not the fast one.
        self.zero = self.world > 1 and os.environ.get("DPI_B200_ZERO", "") == "1"
        if self.zero:
            from dpi_b200.dist import shard_layout
            self.rank = torch.distributed.get_rank(process_group)
            self.shard_n, n_all = shard_layout(P + 32, self.world)
            # ONE buffer [flow flat | log_scale (32) | pad]: the flow's parameter becomes a view of it
            self.pall = torch.zeros(n_all, dtype=f32, device=dev)
            self.pall[:P].copy_(self.flow.flat.data)
            self.flow.flat.data = self.pall[:P]
            self.params = self.flow.flat.data
            self.log_scale = self.pall[P:P + 32]
            self.log_scale[0] = float(log_scale_init)
            self.grads_all = torch.zeros(n_all, dtype=f32, device=dev)
            self.p_shard = self.pall[self.rank * self.shard_n:(self.rank + 1) * self.shard_n]
            self.g_shard = torch.zeros(self.shard_n, dtype=f32, device=dev)
            self.m = torch.zeros(self.shard_n, dtype=f32, device=dev)     # Adam state of this rank's shard only
            self.v = torch.zeros(self.shard_n, dtype=f32, device=dev)
        else:
            # parameters: [flow flat | log_scale | pad]; gradients share the layout so ONE all-reduce covers both
            self.params = self.flow.flat.data
            self.log_scale = torch.full((32,), 0.0, dtype=f32, device=dev)
            self.log_scale[0] = float(log_scale_init)
            self.grads_all = torch.zeros(P + 32, dtype=f32, device=dev)
            self.m = torch.zeros(P, dtype=f32, device=dev)
            self.v = torch.zeros(P, dtype=f32, device=dev)
        self.grads = self.grads_all[:P]
        self.g_ls = self.grads_all[P:P + 1]
        self.m_ls = torch.zeros(32, dtype=f32, device=dev)
        self.v_ls = torch.zeros(32, dtype=f32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sq = torch.zeros(2, dtype=torch.float64, device=dev)
        self.norm_out = torch.zeros(2, dtype=f32, device=dev)
        self.terms = torch.zeros(16, dtype=f32, device=dev)
        B, D = self.B, self.D
        self.z = torch.zeros(B, D, dtype=f32, device=dev)
        self.x = torch.empty(B, D, dtype=f32, device=dev)
        self.logdet_flow = torch.empty(B, dtype=f32, device=dev)
        self.img = torch.empty(B, D, dtype=f32, device=dev)
        self.det = torch.empty(B, dtype=f32, device=dev)
        self.prior_vals = torch.empty(B, 6, dtype=f32, device=dev)
        self.g_img_prior = torch.empty(B, D, dtype=f32, device=dev)
        self.g_img_data = torch.empty(B, D, dtype=f32, device=dev)
        self.g_x = torch.empty(B, D, dtype=f32, device=dev)
        self.g_ls_partial = torch.empty(B, dtype=f32, device=dev)
        self.handle = self.flow._handle(dev)
        nbytes = self.lib.dpi_flow_workspace_bytes(self.handle, B)
        if nbytes == 0:
            _lib.check(-1, "dpi_flow_workspace_bytes")
        self.ws_flow = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.ws_opt = torch.zeros(self.lib.dpi_clip_adam_workspace_bytes(P), dtype=torch.uint8, device=dev)
        # Data parallel on the tcgen05 path: run the flow backward in stage ranges and all-reduce each range's slice of
        # the flat gradient (the backward visits the flow blocks in parameter order) while the next range runs.
        self._buckets = self._plan_buckets()
        if self.world > 1:
            # replicas must START identical: every rank takes rank 0's parameters, running statistics, scale
            from dpi_b200.dist import broadcast_state_
            broadcast_state_([self.pall] if self.zero else [self.params, self.log_scale], 0, self.pg)
            broadcast_state_([self.flow.bn_running], 0, self.pg)
        self.use_graph = use_graph
        self._prof = None
        self._graphs = None
        self._warm = 0
        self.launches_per_step = None

    def _plan_buckets(self):
        # tcgen05 path: 4 ranges of a many-launch backward; small-batch kernel: 2 persistent launches (each costs a launch
        # + pipeline fill; measured on C1 at 8 ranks: 2 ranges 1.442 ms, 3 ranges 1.549 ms, 4 ranges 1.579 ms)
        kind = int(self.lib.dpi_flow_range_ok(self.handle, self.B))
        default = ("4" if kind == 1 else "2") if self.world > 1 else "0"
        if kind == 2 and self.zero:
            # measured at 8 ranks on C1: one reduce-scatter + shard Adam + all-gather 1.459 ms; two bucket all-reduces +
            # shard Adam + all-gather 1.535 ms (half again the traffic)
            default = "0"
        nb = int(os.environ.get("DPI_B200_BUCKETS", default))
        if nb < 2 or kind == 0:
            return None
        from dpi_b200.dist import gradient_buckets
        nf = self.flow.n_flow
        offs = [int(self.lib.dpi_flow_param_offset(self.handle, i, 0)) for i in range(nf)]
        return gradient_buckets(nf, offs, self.P + 32, nb)

    def _flow_backward(self, st):
        lib = self.lib
        if self._buckets is None:
            _lib.check(lib.dpi_flow_backward(self.handle, self.B, self.params.data_ptr(), self.grads.data_ptr(),
                                             self.z.data_ptr(), self.g_x.data_ptr(), self.g_det.data_ptr(), None,
                                             self.ws_flow.data_ptr(), self.ws_flow.numel(), st), "dpi_flow_backward")
            self._reduced = False
            return
        works = []
        for s_lo, s_hi, f_lo, f_hi in self._buckets:
            _lib.check(lib.dpi_flow_backward_range(self.handle, self.B, self.params.data_ptr(), self.grads.data_ptr(),
                                                   self.z.data_ptr(), self.g_x.data_ptr(), self.g_det.data_ptr(), None,
                                                   self.ws_flow.data_ptr(), self.ws_flow.numel(), s_lo, s_hi, st),
                       "dpi_flow_backward_range")
            if self.world > 1:
                from dpi_b200.dist import allreduce_buckets_
                works += allreduce_buckets_(self.grads_all, [(s_lo, s_hi, f_lo, f_hi)], self.pg)
        for w in works:
            w.wait()
        self._reduced = True

    # -- subclass hooks -------------------------------------------------------------------------
    def _data_forward_backward(self, st):
        raise NotImplementedError

    def _terms(self, st):
        raise NotImplementedError

    # -- the step, as two enqueue halves around the (optional) gradient all-reduce ------------------
    def _mark(self, name):
        if self._prof is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._prof.append((name, ev))

    def _enqueue_loss_grad(self):
        lib, st, B = self.lib, _lib.stream_ptr(), self.B
        fl = self.flow
        self._mark("start")
        _lib.check(lib.dpi_flow_run(self.handle, 0, 1, B, self.params.data_ptr(), fl.bn_running.data_ptr(),
                                    self.z.data_ptr(), self.x.data_ptr(), self.logdet_flow.data_ptr(),
                                    self.ws_flow.data_ptr(), self.ws_flow.numel(), st), "dpi_flow_run")
        self._mark("flow_reverse_fwd")
        _lib.check(lib.dpi_image_head(B, self.npix, self.x.data_ptr(), self.log_scale.data_ptr(), self.img.data_ptr(),
                                      self.det.data_ptr(), _lib.ptr(self.prior_img), self.flux_target, self.center,
                                      self.prior_vals.data_ptr(), self.gw_prior.data_ptr(),
                                      self.g_img_prior.data_ptr(), st), "dpi_image_head")
        self._mark("image_head")
        self._data_forward_backward(st)
        self._mark("measurement_fwd_bwd")
        _lib.check(lib.dpi_positivity_backward(B, self.npix, self.x.data_ptr(), self.log_scale.data_ptr(),
                                               self.img.data_ptr(), self.g_img_prior.data_ptr(),
                                               self.g_img_data.data_ptr(), self.g_det.data_ptr(), self.g_x.data_ptr(),
                                               self.g_ls_partial.data_ptr(), st), "dpi_positivity_backward")
        self._mark("positivity_bwd")
        # the log_scale gradient rides in the tail of the gradient buffer: finish it before the last range is reduced
        _lib.check(lib.dpi_reduce_sum(B, self.g_ls_partial.data_ptr(), self.g_ls.data_ptr(), None, st),
                   "dpi_reduce_sum")
        self._flow_backward(st)
        self._mark("flow_reverse_bwd")
        self._terms(st)
        self._mark("loss_terms")

    def _enqueue_update(self):
        lib, st = self.lib, _lib.stream_ptr()
        gs = 1.0 / self.world
        if self.zero:
            from dpi_b200.dist import all_gather_shards_
            b1, b2 = self.betas
            # g_shard holds this rank's slice of the summed gradient (reduce-scatter, or a slice of the bucket-reduced buffer)
            g = self.g_shard if not self._reduced else self.grads_all[self.rank * self.shard_n:(self.rank + 1) * self.shard_n]
            _lib.check(lib.dpi_grad_sqnorm(self.shard_n, g.data_ptr(), self.sq.data_ptr(), self.ws_opt.data_ptr(),
                                           self.ws_opt.numel(), st), "dpi_grad_sqnorm")
            torch.distributed.all_reduce(self.sq[:1], op=torch.distributed.ReduceOp.SUM, group=self.pg)
            _lib.check(lib.dpi_step_increment(self.step_dev.data_ptr(), st), "dpi_step_increment")
            _lib.check(lib.dpi_clip_adam(self.shard_n, self.p_shard.data_ptr(), g.data_ptr(), self.m.data_ptr(),
                                         self.v.data_ptr(), self.sq.data_ptr(), 1, self.clip, self.lr, b1, b2, self.eps,
                                         self.step_dev.data_ptr(), gs, self.norm_out.data_ptr(), st), "dpi_clip_adam")
            all_gather_shards_(self.pall, self.p_shard, self.pg)
            self._mark("clip_adam")
            return
        _lib.check(lib.dpi_grad_sqnorm(self.P + 1, self.grads_all.data_ptr(), self.sq.data_ptr(),
                                       self.ws_opt.data_ptr(), self.ws_opt.numel(), st), "dpi_grad_sqnorm")
        _lib.check(lib.dpi_step_increment(self.step_dev.data_ptr(), st), "dpi_step_increment")
        b1, b2 = self.betas
        _lib.check(lib.dpi_clip_adam(self.P, self.params.data_ptr(), self.grads.data_ptr(), self.m.data_ptr(),
                                     self.v.data_ptr(), self.sq.data_ptr(), 1, self.clip, self.lr, b1, b2, self.eps,
                                     self.step_dev.data_ptr(), gs, self.norm_out.data_ptr(), st), "dpi_clip_adam")
        _lib.check(lib.dpi_clip_adam(1, self.log_scale.data_ptr(), self.g_ls.data_ptr(), self.m_ls.data_ptr(),
                                     self.v_ls.data_ptr(), self.sq.data_ptr(), 1, self.clip, self.lr, b1, b2, self.eps,
                                     self.step_dev.data_ptr(), gs, None, st), "dpi_clip_adam")
        self._mark("clip_adam")

    def _allreduce(self):
        self._allreduce_impl()
        self._mark("grad_allreduce")

    def _allreduce_impl(self):
        if self.world > 1 and not getattr(self, "_reduced", False):
            if self.zero:
                from dpi_b200.dist import reduce_scatter_sum_
                reduce_scatter_sum_(self.g_shard, self.grads_all, self.pg)
            else:
                from dpi_b200.dist import allreduce_sum_
                allreduce_sum_(self.grads_all, self.pg)

    def loss_and_grad(self, z: Optional[torch.Tensor] = None):
        """Forward + backward only (no update): fills self.grads / self.g_ls / self.terms."""
        if z is not None:
            self.z.copy_(z, non_blocking=True)
        with torch.cuda.device(self.device):
            self._enqueue_loss_grad()
        self.flow._bump_bn()
        return self.terms

    def step(self, z: Optional[torch.Tensor] = None):
        """One optimisation step. z [B, D] (host pinned or device) is copied into the static input
        buffer; None re-uses whatever self.z holds (e.g. torch.randn written in place)."""
        if z is not None:
            self.z.copy_(z, non_blocking=True)
        with torch.cuda.device(self.device):
            if not self.use_graph:
                n0 = _lib.launch_count()
                self._enqueue_loss_grad()
                self._allreduce()
                self._enqueue_update()
                self.launches_per_step = _lib.launch_count() - n0
            elif self._graphs is None:
                n0 = _lib.launch_count()
                self._enqueue_loss_grad()
                self._allreduce()
                self._enqueue_update()
                self.launches_per_step = _lib.launch_count() - n0
                self._warm += 1
                if self._warm >= 2:
                    self._capture()
            elif self._graphs[1] is None:   # one graph holds the whole step, the NCCL all-reduce included
                self._graphs[0].replay()
            else:
                ga, gb = self._graphs
                ga.replay()
                self._allreduce()
                gb.replay()
        self.flow._bump_bn()
        return self.terms

    def _capture(self):
        torch.cuda.synchronize(self.device)
        ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        # the captured launches must not change any state: save / restore what a step mutates
        saved = [t.clone() for t in (self.params, self.log_scale, self.m, self.v, self.m_ls, self.v_ls, self.step_dev,
                                     self.flow.bn_running)]
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        single = None
        if self.world > 1 and os.environ.get("DPI_B200_GRAPH_NCCL", "1") != "0":
            # Data parallel: try to capture the gradient all-reduce inside ONE graph with both halves of the step (no
            # host round trip between backward and Adam). NCCL collectives are capturable; if this build refuses,
            # fall back to two graphs with the all-reduce issued eagerly between them.
            try:
                single = torch.cuda.CUDAGraph()
                with torch.cuda.stream(s):
                    with torch.cuda.graph(single, stream=s):
                        self._enqueue_loss_grad()
                        self._allreduce_impl()
                        self._enqueue_update()
            except Exception:
                single = None
                torch.cuda.synchronize(self.device)
        if single is not None:
            ga, gb = single, None
        else:
            with torch.cuda.stream(s):
                with torch.cuda.graph(ga, stream=s):
                    self._enqueue_loss_grad()
                with torch.cuda.graph(gb, stream=s):
                    self._enqueue_update()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        for t, sv in zip((self.params, self.log_scale, self.m, self.v, self.m_ls, self.v_ls, self.step_dev,
                          self.flow.bn_running), saved):
            t.copy_(sv)
        self._graphs = (ga, gb)

    def close(self):
        """Drop the captured step graphs (they pin the NCCL communicator under data parallel); the next step() would
        re-warm and re-capture."""
        if self._graphs is not None:
            torch.cuda.synchronize(self.device)
            self._graphs = None
            self._warm = 0

    # -- conveniences ---------------------------------------------------------------------------
    def sample_z(self, generator=None):
        """z ~ N(0, I) drawn on the device into the static input buffer."""
        self.z.normal_(generator=generator)
        return self.z

    def log_scale_value(self) -> float:
        return float(self.log_scale[0].item())

    def kernel_bytes(self):
        """Algorithmic bytes per launch group, used for per-kernel roofline lines."""
        fl = self.flow
        S, B, D, H = 2 * fl.n_flow, self.B, self.D, fl.hidden
        P = fl.flat.numel()
        act = B * (D + 2 * H + 2 * (D // 2))          # per stage: y, l1, l2, o
        return {
            "flow_reverse_fwd": 4 * (P + B * D + S * act + B * D),
            "image_head": 4 * (3 * B * D + D),
            "measurement_fwd_bwd": self._data_bytes(),
            "positivity_bwd": 4 * 5 * B * D,
            "flow_reverse_bwd": 4 * (2 * P + S * act + B * D),
            "loss_terms": 4 * 10 * B,
            "clip_adam": 4 * 8 * P,                      # norm pass (1 read) + Adam (4 reads, 3 writes)
            "grad_allreduce": 4 * P,
        }

    def _data_bytes(self):
        return 4 * 2 * self.B * self.D

    def profile_kernels(self, steps=5, flush=None):
        """Average CUDA-event duration of each launch group over `steps` eager (non-graph) steps."""
        acc, kb = {}, self.kernel_bytes()
        for _ in range(steps):
            if flush is not None:
                flush.fill_(1)
            self.sample_z()
            self._prof = []
            with torch.cuda.device(self.device):
                self._enqueue_loss_grad()
                self._allreduce()
                self._enqueue_update()
            torch.cuda.synchronize(self.device)
            prev = None
            for name, ev in self._prof:
                if prev is not None and name != "start":
                    acc.setdefault(name, []).append(prev.elapsed_time(ev))
                prev = ev
            self._prof = None
            self.flow._bump_bn()
        out = {}
        for name, v in acc.items():
            if name == "grad_allreduce" and self.world == 1:
                continue
            out[name] = dict(ms=sum(v) / len(v), bytes=kb.get(name, 0))
        return out

    def algorithmic_bytes_per_step(self) -> int:
        """40 P (weights fwd + bwd, grad write, Adam 4 reads + 3 writes) + saved activations
        once written and once read + constants."""
        C_ = 2 * self.flow.n_flow
        return int(40 * self.flow.flat.numel() + 4 * self.B * self.D * (2 * C_ + 4) + self._const_bytes())

    def _const_bytes(self):
        return 0


class InterferometryTrainer(_FusedTrainer):
    """DPI_interferometry.py:191-235 (closure phase + log closure amplitude + 5 image priors)."""

    def __init__(self, flow: RealNVP, wl: dict, batch: int, lr=1e-4, clip=1e-3, device="cuda", **kw):
        super().__init__(flow, wl["log_scale_init"], batch, wl["npix"], lr, clip, device, **kw)
        from dpi_b200.interferometry_helpers import _EhtOp
        dev, f32, B = self.device, torch.float32, self.B
        self.wl, w = wl, wl["w"]
        self.op = _EhtOp(wl["npix"], wl["dft_mat"], wl["cphase_ind"], wl["cphase_sign"], wl["camp_ind"], dev)
        self.cphase_true = wl["cphase_true"].to(dev, f32).contiguous()
        self.logcamp_true = wl["logcamp_true"].to(dev, f32).contiguous()
        self.sigma_cp = wl["sigma_cp"].to(dev, f32).contiguous()
        self.sigma_ca = wl["sigma_ca"].to(dev, f32).contiguous()
        self.prior_img = wl["prior"].to(dev, f32).contiguous()
        self.flux_target, self.center = float(wl["flux"]), float(self.npix / 2 - 0.5)
        self.w = w
        wp = [w.get(k, 0.0) for k in _PRIOR_ORDER]
        self.wp_host = (C.c_float * 6)(*wp)
        self.gw_prior = (torch.tensor(wp, dtype=f32, device=dev) / B).repeat(B, 1).contiguous()
        self.g_det = torch.full((B,), -w["logdet"] / B, dtype=f32, device=dev)
        self.vis = torch.empty(B, 2, self.op.n_vis, dtype=f32, device=dev)
        self.loss_cp = torch.empty(B, dtype=torch.float64, device=dev)
        self.loss_ca = torch.empty(B, dtype=f32, device=dev)
        self.ws_eht = self.op.workspace(B)

    def _data_forward_backward(self, st):
        lib, B, op = self.lib, self.B, self.op
        _lib.check(lib.dpi_eht_forward(op.handle, B, self.img.data_ptr(), self.vis.data_ptr(), None, None, None,
                                       self.ws_eht.data_ptr(), self.ws_eht.numel(), st), "dpi_eht_forward")
        _lib.check(lib.dpi_eht_data_grad(op.handle, B, self.vis.data_ptr(), self.cphase_true.data_ptr(),
                                         self.sigma_cp.data_ptr(), self.logcamp_true.data_ptr(),
                                         self.sigma_ca.data_ptr(), self.w["cphase"] / B, self.w["camp"] / B,
                                         self.loss_cp.data_ptr(), self.loss_ca.data_ptr(), self.g_img_data.data_ptr(),
                                         self.ws_eht.data_ptr(), self.ws_eht.numel(), st), "dpi_eht_data_grad")

    def _terms(self, st):
        _lib.check(self.lib.dpi_step_terms(self.B, self.loss_cp.data_ptr(), self.w["cphase"], self.loss_ca.data_ptr(),
                                           self.w["camp"], self.prior_vals.data_ptr(), self.wp_host,
                                           self.logdet_flow.data_ptr(), self.det.data_ptr(), self.w["logdet"],
                                           self.terms.data_ptr(), st), "dpi_step_terms")

    def _const_bytes(self):
        return 2 * 4 * self.op.P * 2 * self.op.n_vis  # F read forward and backward

    def _data_bytes(self):
        nv2 = 2 * self.op.n_vis
        return 4 * (2 * self.op.P * nv2 + 2 * self.B * self.D + 4 * self.B * nv2)


class MRITrainer(_FusedTrainer):
    """DPI_MRI.py:160-207 (masked ortho FFT2 k-space chi^2 + TV/L1 priors)."""

    def __init__(self, flow: RealNVP, wl: dict, batch: int, lr=1e-5, clip=1e-2, device="cuda", **kw):
        super().__init__(flow, wl["log_scale_init"], batch, wl["npix"], lr, clip, device, **kw)
        dev, f32, B = self.device, torch.float32, self.B
        self.wl, w = wl, wl["w"]
        self.w = w
        self.mask = wl["mask"].to(dev, f32).contiguous()
        self.kspace_true = wl["kspace_true"].to(dev, f32).contiguous()
        self.sigma, self.mean_mask = float(wl["sigma"]), float(wl["mean_mask"])
        self.prior_img = None
        self.flux_target, self.center = 0.0, 0.0
        wp = [w.get(k, 0.0) for k in _PRIOR_ORDER]
        self.wp_host = (C.c_float * 6)(*wp)
        self.gw_prior = (torch.tensor(wp, dtype=f32, device=dev) / B).repeat(B, 1).contiguous()
        self.g_det = torch.full((B,), -w["logdet"] / B, dtype=f32, device=dev)
        self.gw_data = torch.full((B,), w["kspace"] / B, dtype=f32, device=dev)
        self.loss_k = torch.empty(B, dtype=f32, device=dev)

    def _data_forward_backward(self, st):
        _lib.check(self.lib.dpi_mri_data_loss(self.B, self.npix, self.img.data_ptr(), self.mask.data_ptr(),
                                              self.kspace_true.data_ptr(), self.sigma, self.mean_mask,
                                              self.loss_k.data_ptr(), self.gw_data.data_ptr(),
                                              self.g_img_data.data_ptr(), st), "dpi_mri_data_loss")

    def _terms(self, st):
        _lib.check(self.lib.dpi_step_terms(self.B, None, 0.0, self.loss_k.data_ptr(), self.w["kspace"],
                                           self.prior_vals.data_ptr(), self.wp_host, self.logdet_flow.data_ptr(),
                                           self.det.data_ptr(), self.w["logdet"], self.terms.data_ptr(), st),
                   "dpi_step_terms")


class DPIxTrainer:
    """alpha-DPI step, DPIx_interferometry.py:329-386 (geometric_model 'simple_crescent_floor_nuisance',
    data_product 'cphase_logcamp'): flow over the model parameters -> sigmoid -> crescent renderer -> closure
    operator -> per-sample objective -> self-normalised importance weights over the whole batch -> backward ->
    clip -> Adam. All device work is libdpi_b200 kernels; under data parallel the softmax statistics (max, sum)
    and the flat gradient are all-reduced."""

    def __init__(self, flow: RealNVP, wl: dict, batch: int, lr=1e-4, clip=1e-4, device="cuda", start_order=4,
                 decay_rate=2000, betas=(0.9, 0.999), eps=1e-8, process_group=None, use_graph=True):
        from dpi_b200.interferometry_helpers import _EhtOp
        self.use_graph, self._graphs, self._warm = use_graph, None, 0
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("dpi_b200 trainers need a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        dev, f32 = self.device, torch.float32
        self.flow = flow.to(dev)
        self.flow.train()
        self.flow._init_reverse()
        self.B, self.npix, self.np_, self.ng = int(batch), int(wl["npix"]), int(wl["nparams"]), int(wl["n_gaussian"])
        assert self.np_ == flow.ndim
        self.fov, self.alpha, self.w = float(wl["fov"]), float(wl["alpha"]), wl["w"]
        self.lr, self.clip, self.betas, self.eps = float(lr), float(clip), betas, float(eps)
        self.start_order, self.decay_rate, self.k = start_order, decay_rate, 0
        self.pg, self.world = process_group, 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.op = _EhtOp(self.npix, wl["dft_mat"], wl["cphase_ind"], wl["cphase_sign"], wl["camp_ind"], dev)
        self.cphase_true = wl["cphase_true"].to(dev, f32).contiguous()
        self.logcamp_true = wl["logcamp_true"].to(dev, f32).contiguous()
        self.sigma_cp = wl["sigma_cp"].to(dev, f32).contiguous()
        self.sigma_ca = wl["sigma_ca"].to(dev, f32).contiguous()
        B, D, P = self.B, self.npix * self.npix, self.flow.flat.numel()
        self.P = P
        self.params = self.flow.flat.data
        self.grads = torch.zeros(P, dtype=f32, device=dev)
        self.m, self.v = torch.zeros(P, dtype=f32, device=dev), torch.zeros(P, dtype=f32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sq = torch.zeros(2, dtype=torch.float64, device=dev)
        self.norm_out = torch.zeros(2, dtype=f32, device=dev)
        self.terms = torch.zeros(8, dtype=f32, device=dev)
        self.stats = torch.zeros(2, dtype=f32, device=dev)
        self.z = torch.zeros(B, self.np_, dtype=f32, device=dev)
        self.ps = torch.empty(B, self.np_, dtype=f32, device=dev)
        self.prm = torch.empty(B, self.np_, dtype=f32, device=dev)
        self.logdet_flow, self.det = torch.empty(B, dtype=f32, device=dev), torch.empty(B, dtype=f32, device=dev)
        self.img = torch.empty(B, D, dtype=f32, device=dev)
        self.vis = torch.empty(B, 2, self.op.n_vis, dtype=f32, device=dev)
        self.loss_cp = torch.empty(B, dtype=torch.float64, device=dev)
        self.loss_ca = torch.empty(B, dtype=f32, device=dev)
        self.loss_i = torch.empty(B, dtype=f32, device=dev)
        self.g_img = torch.empty(B, D, dtype=f32, device=dev)
        self.gscale, self.g_logdet = torch.empty(B, dtype=f32, device=dev), torch.empty(B, dtype=f32, device=dev)
        self.g_prm = torch.empty(B, self.np_, dtype=f32, device=dev)
        self.g_ps = torch.empty(B, self.np_, dtype=f32, device=dev)
        self.handle = self.flow._handle(dev)
        nbytes = self.lib.dpi_flow_workspace_bytes(self.handle, B)
        if nbytes == 0:
            _lib.check(-1, "dpi_flow_workspace_bytes")
        self.ws_flow = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.ws_eht = self.op.workspace(B)
        self.ws_opt = torch.zeros(self.lib.dpi_clip_adam_workspace_bytes(P), dtype=torch.uint8, device=dev)
        self.launches_per_step = None
        if self.world > 1:      # replicas must start identical
            from dpi_b200.dist import broadcast_state_
            broadcast_state_([self.params, self.flow.bn_running], 0, self.pg)

    def data_weight(self):
        return min(10.0 ** (-self.start_order + self.k / self.decay_rate), 1.0)      # DPIx_interferometry.py:331

    def sample_z(self, generator=None):
        self.z.normal_(generator=generator)
        return self.z

    def _objective(self, mode, dw, st):
        _lib.check(self.lib.dpi_alpha_objective(self.B, self.np_, self.z.data_ptr(), self.loss_cp.data_ptr(),
                                                self.loss_ca.data_ptr(), self.logdet_flow.data_ptr(), self.det.data_ptr(),
                                                self.w["cphase"], self.w["camp"], self.w["scale_factor"], dw, self.alpha,
                                                mode, self.B * self.world, self.loss_i.data_ptr(), self.gscale.data_ptr(),
                                                self.g_logdet.data_ptr(), self.terms.data_ptr(), self.stats.data_ptr(), st),
                   "dpi_alpha_objective")

    # -- the step in three enqueue parts: [flow -> renderer -> closure operator and its unit-weight gradient],
    #    the objective (takes the annealed data weight as a host scalar, and two scalar all-reduces under data
    #    parallel), [renderer / flow backward], then the update. Parts 1, 3 and the update are CUDA-graph replays
    #    after warm-up.
    def _enqueue_forward(self):
        lib, B, st, fl = self.lib, self.B, _lib.stream_ptr(), self.flow
        _lib.check(lib.dpi_flow_run(self.handle, 0, 1, B, self.params.data_ptr(), fl.bn_running.data_ptr(),
                                    self.z.data_ptr(), self.ps.data_ptr(), self.logdet_flow.data_ptr(),
                                    self.ws_flow.data_ptr(), self.ws_flow.numel(), st), "dpi_flow_run")
        _lib.check(lib.dpi_sigmoid_forward(B, self.np_, self.ps.data_ptr(), self.prm.data_ptr(), self.det.data_ptr(), st),
                   "dpi_sigmoid_forward")
        _lib.check(lib.dpi_crescent_forward(B, self.npix, self.ng, self.fov, self.prm.data_ptr(), self.img.data_ptr(), st),
                   "dpi_crescent_forward")
        _lib.check(lib.dpi_eht_forward(self.op.handle, B, self.img.data_ptr(), self.vis.data_ptr(), None, None, None,
                                       self.ws_eht.data_ptr(), self.ws_eht.numel(), st), "dpi_eht_forward")
        # unit-weight data gradient d(w_ca l_ca + w_cp l_cp)/d img; the per-sample importance weight enters as gscale
        _lib.check(lib.dpi_eht_data_grad(self.op.handle, B, self.vis.data_ptr(), self.cphase_true.data_ptr(),
                                         self.sigma_cp.data_ptr(), self.logcamp_true.data_ptr(), self.sigma_ca.data_ptr(),
                                         self.w["cphase"], self.w["camp"], self.loss_cp.data_ptr(), self.loss_ca.data_ptr(),
                                         self.g_img.data_ptr(), self.ws_eht.data_ptr(), self.ws_eht.numel(), st),
                   "dpi_eht_data_grad")

    def _enqueue_objective(self, dw):
        st = _lib.stream_ptr()
        if self.world > 1 and self.alpha != 1.0:
            import torch.distributed as dist
            self._objective(1, dw, st)
            dist.all_reduce(self.stats[0:1], op=dist.ReduceOp.MAX, group=self.pg)
            self._objective(2, dw, st)
            dist.all_reduce(self.stats[1:2], op=dist.ReduceOp.SUM, group=self.pg)
            self._objective(3, dw, st)
        else:
            self._objective(0, dw, st)

    def _enqueue_backward(self):
        lib, B, st = self.lib, self.B, _lib.stream_ptr()
        _lib.check(lib.dpi_crescent_backward(B, self.npix, self.ng, self.fov, self.prm.data_ptr(), self.g_img.data_ptr(),
                                             self.gscale.data_ptr(), self.g_prm.data_ptr(), st), "dpi_crescent_backward")
        _lib.check(lib.dpi_sigmoid_backward(B, self.np_, self.ps.data_ptr(), self.g_prm.data_ptr(), self.g_logdet.data_ptr(),
                                            self.g_ps.data_ptr(), st), "dpi_sigmoid_backward")
        _lib.check(lib.dpi_flow_backward(self.handle, B, self.params.data_ptr(), self.grads.data_ptr(), self.z.data_ptr(),
                                         self.g_ps.data_ptr(), self.g_logdet.data_ptr(), None, self.ws_flow.data_ptr(),
                                         self.ws_flow.numel(), st), "dpi_flow_backward")

    def _enqueue_update(self):
        lib, st = self.lib, _lib.stream_ptr()
        b1, b2 = self.betas
        _lib.check(lib.dpi_grad_sqnorm(self.P, self.grads.data_ptr(), self.sq.data_ptr(), self.ws_opt.data_ptr(),
                                       self.ws_opt.numel(), st), "dpi_grad_sqnorm")
        _lib.check(lib.dpi_step_increment(self.step_dev.data_ptr(), st), "dpi_step_increment")
        # under data parallel every rank's weights already sum to 1 over the GLOBAL batch: plain sum of gradients
        _lib.check(lib.dpi_clip_adam(self.P, self.params.data_ptr(), self.grads.data_ptr(), self.m.data_ptr(),
                                     self.v.data_ptr(), self.sq.data_ptr(), 1, self.clip, self.lr, b1, b2, self.eps,
                                     self.step_dev.data_ptr(), 1.0, self.norm_out.data_ptr(), st), "dpi_clip_adam")

    def loss_and_grad(self, z=None, data_weight=None):
        if z is not None:
            self.z.copy_(z, non_blocking=True)
        dw = self.data_weight() if data_weight is None else float(data_weight)
        with torch.cuda.device(self.device):
            self._enqueue_forward()
            self._enqueue_objective(dw)
            self._enqueue_backward()
        self.flow._bump_bn()
        return self.terms

    def _allreduce_grads(self):
        if self.world > 1:
            from dpi_b200.dist import allreduce_sum_
            allreduce_sum_(self.grads, self.pg)

    def step(self, z=None, data_weight=None):
        if z is not None:
            self.z.copy_(z, non_blocking=True)
        dw = self.data_weight() if data_weight is None else float(data_weight)
        with torch.cuda.device(self.device):
            if self._graphs is None:
                n0 = _lib.launch_count()
                self._enqueue_forward()
                self._enqueue_objective(dw)
                self._enqueue_backward()
                self._allreduce_grads()
                self._enqueue_update()
                self.launches_per_step = _lib.launch_count() - n0
                self._warm += 1
                if self.use_graph and self._warm >= 2:
                    self._capture()
            else:
                ga, gb, gc = self._graphs
                ga.replay()
                self._enqueue_objective(dw)
                gb.replay()
                self._allreduce_grads()
                gc.replay()
        self.flow._bump_bn()
        self.k += 1
        return self.terms

    def _capture(self):
        """Capture the three launch sequences (they change no state while being captured: nothing executes)."""
        torch.cuda.synchronize(self.device)
        graphs = [torch.cuda.CUDAGraph() for _ in range(3)]
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for g, fn in zip(graphs, (self._enqueue_forward, self._enqueue_backward, self._enqueue_update)):
                with torch.cuda.graph(g, stream=s):
                    fn()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self._graphs = tuple(graphs)

    def close(self):
        if self._graphs is not None:
            torch.cuda.synchronize(self.device)
            self._graphs = None
            self._warm = 0

    # ------------------------------------------------------------------ posterior sampling
    @torch.no_grad()
    def sample_posterior(self, n_concat=1, rejsamp=True, generator=None, z=None):
        """`Gen_samples` of DPIx_interferometry.py:433-470: draw z, run the trained flow in reverse with the BatchNorm
        running statistics (the script calls `.eval()` first, :899), render, evaluate loss_orig = 0.5 loss_data +
        logprob per sample, importance weights softmax_i(-loss_orig), rejection step U < w / max w. Forward-only reuse
        of the training kernels. Returns dict(params, img, weights, accepted (bool), img_mean, img_std, img_map)."""
        lib, B, st = self.lib, self.B, None
        was_training = self.flow.training
        self.flow.eval()
        out = dict(params=[], img=[], weights=[], accepted=[], loss_data=[])
        try:
            for k in range(n_concat):
                if z is not None:
                    self.z.copy_(z[k] if z.dim() == 3 else z)
                else:
                    self.sample_z(generator)
                with torch.cuda.device(self.device):
                    st = _lib.stream_ptr()
                    _lib.check(lib.dpi_flow_run(self.handle, 0, 0, B, self.params.data_ptr(), self.flow.bn_running.data_ptr(),
                                                self.z.data_ptr(), self.ps.data_ptr(), self.logdet_flow.data_ptr(),
                                                self.ws_flow.data_ptr(), self.ws_flow.numel(), st), "dpi_flow_run")
                    _lib.check(lib.dpi_sigmoid_forward(B, self.np_, self.ps.data_ptr(), self.prm.data_ptr(),
                                                       self.det.data_ptr(), st), "dpi_sigmoid_forward")
                    _lib.check(lib.dpi_crescent_forward(B, self.npix, self.ng, self.fov, self.prm.data_ptr(),
                                                        self.img.data_ptr(), st), "dpi_crescent_forward")
                    _lib.check(lib.dpi_eht_forward(self.op.handle, B, self.img.data_ptr(), self.vis.data_ptr(), None, None,
                                                   None, self.ws_eht.data_ptr(), self.ws_eht.numel(), st), "dpi_eht_forward")
                    _lib.check(lib.dpi_eht_data_grad(self.op.handle, B, self.vis.data_ptr(), self.cphase_true.data_ptr(),
                                                     self.sigma_cp.data_ptr(), self.logcamp_true.data_ptr(),
                                                     self.sigma_ca.data_ptr(), self.w["cphase"], self.w["camp"],
                                                     self.loss_cp.data_ptr(), self.loss_ca.data_ptr(), self.g_img.data_ptr(),
                                                     self.ws_eht.data_ptr(), self.ws_eht.numel(), st), "dpi_eht_data_grad")
                    # loss_orig = 0.5 * loss_data + logprob (:466): scale_factor 1, data weight 1; softmax(-loss_orig) = alpha 0
                    _lib.check(lib.dpi_alpha_objective(B, self.np_, self.z.data_ptr(), self.loss_cp.data_ptr(),
                                                       self.loss_ca.data_ptr(), self.logdet_flow.data_ptr(), self.det.data_ptr(),
                                                       self.w["cphase"], self.w["camp"], 1.0, 1.0, 0.0, 0, B,
                                                       self.loss_i.data_ptr(), self.gscale.data_ptr(), self.g_logdet.data_ptr(),
                                                       self.terms.data_ptr(), self.stats.data_ptr(), st), "dpi_alpha_objective")
                w = -self.g_logdet.clone()                       # g_logdet = -w_i * scale_factor (= 1)
                out["weights"].append(w)
                out["params"].append(self.prm.clone())
                out["img"].append(self.img.view(B, self.npix, self.npix).clone())
                out["loss_data"].append(self.w["camp"] * self.loss_ca + self.w["cphase"] * self.loss_cp.float())
                if rejsamp:
                    u = torch.rand(B, device=self.device, generator=generator)
                    out["accepted"].append(w / w.max() > u)                   # rej_prob / max(rej_prob) > U (:471-474)
                else:
                    out["accepted"].append(torch.ones(B, dtype=torch.bool, device=self.device))
        finally:
            if was_training:
                self.flow.train()
        res = {k: torch.cat(v) for k, v in out.items()}
        res["img_mean"], res["img_std"] = res["img"].mean(0), res["img"].std(0)
        res["img_map"] = res["img"][torch.argmin(res["loss_data"])]
        return res

    def sample_posterior2(self, n_concat=1, generator=None, z=None):
        """`Gen_samples2` of DPIx_interferometry.py:665-709: the same pass without the rejection step - every sample is
        kept and the posterior summaries are importance-weighted: per batch of n_batch samples w = softmax_i(-loss_orig),
        img_mean = sum_i w_i img_i, img_std = sqrt(sum_i w_i (img_i - img_mean)^2), and the log importance weights
        -loss_orig (:699-707). Returns dict(params, img, weights, log_weights, img_mean [n_concat, npix, npix], img_std,
        img_map)."""
        res = self.sample_posterior(n_concat=n_concat, rejsamp=False, generator=generator, z=z)
        B, n = self.B, self.npix
        w = res["weights"].view(-1, B, 1, 1)
        img = res["img"].view(-1, B, n, n)
        mean = (w * img).sum(1)
        res["img_mean"] = mean
        res["img_std"] = torch.sqrt((w * (img - mean[:, None]) ** 2).sum(1))
        # softmax(-loss_orig) = exp(-loss_orig - logsumexp(-loss_orig)): the script keeps the un-normalised exponent; the
        # normalised one differs by a per-batch constant
        res["log_weights"] = torch.log(res["weights"].clamp_min(1e-45))
        res.pop("accepted")
        return res
