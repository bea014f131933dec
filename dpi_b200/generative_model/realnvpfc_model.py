"""Drop-in for DPItorch/generative_model/realnvpfc_model.py backed by libdpi_b200 (sm_100a).

Public surface kept from the reference:

    RealNVP(ndim, n_flow, affine=True, seqfrac=4, permute='random', batch_norm=True)
        .forward(x)  -> (z, logdet[B])          realnvpfc_model.py:583-592
        .reverse(z)  -> (x, logdet[B])          realnvpfc_model.py:594-603
        .orders / .inverse_orders               lists of numpy int arrays (:568-581)
        .blocks[i].actnorm/.actnorm2/.coupling/.coupling2   (read-only views here)
        .train() / .eval()                      BatchNorm batch vs running statistics
        .state_dict() / .load_state_dict()      the reference's keys and shapes, both ways

What differs, by design: every trainable tensor lives in ONE flat fp32 parameter (`self.flat`,
laid out in the reference's `parameters()` order, tensor starts padded to 32 floats) so that the
optimiser, the gradient all-reduce and the clip touch one buffer. `parameters()` therefore
yields that single tensor; `named_parameter_views()` gives the reference-named views.
Compute runs only on CUDA through `torch.autograd.Function`s over the C ABI - no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch
from torch import nn

from dpi_b200 import _lib

# tensor ids of include/dpi_b200.h
_T_ACT = ("actnorm.loc", "actnorm.log_scale_inv", "actnorm2.loc", "actnorm2.log_scale_inv")
_T_CPL_BN = ("net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias", "net.3.weight", "net.3.bias",
             "net.5.weight", "net.5.bias", "net.6.scale", "net.6.fc.weight", "net.6.fc.bias")
_T_CPL_NOBN = ("net.0.weight", "net.0.bias", None, None, "net.2.weight", "net.2.bias",
               None, None, "net.4.scale", "net.4.fc.weight", "net.4.fc.bias")
_N_TENSOR_IDS = 4 + 2 * 11


def Order_inverse(order):
    """Inverse permutation (reference: O(D^2) double loop at realnvpfc_model.py:553-559)."""
    order = np.asarray(order)
    inv = np.empty_like(order)
    inv[order] = np.arange(len(order), dtype=order.dtype)
    return inv


class _FlowFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, direction, train, inp, flat):
        lib = _lib.load()
        if not inp.is_cuda:
            raise RuntimeError("dpi_b200.RealNVP runs on CUDA only (no CPU fallback); got a CPU tensor")
        if inp.dtype != torch.float32 or flat.dtype != torch.float32:
            raise RuntimeError("dpi_b200.RealNVP computes in float32")
        if inp.dim() != 2 or inp.shape[1] != model.ndim:
            raise RuntimeError(f"expected input [B, {model.ndim}], got {tuple(inp.shape)}")
        inp = inp.contiguous()
        B = inp.shape[0]
        h = model._handle(inp.device)
        nbytes = lib.dpi_flow_workspace_bytes(h, B)
        if nbytes == 0:
            _lib.check(-1, "dpi_flow_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=inp.device)
        out = torch.empty_like(inp)
        logdet = torch.empty(B, dtype=torch.float32, device=inp.device)
        _lib.check(lib.dpi_flow_run(h, direction, int(train), B, flat.data_ptr(), model.bn_running.data_ptr(),
                                    inp.data_ptr(), out.data_ptr(), logdet.data_ptr(), ws.data_ptr(), nbytes,
                                    _lib.stream_ptr()), "dpi_flow_run")
        ctx.model, ctx.direction, ctx.train, ctx.ws, ctx.handle = model, direction, train, ws, h
        ctx.save_for_backward(inp, flat)
        model._last_ws = ws
        return out, logdet

    @staticmethod
    def backward(ctx, g_out, g_logdet):
        lib = _lib.load()
        inp, flat = ctx.saved_tensors
        if ctx.direction != 0 or not ctx.train:
            raise NotImplementedError("dpi_b200: gradients are implemented for RealNVP.reverse in train() mode "
                                      "(the DPI training path); forward()/eval() are inference-only")
        B = inp.shape[0]
        g_out = torch.zeros_like(inp) if g_out is None else g_out.contiguous().float()
        g_logdet = torch.zeros(B, device=inp.device) if g_logdet is None else g_logdet.contiguous().float()
        grads = torch.zeros_like(flat)
        g_in = torch.empty_like(inp) if ctx.needs_input_grad[3] else None
        with torch.cuda.device(inp.device):
            _lib.check(lib.dpi_flow_backward(ctx.handle, B, flat.data_ptr(), grads.data_ptr(), inp.data_ptr(),
                                             g_out.data_ptr(), g_logdet.data_ptr(), _lib.ptr(g_in),
                                             ctx.ws.data_ptr(), ctx.ws.numel(), _lib.stream_ptr()),
                       "dpi_flow_backward")
        return None, None, None, g_in, grads


class _View:
    """Read-only attribute tree mirroring the reference modules (`.blocks[i].coupling.net[0].weight`)."""

    def __init__(self, model, prefix):
        self._m, self._p = model, prefix

    def _get(self, name):
        return self._m._view(self._p + name)


class ActNorm(_View):
    """View of one scalar ActNorm (realnvpfc_model.py:8-66)."""
    loc = property(lambda s: s._get("loc"))
    log_scale_inv = property(lambda s: s._get("log_scale_inv"))
    initialized = property(lambda s: s._m._initialized_tensor(s._p))


class _Linear(_View):
    weight = property(lambda s: s._get("weight"))
    bias = property(lambda s: s._get("bias"))


class _BatchNorm(_View):
    weight = property(lambda s: s._get("weight"))
    bias = property(lambda s: s._get("bias"))
    running_mean = property(lambda s: s._get("running_mean"))
    running_var = property(lambda s: s._get("running_var"))


class ZeroFC(_View):
    """View of the zero-initialised output layer (realnvpfc_model.py:128-141)."""
    scale = property(lambda s: s._get("scale"))

    @property
    def fc(self):
        return _Linear(self._m, self._p + "fc.")


class AffineCoupling(_View):
    """View of one coupling net (realnvpfc_model.py:143-257)."""

    @property
    def net(self):
        m, p, bn = self._m, self._p + "net.", self._m.batch_norm
        if bn:
            return {0: _Linear(m, p + "0."), 2: _BatchNorm(m, p + "2."), 3: _Linear(m, p + "3."),
                    5: _BatchNorm(m, p + "5."), 6: ZeroFC(m, p + "6.")}
        return {0: _Linear(m, p + "0."), 2: _Linear(m, p + "2."), 4: ZeroFC(m, p + "4.")}


class Flow(_View):
    """View of one flow block (realnvpfc_model.py:427-474)."""
    actnorm = property(lambda s: ActNorm(s._m, s._p + "actnorm."))
    actnorm2 = property(lambda s: ActNorm(s._m, s._p + "actnorm2."))
    coupling = property(lambda s: AffineCoupling(s._m, s._p + "coupling."))
    coupling2 = property(lambda s: AffineCoupling(s._m, s._p + "coupling2."))


class RealNVP(nn.Module):
    def __init__(self, ndim, n_flow, affine=True, seqfrac=4, permute='random', batch_norm=True):
        super().__init__()
        if not affine:
            raise NotImplementedError("dpi_b200.RealNVP implements the affine coupling used by every DPI script")
        self.ndim, self.n_flow, self.batch_norm = int(ndim), int(n_flow), bool(batch_norm)
        self.hidden = int(ndim / (2 * seqfrac))  # realnvpfc_model.py:176
        if self.hidden < 1 or self.ndim < 2:
            raise ValueError("ndim/seqfrac too small")
        self.orders = []
        for i in range(n_flow):  # :568-577
            if permute == 'random':
                self.orders.append(np.random.RandomState(seed=i).permutation(ndim))
            elif permute == 'reverse':
                self.orders.append(np.arange(ndim - 1, -1, -1))
            else:
                print('We can only do no permutation, random permutation or reverse permutation in affine '
                      'coupling layer. Using no permutation by default!')
                self.orders.append(np.arange(ndim))
        self.inverse_orders = [Order_inverse(o) for o in self.orders]

        lib = _lib.load()
        offs = np.full((n_flow, _N_TENSOR_IDS), -1, dtype=np.int64)
        bn_offs = np.full((n_flow, 2, 4), -1, dtype=np.int64)
        n_params, n_bn = C.c_int64(0), C.c_int64(0)
        _lib.check(lib.dpi_flow_layout(self.ndim, self.n_flow, self.hidden, int(self.batch_norm),
                                       offs.ctypes.data, bn_offs.ctypes.data, C.addressof(n_params),
                                       C.addressof(n_bn)), "dpi_flow_layout")
        self._index: "OrderedDict[str, Tuple[str, int, Tuple[int, ...]]]" = OrderedDict()
        Da, Dout, H = ndim - ndim // 2, 2 * (ndim // 2), self.hidden
        cpl_names = _T_CPL_BN if self.batch_norm else _T_CPL_NOBN
        cpl_shapes = ((H, Da), (H,), (H,), (H,), (H, H), (H,), (H,), (H,), (Dout,), (Dout, H), (Dout,))
        bn_layers = ("2", "5")
        for i in range(n_flow):  # state_dict order of the reference module tree
            for a, an in enumerate(("actnorm", "actnorm2")):
                self._index[f"blocks.{i}.{an}.loc"] = ("p", int(offs[i, 2 * a]), (1,))
                self._index[f"blocks.{i}.{an}.log_scale_inv"] = ("p", int(offs[i, 2 * a + 1]), (1,))
                self._index[f"blocks.{i}.{an}.initialized"] = ("init", (i, a), ())
            for c, cn in enumerate(("coupling", "coupling2")):
                base = f"blocks.{i}.{cn}."
                for t, (name, shape) in enumerate(zip(cpl_names, cpl_shapes)):
                    if name is None:
                        continue
                    self._index[base + name] = ("p", int(offs[i, 4 + 11 * c + t]), shape)
                    if self.batch_norm and name in ("net.2.bias", "net.5.bias"):
                        l = bn_layers.index(name.split(".")[1])
                        self._index[base + f"net.{bn_layers[l]}.running_mean"] = ("r", int(bn_offs[i, c, 2 * l]), (H,))
                        self._index[base + f"net.{bn_layers[l]}.running_var"] = ("r", int(bn_offs[i, c, 2 * l + 1]), (H,))
                        self._index[base + f"net.{bn_layers[l]}.num_batches_tracked"] = ("nbt", (i, c, l), ())

        flat = torch.zeros(int(n_params.value), dtype=torch.float32)
        bn = torch.zeros(int(n_bn.value), dtype=torch.float32)
        self.flat = nn.Parameter(flat)
        self.register_buffer("bn_running", bn)
        self._initialized = np.zeros((n_flow, 2), dtype=np.uint8)
        self._nbt = np.zeros((n_flow, 2, 2), dtype=np.int64)
        self._handles: Dict[int, int] = {}
        self._last_ws = None
        self._flow_mode = None
        self.reset_parameters()
        object.__setattr__(self, "blocks", [Flow(self, f"blocks.{i}.") for i in range(n_flow)])

    # ------------------------------------------------------------------ parameters / state
    def reset_parameters(self):
        """Reference initialisation: N(0, 0.05) first two Linear weights, zero biases, BatchNorm (1, 0),
        zero ZeroFC and ActNorm (realnvpfc_model.py:11-12, 131-135, 189-193)."""
        with torch.no_grad():
            self.flat.zero_()
            self.bn_running.zero_()
            for key, (kind, off, shape) in self._index.items():
                if kind == "p":
                    v = self._view(key)
                    tail = key.split("coupling")[-1] if "coupling" in key else ""
                    if tail.endswith("net.0.weight") or (tail.endswith("net.3.weight") and self.batch_norm) or \
                            (tail.endswith("net.2.weight") and not self.batch_norm):
                        v.normal_(0, 0.05)
                    elif self.batch_norm and (tail.endswith("net.2.weight") or tail.endswith("net.5.weight")):
                        v.fill_(1.0)
                elif kind == "r" and key.endswith("running_var"):
                    self._view(key).fill_(1.0)
        self._initialized[:] = 0
        self._nbt[:] = 0

    def _view(self, key: str) -> torch.Tensor:
        kind, off, shape = self._index[key]
        n = int(np.prod(shape)) if shape else 1
        if kind == "p":
            return self.flat.detach()[off:off + n].view(shape)
        if kind == "r":
            return self.bn_running[off:off + n].view(shape)
        raise KeyError(key)

    def _initialized_tensor(self, prefix: str) -> torch.Tensor:
        kind, (i, a), _ = self._index[prefix + "initialized"]
        return torch.tensor(int(self._initialized[i, a]), dtype=torch.uint8)

    def named_parameter_views(self):
        """(reference parameter name, view into the flat parameter) in `parameters()` order."""
        for key, (kind, off, shape) in self._index.items():
            if kind == "p":
                yield key, self._view(key)

    def named_grad_views(self, grad: torch.Tensor = None):
        """(reference parameter name, view into a flat gradient) - default: `self.flat.grad`."""
        g = self.flat.grad if grad is None else grad
        for key, (kind, off, shape) in self._index.items():
            if kind == "p":
                n = int(np.prod(shape))
                yield key, g[off:off + n].view(shape)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for key, (kind, off, shape) in self._index.items():
            if kind in ("p", "r"):
                v = self._view(key)
                destination[prefix + key] = v if keep_vars else v.detach().clone()
            elif kind == "init":
                destination[prefix + key] = torch.tensor(int(self._initialized[off]), dtype=torch.uint8)
            else:
                destination[prefix + key] = torch.tensor(int(self._nbt[off]), dtype=torch.int64)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        if prefix + "flat" in state_dict:  # our own flat format
            with torch.no_grad():
                self.flat.copy_(state_dict[prefix + "flat"])
                if prefix + "bn_running" in state_dict:
                    self.bn_running.copy_(state_dict[prefix + "bn_running"])
            self._initialized[:] = 1
            return
        seen = set()
        with torch.no_grad():
            for key, (kind, off, shape) in self._index.items():
                full = prefix + key
                if full not in state_dict:
                    missing_keys.append(full)
                    continue
                seen.add(full)
                val = state_dict[full]
                if kind in ("p", "r"):
                    if tuple(val.shape) != tuple(shape):
                        error_msgs.append(f"size mismatch for {full}: copying a param with shape "
                                          f"{tuple(val.shape)} from checkpoint, the shape in current model is {shape}.")
                        continue
                    self._view(key).copy_(val)
                elif kind == "init":
                    self._initialized[off] = int(val)
                else:
                    self._nbt[off] = int(val)
        if strict:
            for k in state_dict:
                if k.startswith(prefix) and k not in seen:
                    unexpected_keys.append(k)

    # ------------------------------------------------------------------ native handle
    def _handle(self, device: torch.device) -> int:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            lib = _lib.load()
            orders = np.ascontiguousarray(np.stack(self.orders).astype(np.int32))
            out = C.c_void_p()
            _lib.check(lib.dpi_flow_create(C.byref(out), idx, self.ndim, self.n_flow, self.hidden,
                                           int(self.batch_norm), orders.ctypes.data), "dpi_flow_create")
            h = out.value
            self._handles[idx] = h
            if self._flow_mode is not None:
                _lib.check(lib.dpi_flow_set_mode(h, *self._flow_mode), "dpi_flow_set_mode")
        return h

    def set_flow_mode(self, persistent: bool, ctas_per_sm: int = 0):
        """persistent=True: one cooperative launch per pass; False: one launch per phase (debug)."""
        self._flow_mode = (int(persistent), int(ctas_per_sm))
        for h in self._handles.values():
            _lib.check(_lib.load().dpi_flow_set_mode(h, *self._flow_mode), "dpi_flow_set_mode")

    def __del__(self):
        try:
            lib = _lib.load()
            for h in self.__dict__.get("_handles", {}).values():
                lib.dpi_flow_destroy(h)
        except Exception:
            pass

    # A native handle is a raw pointer owned by exactly one Python object: copies (copy.deepcopy for an EMA / best-model
    # snapshot) and pickles (torch.save(model)) drop it and re-create their own lazily on first use.
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_handles"] = {}
        state["_last_ws"] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_handles", "_last_ws", "blocks"):
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        new.__dict__["_handles"] = {}
        new.__dict__["_last_ws"] = None
        object.__setattr__(new, "blocks", [Flow(new, f"blocks.{i}.") for i in range(self.n_flow)])
        return new

    def gather_map(self, direction: int, stage: int) -> np.ndarray:
        """The fused permutation map the kernels use for execution stage `stage` (direction 0 = reverse,
        1 = forward, 2 = trailing map of forward)."""
        out = np.empty(self.ndim, dtype=np.int32)
        dev = self.flat.device if self.flat.is_cuda else torch.device("cuda", torch.cuda.current_device())
        _lib.check(_lib.load().dpi_flow_gather_map(self._handle(dev), direction, stage, out.ctypes.data),
                   "dpi_flow_gather_map")
        return out

    # ------------------------------------------------------------------ ActNorm initialisation
    def _init_reverse(self):
        """First `reverse` call on a fresh ActNorm zero-initialises it (realnvpfc_model.py:23-25,52-54)."""
        if self._initialized.all():
            return
        with torch.no_grad():
            for i in range(self.n_flow):
                for a, an in enumerate(("actnorm", "actnorm2")):
                    if not self._initialized[i, a]:
                        self._view(f"blocks.{i}.{an}.loc").zero_()
                        self._view(f"blocks.{i}.{an}.log_scale_inv").zero_()
        self._initialized[:] = 1

    def _init_forward(self, x):
        """Data-dependent init on the first `forward` (realnvpfc_model.py:18-28,33-35): loc = -mean,
        log_scale_inv = log(std + 1e-6) over the whole input of that ActNorm. The input of the ActNorm
        of execution stage e is a permutation of the previous stage's output, so its global mean / std
        are read from the saved stage outputs. One-off host-driven loop, not on the hot path."""
        lib = _lib.load()
        B = x.shape[0]
        h = self._handle(x.device)
        for e in range(2 * self.n_flow):
            i, a = e // 2, e % 2
            if self._initialized[i, a]:
                continue
            an = "actnorm" if a == 0 else "actnorm2"
            with torch.no_grad():
                if e == 0:
                    src = x
                else:
                    saved_bn, saved_nbt = self.bn_running.clone(), self._nbt.copy()
                    _FlowFunction.apply(self, 1, self.training, x, self.flat.detach())
                    self.bn_running.copy_(saved_bn)
                    self._nbt = saved_nbt
                    off = lib.dpi_flow_ws_offset(h, B, 0, e - 1)
                    ld = lib.dpi_flow_ws_offset(h, B, 5, 0)       # row stride of the saved [ndim][ld] matrices
                    src = self._last_ws[off:off + 4 * ld * self.ndim].view(torch.float32).view(self.ndim, ld)[:, :B]
                self._view(f"blocks.{i}.{an}.loc").copy_(-src.mean().reshape(1))
                self._view(f"blocks.{i}.{an}.log_scale_inv").copy_(torch.log(src.std().reshape(1) + 1e-6))
            self._initialized[i, a] = 1

    # ------------------------------------------------------------------ the two flow directions
    def _bump_bn(self):
        if self.training and self.batch_norm:
            self._nbt += 1

    def forward(self, input):
        """x -> (z, logdet): blocks in order, then `out[:, orders[i]]` (realnvpfc_model.py:583-592)."""
        if not self._initialized.all():
            self._init_forward(input)
        out, logdet = _FlowFunction.apply(self, 1, self.training or not self.batch_norm, input, self.flat)
        self._bump_bn()
        return out, logdet

    def reverse(self, out):
        """z -> (x, logdet): blocks in reverse order (realnvpfc_model.py:594-603)."""
        self._init_reverse()
        x, logdet = _FlowFunction.apply(self, 0, self.training or not self.batch_norm, out, self.flat)
        self._bump_bn()
        return x, logdet
