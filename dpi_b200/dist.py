"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The DPI step shards by sample: every rank draws its own z (seed = base + rank), runs the same fused
step, and the flat gradient buffer [flow parameters | log_scale] is summed across ranks once per
step; the clip + Adam kernel then applies 1/world (BatchNorm batch statistics stay per replica -
DDP semantics). Parameters never diverge because every rank applies the same update
to the same initial weights.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str = None, device: torch.device = None):
    """Initialise the default process group from torchrun's environment (no-op for world size 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, **kw)
    return world


def rank_seed(base: int, rank: int) -> int:
    """Per-rank z stream (seed 1234 + rank)."""
    return int(base) + int(rank)


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous sample range of `rank` when a global batch is split (strong scaling)."""
    per = global_batch // world
    rem = global_batch % world
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def gradient_buckets(n_flow: int, block_offsets, total: int, n_buckets: int):
    """Split the flat gradient [flow blocks in parameter order | tail] into `n_buckets` contiguous slices of whole flow
    blocks, in the order the backward pass finishes them. RealNVP.reverse applies block n_flow-1 first, so its
    backward reaches block 0 first: bucket k covers blocks [i0, i1) = execution stages [2 (n_flow - i1), 2 (n_flow - i0))
    and the flat range [block_offsets[i0], block_offsets[i1]) (the last bucket runs to `total`, which includes the
    log_scale tail). Returns [(stage_lo, stage_hi, flat_lo, flat_hi)], or None when fewer than two buckets result."""
    nb = min(int(n_buckets), int(n_flow))
    if nb < 2:
        return None
    per = -(-n_flow // nb)
    out = []
    for i0 in range(0, n_flow, per):
        i1 = min(i0 + per, n_flow)
        lo = int(block_offsets[i0])
        hi = int(block_offsets[i1]) if i1 < n_flow else int(total)
        out.append((2 * (n_flow - i1), 2 * (n_flow - i0), lo, hi))
    return out if len(out) > 1 else None


def allreduce_buckets_(flat: torch.Tensor, buckets, group=None):
    """Asynchronous sum-all-reduce of each bucket's slice (issue order = bucket order); returns the work handles."""
    works = []
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for _, _, lo, hi in buckets:
            works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True))
    return works


def shard_layout(n: int, world: int, align: int = 32):
    """Equal contiguous shards of a flat buffer of `n` floats: (floats per rank - a multiple of `align` -, padded total).
    SURVEY 8(e): reduce-scatter -> per-shard ||g||^2 -> scalar all-reduce -> clip + Adam on the shard -> all-gather."""
    per = -(-int(n) // (int(world) * align)) * align
    return per, per * int(world)


def reduce_scatter_sum_(shard: torch.Tensor, full: torch.Tensor, group=None) -> torch.Tensor:
    """shard <- this rank's slice of the element-wise sum of `full` over all ranks (NCCL reduce-scatter; gloo, which has
    no reduce-scatter, falls back to all-reduce + slice so the host logic can be tested on CPU)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert full.numel() == shard.numel() * world
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(shard, full, op=dist.ReduceOp.SUM, group=group)
    else:
        tmp = full.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        shard.copy_(tmp[rank * shard.numel():(rank + 1) * shard.numel()])
    return shard


def all_gather_shards_(full: torch.Tensor, shard: torch.Tensor, group=None) -> torch.Tensor:
    """full <- concatenation of every rank's shard (in place when `shard` is this rank's slice of `full`)."""
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(full, shard, group=group)
    else:
        parts = [torch.empty_like(shard) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, shard.clone(), group=group)
        full.copy_(torch.cat(parts))
    return full


def broadcast_state_(tensors, src: int = 0, group=None):
    """Make every replica start from rank `src`'s tensors (parameters, BatchNorm running statistics, log_scale): the
    data-parallel step keeps replicas identical only if they start identical."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)


def replicas_max_divergence(flat: torch.Tensor, group=None) -> float:
    """max |flat - rank0's flat| over all ranks (0.0 when the replicas are in sync)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0.0
    ref = flat.clone()
    dist.broadcast(ref, src=0, group=group)
    d = (flat - ref).abs().max().reshape(1)
    dist.all_reduce(d, op=dist.ReduceOp.MAX, group=group)
    return float(d)


def pack_named(model, named: dict, out: torch.Tensor = None) -> torch.Tensor:
    """Pack reference-named tensors (e.g. gradients keyed like state_dict) into the flat layout."""
    flat = torch.zeros_like(model.flat.detach()) if out is None else out
    for key, view in model.named_grad_views(flat):
        view.copy_(named[key])
    return flat
