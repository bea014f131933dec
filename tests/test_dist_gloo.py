"""CPU, world_size 2 over gloo: the data-parallel host logic (flat gradient packing, one all-reduce,
1/world scaling, identical update on every replica) with the oracle supplying per-shard gradients."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_grads(sd, z, cx, cl):
    from oracle import dpi_oracle as O
    sd = {k: v.clone() for k, v in sd.items()}
    keys = O.param_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
    x, ld = O.realnvp_reverse(sd, z, train=True)
    ((x * cx).sum() / z.shape[0] + (ld * cl).sum() / z.shape[0]).backward()
    return {k: sd[k].grad for k in keys}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    from dpi_b200 import dist as ddist
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    from oracle import dpi_oracle as O
    assert ddist.init_from_env("gloo") == world
    D, NF, B = 16, 2, 8
    sd = O.init_state(D, NF, seqfrac=1, seed=5, randomize=True)
    model = RealNVP(D, NF, seqfrac=1)
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(ddist.rank_seed(1234, rank))
    z = torch.randn(B, D, generator=g)
    cx, cl = torch.ones(B, D), torch.full((B,), -0.01)
    flat_g = ddist.pack_named(model, _shard_grads(sd, z, cx, cl))
    ddist.allreduce_sum_(flat_g)
    # what the fused kernel does with grad_scale = 1/world, restated: clip on the averaged gradient, Adam
    gavg = flat_g / world
    norm, coef = O.clip_grad_norm([gavg], 1e-3)
    p = model.flat.detach().clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    O.adam_update(p, gavg * coef, m, v, 1, 1e-4)
    assert ddist.replicas_max_divergence(p) == 0.0           # replicas stay bit-identical
    # bucketed variant (the tcgen05 trainer path): whole flow blocks in the order the backward finishes them, each slice
    # reduced asynchronously - must equal the single all-reduce bit for bit
    offs = [int(model._index[f"blocks.{i}.actnorm.loc"][1]) for i in range(NF)]
    total = flat_g.numel()
    buckets = ddist.gradient_buckets(NF, offs, total, 2)
    assert buckets == [(2, 4, 0, offs[1]), (0, 2, offs[1], total)]
    flat_b = ddist.pack_named(model, _shard_grads(sd, z, cx, cl))
    for w in ddist.allreduce_buckets_(flat_b, buckets):
        w.wait()
    assert torch.equal(flat_b, flat_g)
    # SURVEY 8(e) sharded optimiser: reduce-scatter -> per-shard ||g||^2 -> scalar all-reduce -> clip + Adam on the shard ->
    # all-gather. Must reproduce the replicated update bit for bit (same element-wise arithmetic, the norm is a sum of the
    # per-shard sums) and leave every replica with identical parameters.
    n = flat_g.numel() - 7                                # a length that does not divide evenly: the tail is padding
    per, n_all = ddist.shard_layout(n, world)
    assert per % 32 == 0 and n_all == per * world and n_all >= n and n_all - n < world * 32
    g_full = torch.zeros(n_all)
    g_full[:n] = ddist.pack_named(model, _shard_grads(sd, z, cx, cl))[:n]
    p_full = torch.zeros(n_all)
    p_full[:n] = model.flat.detach()[:n]
    if rank == 1:
        p_full[:n] += 1.0                                   # a replica that starts different ...
    ddist.broadcast_state_([p_full], 0)                     # ... is overwritten with rank 0's state
    g_sh = torch.empty(per)
    ddist.reduce_scatter_sum_(g_sh, g_full)
    kk = max(0, min(per, n - rank * per))
    assert torch.equal(g_sh[:kk], flat_g[rank * per:rank * per + kk]) and float(g_sh[kk:].abs().sum()) == 0.0
    sq = (g_sh.double() ** 2).sum().reshape(1)
    dist.all_reduce(sq)
    norm_z = float(torch.sqrt(sq)) / world
    assert abs(norm_z - float(norm)) <= 1e-6 * float(norm)
    coef_z = min(1.0, 1e-3 / (norm_z + 1e-6))
    p_sh = p_full[rank * per:(rank + 1) * per]
    m_sh, v_sh = torch.zeros(per), torch.zeros(per)
    O.adam_update(p_sh, g_sh / world * coef_z, m_sh, v_sh, 1, 1e-4)
    ddist.all_gather_shards_(p_full, p_sh)
    assert ddist.replicas_max_divergence(p_full) == 0.0
    assert float((p_full[:n] - p[:n]).abs().max()) <= 2e-7 * float(p.abs().max())
    assert float(p_full[n:].abs().max()) == 0.0             # padding never moves
    lo, hi = ddist.shard_range(10, rank, world)
    assert (hi - lo) == 5 and lo == 5 * rank
    if rank == 0:
        torch.save(dict(p=p, gavg=gavg, norm=norm), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average_matches_single_process(tmp_path):
    from dpi_b200 import dist as ddist
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    from oracle import dpi_oracle as O
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    D, NF, B = 16, 2, 8
    sd = O.init_state(D, NF, seqfrac=1, seed=5, randomize=True)
    model = RealNVP(D, NF, seqfrac=1)
    model.load_state_dict(sd)
    cx, cl = torch.ones(B, D), torch.full((B,), -0.01)
    tot = torch.zeros_like(model.flat.detach())
    for r in range(2):
        g = torch.Generator().manual_seed(ddist.rank_seed(1234, r))
        tot += ddist.pack_named(model, _shard_grads(sd, torch.randn(B, D, generator=g), cx, cl))
    err = float((got["gavg"] - tot / 2).abs().max() / (tot / 2).abs().max())
    assert err < 1e-5, err   # thread-count dependent summation order inside the CPU oracle
    # padding of the flat layout stays zero, so the flat norm equals the norm over named tensors
    named = dict(model.named_grad_views(got["gavg"]))
    assert np.isclose(float(got["norm"]), float(torch.sqrt(sum((t.double() ** 2).sum() for t in named.values()))),
                      rtol=1e-6)


def test_gradient_buckets_cover_the_buffer_in_backward_order():
    from dpi_b200.dist import gradient_buckets
    offs = [0, 100, 250, 400, 900]                      # 5 flow blocks, tail up to 1000
    b = gradient_buckets(5, offs, 1000, 4)              # ceil(5/4) = 2 blocks per bucket -> 3 buckets
    assert b == [(6, 10, 0, 250), (2, 6, 250, 900), (0, 2, 900, 1000)]
    assert b[0][1] == 10 and b[-1][0] == 0              # stages run from S = 10 down to 0 without gaps
    assert all(x[0] == y[1] for x, y in zip(b[:-1], b[1:]))
    assert all(x[3] == y[2] for x, y in zip(b[:-1], b[1:]))   # flat slices tile [0, total)
    assert gradient_buckets(5, offs, 1000, 1) is None and gradient_buckets(1, [0], 10, 4) is None
    assert gradient_buckets(3, [0, 10, 20], 30, 8) == [(4, 6, 0, 10), (2, 4, 10, 20), (0, 2, 20, 30)]
