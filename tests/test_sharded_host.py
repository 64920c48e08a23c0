"""Host logic of the row-sharded tables (K9 protocol) on CPU: pure-function
checks for G in {1,2,4,8} and a real 2-process gloo run of the exchange."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.shard_cpu import CpuShardOps


def _router(world, rank, V, group=None):
    import etr_b200  # noqa: F401
    from etr_b200.sharded import ShardRouter
    return ShardRouter(world, rank, V, CpuShardOps(), group)


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_partition_is_stable_and_invertible(G):
    rng = np.random.default_rng(G)
    V = 1000
    ids = torch.tensor(rng.integers(0, V, size=500))
    send_rows, send_pos, inv_pos, counts = CpuShardOps().partition(ids, G, V)
    assert int(counts.sum()) == ids.numel()
    owner = ids[send_pos] % G
    assert torch.all(owner[1:] >= owner[:-1])                               # grouped by owner
    for g in range(G):                                                      # stable inside a group
        pos = send_pos[owner == g]
        assert torch.all(pos[1:] > pos[:-1])
    assert torch.equal(send_rows * G + owner, ids[send_pos])                # local row <-> global id
    assert torch.equal(send_pos[inv_pos], torch.arange(ids.numel()))       # exact inverse
    assert counts.tolist() == [int((ids % G == g).sum()) for g in range(G)]


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_local_rows_cover_the_table(G):
    V = 1003
    rows = [_router(G, r, V).local_rows() for r in range(G)]
    assert sum(rows) == V
    assert rows == [len(range(r, V, G)) for r in range(G)]


def _worker(rank, world, port, V, width, seed, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(seed)
        full = torch.randn(V, width, generator=g)                           # same on every rank
        shard = full[rank::world].contiguous()                              # this rank's rows
        rng = np.random.default_rng(100 + rank)
        ids = torch.tensor((rng.random(300) ** 3 * V).astype(np.int64))     # Zipf-like, duplicates
        router = _router(world, rank, V)
        route = router.dispatch(ids)
        served = CpuShardOps().take_rows(shard, shard.shape[0], width, width, 0, route.recv_rows)
        rows = router.return_rows(route, served)
        looked_up = rows[route.inv_pos]                                     # un-permute
        ok_fwd = torch.equal(looked_up, full[ids])                          # bit-exact vs the unsharded gather
        # backward: per-slot grads -> owners; every owner scatter-adds into its shard
        grads = torch.randn(ids.numel(), width, generator=torch.Generator().manual_seed(7 + rank))
        recv = router.send_grads(route, grads[route.send_pos])
        local = torch.zeros(shard.shape[0], width, dtype=torch.float64)
        local.index_add_(0, route.recv_rows, recv.double())
        # reference: gather everyone's (ids, grads) and scatter-add densely
        all_ids = [torch.empty_like(ids) for _ in range(world)]
        all_g = [torch.empty_like(grads) for _ in range(world)]
        dist.all_gather(all_ids, ids)
        dist.all_gather(all_g, grads)
        dense = torch.zeros(V, width, dtype=torch.float64)
        for i_, g_ in zip(all_ids, all_g):
            dense.index_add_(0, i_, g_.double())
        ok_bwd = torch.allclose(local, dense[rank::world], atol=1e-12)
        results[rank] = (ok_fwd, ok_bwd, int(sum(route.send_counts)), route.n_recv)
    finally:
        dist.destroy_process_group()


def test_exchange_protocol_over_gloo_world2():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, 29650 + os.getpid() % 200, 997, 20, 3, results), nprocs=world, join=True)
    assert len(results) == world
    for r in range(world):
        ok_fwd, ok_bwd, n_sent, n_recv = results[r]
        assert ok_fwd and ok_bwd, (r, ok_fwd, ok_bwd)
        assert n_sent == 300
    assert sum(results[r][3] for r in range(world)) == 600


# ------------------------------------------------------------------ peer form (request / serve / deferred gradient)
def _peer_worker(rank, world, port, V, k, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.peer_exchange_cpu import PeerExchangeCpu
        F, B = 5, 64
        full = torch.randn(V, k + 1, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
        px = PeerExchangeCpu(world, rank, full[rank::world].clone(), k)
        rng = np.random.default_rng(50 + rank)
        ok = True
        for step in range(3):                                              # 3 steps: the accumulator is never cleared
            X = torch.tensor((rng.random((B, F)) ** 3 * V).astype(np.int64))
            g = torch.tensor(rng.normal(size=B))
            dflat = torch.tensor(rng.normal(size=(B, F, k)))
            rows, st = px.exchange_forward(X)
            ok &= torch.equal(rows, full[X])                               # gathered rows bit-exact
            deferred = px.export_deferred(st, rows, g, dflat)
            t, grad = px.push_and_apply(st, deferred)
            # the request-driven owner pass (leaders picked at random, as the device race would) gives the SAME bits
            t2, grad2 = px.push_and_apply_paired(st, deferred, rng)
            o1, o2 = torch.argsort(t), torch.argsort(t2)
            ok &= torch.equal(t[o1], t2[o2]) and torch.equal(grad[o1], grad2[o2])
            # reference: autograd on the GLOBAL batch through the unsharded FM expression
            Xs, gs, ds = [torch.empty_like(X) for _ in range(world)], [torch.empty_like(g) for _ in range(world)], \
                [torch.empty_like(dflat) for _ in range(world)]
            dist.all_gather(Xs, X); dist.all_gather(gs, g); dist.all_gather(ds, dflat)
            XA, gA, dA = torch.cat(Xs), torch.cat(gs), torch.cat(ds)
            tab = full.clone().requires_grad_(True)
            e = tab[XA]                                                    # [B*,F,k+1]
            v, w = e[..., :k], e[..., k]
            z = w.sum(1) + 0.5 * ((v.sum(1) ** 2).sum(1) - (v ** 2).sum((1, 2)))
            ((gA * z).sum() + (dA * v).sum()).backward()
            ref = tab.grad[rank::world]
            mine = torch.zeros_like(ref)
            mine[t] = grad
            ok &= torch.allclose(mine, ref, atol=1e-10)
            ok &= int((ref.abs().sum(1) > 0).sum()) == t.numel()           # touched list == rows with a gradient
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_peer_exchange_protocol_over_gloo_world2():
    """request / serve / virtual ids / deferred FM gradient / rank-ordered stamped accumulation, stated on CPU
    (oracle/peer_exchange_cpu.py) and run with 2 gloo ranks against the unsharded FM expression under autograd."""
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_peer_worker, args=(world, 29850 + os.getpid() % 100, 499, 4, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}
