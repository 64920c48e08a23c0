"""Sharded-table manager: embedding tables too big for one
GPU are row-sharded across the G GPUs of one box -- owner(id) = id mod G, local
row = id div G (NVSwitch gives uniform bandwidth to every peer, so plain modulo
sharding is balanced and Zipf heads spread over the ranks).

Forward:  stable partition of the rank's ids by owner (K9 kernel) -> counts
all-to-all -> ids all-to-all -> owners gather their rows (bit-exact row dump)
-> rows all-to-all back -> the received buffer IS the table of the fused
gather+interaction kernel, indexed by the inverse permutation (no un-permute
pass, routing deterministic).
Backward: per-slot gradient rows -> grouped by owner -> all-to-all -> each owner
runs the ordinary sorted-ID segment reduction + Adam on its local shard.

The collectives are ``torch.distributed.all_to_all_single`` (NCCL over NVLink on
GPUs; gloo in the CPU protocol tests).  The device work is injected as ``ops``
(``CudaShardOps`` in the product; the tests inject ``oracle.shard_cpu`` to run the
very same protocol over gloo without a GPU).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


class Route:
    """Everything one dispatch decided: needed again for the return trip and the backward."""

    def __init__(self, n, send_pos, inv_pos, send_counts: List[int], recv_counts: List[int], recv_rows):
        self.n, self.send_pos, self.inv_pos = n, send_pos, inv_pos
        self.send_counts, self.recv_counts, self.recv_rows = send_counts, recv_counts, recv_rows

    @property
    def n_recv(self) -> int:
        return int(sum(self.recv_counts))


class ShardRouter:
    """The exchange protocol, independent of where the tensors live."""

    def __init__(self, world: int, rank: int, rows_global: int, ops, group=None):
        assert 0 <= rank < world
        self.world, self.rank, self.rows_global, self.ops, self.group = world, rank, int(rows_global), ops, group

    def local_rows(self, rank: Optional[int] = None) -> int:
        r = self.rank if rank is None else rank
        return (self.rows_global - r + self.world - 1) // self.world

    def _a2a(self, out, inp, out_splits, in_splits):
        if self.world == 1:
            out.copy_(inp)
        else:
            dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)
        return out

    def dispatch(self, ids_flat: torch.Tensor) -> Route:
        """ids_flat int64 [n] (global ids) -> Route; afterwards ``route.recv_rows`` holds the
        LOCAL rows this rank must serve, grouped by requesting rank."""
        n = ids_flat.numel()
        send_rows, send_pos, inv_pos, counts = self.ops.partition(ids_flat, self.world, self.rows_global)
        if self.world == 1:
            return Route(n, send_pos, inv_pos, [n], [n], send_rows)
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts, group=self.group)
        sc = [int(x) for x in counts.tolist()]             # host sync: 2*G integers per step
        rc = [int(x) for x in recv_counts.tolist()]
        recv_rows = send_rows.new_empty((sum(rc),))
        self._a2a(recv_rows, send_rows, rc, sc)
        return Route(n, send_pos, inv_pos, sc, rc, recv_rows)

    def return_rows(self, route: Route, rows_local: torch.Tensor) -> torch.Tensor:
        """rows_local [n_recv, ld] (served rows, request order) -> [n, ld] in this rank's send order."""
        out = rows_local.new_empty((route.n, rows_local.shape[1]))
        return self._a2a(out, rows_local, route.send_counts, route.recv_counts)

    def send_grads(self, route: Route, grads_sorted: torch.Tensor) -> torch.Tensor:
        """grads_sorted [n, ld] in send order -> [n_recv, ld] aligned with ``route.recv_rows``."""
        out = grads_sorted.new_empty((route.n_recv, grads_sorted.shape[1]))
        return self._a2a(out, grads_sorted, route.recv_counts, route.send_counts)


class CudaShardOps:
    """Device work of the protocol: the K9 partition kernel and bit-exact row gathers."""

    def __init__(self, rt):
        self.rt = rt

    def partition(self, ids_flat: torch.Tensor, world: int, rows_global: int):
        rt = self.rt
        ids_flat = ids_flat.contiguous()
        n = ids_flat.numel()
        send_rows = rt.empty((n,), torch.int64)
        send_pos = rt.empty((n,), torch.int64)
        inv_pos = rt.empty((n,), torch.int64)
        counts = rt.empty((world,), torch.int32)
        check(rt.lib.etr_shard_partition(rt.ctx, ids_flat.data_ptr(), n, world, rows_global, send_rows.data_ptr(),
                                         send_pos.data_ptr(), inv_pos.data_ptr(), counts.data_ptr(), rt.stream))
        return send_rows, send_pos, inv_pos, counts

    def take_rows(self, src: torch.Tensor, rows: int, width: int, stride: int, dtype_code: int,
                  idx: torch.Tensor) -> torch.Tensor:
        """out[j, 0:width] = src[idx[j], 0:width] as fp32 [m, stride_out] (padding columns zero)."""
        rt = self.rt
        m = idx.numel()
        ld = (width + 3) // 4 * 4
        out = rt.zeros((m, ld)) if ld != width else rt.empty((m, ld))
        t = _lib.etr_table(src.data_ptr(), rows, width, stride, dtype_code, 0)
        check(rt.lib.etr_embedding_gather(rt.ctx, C.byref(t), idx.data_ptr(), m, out.data_ptr(), ld, rt.stream))
        return out


class VirtualTable:
    """The rows a rank received for its batch, presented as an fp32 table [n, ld] so the fused
    gather + interaction kernels run on it unchanged (ids = inverse permutation)."""

    def __init__(self, rt, rows_buf: torch.Tensor, width: int):
        self.rt, self.data, self.rows, self.width, self.stride = rt, rows_buf, rows_buf.shape[0], width, rows_buf.shape[1]
        self.dtype = torch.float32

    def desc(self) -> _lib.etr_table:
        return _lib.etr_table(self.data.data_ptr(), self.rows, self.width, self.stride, _lib.ETR_F32, 0)

    @property
    def grad_ld(self) -> int:
        return self.stride


class ShardedTable:
    """A [rows_global, width] table of which this rank holds rows rank, rank+G, rank+2G, ..."""

    def __init__(self, rt, rows_global: int, width: int, world: int, rank: int, dtype=torch.float32, group=None):
        from .runtime import EmbeddingTable
        self.rt, self.world, self.rank, self.rows_global, self.width = rt, world, rank, int(rows_global), width
        self.router = ShardRouter(world, rank, rows_global, CudaShardOps(rt), group)
        self.local = EmbeddingTable(rt, max(self.router.local_rows(), 1), width, dtype)

    def load_global(self, full: torch.Tensor, c0: int = 0):
        """Take this rank's rows out of a full [V, w] tensor (tests / checkpoint import)."""
        mine = full[self.rank:: self.world].to(self.rt.device)
        self.local.data[: mine.shape[0], c0:c0 + mine.shape[1]] = mine.to(self.local.dtype)

    def lookup(self, ids):
        """IdsBatch (single-hot) -> (VirtualTable, virtual IdsBatch, Route)."""
        from .runtime import IdsBatch, _TORCH2ETR
        assert not ids.is_bag, "sharded tables: single-hot ids only in this round"
        rt, ops = self.rt, self.router.ops
        # flatten the batch's ids in (b, f) order whatever the strides are
        X = torch.as_strided(ids.ids, (ids.B, ids.F), (ids.sb, ids.sf)).reshape(-1)
        route = self.router.dispatch(X)
        served = ops.take_rows(self.local.data, self.local.rows, self.local.width, self.local.stride,
                               _TORCH2ETR[self.local.dtype], route.recv_rows)
        rows = self.router.return_rows(route, served)                   # [n, ld], send order
        vt = VirtualTable(rt, rows, self.width)
        vids = IdsBatch(rt, route.inv_pos, ids.B, ids.F, 1, ids.F, 1, 1)
        return vt, vids, route

    def sparse_grad(self, route: Route, bag_grad: torch.Tensor):
        """bag_grad [n, ld] (one row per slot, original order) -> SparseGrad on the LOCAL shard."""
        from .runtime import IdsBatch, SparseGrad
        ops = self.router.ops
        ld = bag_grad.shape[1]
        grouped = ops.take_rows(bag_grad, bag_grad.shape[0], ld, ld, _lib.ETR_F32, route.send_pos)
        recv = self.router.send_grads(route, grouped)                   # [n_recv, ld]
        if recv.shape[1] != self.local.grad_ld:
            pad = self.rt.zeros((recv.shape[0], self.local.grad_ld))
            pad[:, : recv.shape[1]] = recv
            recv = pad
        lids = IdsBatch(self.rt, route.recv_rows, route.n_recv, 1, 1, 1, 1, 1)
        return SparseGrad(self.local, lids, recv)
