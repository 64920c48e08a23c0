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
import os
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


# ===========================================================================
# Peer-memory form (NVLink / NVSwitch, CUDA IPC): no all-to-all at all.
class _CudaArray:
    """__cuda_array_interface__ shim: lets torch view a raw device pointer without a copy."""

    def __init__(self, ptr: int, shape, typestr: str, owner=None):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr,
                                         "version": 2, "strides": None}
        self._owner = owner


class PeerBuffer:
    """A cudaMalloc'ed, zero-initialised buffer of this rank, exported over CUDA IPC and mapped by
    every rank of the group: ``peer[g]`` is rank g's buffer as seen from here (``peer[rank]`` = own)."""

    def __init__(self, rt, nbytes: int, world: int, rank: int, group=None):
        self.rt, self.nbytes, self.world, self.rank = rt, int(nbytes), world, rank
        ptr, handle = C.c_void_p(), (C.c_char * 64)()
        check(rt.lib.etr_peer_alloc(rt.ctx, self.nbytes, C.byref(ptr), handle))
        self.ptr = int(ptr.value)
        handles: List[Optional[bytes]] = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.peer: List[int] = []
        for g in range(world):
            if g == rank:
                self.peer.append(self.ptr)
            else:
                q = C.c_void_p()
                check(rt.lib.etr_peer_open(rt.ctx, handles[g], C.byref(q)))
                self.peer.append(int(q.value))

    def tensor(self, shape, dtype: torch.dtype) -> torch.Tensor:
        typestr = {torch.float32: "<f4", torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        return torch.as_tensor(_CudaArray(self.ptr, shape, typestr, self), device=self.rt.device)

    def peer_array(self, byte_offset: int = 0):
        return (C.c_void_p * self.world)(*[q + byte_offset for q in self.peer])


class PeerShardedTable:
    """A [rows_global, width] fp32 table row-sharded over the GPUs of one box (owner = id mod G,
    local row = id div G) whose shards are mapped into every rank through CUDA IPC.

    forward : the fused gather + FM kernel reads row ``id`` straight from its owner's HBM over
              NVLink (etr_table.reserved = shard-set id) -- gather and exchange are ONE kernel;
    backward: the fused backward exports this rank's de-duplicated gradient rows in DEFERRED form
              [P, sum_g] (dv = P - v * sum_g is finished by the owner, so no row is re-read over
              NVLink), ``etr_shard_push`` stores them into the owners' mailboxes;
    apply   : after a device-side barrier each owner adds the G source regions of its mailbox into a
              dense accumulator (rank order, no sort, no atomics) and runs Adam on the touched rows of
              its own shard; a second barrier closes the step.
    Nothing is read back to the host, so the whole data-parallel step is one CUDA graph."""

    def __init__(self, rt, rows_global: int, width: int, world: int, rank: int, group=None, record: Optional[bool] = None):
        """``record`` (default: whenever a row is 17..20 floats, the FM table of k = 16): the shard holds 256-byte records [var | m | v | pad] like the
        unsharded FM table (runtime.EmbeddingTable record=True), so the owner's apply moves a row's whole state as two
        full lines; mailbox / response rows stay densely packed (``ld`` floats)."""
        from .runtime import EmbeddingTable
        assert world & (world - 1) == 0, "peer-sharded tables need a power-of-two number of ranks"
        self.rt, self.world, self.rank, self.group = rt, world, rank, group
        self.rows, self.width, self.dtype = int(rows_global), int(width), torch.float32
        self.ld = (self.width + 3) // 4 * 4                             # mailbox / response / gradient row
        self.record = (self.ld == 20 and os.environ.get("ETR_PEER_RECORD", "1") != "0") if record is None else bool(record)
        self.stride = 64 if self.record else self.ld
        if self.record:
            self.ld = 20
        self.local_rows_max = (self.rows + world - 1) // world          # same allocation on every rank
        self.local_rows = (self.rows - rank + world - 1) // world
        self._shard = PeerBuffer(rt, max(self.local_rows_max, 1) * self.stride * 4, world, rank, group)
        data = self._shard.tensor((max(self.local_rows_max, 1), self.stride), torch.float32)
        if self.record:
            self.local = EmbeddingTable(rt, max(self.local_rows, 1), self.width, torch.float32, record=True,
                                        rec=data[: max(self.local_rows, 1)])
        else:
            self.local = EmbeddingTable(rt, max(self.local_rows, 1), self.width, torch.float32,
                                        data=data[: max(self.local_rows, 1)])
        sid = C.c_int32(0)
        check(rt.lib.etr_shard_set_create(rt.ctx, self._shard.peer_array(), world, rank, self.rows, C.byref(sid)))
        self.shard_set = int(sid.value)
        # barrier state + dense all-reduce slots + gradient mailbox (allocated on first use)
        self._flags = PeerBuffer(rt, 64 * 4, world, rank, group)
        self._epoch = rt.zeros((1,), torch.int32)
        self._flag_ptrs = self._flags.peer_array()
        self._ar = None
        self._mb = None
        self._gacc = None
        self._own = None                                  # owner-side prep state (map, step, mask, others)
        self._prep_stream = None
        self._prep_done = None                            # event of the prep that the next apply_mailbox consumes
        self.owner_prep = os.environ.get("ETR_PEER_OWNER_PREP", "1") != "0"
        self.cap = 0
        # request mailboxes (ids + counts) come in ``n_req_sets`` sets: set 0 serves requests made inside the step, the
        # Trainer's plan-ahead sends the requests of the batch staged in buffer set s into set s one step early
        self.n_req_sets, self._mb_sets, self._cur_set = 1, 0, 0
        if world > 1:
            dist.barrier(group=group)

    # -- EmbeddingTable-like surface for the kernels ------------------------
    @property
    def data(self) -> torch.Tensor:
        return self.local.data

    @property
    def m(self):
        return self.local.m

    @property
    def v(self):
        return self.local.v

    @property
    def grad_ld(self) -> int:
        return self.ld

    def desc(self) -> _lib.etr_table:
        return _lib.etr_table(self._shard.ptr, self.rows, self.width, self.stride, _lib.ETR_F32, self.shard_set)

    def cols(self, c0: int, c1: int) -> torch.Tensor:
        return self.local.data[:, c0:c1]

    def init_uniform(self, lo, hi, gen, c0=0, c1=None):
        self.local.init_uniform(lo, hi, gen, c0, c1)

    def load_global(self, full: torch.Tensor, c0: int = 0):
        mine = full[self.rank:: self.world].to(self.rt.device)
        self.local.data[: mine.shape[0], c0:c0 + mine.shape[1]] = mine.to(torch.float32)

    # -- collectives over peer memory ---------------------------------------
    def barrier(self):
        rt = self.rt
        check(rt.lib.etr_peer_barrier(rt.ctx, self._flag_ptrs, self._flags.ptr, self._epoch.data_ptr(), self.world,
                                      self.rank, rt.stream))

    def _ensure_ar(self, n: int):
        if self._ar is None or self._ar[1] < n:
            torch.cuda.synchronize(self.rt.device)
            buf = PeerBuffer(self.rt, self.world * n * 4, self.world, self.rank, self.group)
            self._ar = (buf, n, buf.peer_array())
            if self.world > 1:
                dist.barrier(group=self.group)
        return self._ar

    def allreduce_push(self, vec: torch.Tensor):
        """stage 1 of the dense-gradient all-reduce (before the mid-step barrier)"""
        assert vec.dtype == torch.float32 and vec.is_contiguous()
        buf, n, ptrs = self._ensure_ar(vec.numel())
        assert n == vec.numel(), "the dense-gradient vector must keep its size"
        rt = self.rt
        check(rt.lib.etr_peer_allreduce_push(rt.ctx, vec.data_ptr(), n, ptrs, self.world, self.rank, rt.stream))

    def allreduce_sum(self, vec: torch.Tensor):
        """stage 2 (after the barrier): vec = sum over ranks, summed in rank order on every rank"""
        buf, n, _ = self._ar
        rt = self.rt
        check(rt.lib.etr_peer_allreduce_sum(rt.ctx, buf.ptr, n, self.world, vec.data_ptr(), rt.stream))

    # -- gradient mailbox -----------------------------------------------------
    def ensure_mailbox(self, n_slots: int):
        """[world][cap] regions per owner; cap = 1.5x the balanced share of a batch's slots (+1024).
        Collective (every rank calls it with the same n_slots)."""
        cap = (3 * n_slots // (2 * self.world) + 1024 + 63) // 64 * 64
        if self._mb is not None and cap <= self.cap and self._mb_sets >= self.n_req_sets:
            return
        cap = max(cap, self.cap)
        assert not torch.cuda.is_current_stream_capturing(), "size the mailbox with an eager step before capture"
        torch.cuda.synchronize(self.rt.device)
        W, ld, rt, S = self.world, self.ld, self.rt, self.n_req_sets
        ids = PeerBuffer(rt, S * W * cap * 8, W, self.rank, self.group)
        grads = PeerBuffer(rt, W * cap * ld * 4, W, self.rank, self.group)
        counts = PeerBuffer(rt, S * 64 * 4, W, self.rank, self.group)
        ids_all, counts_all = ids.tensor((S, W * cap), torch.int64), counts.tensor((S, 64), torch.int32)
        self._mb = {
            "ids": ids, "grads": grads, "counts": counts,
            "ids_t": [ids_all[q] for q in range(S)], "grads_t": grads.tensor((W * cap, ld), torch.float32),
            "counts_t": [counts_all[q, :W] for q in range(S)],
            "ids_ptrs": [ids.peer_array((q * W + self.rank) * cap * 8) for q in range(S)],
            "grads_ptrs": grads.peer_array(self.rank * cap * ld * 4),
            "counts_ptrs": [counts.peer_array((q * 64 + self.rank) * 4) for q in range(S)],
            "local_cnt": [rt.zeros((W,), torch.int32) for _ in range(S)],
            "touched": rt.empty((W * cap,), torch.int32), "n_touched": rt.zeros((1,), torch.int32),
        }
        self._mb_sets = S
        self._own, self._prep_done = None, None           # sized by cap: rebuilt on the next prep
        # response buffer of the de-duplicated forward exchange: [owner][cap][ld] on THIS rank; owner g writes its
        # region through the pointer resp_ptrs[source] = source's buffer + g * cap * ld * 4
        resp = PeerBuffer(rt, W * cap * ld * 4, W, self.rank, self.group)
        self._mb["resp"] = resp
        self._mb["resp_t"] = resp.tensor((W * cap, ld), torch.float32)
        self._mb["resp_ptrs"] = resp.peer_array(self.rank * cap * ld * 4)
        self.cap = cap
        if W > 1:
            dist.barrier(group=self.group)

    def push(self, unique_ids: torch.Tensor, n_unique: torch.Tensor, max_unique: int, unique_grad: torch.Tensor):
        rt, mb = self.rt, self._mb
        self._cur_set = 0
        check(rt.lib.etr_shard_push(rt.ctx, unique_ids.data_ptr(), n_unique.data_ptr(), max_unique,
                                    unique_grad.data_ptr(), unique_grad.shape[1], self.world, self.cap, mb["ids_ptrs"][0],
                                    mb["grads_ptrs"], mb["counts_ptrs"][0], mb["local_cnt"][0].data_ptr(), rt.stream))

    # -- de-duplicated forward exchange ------------------------------------------
    def request(self, plan, B: int, F: int, req_set: int = 0, stream=None):
        """unique ids -> the owners' request mailboxes (set ``req_set``), and the occurrence -> response-row map.  Both
        depend on the plan only, so the Trainer's plan-ahead runs them one step early (``stream`` = its side stream)."""
        rt, W = self.rt, self.world
        self.ensure_mailbox(plan.n_slots)
        mb, cap = self._mb, self.cap
        if getattr(plan, "slot_of_u", None) is None:
            plan.slot_of_u = rt.empty((max(plan.n_slots, 1),), torch.int32)
            plan.vid = rt.empty((B * F,), torch.int64)
        st = stream if stream is not None else rt.stream
        check(rt.lib.etr_shard_request(rt.ctx, plan.unique_ids.data_ptr(), plan.counts.data_ptr(), plan.n_slots, W, cap,
                                       mb["ids_ptrs"][req_set], mb["counts_ptrs"][req_set], mb["local_cnt"][req_set].data_ptr(),
                                       plan.slot_of_u.data_ptr(), st))
        check(rt.lib.etr_shard_vid_map(rt.ctx, plan.sorted_bag.data_ptr(), plan.seg_start.data_ptr(),
                                       plan.counts.data_ptr(), plan.n_slots, plan.slot_of_u.data_ptr(), plan.vid.data_ptr(), st))
        plan.req_set = req_set

    def exchange_forward(self, plan, B: int, F: int):
        """[request -> virtual ids -> barrier ->] serve -> barrier.  Returns (VirtualTable over the response buffer,
        IdsBatch of response-buffer rows [B,F], slot_of_u) -- the fused gather kernel then runs on local memory.  The
        bracketed part is skipped when the plan's requests were sent ahead of time (``request``; the barrier that closed
        the previous step has ordered them)."""
        from .runtime import IdsBatch
        rt, W = self.rt, self.world
        if getattr(plan, "req_set", None) is None:
            self.request(plan, B, F, 0)
            self.barrier()
        q, plan.req_set = plan.req_set, None
        self._cur_set = q
        mb, cap, ld = self._mb, self.cap, self.ld
        if self.owner_prep:
            self._launch_owner_prep(q)
        t = self.local.desc()
        check(rt.lib.etr_shard_serve(rt.ctx, C.byref(t), mb["ids_t"][q].data_ptr(), mb["counts_t"][q].data_ptr(), W, cap,
                                     mb["resp_ptrs"], ld, rt.stream))
        self.barrier()
        vt = VirtualTable(rt, mb["resp_t"], self.width)
        return vt, IdsBatch(rt, plan.vid, B, F, 1, F, 1, 1), plan.slot_of_u

    def _launch_owner_prep(self, q: int):
        """The requests of set ``q`` have arrived (a barrier has ordered them): on a side stream, while the forward
        runs, every entry claims its row (etr_shard_owner_prep) so that the apply after the gradient barrier is one pass."""
        rt, mb, W, cap = self.rt, self._mb, self.world, self.cap
        if self._own is None:
            assert not torch.cuda.is_current_stream_capturing(), "run one eager step before capture (owner-prep buffers)"
            self._own = {"map": rt.zeros((max(self.local_rows, 1),), torch.int64), "step": rt.zeros((1,), torch.int32),
                         "mask": rt.zeros((W * cap,), torch.int32), "others": rt.empty((W * cap, W), torch.int32)}
        own = self._own
        if self._prep_stream is None:
            self._prep_stream = torch.cuda.Stream(device=rt.device)
        cur = torch.cuda.current_stream(rt.device)
        if self._prep_done is not None:                   # a prep that no apply consumed: its claims are stale
            cur.wait_event(self._prep_done)
            own["mask"].zero_()
        ps = self._prep_stream
        ps.wait_stream(cur)                               # after the previous apply (mask / map free) and the requests' barrier
        check(rt.lib.etr_shard_owner_prep(rt.ctx, mb["ids_t"][q].data_ptr(), mb["counts_t"][q].data_ptr(), W, cap,
                                          max(self.local_rows, 1), own["map"].data_ptr(), own["step"].data_ptr(),
                                          own["mask"].data_ptr(), own["others"].data_ptr(), ps.cuda_stream))
        self._prep_done = torch.cuda.Event()
        self._prep_done.record(ps)

    def apply_mailbox(self, d_lr_t: torch.Tensor, b1: float, b2: float, eps: float, mode: int):
        """owner side, no sort: the G source regions are added into the dense accumulator in rank order
        (rows are unique within a region: no atomics, deterministic), first-touch rows go on a list, and
        row-wise Adam runs over that list, finishing the deferred FM gradient dv = P - v * sum_g."""
        if mode != _lib.ADAM_ROWWISE:
            raise NotImplementedError("peer-sharded tables: row-wise Adam only (keras_dense decays every row of "
                                      "every shard each step; use the all-to-all form for that parity mode)")
        rt, mb, W, cap, ld = self.rt, self._mb, self.world, self.cap, self.ld
        if self._prep_done is not None:
            # the rows came back through the request slots and the prep pass has paired them up: one pass
            torch.cuda.current_stream(rt.device).wait_event(self._prep_done)
            self._prep_done = None
            q, own, t = self._cur_set, self._own, self.local.desc()
            check(rt.lib.etr_shard_owner_apply(rt.ctx, C.byref(t), self.local.m.data_ptr(), self.local.v.data_ptr(),
                                               mb["ids_t"][q].data_ptr(), mb["counts_t"][q].data_ptr(),
                                               mb["grads_t"].data_ptr(), W, cap, ld, own["mask"].data_ptr(),
                                               own["others"].data_ptr(), self.width - 1, d_lr_t.data_ptr(), b1, b2, eps,
                                               rt.stream))
            return
        if self._gacc is None:
            assert not torch.cuda.is_current_stream_capturing()
            assert ld >= self.width + 1, "the accumulator row needs a spare last column for its stamp"
            self._gacc = rt.zeros((max(self.local_rows, 1), ld))
        q = self._cur_set                                 # the request set of the step being applied (the owner kept the ids)
        check(rt.lib.etr_shard_mailbox_accumulate(rt.ctx, mb["ids_t"][q].data_ptr(), mb["grads_t"].data_ptr(),
                                                  mb["counts_t"][q].data_ptr(), W, cap, ld, self._gacc.data_ptr(),
                                                  self._epoch.data_ptr(),
                                                  mb["touched"].data_ptr(), mb["n_touched"].data_ptr(), W * cap, rt.stream))
        t = self.local.desc()
        check(rt.lib.etr_shard_touched_adam(rt.ctx, C.byref(t), self.local.m.data_ptr(), self.local.v.data_ptr(),
                                            self._gacc.data_ptr(), ld, mb["touched"].data_ptr(),
                                            mb["n_touched"].data_ptr(), W * cap, self.width - 1, d_lr_t.data_ptr(),
                                            b1, b2, eps, rt.stream))


class PeerFMGrad:
    """FM-family table gradient of a peer-sharded table: ``push`` exports this rank's de-duplicated
    rows (fused backward, apply = 0) into the owners' mailboxes; ``apply`` is the owner-side update."""

    def __init__(self, table: PeerShardedTable, fused):
        self.table, self.fused = table, fused            # fused: runtime.FusedFMGrad built on the sharded desc

    def push(self):
        f, T = self.fused, self.table
        f.reduce()                                        # unique rows of this rank's batch (global ids)
        T.ensure_mailbox(f.plan.n_slots)
        T.push(f.plan.unique_ids, f.plan.counts, f.plan.n_slots, f.unique_grad)

    def apply(self, d_lr_t, b1, b2, eps, mode=_lib.ADAM_ROWWISE):
        self.table.apply_mailbox(d_lr_t, b1, b2, eps, mode)


class PeerSlotGrad(PeerFMGrad):
    """Gradient of a peer-sharded table whose forward went through the de-duplicated exchange: the rows return
    through the request slots (``slot_of_u``), no ids and no atomics."""

    def __init__(self, table: PeerShardedTable, fused, slot_of_u: torch.Tensor):
        super().__init__(table, fused)
        self.slot_of_u = slot_of_u

    def push(self):
        """fused FM backward (deferred form) writing every unique row straight into its owner's mailbox slot"""
        from .runtime import _TORCH2ETR, _p
        f, T = self.fused, self.table
        rt, t, plan, df = T.rt, f.table.desc(), f.plan.join(), f.dflat
        prep = getattr(plan, "fm_prep", None)             # row descriptors / long-run items built with the plan, if any
        check(rt.lib.etr_fm_fused_backward_push(
            rt.ctx, C.byref(t), f.k, f.ids.F, f.ids.B, plan.sorted_bag.data_ptr(), plan.seg_start.data_ptr(),
            plan.unique_ids.data_ptr(), plan.counts.data_ptr(), plan.n_slots, f.dlogit.data_ptr(), f.sumv.data_ptr(),
            _p(df), _TORCH2ETR[df.dtype] if df is not None else 0, df.stride(0) if df is not None else 0, f.flat_col0,
            self.slot_of_u.data_ptr(), T.cap, T._mb["grads_ptrs"], prep.data_ptr() if prep is not None else None,
            prep.numel() if prep is not None else 0, rt.stream))
