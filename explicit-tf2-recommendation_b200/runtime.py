"""Host-side runtime: device context, HBM-resident tables, id batches, and thin
typed wrappers over the C ABI.  torch is used for device memory, streams and
DLPack only (plumbing); all arithmetic happens in libetr.so.
"""
from __future__ import annotations

import ctypes as C
import math
import contextlib
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check

_TORCH2ETR = {torch.float32: _lib.ETR_F32, torch.bfloat16: _lib.ETR_BF16}


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Runtime:
    """One etr_ctx per CUDA device (created lazily, shared by all layers)."""

    _instances: Dict[int, "Runtime"] = {}

    @classmethod
    def get(cls, device=None) -> "Runtime":
        if not torch.cuda.is_available():
            raise RuntimeError("explicit-tf2-recommendation_b200 needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback")
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        if idx is None:
            idx = torch.cuda.current_device()
        if idx not in cls._instances:
            cls._instances[idx] = Runtime(idx)
        return cls._instances[idx]

    def __init__(self, index: int):
        self.lib = _lib.load()
        self.index = index
        self.device = torch.device("cuda", index)
        torch.cuda.set_device(index)
        torch.zeros(1, device=self.device)          # make sure the primary context exists
        h = C.c_void_p()
        check(self.lib.etr_ctx_create(index, C.byref(h)))
        self.ctx = h
        # a second context (own workspace + error word) for work that runs CONCURRENTLY on the side
        # stream: the sorted-id plan of a batch depends only on the ids, so it overlaps forward/backward
        h2 = C.c_void_p()
        check(self.lib.etr_ctx_create(index, C.byref(h2)))
        self.side_ctx = h2
        self._side_stream = None

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def side_stream(self) -> "torch.cuda.Stream":
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        return self._side_stream

    @property
    def launches(self) -> int:
        return int(self.lib.etr_ctx_launch_count(self.ctx)) + int(self.lib.etr_ctx_launch_count(self.side_ctx))

    def poll_error(self) -> None:
        """Raise EtrIdRangeError if a kernel saw an out-of-range id (synchronises)."""
        bad = C.c_int64(0)
        check(self.lib.etr_ctx_poll_error(self.ctx, self.stream, C.byref(bad)))
        check(self.lib.etr_ctx_poll_error(self.side_ctx, self.stream, C.byref(bad)))

    def peek_error_async(self) -> None:
        """Enqueue a copy of both contexts' error words into pinned host memory (no synchronisation)."""
        if getattr(self, "_err_host", None) is None:
            self._err_host = torch.zeros(4, dtype=torch.int64).pin_memory()
            self._err_event = torch.cuda.Event()
            self._err_pending = False
        check(self.lib.etr_ctx_peek_error_async(self.ctx, self.stream, self._err_host.data_ptr()))
        check(self.lib.etr_ctx_peek_error_async(self.side_ctx, self.stream, self._err_host[2:].data_ptr()))
        self._err_event.record(torch.cuda.current_stream(self.device))
        self._err_pending = True

    def check_peeked_error(self, wait: bool = False) -> None:
        """Raise if the last peeked error word was set (EtrIdRangeError / EtrOverflowError / EtrPeerTimeoutError)."""
        if not getattr(self, "_err_pending", False):
            return
        if wait:
            self._err_event.synchronize()
        elif not self._err_event.query():
            return
        self._err_pending = False
        bad = C.c_int64(0)
        for ctx_, off in ((self.ctx, 0), (self.side_ctx, 2)):
            if int(self._err_host[off]) != 0:
                word = (C.c_uint64 * 2)(int(self._err_host[off]) & (2 ** 64 - 1), int(self._err_host[off + 1]) & (2 ** 64 - 1))
                check(self.lib.etr_ctx_decode_error(ctx_, word, self.stream, C.byref(bad)))

    # ------------------------------------------------------------ tensors
    def to_device(self, x, dtype: torch.dtype) -> torch.Tensor:
        """Accept a torch tensor (any device), a numpy array, or any DLPack
        producer (e.g. ``tf.Tensor`` via ``__dlpack__``) -> torch tensor on this
        device (zero-copy when it already lives there)."""
        if isinstance(x, torch.Tensor):
            t = x
        elif isinstance(x, np.ndarray):
            t = torch.from_numpy(x)
        elif hasattr(x, "__dlpack__"):
            t = torch.from_dlpack(x)
        else:
            t = torch.as_tensor(x)
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        if t.dtype != dtype:
            t = t.to(dtype)
        return t

    def empty(self, shape, dtype=torch.float32) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype=torch.float32) -> torch.Tensor:
        return torch.zeros(shape, dtype=dtype, device=self.device)


# ---------------------------------------------------------------------------
class EmbeddingTable:
    """HBM-resident table [rows, stride]; the first ``width`` columns of a row
    are meaningful, the rest are zero padding so that a row is a whole number of
    16-byte chunks.  Replaces the tf.Variable of tf.keras.layers.Embedding
    (2.FM/CustomLayers.py:129-134).  For the FM family the [V,1] ``w`` Embedding
    is fused in as column ``k`` (one DRAM burst serves v and w)."""

    def __init__(self, rt: Runtime, rows: int, width: int, dtype: torch.dtype = torch.float32,
                 data: Optional[torch.Tensor] = None, row_align: int = 16, record: bool = False,
                 rec: Optional[torch.Tensor] = None):
        """``row_align`` (bytes, a multiple of 16): rows start on that boundary.

        ``record=True`` (fp32, width <= 20): the row and its two Adam slots are interleaved in ONE
        256-byte, 256-byte-aligned record ``[var 0..19 | m 20..39 | v 40..59 | pad]``: ``data``, ``m`` and
        ``v`` are column views of the same buffer with row stride 64, so every kernel that addresses them
        through (pointer, stride) works unchanged, the gather still finds ``[v_0..v_15, w]`` inside one
        128-byte line, and the fused apply moves a row's whole state as two full lines
        (csrc/fm_fused_apply.cu: fm_fused_record_kernel)."""
        assert dtype in _TORCH2ETR
        self.rt, self.rows, self.width, self.dtype = rt, int(rows), int(width), dtype
        esz = torch.empty((), dtype=dtype).element_size()
        assert row_align % 16 == 0
        epc = row_align // esz                                        # elements per alignment unit
        self.row_width = ((self.width + epc - 1) // epc) * epc        # meaningful columns, whole 16-byte chunks
        self.record = bool(record)
        self._m = self._v = None
        if self.record:
            assert dtype == torch.float32 and data is None and self.row_width <= 20
            self.row_width = 20
            self.stride = 64
            if rec is not None:       # caller-owned record storage (a peer-mapped shard), already zeroed
                assert rec.shape == (self.rows, 64) and rec.dtype == torch.float32 and rec.is_contiguous()
            self.rec = rt.zeros((self.rows, 64), torch.float32) if rec is None else rec
            assert self.rec.data_ptr() % 256 == 0
            self.data = self.rec[:, 0:20]
            self._m, self._v = self.rec[:, 20:40], self.rec[:, 40:60]
            return
        self.stride = self.row_width
        if data is not None:          # caller-owned storage (e.g. a peer-mapped shard), already zeroed
            assert data.shape == (self.rows, self.stride) and data.dtype == dtype and data.is_contiguous()
        self.data = rt.zeros((self.rows, self.stride), dtype) if data is None else data

    def desc(self) -> _lib.etr_table:
        return _lib.etr_table(self.data.data_ptr(), self.rows, self.width, self.stride, _TORCH2ETR[self.dtype],
                              _lib.ETR_TABLE_RECORD if self.record else 0)

    def cols(self, c0: int, c1: int) -> torch.Tensor:
        """Strided view of columns [c0,c1) -- e.g. ``embed/embeddings`` = cols(0,k),
        ``w/embeddings`` = cols(k,k+1).  Exportable with ``__dlpack__``."""
        return self.data[:, c0:c1]

    def init_uniform(self, lo: float, hi: float, gen: torch.Generator, c0: int = 0, c1: Optional[int] = None):
        c1 = self.width if c1 is None else c1
        vals = torch.empty((self.rows, c1 - c0), dtype=torch.float32, device=self.rt.device)
        vals.uniform_(lo, hi, generator=gen)
        self.data[:, c0:c1] = vals.to(self.dtype)

    @property
    def m(self) -> torch.Tensor:
        if self._m is None:
            self._m = self.rt.zeros((self.rows, self.stride), torch.float32)
        return self._m

    @property
    def v(self) -> torch.Tensor:
        if self._v is None:
            self._v = self.rt.zeros((self.rows, self.stride), torch.float32)
        return self._v

    @property
    def grad_ld(self) -> int:
        return self.row_width         # gradient rows use the table's column layout (fp32), densely packed


# ---------------------------------------------------------------------------
class IdsBatch:
    """The categorical input of one batch on the device (a1):
    dict name -> [B] / [B,1] / [B,L]   or   X [B,F] / [B,F,L]   or   CSR."""

    def __init__(self, rt: Runtime, ids: torch.Tensor, B: int, F: int, L: int, sb: int, sf: int, sl: int,
                 pad_id: Optional[int] = None, pooling: str = "sum", csr_offsets: Optional[torch.Tensor] = None):
        self.rt, self.ids, self.B, self.F, self.L = rt, ids, B, F, L
        self.sb, self.sf, self.sl = sb, sf, sl
        self.pad_id, self.pooling, self.csr = pad_id, pooling, csr_offsets
        self.nnz = int(ids.numel()) if csr_offsets is not None else None

    @property
    def is_bag(self) -> bool:
        return self.csr is not None or self.L != 1

    @property
    def n_slots(self) -> int:
        return self.nnz if self.csr is not None else self.B * self.F * self.L

    def desc(self) -> _lib.etr_ids:
        return _lib.etr_ids(self.ids.data_ptr(), _p(self.csr), self.B, self.F, self.L, self.sb, self.sf, self.sl,
                            0 if self.pad_id is None else int(self.pad_id), 0 if self.pad_id is None else 1,
                            _lib.POOL_MEAN if self.pooling == "mean" else _lib.POOL_SUM)

    # -- constructors -------------------------------------------------------
    @staticmethod
    def from_matrix(rt: Runtime, X, pad_id=None, pooling="sum") -> "IdsBatch":
        t = rt.to_device(X, torch.int64)
        if t.dim() == 2:
            t = t.contiguous()
            B, F = t.shape
            return IdsBatch(rt, t, B, F, 1, F, 1, 1, pad_id, pooling)
        assert t.dim() == 3, "ids must be [B,F] or [B,F,L]"
        t = t.contiguous()
        B, F, L = t.shape
        return IdsBatch(rt, t, B, F, L, F * L, L, 1, pad_id, pooling)

    @staticmethod
    def from_csr(rt: Runtime, values, offsets, B: int, F: int, pad_id=None, pooling="sum") -> "IdsBatch":
        v = rt.to_device(values, torch.int64).contiguous()
        o = rt.to_device(offsets, torch.int32).contiguous()
        assert o.numel() == B * F + 1
        return IdsBatch(rt, v, B, F, 1, 0, 0, 0, pad_id, pooling, csr_offsets=o)

    @staticmethod
    def from_dict(rt: Runtime, inputs, names: Sequence[str], pad_id=None, pooling="sum") -> "IdsBatch":
        """The reference input idiom (2.FM/CustomLayers.py:138-144): every
        feature is [B] or [B,1] int64.  Produces a field-major [F,B] buffer: one
        H2D copy per host column, or ONE assemble kernel for device columns.  A
        [B,L] feature makes the whole batch a padded-bag batch [B,F,L]."""
        first = inputs[names[0]]
        cols = [inputs[n] for n in names]
        shapes = [tuple(c.shape) for c in cols]
        L = max((s[1] if len(s) == 2 else 1) for s in shapes)
        B = shapes[0][0]
        F = len(names)
        if L > 1:
            ts = []
            for c, s in zip(cols, shapes):
                t = rt.to_device(c, torch.int64).reshape(B, -1)
                if t.shape[1] != L:
                    assert pad_id is not None, "ragged bag widths need a pad_id"
                    t = torch.nn.functional.pad(t, (0, L - t.shape[1]), value=pad_id)
                ts.append(t)
            X = torch.stack(ts, dim=1).contiguous()
            return IdsBatch(rt, X, B, F, L, F * L, L, 1, pad_id, pooling)
        out = rt.empty((F, B), torch.int64)
        on_device = all(isinstance(c, torch.Tensor) and c.device == rt.device and c.dtype == torch.int64
                        for c in cols)
        if on_device:
            flat = [c.reshape(-1).contiguous() for c in cols]
            for f0 in range(0, F, 64):
                chunk = flat[f0:f0 + 64]
                arr = (C.c_void_p * len(chunk))(*[c.data_ptr() for c in chunk])
                check(rt.lib.etr_assemble_ids(rt.ctx, arr, len(chunk), B, out[f0:].data_ptr(), rt.stream))
        else:
            for f, c in enumerate(cols):
                if isinstance(c, np.ndarray):
                    c = torch.from_numpy(c)
                elif not isinstance(c, torch.Tensor):
                    c = torch.from_dlpack(c) if hasattr(c, "__dlpack__") else torch.as_tensor(c)
                out[f].copy_(c.reshape(-1), non_blocking=True)
        del first
        return IdsBatch(rt, out, B, F, 1, 1, B, 1, pad_id, pooling)

    @staticmethod
    def make(rt: Runtime, inputs, names: Sequence[str], pad_id=None, pooling="sum") -> "IdsBatch":
        if isinstance(inputs, IdsBatch):
            return inputs
        if isinstance(inputs, dict):
            return IdsBatch.from_dict(rt, inputs, names, pad_id, pooling)
        return IdsBatch.from_matrix(rt, inputs, pad_id, pooling)


# ---------------------------------------------------------------------------
class SparsePlan:
    """Sorted-ID plan of one batch (shared by every table indexed by the same ids)."""

    def __init__(self, rt: Runtime, ids: IdsBatch, table_rows: int, overlap: bool = False):
        """``overlap=True``: the sort runs on the runtime's side stream (forked from the current
        stream, using the side ctx's workspace) and overlaps whatever is enqueued next on the current
        stream; consumers call ``join()`` first."""
        n = ids.n_slots
        self.n_slots = n
        self.rt = rt
        self._pending = None
        self.fm_prep = None
        cur = torch.cuda.current_stream(rt.device)
        if overlap:
            side = rt.side_stream
            side.wait_stream(cur)                        # the ids are produced on the current stream
            ctx_, stream_ctx = rt.side_ctx, torch.cuda.stream(side)
        else:
            ctx_, stream_ctx = rt.ctx, torch.cuda.stream(cur)
        with stream_ctx:
            self.sorted_bag = rt.empty((max(n, 1),), torch.int32)
            self.unique_ids = rt.empty((max(n, 1),), torch.int64)
            self.seg_start = rt.empty((n + 1,), torch.int32)
            self.counts = rt.zeros((2,), torch.int32)        # [n_unique, n_valid]
            self.sorted_key = rt.empty((max(n, 1),), torch.int32)   # uint32 ids of the sorted occurrences
            d = ids.desc()
            check(rt.lib.etr_sparse_plan_keys(ctx_, C.byref(d), ids.nnz or 0, table_rows, self.sorted_bag.data_ptr(),
                                              self.unique_ids.data_ptr(), self.seg_start.data_ptr(),
                                              self.counts[0:].data_ptr(), self.counts[1:].data_ptr(),
                                              self.sorted_key.data_ptr(),
                                              torch.cuda.current_stream(rt.device).cuda_stream))
        if overlap:
            self._pending = rt.side_stream
            for t in (self.sorted_bag, self.unique_ids, self.seg_start, self.counts, self.sorted_key):
                t.record_stream(cur)                     # allocated on the side stream, consumed on the current one

    def rebuild(self, ids: IdsBatch, table_rows: int) -> "SparsePlan":
        """Plan another batch of the same shape INTO THE SAME BUFFERS, on the side stream (forked from the current one):
        the plan of the NEXT batch is sorted while the current step runs (Trainer ``next_batch=``); consumers ``join()``."""
        assert ids.n_slots == self.n_slots
        rt = self.rt
        cur = torch.cuda.current_stream(rt.device)
        side = rt.side_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.counts.zero_()
            d = ids.desc()
            check(rt.lib.etr_sparse_plan_keys(rt.side_ctx, C.byref(d), ids.nnz or 0, table_rows, self.sorted_bag.data_ptr(),
                                              self.unique_ids.data_ptr(), self.seg_start.data_ptr(),
                                              self.counts[0:].data_ptr(), self.counts[1:].data_ptr(),
                                              self.sorted_key.data_ptr(), side.cuda_stream))
            if self.fm_prep is not None:
                self._run_prepare(side)
        self._pending = side
        return self

    def prepare_fm(self) -> "SparsePlan":
        """Row descriptors + long-run items for the tiled fused FM apply / push (etr_fm_fused_prepare): depend on the ids
        only, so they are built once per plan -- right behind the sort, on the stream the plan was built on."""
        if self.fm_prep is not None:
            return self
        rt = self.rt
        side = self._pending
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            nbytes = int(rt.lib.etr_fm_fused_prepare_bytes(self.n_slots))
            self.fm_prep = rt.empty((nbytes,), torch.uint8)
            assert self.fm_prep.data_ptr() % 256 == 0
            self._run_prepare(side)
        if side is not None:
            self.fm_prep.record_stream(torch.cuda.current_stream(rt.device))
        return self

    def _run_prepare(self, side) -> None:
        rt = self.rt
        check(rt.lib.etr_fm_fused_prepare(rt.side_ctx if side is not None else rt.ctx, self.seg_start.data_ptr(),
                                          self.unique_ids.data_ptr(), self.counts.data_ptr(), self.n_slots,
                                          self.fm_prep.data_ptr(), self.fm_prep.numel(),
                                          torch.cuda.current_stream(rt.device).cuda_stream))

    def join(self) -> "SparsePlan":
        """Make the current stream wait for an overlapped plan (no-op otherwise)."""
        if self._pending is not None:
            torch.cuda.current_stream(self.rt.device).wait_stream(self._pending)
            self._pending = None
        return self

    @property
    def n_unique(self) -> int:        # synchronises; tests / export only
        return int(self.counts[0].item())


class SparseGrad:
    """IndexedSlices analogue: per-bag gradient rows (table row layout) + ids."""

    def __init__(self, table: EmbeddingTable, ids: IdsBatch, bag_grad: torch.Tensor):
        self.table, self.ids, self.bag_grad = table, ids, bag_grad
        self.plan: Optional[SparsePlan] = None
        self.unique_grad: Optional[torch.Tensor] = None

    def reduce(self, plan: Optional[SparsePlan] = None) -> "SparseGrad":
        """sort + segment-reduce -> (unique ids, summed rows) = deduplicated IndexedSlices."""
        rt = self.table.rt
        self.plan = (plan or self.plan or SparsePlan(rt, self.ids, self.table.rows)).join()
        ld = self.bag_grad.shape[1]
        self.unique_grad = rt.empty((max(self.plan.n_slots, 1), ld), torch.float32)
        check(rt.lib.etr_sparse_segment_reduce(rt.ctx, self.plan.sorted_bag.data_ptr(), self.plan.seg_start.data_ptr(),
                                               self.plan.counts.data_ptr(), self.plan.n_slots,
                                               self.bag_grad.data_ptr(), ld, self.unique_grad.data_ptr(), rt.stream))
        return self

    def indexed_slices(self):
        """(unique ids [U], grads [U,width]) -- synchronises; for tests/export."""
        if self.unique_grad is None:
            self.reduce()
        u = self.plan.n_unique
        return self.plan.unique_ids[:u], self.unique_grad[:u, : self.table.width]


class FlatSparseGrad(SparseGrad):
    """SparseGrad whose per-occurrence rows are slices of a dense fp32 matrix ``flat`` [B, ld] (occurrence (b, f) =
    flat[b, col0 + f*k : col0 + (f+1)*k], single-hot ids): the segment reduction reads them in place
    (etr_sparse_segment_reduce_flat) instead of a [B*F, k] re-layout -- the table gradient of DCN / PNN-style models,
    whose only use of the rows is Flatten(embeddings) (3.DCN/CustomLayers.py:1096-1100)."""

    def __init__(self, table: EmbeddingTable, ids: IdsBatch, flat: torch.Tensor, col0: int, k: int):
        assert flat.dtype == torch.float32 and flat.stride(1) == 1 and not ids.is_bag and table.grad_ld == k
        super().__init__(table, ids, None)
        self.flat, self.col0, self.k = flat, int(col0), int(k)

    @staticmethod
    def eligible(table, ids: IdsBatch, flat: torch.Tensor, col0: int, k: int) -> bool:
        return (isinstance(table, EmbeddingTable) and flat.dtype == torch.float32 and flat.stride(1) == 1 and not ids.is_bag
                and table.grad_ld == k and k % 4 == 0 and col0 % 4 == 0 and flat.stride(0) % 4 == 0
                and flat.data_ptr() % 16 == 0 and os.environ.get("ETR_FLAT_SEGRED", "1") != "0")

    def reduce(self, plan: Optional[SparsePlan] = None) -> "SparseGrad":
        rt = self.table.rt
        self.plan = (plan or self.plan or SparsePlan(rt, self.ids, self.table.rows)).join()
        self.unique_grad = rt.empty((max(self.plan.n_slots, 1), self.k), torch.float32)
        check(rt.lib.etr_sparse_segment_reduce_flat(rt.ctx, self.plan.sorted_bag.data_ptr(), self.plan.seg_start.data_ptr(),
                                                    self.plan.counts.data_ptr(), self.plan.n_slots, self.flat.data_ptr(),
                                                    self.flat.stride(0), self.col0, self.ids.F, self.k,
                                                    self.unique_grad.data_ptr(), rt.stream))
        return self


class FusedFMGrad:
    """Table gradient of the FM family in its un-materialised form: the plan's runs plus
    (dlogit, sum_v, dflat).  ``apply`` runs backward + segment reduction + Adam in one pass
    (etr_fm_fused_backward_apply); ``indexed_slices`` exports the deduplicated rows."""

    def __init__(self, table: EmbeddingTable, ids: IdsBatch, k: int, dlogit: torch.Tensor, sumv: torch.Tensor,
                 dflat: Optional[torch.Tensor] = None, flat_col0: int = 0, plan: Optional["SparsePlan"] = None):
        self.table, self.ids, self.k = table, ids, k
        self.dlogit, self.sumv, self.dflat, self.flat_col0 = dlogit, sumv, dflat, flat_col0
        self.plan: Optional[SparsePlan] = plan
        self.unique_grad: Optional[torch.Tensor] = None

    @staticmethod
    def eligible(table: EmbeddingTable, ids: IdsBatch, k: int) -> bool:
        lpr = k // 4
        return (table.dtype == torch.float32 and not ids.is_bag and ids.pad_id is None and k % 4 == 0 and
                1 <= lpr <= 32 and (lpr & (lpr - 1)) == 0 and table.width == k + 1)

    def _run(self, apply: bool, lr_t: float, d_lr_t, b1: float, b2: float, eps: float, out):
        rt, t = self.table.rt, self.table.desc()
        if self.plan is None:
            self.plan = SparsePlan(rt, self.ids, self.table.rows)
        self.plan.join()
        df = self.dflat
        check(rt.lib.etr_fm_fused_backward_apply(
            rt.ctx, C.byref(t), _p(self.table.m) if apply else None, _p(self.table.v) if apply else None, self.k,
            self.ids.F, self.ids.B, self.plan.sorted_bag.data_ptr(), self.plan.seg_start.data_ptr(),
            self.plan.unique_ids.data_ptr(), self.plan.counts.data_ptr(), self.plan.n_slots, self.dlogit.data_ptr(),
            self.sumv.data_ptr(), _p(df), _TORCH2ETR[df.dtype] if df is not None else 0,
            df.stride(0) if df is not None else 0, self.flat_col0, lr_t, _p(d_lr_t), b1, b2, eps, int(apply), _p(out),
            rt.stream))

    # "tile" (default) = tiled kernels (csrc/fm_fused_tile.cu; RECORD tables, picked inside etr_fm_fused_backward_apply),
    # "flat" = occurrence-parallel kernel (csrc/fm_fused_flat.cu; RECORD tables), "rows" = row-parallel kernels
    apply_kernel = os.environ.get("ETR_FUSED_APPLY", "tile")

    def apply(self, d_lr_t: torch.Tensor, b1: float, b2: float, eps: float) -> None:
        if self.table.record and self.k == 16 and FusedFMGrad.apply_kernel == "flat":
            rt, t = self.table.rt, self.table.desc()
            if self.plan is None:
                self.plan = SparsePlan(rt, self.ids, self.table.rows)
            self.plan.join()
            df = self.dflat
            check(rt.lib.etr_fm_fused_flat_apply(
                rt.ctx, C.byref(t), self.k, self.ids.F, self.ids.B, self.plan.sorted_key.data_ptr(),
                self.plan.sorted_bag.data_ptr(), self.plan.n_slots, self.dlogit.data_ptr(), self.sumv.data_ptr(), _p(df),
                _TORCH2ETR[df.dtype] if df is not None else 0, df.stride(0) if df is not None else 0, self.flat_col0,
                0.0, _p(d_lr_t), b1, b2, eps, rt.stream))
            return
        if self.table.record and self.k == 16 and FusedFMGrad.apply_kernel == "tile" and \
                (self.dflat is None or self.dflat.dtype == torch.bfloat16):
            rt, t = self.table.rt, self.table.desc()
            if self.plan is None:
                self.plan = SparsePlan(rt, self.ids, self.table.rows)
            self.plan.prepare_fm().join()
            df, pl = self.dflat, self.plan
            check(rt.lib.etr_fm_fused_backward_apply_prepared(
                rt.ctx, C.byref(t), _p(self.table.m), _p(self.table.v), self.k, self.ids.F, self.ids.B,
                pl.sorted_bag.data_ptr(), pl.seg_start.data_ptr(), pl.unique_ids.data_ptr(), pl.counts.data_ptr(), pl.n_slots,
                self.dlogit.data_ptr(), self.sumv.data_ptr(), _p(df), _TORCH2ETR[df.dtype] if df is not None else 0,
                df.stride(0) if df is not None else 0, self.flat_col0, 0.0, _p(d_lr_t), b1, b2, eps,
                pl.fm_prep.data_ptr(), pl.fm_prep.numel(), rt.stream))
            return
        self._run(True, 0.0, d_lr_t, b1, b2, eps, None)

    def reduce(self, plan: Optional[SparsePlan] = None) -> "FusedFMGrad":
        """materialise the deduplicated gradient rows (tests, export, the keras_dense apply)"""
        rt = self.table.rt
        self.plan = (plan or self.plan or SparsePlan(rt, self.ids, self.table.rows)).join()
        self.unique_grad = rt.empty((max(self.plan.n_slots, 1), self.table.grad_ld))
        self._run(False, 0.0, None, 0.0, 0.0, 0.0, self.unique_grad)
        return self

    def indexed_slices(self):
        if self.unique_grad is None:
            self.reduce()
        u = self.plan.n_unique
        return self.plan.unique_ids[:u], self.unique_grad[:u, : self.table.width]


# ---------------------------------------------------------------------------
# thin op wrappers
def gather_fm_forward(table: EmbeddingTable, k: int, has_w: bool, ids: IdsBatch, bias=None, logit=None, prob=None,
                      sumv=None, flat=None, flat_col0: int = 0, cont: Optional[torch.Tensor] = None):
    """``cont`` [B,C] fp32 (any strides): the kernel also writes the front columns
    [0, flat_col0) of ``flat`` = zero padding then the dense features."""
    rt = table.rt
    t, d = table.desc(), ids.desc()
    flat_dtype = _TORCH2ETR[flat.dtype] if flat is not None else 0
    flat_ld = flat.stride(0) if flat is not None else 0
    if cont is not None:
        assert cont.dtype == torch.float32 and cont.shape[0] == ids.B
        cn, csb, csc = cont.shape[1], cont.stride(0), cont.stride(1)
    else:
        cn, csb, csc = 0, 0, 0
    check(rt.lib.etr_gather_fm_forward(rt.ctx, C.byref(t), k, int(has_w), C.byref(d), _p(bias), _p(logit), _p(prob),
                                       _p(sumv), _p(flat), flat_dtype, flat_ld, flat_col0, _p(cont), cn, csb, csc,
                                       rt.stream))


def gather_fm_backward(table: EmbeddingTable, k: int, has_w: bool, ids: IdsBatch, dlogit=None, dflat=None,
                       flat_col0: int = 0) -> torch.Tensor:
    rt = table.rt
    t, d = table.desc(), ids.desc()
    bag_grad = rt.empty((ids.B * ids.F, table.grad_ld), torch.float32)
    flat_dtype = _TORCH2ETR[dflat.dtype] if dflat is not None else 0
    flat_ld = dflat.stride(0) if dflat is not None else 0
    check(rt.lib.etr_gather_fm_backward(rt.ctx, C.byref(t), k, int(has_w), C.byref(d), _p(dlogit), _p(dflat),
                                        flat_dtype, flat_ld, flat_col0, bag_grad.data_ptr(), table.grad_ld,
                                        rt.stream))
    return bag_grad


def embedding_gather(table: EmbeddingTable, ids) -> torch.Tensor:
    """Generic Embedding.call: ids [...] -> [..., width] fp32 (bit-exact rows)."""
    rt = table.rt
    t = rt.to_device(ids, torch.int64).contiguous()
    out = rt.empty(tuple(t.shape) + (table.width,), torch.float32)
    d = table.desc()
    check(rt.lib.etr_embedding_gather(rt.ctx, C.byref(d), t.data_ptr(), t.numel(), out.data_ptr(), table.width,
                                      rt.stream))
    return out


def gemm_f32(rt: Runtime, A, B, C_out, M, N, K, lda, ldb, ldc, trans_a=False, trans_b=False, alpha=1.0, beta=0.0,
             bias=None, act=None):
    check(rt.lib.etr_gemm_f32(rt.ctx, int(trans_a), int(trans_b), M, N, K, alpha, A.data_ptr(), lda, B.data_ptr(),
                              ldb, beta, C_out.data_ptr(), ldc, _p(bias), _lib.ACT[act], rt.stream))


def bce_forward_backward(rt: Runtime, prob: torch.Tensor, label: torch.Tensor, want_grad=True):
    B = prob.numel()
    loss = rt.empty((1,), torch.float32)
    dlogit = rt.empty((B,), torch.float32) if want_grad else None
    check(rt.lib.etr_bce_forward_backward(rt.ctx, prob.data_ptr(), label.data_ptr(), B, loss.data_ptr(), _p(dlogit),
                                          rt.stream))
    return loss, dlogit


def lr_t(lr: float, b1: float, b2: float, t: int) -> float:
    return lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)


# ---------------------------------------------------------------------------
# tensor-core (tcgen05) wrappers
def gemm_bf16_tn(rt: Runtime, A: torch.Tensor, B: torch.Tensor, C_out: torch.Tensor, M: int, N: int, K: int,
                 bias: Optional[torch.Tensor] = None, act=None):
    """C[M,N] = act(A[M,K] B[N,K]^T + bias): A, B bf16 with unit column stride and
    row strides that are multiples of 8 elements; C fp32 or bf16."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.stride(1) == 1 and B.stride(1) == 1
    check(rt.lib.etr_gemm_bf16_tn(rt.ctx, M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                  C_out.data_ptr(), C_out.stride(0), _TORCH2ETR[C_out.dtype], _p(bias), _lib.ACT[act],
                                  rt.stream))


def gemm_bf16_wgrad(rt: Runtime, A: torch.Tensor, B: torch.Tensor, C_out: torch.Tensor, M: int, N: int, K: int):
    """C[M,N] (fp32) = A^T B: A [K,M], B [K,N] bf16 row-major (K = batch) -- dW = dY^T X without transposing either."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.stride(1) == 1 and B.stride(1) == 1
    assert C_out.dtype == torch.float32 and C_out.stride(1) == 1
    check(rt.lib.etr_gemm_bf16_nn_wgrad(rt.ctx, M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                        C_out.data_ptr(), C_out.stride(0), rt.stream))


def gemm_bf16_tn_accumulate(rt: Runtime, A: torch.Tensor, B: torch.Tensor, C_io: torch.Tensor, M: int, N: int, K: int):
    """C[M,N] (fp32) += A[M,K] B[N,K]^T (bf16 operands, fp32 accumulate)."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.stride(1) == 1 and B.stride(1) == 1
    assert C_io.dtype == torch.float32 and C_io.stride(1) == 1
    check(rt.lib.etr_gemm_bf16_tn_accumulate(rt.ctx, M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                             C_io.data_ptr(), C_io.stride(0), rt.stream))


def cast_bf16(rt: Runtime, src: torch.Tensor, transpose: bool = False, ld_dst: Optional[int] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 [rows, ld] (or [cols, ld] transposed), ld rounded up to 8, zero padded."""
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    rows, cols = src.shape
    inner = rows if transpose else cols
    ld = ld_dst or ((inner + 7) // 8 * 8)
    dst = rt.empty((cols if transpose else rows, ld), torch.bfloat16)
    check(rt.lib.etr_cast_bf16(rt.ctx, src.data_ptr(), rows, cols, src.stride(0), dst.data_ptr(), ld, int(transpose),
                               rt.stream))
    return dst


def transpose_bf16(rt: Runtime, src: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """bf16 [rows, cols] (row stride src.stride(0)) -> bf16 [cols, ld], ld = rows rounded up to 8, zero padded."""
    assert src.dtype == torch.bfloat16 and src.stride(1) == 1
    ld = (rows + 7) // 8 * 8
    dst = rt.empty((cols, ld), torch.bfloat16)
    check(rt.lib.etr_transpose_bf16(rt.ctx, src.data_ptr(), rows, cols, src.stride(0), dst.data_ptr(), ld, rt.stream))
    return dst
