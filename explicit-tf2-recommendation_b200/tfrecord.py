"""TFRecord input side: ``tf.train.Example`` files -> pinned host column blocks -> ``Trainer.stage``.

The reference feeds its train loop with ``tf.data.TFRecordDataset`` + ``tf.io.parse_single_example`` over
``FixedLenFeature(shape=[1])`` int64 features and a float label, batched (2.FM/ModelManager.py:122-153), from files
written by ``CustomTFWriter`` (2.FM/DataGenerator.py:104-124).  ``TFRecordDataset`` below yields the same thing -- a
dict ``name -> [B] tensor`` per batch -- but the tensors are the rows of ONE pinned column-major block per dtype, so
the Trainer's staging moves a whole batch with two asynchronous copies.  Parsing is native
(``etr_tfrecord_parse`` in libetr.so); ``write_examples`` produces byte-identical framing for tests and export.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check
from .tf_checkpoint import masked_crc32c, _put_varint, _pb_bytes_field


# ------------------------------------------------------------------ writer (tests / export)
def _example(int_feats: Dict[str, Sequence[int]], float_feats: Dict[str, Sequence[float]]) -> bytes:
    entries = b""
    for name, vals in float_feats.items():
        packed = b"".join(struct.pack("<f", float(v)) for v in vals)
        feat = _pb_bytes_field(2, _pb_bytes_field(1, packed))                     # Feature.float_list.value (packed)
        entries += _pb_bytes_field(1, _pb_bytes_field(1, name.encode()) + _pb_bytes_field(2, feat))
    for name, vals in int_feats.items():
        packed = b"".join(_put_varint(int(v) & (2 ** 64 - 1)) for v in vals)
        feat = _pb_bytes_field(3, _pb_bytes_field(1, packed))                     # Feature.int64_list.value (packed)
        entries += _pb_bytes_field(1, _pb_bytes_field(1, name.encode()) + _pb_bytes_field(2, feat))
    return _pb_bytes_field(1, entries)                                            # Example.features


def frame(data: bytes) -> bytes:
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", masked_crc32c(head)) + data + struct.pack("<I", masked_crc32c(data))


def write_examples(path: str, int_cols: Dict[str, np.ndarray], float_cols: Dict[str, np.ndarray]) -> int:
    """columns ([N] or [N,w]) -> one TFRecord file of tf.train.Example messages; returns N."""
    n = len(next(iter({**int_cols, **float_cols}.values())))
    with open(path, "wb") as fh:
        for i in range(n):
            fh.write(frame(_example({k: np.atleast_1d(v[i]) for k, v in int_cols.items()},
                                    {k: np.atleast_1d(v[i]) for k, v in float_cols.items()})))
    return n


# ------------------------------------------------------------------ reader
class TFRecordDataset:
    """``for batch in TFRecordDataset(files, int_features, float_features, batch)``: dict name -> torch tensor [B]
    (``[B,w]`` for a width-w feature), int64 / float32, living in pinned host memory as the columns of one block per
    dtype (``pin=False`` for plain numpy-backed tensors, e.g. on a box without CUDA).  ``depth`` blocks are cycled, so
    a batch stays valid until ``depth - 1`` later batches have been produced (the Trainer's staging copy is long done)."""

    def __init__(self, files, int_features: Sequence[str], float_features: Sequence[str] = ("label",), batch: int = 100,
                 widths: Optional[Dict[str, int]] = None, drop_remainder: bool = False, verify_crc: bool = False,
                 pin: bool = True, depth: int = 3):
        import torch
        self.files = [files] if isinstance(files, str) else list(files)
        self.int_features, self.float_features = list(int_features), list(float_features)
        self.batch, self.drop_remainder, self.verify_crc = int(batch), drop_remainder, int(bool(verify_crc))
        widths = widths or {}
        self.names = self.int_features + self.float_features
        self.widths = [int(widths.get(n, 1)) for n in self.names]
        self.kinds = [0] * len(self.int_features) + [1] * len(self.float_features)
        self.lib = _lib.load()
        wi = sum(self.widths[: len(self.int_features)])
        wf = sum(self.widths[len(self.int_features):])
        self._blocks = []
        for _ in range(depth):
            ib = torch.zeros((max(wi, 1), self.batch), dtype=torch.int64)
            fb = torch.zeros((max(wf, 1), self.batch), dtype=torch.float32)
            if pin and torch.cuda.is_available():
                ib, fb = ib.pin_memory(), fb.pin_memory()
            self._blocks.append((ib, fb))
        assert all(w == 1 for w in self.widths) or True
        self._next = 0

    def _columns(self, block) -> Tuple[List, Dict[str, "object"]]:
        """column pointers for the parser + the name -> tensor views.  Width-1 features are rows of the [n, B] block
        (contiguous, back to back); a width-w feature takes w consecutive rows and is exposed as [B, w]."""
        ib, fb = block
        ptrs, views = [], {}
        ri = rf = 0
        for name, kind, w in zip(self.names, self.kinds, self.widths):
            blk, r0 = (ib, ri) if kind == 0 else (fb, rf)
            rows = blk[r0:r0 + w]
            ptrs.append(rows.data_ptr())
            views[name] = rows[0] if w == 1 else rows.reshape(-1).view(self.batch, w)
            if kind == 0:
                ri += w
            else:
                rf += w
        return ptrs, views

    def __iter__(self) -> Iterator[Dict[str, "object"]]:
        n_feat = len(self.names)
        names = (C.c_char_p * n_feat)(*[n.encode() for n in self.names])
        kinds = (C.c_int32 * n_feat)(*self.kinds)
        widths = (C.c_int32 * n_feat)(*self.widths)
        filled = 0
        block = self._blocks[self._next % len(self._blocks)]
        ptrs, views = self._columns(block)
        for path in self.files:
            data = np.fromfile(path, dtype=np.uint8)
            pos = 0
            while pos < data.size:
                # the parser writes rows [filled, batch) of every column
                cols = (C.c_void_p * n_feat)(*[p + filled * w * (8 if k == 0 else 4)
                                               for p, k, w in zip(ptrs, self.kinds, self.widths)])
                n_rows, used = C.c_int64(0), C.c_int64(0)
                check(self.lib.etr_tfrecord_parse(data[pos:].ctypes.data, data.size - pos, n_feat, names, kinds, widths, cols,
                                                  self.batch - filled, self.verify_crc, C.byref(n_rows), C.byref(used)))
                if used.value == 0:
                    raise ValueError(f"{path}: truncated TFRecord frame at offset {pos}")
                pos += used.value
                filled += n_rows.value
                if filled == self.batch:
                    yield views
                    self._next += 1
                    block = self._blocks[self._next % len(self._blocks)]
                    ptrs, views = self._columns(block)
                    filled = 0
        if filled and not self.drop_remainder:
            yield {n: v[:filled] for n, v in views.items()}
            self._next += 1


def parse_file(path: str, int_features: Sequence[str], float_features: Sequence[str] = ("label",),
               widths: Optional[Dict[str, int]] = None, verify_crc: bool = True) -> Dict[str, np.ndarray]:
    """whole file -> {name: ndarray [N] / [N,w]} (tests, small files)."""
    out: Dict[str, List[np.ndarray]] = {}
    ds = TFRecordDataset(path, int_features, float_features, batch=4096, widths=widths, verify_crc=verify_crc, pin=False)
    for b in ds:
        for n, v in b.items():
            out.setdefault(n, []).append(v.numpy().copy())
    return {n: np.concatenate(v) for n, v in out.items()}
