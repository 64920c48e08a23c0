"""Reader for the TensorFlow checkpoints the reference ships.

``tf.train.Checkpoint(model=..., optimizer=...).save`` (2.FM/ModelManager.py:112-119) writes a *tensor
bundle*: ``ckpt-N.index`` is an SSTable (the LevelDB table format: prefix-compressed key/value blocks,
an index block, a 48-byte footer ending in the magic 0xdb4775248b80fb57) whose values are
``BundleEntryProto`` messages (dtype, shape, shard, offset, size), and ``ckpt-N.data-00000-of-00001`` holds
the raw little-endian tensor bytes.  This module parses both with numpy + a 40-line protobuf wire
decoder -- no TensorFlow -- so the weights and Adam slots of ``2.FM/ranking_model/checkpoint/ckpt-*``
can be loaded into the B200 layers (``load_deepfm``) and used as realistic fixtures.

Host-side I/O only: nothing here is on the hot path and nothing here computes.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 14: None, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}   # 14 = bfloat16


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _pb_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yield (field number, wire type, value) of one protobuf message (wire types 0, 1, 2, 5)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _read_block(data: bytes, offset: int, size: int) -> bytes:
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if ctype != 0:
        raise ValueError("compressed SSTable block (snappy) -- tensor-bundle indexes are written uncompressed")
    return raw


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(index_path: str) -> Dict[str, dict]:
    """``ckpt-N.index`` -> {tensor key: {dtype, shape, shard_id, offset, size, crc32c}} ('' = bundle header)."""
    data = open(index_path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _MAGIC:
        raise ValueError(f"{index_path}: not an SSTable (bad footer magic)")
    footer = data[-48:]
    _, p = _varint(footer, 0)           # metaindex handle (offset, size): unused
    _, p = _varint(footer, p)
    ioff, p = _varint(footer, p)
    isize, p = _varint(footer, p)
    entries: Dict[str, dict] = {}
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, q = _varint(handle, 0)
        bsize, q = _varint(handle, q)
        for key, val in _block_entries(_read_block(data, boff, bsize)):
            if key == b"":
                entries[""] = {"header": dict((f, v) for f, _, v in _pb_fields(val))}
                continue
            e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
            for fno, _, v in _pb_fields(val):
                if fno == 1:
                    e["dtype"] = v
                elif fno == 2:
                    for f2, _, dim in _pb_fields(v):
                        if f2 == 2:                          # TensorShapeProto.dim
                            size = 0
                            for f3, _, s in _pb_fields(dim):
                                if f3 == 1:
                                    size = s
                            e["shape"].append(size)
                elif fno == 3:
                    e["shard_id"] = v
                elif fno == 4:
                    e["offset"] = v
                elif fno == 5:
                    e["size"] = v
                elif fno == 6:
                    e["crc32c"] = v
            entries[key.decode("utf-8")] = e
    return entries


def _crc32c_table():
    poly, tab = 0x82F63B78, []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TAB: Optional[List[int]] = None


def masked_crc32c(buf: bytes) -> int:
    """The checksum TensorFlow stores beside every tensor (crc32c, rotated and offset)."""
    global _CRC_TAB
    if _CRC_TAB is None:
        _CRC_TAB = _crc32c_table()
    c = 0xFFFFFFFF
    for b in buf:
        c = _CRC_TAB[(c ^ b) & 0xFF] ^ (c >> 8)
    c ^= 0xFFFFFFFF
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def load_checkpoint(prefix: str, verify_crc: bool = False) -> Dict[str, np.ndarray]:
    """``prefix`` = path without the ``.index`` suffix -> {key: ndarray}; string / variant tensors (the
    object-graph proto) are skipped."""
    entries = read_index(prefix + ".index")
    nshards = 1
    hdr = entries.pop("", None)
    if hdr and 1 in hdr["header"]:
        nshards = int(hdr["header"][1])
    shards = {}
    out: Dict[str, np.ndarray] = {}
    for key, e in entries.items():
        np_dt = _DTYPES.get(e["dtype"], None)
        if np_dt is None:
            continue
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open(f"{prefix}.data-{sid:05d}-of-{nshards:05d}", "rb").read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if verify_crc and e["crc32c"] is not None and masked_crc32c(raw) != e["crc32c"]:
            raise ValueError(f"{prefix}: checksum mismatch for {key}")
        out[key] = np.frombuffer(raw, dtype=np_dt).reshape(e["shape"]).copy()
    return out


_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def variables(prefix: str) -> Dict[str, np.ndarray]:
    """Checkpoint keys with the object-graph suffix stripped: ``model/layer/embed/embeddings`` ..."""
    return {k[:-len(_SUFFIX)] if k.endswith(_SUFFIX) else k: v for k, v in load_checkpoint(prefix).items()}


def deepfm_weights(prefix: str) -> Dict[str, np.ndarray]:
    """The DeepFM ranking checkpoint (2.FM/ranking_model/checkpoint/ckpt-N) as reference-named arrays:
    bias (1,), embed [V,k], w [V,1], MLP kernels / biases, the Adam slots of each (``<name>/m``,
    ``<name>/v``) and the optimizer scalars (iter, beta_1, beta_2, learning_rate)."""
    vs = variables(prefix)
    out: Dict[str, np.ndarray] = {}
    for key, arr in vs.items():
        parts = key.split("/")
        slot = None
        if ".OPTIMIZER_SLOT" in parts:
            i = parts.index(".OPTIMIZER_SLOT")
            slot = parts[i + 2]                     # .OPTIMIZER_SLOT/optimizer/{m,v}
            parts = parts[:i]
        name = "/".join(p for p in parts if p not in ("model", "layer") and not p.startswith("layer_with_weights-"))
        if parts and parts[0] == "optimizer":
            name = "optimizer/" + parts[-1]
        out[name + (f"/{slot}" if slot else "")] = arr
    return out


def find_by_shape(ws: Dict[str, np.ndarray], shape, slot: Optional[str] = None) -> List[str]:
    return [k for k, v in ws.items() if tuple(v.shape) == tuple(shape) and
            ((slot is None and not k.endswith(("/m", "/v"))) or (slot is not None and k.endswith("/" + slot)))]


def load_deepfm(layer, prefix: str, trainer=None) -> Dict[str, np.ndarray]:
    """Copy a shipped DeepFM checkpoint into a ``DeepFMRankingLayer`` (and, with ``trainer``, its Adam
    slots + iteration count): variables are matched by the reference's variable order and shapes
    (bias, embed, w, MLP_layer1 kernels/biases, MLP_layer2) -- 2.FM/CustomLayers.py:263-277."""
    import torch
    ws = deepfm_weights(prefix)
    V, k = layer.feature_dims, layer.embedding_dims
    emb_key = find_by_shape(ws, (V, k))
    w_key = find_by_shape(ws, (V, 1))
    bias_key = [n for n in find_by_shape(ws, (1,)) if "bias" in n.split("/")[-1] and "MLP" not in n and "mlp" not in n]
    assert len(emb_key) == 1 and len(w_key) == 1, (emb_key, w_key, {n: a.shape for n, a in ws.items()})
    dev = layer.rt.device
    layer.embed.copy_(torch.from_numpy(ws[emb_key[0]]).to(dev))
    layer.w.copy_(torch.from_numpy(ws[w_key[0]]).to(dev))
    if bias_key:
        layer.params.set("bias", torch.from_numpy(ws[bias_key[0]]))
    if trainer is not None:
        t = layer.table
        for slot, dst in (("m", t.m), ("v", t.v)):
            dst[:, :k].copy_(torch.from_numpy(ws[emb_key[0] + "/" + slot]).to(dev))
            dst[:, k:k + 1].copy_(torch.from_numpy(ws[w_key[0] + "/" + slot]).to(dev))
    return ws


# ------------------------------------------------------------------ writer
def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_varint_field(fno: int, v: int) -> bytes:
    return _put_varint(fno << 3) + _put_varint(v)


def _pb_bytes_field(fno: int, b: bytes) -> bytes:
    return _put_varint((fno << 3) | 2) + _put_varint(len(b)) + b


_NP2DT = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def _block(entries: List[Tuple[bytes, bytes]]) -> bytes:
    """One SSTable block without prefix compression (every entry is a restart point)."""
    body, restarts = bytearray(), []
    for key, val in entries:
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(val)) + key + val
    for r in restarts or [0]:
        body += struct.pack("<I", r)
    body += struct.pack("<I", max(len(restarts), 1))
    return bytes(body)


def write_checkpoint(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write {key: ndarray} as a single-shard tensor bundle (``prefix.index`` +
    ``prefix.data-00000-of-00001``) that ``tf.train.load_checkpoint`` / this module's reader accept:
    the export direction of the drop-in (weights trained here -> the reference's tooling)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data = bytearray()
    items: List[Tuple[bytes, bytes]] = []
    header = _pb_varint_field(1, 1) + _pb_bytes_field(3, _pb_varint_field(1, 1))      # num_shards=1, version.producer=1
    items.append((b"", header))
    for key in sorted(tensors):
        arr = np.asarray(tensors[key])
        if arr.dtype not in _NP2DT:
            raise ValueError(f"{key}: dtype {arr.dtype} not supported by the bundle writer")
        raw = arr.tobytes(order="C")
        shape = b"".join(_pb_bytes_field(2, _pb_varint_field(1, int(d))) for d in arr.shape)
        entry = (_pb_varint_field(1, _NP2DT[arr.dtype]) + _pb_bytes_field(2, shape) +
                 (_pb_varint_field(4, len(data)) if len(data) else b"") + _pb_varint_field(5, len(raw)) +
                 _put_varint((6 << 3) | 5) + struct.pack("<I", masked_crc32c(raw)))
        items.append((key.encode("utf-8"), entry))
        data += raw
    out = bytearray()

    def emit(block: bytes) -> Tuple[int, int]:
        off = len(out)
        out.extend(block)
        out.append(0)                                                     # kNoCompression
        out.extend(struct.pack("<I", masked_crc32c(block + b"\x00")))
        return off, len(block)

    d_off, d_size = emit(_block(items))
    m_off, m_size = emit(_block([]))
    last_key = items[-1][0] + b"\x00"
    i_off, i_size = emit(_block([(last_key, _put_varint(d_off) + _put_varint(d_size))]))
    footer = _put_varint(m_off) + _put_varint(m_size) + _put_varint(i_off) + _put_varint(i_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(out))
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(bytes(data))
