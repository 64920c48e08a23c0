"""Pure-Python TFRecord / tf.train.Example reader -- TEST INFRASTRUCTURE ONLY (the checker for etr_tfrecord_parse).

Restates the published formats independently of the native parser: TFRecord framing (uint64 length, masked crc32c of the
length, payload, masked crc32c of the payload -- tensorflow/core/lib/io/record_writer.cc) and the protobuf wire format of
tf.train.Example (tensorflow/core/example/{example,feature}.proto), which is what tf.io.parse_single_example decodes in
2.FM/ModelManager.py:127-133."""
import struct
from typing import Dict, Iterator, List


def _varint(b: bytes, p: int):
    v = s = 0
    while True:
        x = b[p]
        p += 1
        v |= (x & 0x7F) << s
        if not x & 0x80:
            return v, p
        s += 7


def _fields(b: bytes):
    p = 0
    while p < len(b):
        tag, p = _varint(b, p)
        fno, wt = tag >> 3, tag & 7
        if wt == 0:
            v, p = _varint(b, p)
        elif wt == 1:
            v, p = b[p:p + 8], p + 8
        elif wt == 2:
            n, p = _varint(b, p)
            v, p = b[p:p + n], p + n
        elif wt == 5:
            v, p = b[p:p + 4], p + 4
        else:
            raise ValueError(wt)
        yield fno, wt, v


def records(path: str) -> Iterator[bytes]:
    data = open(path, "rb").read()
    p = 0
    while p < len(data):
        (n,) = struct.unpack_from("<Q", data, p)
        yield data[p + 12:p + 12 + n]
        p += 12 + n + 4


def parse_example(rec: bytes) -> Dict[str, List]:
    out: Dict[str, List] = {}
    for fno, _, feats in _fields(rec):
        if fno != 1:
            continue
        for f2, _, entry in _fields(feats):
            if f2 != 1:
                continue
            key, feat = None, b""
            for f3, _, v in _fields(entry):
                if f3 == 1:
                    key = v.decode()
                elif f3 == 2:
                    feat = v
            vals: List = []
            for f4, _, lst in _fields(feat):
                for f5, wt, v in _fields(lst):
                    if f5 != 1:
                        continue
                    if f4 == 3:                                   # Int64List
                        if wt == 2:
                            q = 0
                            while q < len(v):
                                x, q = _varint(v, q)
                                vals.append(x - (1 << 64) if x >> 63 else x)
                        else:
                            vals.append(v - (1 << 64) if v >> 63 else v)
                    elif f4 == 2:                                 # FloatList
                        if wt == 2:
                            vals += list(struct.unpack("<%df" % (len(v) // 4), v))
                        else:
                            vals.append(struct.unpack("<f", v)[0])
            out[key] = vals
    return out
