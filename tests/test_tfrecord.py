"""TFRecord input side: the native tf.train.Example parser against an independent pure-Python reader of
the same wire format, the reference's schema (int64 FixedLenFeature(shape=[1]) features + float label,
2.FM/ModelManager.py:127-133), error behaviour, and the pinned column-block layout the Trainer stages from.  CPU only
(the parser is host code inside libetr.so)."""
import os
import struct

import numpy as np
import pytest

import etr_b200  # noqa: F401
from etr_b200 import tfrecord as T
from etr_b200._lib import EtrError
from oracle import tfrecord_py as O

FEATS = ['user_tag0', 'user_tag1', 'item_tag1', 'item_tag2', 'item_tag3']


@pytest.fixture()
def sample(tmp_path, lib_path):
    rng = np.random.default_rng(7)
    N = 777
    ints = {f: rng.integers(0, 5547, size=N) for f in FEATS}
    ints["item_tag3"][:5] = [0, 1, 127, 128, 2 ** 40]                # varint width boundaries
    flt = {"label": (rng.random(N) < 0.3).astype(np.float32)}
    path = str(tmp_path / "data_train_0.tfrecord")
    T.write_examples(path, ints, flt)
    return path, ints, flt, N


def test_native_parser_matches_python_reader(sample):
    path, ints, flt, N = sample
    got = T.parse_file(path, FEATS, ["label"], verify_crc=True)
    ref = [O.parse_example(r) for r in O.records(path)]
    assert len(ref) == N
    for f in FEATS:
        assert np.array_equal(got[f], np.array([r[f][0] for r in ref], dtype=np.int64))
        assert np.array_equal(got[f], ints[f])
    assert np.array_equal(got["label"], np.array([r["label"][0] for r in ref], dtype=np.float32))


def test_batches_are_column_blocks(sample):
    path, ints, flt, N = sample
    ds = T.TFRecordDataset(path, FEATS, ["label"], batch=256, pin=False)
    seen = 0
    for b in ds:
        n = len(b["label"])
        for i, f in enumerate(FEATS):
            assert np.array_equal(b[f].numpy(), ints[f][seen:seen + n])
            if n == 256 and i:
                # columns of one dtype sit back to back in one allocation: the Trainer copies them in ONE transfer
                assert b[f].data_ptr() == b[FEATS[i - 1]].data_ptr() + 256 * 8
        assert np.array_equal(b["label"].numpy(), flt["label"][seen:seen + n])
        seen += n
    assert seen == N
    assert sum(len(b["label"]) for b in T.TFRecordDataset(path, FEATS, ["label"], batch=256, pin=False, drop_remainder=True)) == 768


def test_multi_file_and_fixed_width_features(tmp_path, lib_path):
    rng = np.random.default_rng(1)
    bags = rng.integers(1, 1000, size=(50, 4))
    y = rng.random(50).astype(np.float32)
    p1, p2 = str(tmp_path / "a.tfrecord"), str(tmp_path / "b.tfrecord")
    T.write_examples(p1, {"bag": bags[:20]}, {"label": y[:20]})
    T.write_examples(p2, {"bag": bags[20:]}, {"label": y[20:]})
    out = [b for b in T.TFRecordDataset([p1, p2], ["bag"], ["label"], batch=50, widths={"bag": 4}, pin=False)]
    assert len(out) == 1 and np.array_equal(out[0]["bag"].numpy(), bags) and np.array_equal(out[0]["label"].numpy(), y)


def test_errors_like_tf(sample, tmp_path):
    path, ints, flt, N = sample
    with pytest.raises(EtrError, match="missing"):                    # FixedLenFeature without default
        T.parse_file(path, FEATS + ["no_such_feature"], ["label"])
    with pytest.raises(EtrError, match="wrong list type"):
        T.parse_file(path, ["label"], [])
    with pytest.raises(EtrError, match="values"):
        T.parse_file(path, FEATS, ["label"], widths={"user_tag0": 2})
    raw = bytearray(open(path, "rb").read())
    raw[40] ^= 0x55
    bad = str(tmp_path / "bad.tfrecord")
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(EtrError, match="checksum"):
        T.parse_file(bad, FEATS, ["label"], verify_crc=True)
    open(bad, "wb").write(bytes(raw[:100]))                            # truncated inside the first frame
    with pytest.raises(ValueError, match="truncated"):
        T.parse_file(bad, FEATS, ["label"], verify_crc=False)


def test_unpacked_lists_and_unknown_fields(tmp_path, lib_path):
    """protobuf allows repeated scalars unpacked and unknown fields anywhere: both must parse like TF does."""
    from etr_b200.tf_checkpoint import _pb_bytes_field, _pb_varint_field, _put_varint
    f_int = _pb_bytes_field(3, _pb_varint_field(1, 300))                                    # Int64List.value unpacked
    f_flt = _pb_bytes_field(2, _put_varint((1 << 3) | 5) + struct.pack("<f", 0.5))          # FloatList.value unpacked
    entries = _pb_bytes_field(1, _pb_bytes_field(1, b"x") + _pb_bytes_field(2, f_int))
    entries += _pb_bytes_field(1, _pb_bytes_field(1, b"label") + _pb_bytes_field(2, f_flt))
    entries += _pb_bytes_field(1, _pb_bytes_field(1, b"other") + _pb_bytes_field(2, _pb_bytes_field(1, _pb_bytes_field(1, b"bytes"))))
    ex = _pb_varint_field(9, 7) + _pb_bytes_field(1, entries)                               # unknown field 9 in front
    path = str(tmp_path / "u.tfrecord")
    open(path, "wb").write(T.frame(ex))
    out = T.parse_file(path, ["x"], ["label"])
    assert out["x"].tolist() == [300] and out["label"].tolist() == [0.5]
