"""
TFRecord intermediates: the native writer/reader in libgnm.so against (a) the RFC 3720 CRC-32C
known answers, (b) the pure-Python restatement in oracle/tfrecord.py and (c) the real google.protobuf runtime
serialising tf.train.Example from a dynamically built copy of TensorFlow's example.proto / feature.proto schema.
"""
import struct
from pathlib import Path

import numpy as np
import pytest

from genomad_b200 import tfrecord
from oracle import tfrecord as otf

# RFC 3720 appendix B.4
CRC_KAT = [
    (b"", 0x00000000),
    (b"123456789", 0xE3069283),
    (bytes(32), 0x8A9136AA),
    (b"\xff" * 32, 0x62A8AB43),
    (bytes(range(32)), 0x46DD794E),
    (bytes(range(31, -1, -1)), 0x113FDB5C),
]


@pytest.mark.parametrize("data,want", CRC_KAT)
def test_crc32c_known_answers(data, want):
    assert otf.crc32c(data) == want
    assert tfrecord.crc32c(data) == want


def test_crc32c_native_matches_oracle_on_odd_lengths():
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 15, 16, 17, 63, 64, 65, 1000, 7013):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tfrecord.crc32c(b) == otf.crc32c(b)


def _rows(n, seed=1):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 257, (n, tfrecord.TOKENS)).astype(np.uint16)
    if n:
        t[0, :6] = [0, 1, 127, 128, 255, 256]           # the one-byte / two-byte varint boundary
    return t


def _example_class():
    """tf.train.Example rebuilt with the protobuf runtime from the published schema (package tensorflow)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="gnm_test_example.proto", package="gnmtest", syntax="proto3")

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, number, ftype, label=F.LABEL_OPTIONAL, type_name=None, oneof=None, packed=None):
        f = m.field.add(name=name, number=number, type=ftype, label=label)
        if type_name:
            f.type_name = type_name
        if oneof is not None:
            f.oneof_index = oneof
        if packed is not None:
            f.options.packed = packed
        return f

    m = msg("BytesList"); field(m, "value", 1, F.TYPE_BYTES, F.LABEL_REPEATED)
    m = msg("FloatList"); field(m, "value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=True)
    m = msg("Int64List"); field(m, "value", 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
    m = msg("Feature")
    m.oneof_decl.add(name="kind")
    field(m, "bytes_list", 1, F.TYPE_MESSAGE, type_name=".gnmtest.BytesList", oneof=0)
    field(m, "float_list", 2, F.TYPE_MESSAGE, type_name=".gnmtest.FloatList", oneof=0)
    field(m, "int64_list", 3, F.TYPE_MESSAGE, type_name=".gnmtest.Int64List", oneof=0)
    m = msg("Features")
    e = m.nested_type.add(name="FeatureEntry")
    e.options.map_entry = True
    field(e, "key", 1, F.TYPE_STRING)
    field(e, "value", 2, F.TYPE_MESSAGE, type_name=".gnmtest.Feature")
    field(m, "feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".gnmtest.Features.FeatureEntry")
    m = msg("Example"); field(m, "features", 1, F.TYPE_MESSAGE, type_name=".gnmtest.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("gnmtest.Example"))


def test_oracle_example_equals_protobuf_runtime():
    Example = _example_class()
    for row in _rows(3):
        ex = Example()
        ex.features.feature["sequence"].int64_list.value.extend(int(v) for v in row)
        assert ex.SerializeToString(deterministic=True) == otf.serialize_example(row)
        back = Example.FromString(otf.serialize_example(row))
        assert list(back.features.feature["sequence"].int64_list.value) == [int(v) for v in row]


@pytest.mark.parametrize("n", [0, 1, 5, 300])
def test_native_writer_bytes_and_round_trip(tmp_path, n):
    t = _rows(n, seed=n)
    p = tmp_path / f"{n}.tfrec"
    tfrecord.write_tfrecord(p, t, threads=3)
    blob = p.read_bytes()
    if n <= 5:
        assert blob == otf.write_tfrecord_bytes(t)
    else:                                               # spot-check: first/last record + every CRC via the oracle reader
        payloads = otf.read_tfrecord_bytes(blob)
        assert len(payloads) == n
        assert payloads[0] == otf.serialize_example(t[0]) and payloads[-1] == otf.serialize_example(t[-1])
    assert tfrecord.count_records(p) == n
    back = tfrecord.read_tfrecord(p)
    assert back.dtype == np.uint16 and back.shape == (n, tfrecord.TOKENS) and np.array_equal(back, t)


def test_native_reader_accepts_protobuf_runtime_output(tmp_path):
    Example = _example_class()
    t = _rows(4, seed=9)
    blob = b""
    for row in t:
        ex = Example()
        ex.features.feature["sequence"].int64_list.value.extend(int(v) for v in row)
        blob += otf.frame(ex.SerializeToString())
    p = tmp_path / "pb.tfrec"
    p.write_bytes(blob)
    assert np.array_equal(tfrecord.read_tfrecord(p), t)


def test_reader_rejects_corruption(tmp_path):
    t = _rows(2)
    good = otf.write_tfrecord_bytes(t)
    p = tmp_path / "x.tfrec"
    for name, blob, msg in (
        ("data", good[:40] + bytes([good[40] ^ 1]) + good[41:], "data CRC"),
        ("len", bytes([good[0] ^ 1]) + good[1:], "length CRC"),
        ("trunc", good[:-3], "truncated"),
        ("short", good[:5], "truncated"),
    ):
        p.write_bytes(blob)
        with pytest.raises(RuntimeError, match=msg):
            tfrecord.read_tfrecord(p)
    # framing fine, but the payload is not Example{sequence: Int64List[5997]}
    p.write_bytes(otf.frame(otf.serialize_example(t[0][:100])))
    with pytest.raises(RuntimeError, match="Int64List"):
        tfrecord.read_tfrecord(p)
    p.write_bytes(otf.frame(otf.serialize_example(t[0]).replace(b"sequence", b"sequencf")))
    with pytest.raises(RuntimeError, match="Int64List"):
        tfrecord.read_tfrecord(p)
    with pytest.raises(RuntimeError, match="cannot open"):
        tfrecord.read_tfrecord(tmp_path / "missing.tfrec")
    with pytest.raises(ValueError):
        tfrecord.write_tfrecord(p, np.zeros((2, 100), np.uint16))


def test_file_order_is_numeric(tmp_path):
    for stem in (20000, 10000, 100000, 123456):
        (tmp_path / f"{stem}.tfrec").write_bytes(b"")
    assert [int(f.stem) for f in tfrecord.tfrecord_files(tmp_path)] == [10000, 20000, 100000, 123456]


def test_record_size_matches_survey_estimate():
    # "5997 varints (~12 KB) per window"; with tokens uniform on 1..256 half need two bytes -> ~9 KB
    t = _rows(1)
    n = len(otf.write_tfrecord_bytes(t))
    assert 6000 < n < 12100
