"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure-Python statement of the TFRecord files the reference's encoding
stage writes (genomad/modules/nn_classification.py:43-52): ``tf.io.TFRecordWriter`` framing around
``tf.train.Example(features={"sequence": Int64List(tokens)}).SerializeToString()``.

TensorFlow is not installed in this image, so the reference's writer cannot be run; this restatement follows the two
published formats instead (TFRecord framing with masked CRC-32C; the Example/Features/Feature/Int64List protobuf schema)
and is pinned in the tests by (a) the RFC 3720 CRC-32C known-answer vectors and (b) the real ``google.protobuf`` runtime
serialising the same message from a dynamically built copy of the schema.  PARITY vs a TensorFlow-written file: UNPINNED.
"""
import struct

_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = (c >> 8) ^ _TABLE[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def varint(v: int) -> bytes:
    out = bytearray()
    while v >= 128:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _field(tag: int, payload: bytes) -> bytes:
    return bytes([tag]) + varint(len(payload)) + payload


def serialize_example(tokens) -> bytes:
    packed = b"".join(varint(int(t)) for t in tokens)
    int64_list = _field(0x0A, packed)                       # Int64List.value = 1, packed
    feature = _field(0x1A, int64_list)                      # Feature.int64_list = 3
    entry = _field(0x0A, b"sequence") + _field(0x12, feature)   # map entry: key = 1, value = 2
    features = _field(0x0A, entry)                          # Features.feature = 1
    return _field(0x0A, features)                           # Example.features = 1


def frame(data: bytes) -> bytes:
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", masked_crc(head)) + data + struct.pack("<I", masked_crc(data))


def write_tfrecord_bytes(token_rows) -> bytes:
    return b"".join(frame(serialize_example(r)) for r in token_rows)


def read_tfrecord_bytes(blob: bytes):
    """-> list of payloads; raises ValueError on CRC mismatch / truncation."""
    out, p = [], 0
    while p < len(blob):
        if p + 12 > len(blob):
            raise ValueError("truncated header")
        (n,) = struct.unpack_from("<Q", blob, p)
        if struct.unpack_from("<I", blob, p + 8)[0] != masked_crc(blob[p:p + 8]):
            raise ValueError("length crc")
        data = blob[p + 12:p + 12 + n]
        if len(data) != n or p + 16 + n > len(blob):
            raise ValueError("truncated record")
        if struct.unpack_from("<I", blob, p + 12 + n)[0] != masked_crc(data):
            raise ValueError("data crc")
        out.append(data)
        p += 16 + n
    return out
