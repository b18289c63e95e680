"""Wire format of `serialize()` / `deserialize()`.
Byte-compatible restatement of /root/reference/src/serialization.js:27-140: a
little-endian tree of 32-bit tagged records, every record padded to 4 bytes.

    ARRAY_BUFFER 1  [1][byteLength][bytes, zero-padded to a multiple of 4]   (34-39)
    TYPED_ARRAY  2  [2][index into TypedArraySubClasses][ARRAY_BUFFER]       (40-48)
    ARRAY        3  [3][n] then n x ([byteLength][record])                   (49-64)
    STRING       4  [4][TYPED_ARRAY of the UTF-8 bytes]                      (65-70)
    OBJECT       5  [5][ARRAY of [key, ARRAY_BUFFER(record of value)]]       (79-87)
    NULL         6  [6]                                                      (27-29)
    NUMBER       7  [7][float32]      <- every number travels as Float32     (71-74)
    BOOLEAN      8  [8][float32 1|0]                                         (75-78)
    undefined       [0]                                                      (30-32)

Python values map as: None -> NULL, bytes/bytearray/memoryview -> ArrayBuffer, 1-D numpy
array -> TypedArray of the same element type, list/tuple -> Array, str, bool, int/float
-> Number, dict -> plain object (keys in JavaScript's own enumeration order: array-index
keys ascending, then the others in insertion order).  Host-side, O(set cells): the cells
themselves come from / go to the device through olap_store_export_sparse /
olap_store_import_sparse (stream compaction / scatter kernels)."""
from __future__ import annotations

import struct

import numpy as np

ARRAY_BUFFER, TYPED_ARRAY, ARRAY, STRING, OBJECT, NULL, NUMBER, BOOLEAN = 1, 2, 3, 4, 5, 6, 7, 8

# serialization.js:1-13 (index 2, Uint8ClampedArray, has no numpy twin: read back as uint8)
_TYPED = [np.int8, np.uint8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.float32, np.float64,
          np.int64, np.uint64]
_TYPE_INDEX = {np.dtype(t): i for i, t in reversed(list(enumerate(_TYPED)))}


class _Undefined:
    def __repr__(self):
        return "undefined"


undefined = _Undefined()


def _u32(*values):
    return struct.pack(f"<{len(values)}I", *values)


def _js_key_order(keys):
    """Object.entries order: canonical array indexes (0 .. 2^32-2) ascending, then strings."""
    def index_of(k):
        if isinstance(k, str) and k.isascii() and k.isdigit() and (k == "0" or k[0] != "0") and int(k) < 4294967295:
            return int(k)
        return None

    ints = sorted((k for k in keys if index_of(k) is not None), key=int)
    return ints + [k for k in keys if index_of(k) is None]


def toBuffer(obj) -> bytes:
    if obj is None:
        return _u32(NULL)
    if obj is undefined:
        return _u32(0)
    if isinstance(obj, (bytes, bytearray, memoryview)):
        raw = bytes(obj)
        return _u32(ARRAY_BUFFER, len(raw)) + raw + b"\0" * (-len(raw) % 4)
    if isinstance(obj, np.ndarray):
        arr = np.ascontiguousarray(obj.reshape(-1))
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        index = _TYPE_INDEX.get(arr.dtype.newbyteorder("="))
        if index is None:
            raise TypeError(f"no TypedArray for dtype {arr.dtype}")
        return _u32(TYPED_ARRAY, index) + toBuffer(arr.tobytes())
    if isinstance(obj, (list, tuple)):
        parts = [toBuffer(item) for item in obj]
        return _u32(ARRAY, len(parts)) + b"".join(_u32(len(p)) + p for p in parts)
    if isinstance(obj, str):
        return _u32(STRING) + toBuffer(np.frombuffer(obj.encode("utf-8"), dtype=np.uint8))
    if isinstance(obj, (bool, np.bool_)):
        return _u32(BOOLEAN) + struct.pack("<f", 1.0 if obj else 0.0)
    if isinstance(obj, (int, float, np.integer, np.floating)):
        return _u32(NUMBER) + np.float32(obj).tobytes()
    if isinstance(obj, dict):
        keys = _js_key_order([str(k) for k in obj])
        by_name = {str(k): v for k, v in obj.items()}
        return _u32(OBJECT) + toBuffer([[k, toBuffer(by_name[k])] for k in keys])
    raise TypeError(f"cannot serialize {type(obj).__name__}")


def fromBuffer(buffer, offset: int = 0):
    view = memoryview(buffer).cast("B") if not isinstance(buffer, memoryview) else buffer.cast("B")

    def u32(at):
        return struct.unpack_from("<I", view, at)[0]

    header = u32(offset)
    if header == ARRAY_BUFFER:
        size = u32(offset + 4)
        return bytes(view[offset + 8:offset + 8 + size])
    if header == TYPED_ARRAY:
        dtype = np.dtype(_TYPED[u32(offset + 4)]).newbyteorder("<")
        return np.frombuffer(fromBuffer(view, offset + 8), dtype=dtype).astype(dtype.newbyteorder("="))
    if header == ARRAY:
        size = u32(offset + 4)
        result = []
        at = offset + 8
        for _ in range(size):
            item_size = u32(at)
            result.append(fromBuffer(view, at + 4))
            at += 4 + item_size
        return result
    if header == STRING:
        return fromBuffer(view, offset + 4).tobytes().decode("utf-8")
    if header == NULL:
        return None
    if header == NUMBER:
        return float(struct.unpack_from("<f", view, offset + 4)[0])
    if header == BOOLEAN:
        return struct.unpack_from("<f", view, offset + 4)[0] == 1.0
    if header == OBJECT:
        return {entry[0]: fromBuffer(entry[1]) for entry in fromBuffer(view, offset + 4)}
    return undefined


def toArrayBuffer(buf) -> bytes:  # serialization.js:142-149
    return bytes(buf)


# ---- the store record (in-memory.js:75-116), shared by the device store and the CPU oracle -----

def _to_int32(values, unsigned):
    """`new Int32Array(doubles)` / `new Uint32Array(doubles)`: ToInt32 / ToUint32 (NaN, +-Inf -> 0,
    truncate, wrap modulo 2^32)."""
    v = np.asarray(values, dtype=np.float64)
    t = np.where(np.isfinite(v), np.trunc(v), 0.0)
    m = np.mod(t, 4294967296.0).astype(np.uint64).astype(np.uint32)
    return m if unsigned else m.view(np.int32)


def store_to_buffer(size, type, defaultValue, keys, values) -> bytes:
    """in-memory.js:75-101.  `keys` in the store's Map order."""
    keys = np.asarray(keys, dtype=np.int64)
    if keys.size and int(keys.max()) >= 4294967296:
        raise OverflowError("cell index does not fit the wire format's Uint32Array of indexes")
    if type == "int32":
        data = _to_int32(values, unsigned=False)
    elif type == "uint32":
        data = _to_int32(values, unsigned=True)
    elif type == "float32":
        data = np.asarray(values, dtype=np.float32)
    else:
        data = np.asarray(values, dtype=np.float64)
    return toBuffer({
        "size": size,
        "type": type,
        "defaultValue": defaultValue,
        "indexes": keys.astype(np.uint32),
        "dataBuffer": data,
    })


def store_from_buffer(buffer):
    """in-memory.js:103-116 -> (size, type, defaultValue, keys, values).  `size` is what a
    Float32 kept of it (serialization.js:71-74)."""
    data = fromBuffer(buffer)
    size = data["size"]
    return (int(size), data["type"], data["defaultValue"], np.asarray(data["indexes"], dtype=np.int64),
            np.asarray(data["dataBuffer"], dtype=np.float64))
