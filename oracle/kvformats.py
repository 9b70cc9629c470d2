"""CPU restatement of the reference's persisted HNSW value formats (TEST INFRASTRUCTURE, like the rest of oracle/).

  ser_vector / deser_vector    SerializedVector, #[revisioned(revision = 1)] enum {F64,F32,I64,I32,I16}
                               (idx/trees/vector.rs:32-56); byte layout pinned by the KATs of key/index/hv.rs:72-101
                               (dim 3): [revision][variant index][len][little-endian fixed-width elements]
  node_to_val / load_node      UndirectedGraph (idx/trees/graph.rs:104-126): BE u16 count + BE u64 ids; loading
                               inserts one by one into a set -> first occurrence wins
  hnsw_state                   HnswState / LayerState (hnsw/mod.rs:61-72, hnsw/layer.rs:24-29), revisioned structs
  storekey_unescape            the escaping `storekey` applies to a byte slice inside a key (0x00 -> 01 00,
                               0x01 -> 01 01, terminator 00), needed to read the hv.rs KATs

PARITY UNPINNED: varint() for values >= 251 (un-vendored `revision 0.17.0`; recalled: 0xFB+u16le, 0xFC+u32le,
0xFD+u64le).
"""
import numpy as np

VARIANTS = ["F64", "F32", "I64", "I32", "I16"]  # declaration order = wire index (vector.rs:34-41)
_NP = {"F64": "<f8", "F32": "<f4", "I64": "<i8", "I32": "<i4", "I16": "<i2"}


def varint(v):
    if v < 251:
        return bytes([v])
    if v < 1 << 16:
        return b"\xfb" + int(v).to_bytes(2, "little")
    if v < 1 << 32:
        return b"\xfc" + int(v).to_bytes(4, "little")
    return b"\xfd" + int(v).to_bytes(8, "little")


def _read_varint(buf, pos):
    b = buf[pos]
    if b < 251:
        return b, pos + 1
    nb = {251: 2, 252: 4, 253: 8}[b]
    return int.from_bytes(buf[pos + 1:pos + 1 + nb], "little"), pos + 1 + nb


def ser_vector(variant, values):
    a = np.asarray(values).astype(_NP[variant])
    return varint(1) + varint(VARIANTS.index(variant)) + varint(a.size) + a.tobytes()


def deser_vector(val):
    rev, p = _read_varint(val, 0)
    assert rev == 1
    vi, p = _read_varint(val, p)
    n, p = _read_varint(val, p)
    a = np.frombuffer(val[p:], _NP[VARIANTS[vi]])
    assert a.size == n
    return VARIANTS[vi], a


def node_to_val(ids):
    return len(ids).to_bytes(2, "big") + b"".join(int(i).to_bytes(8, "big") for i in ids)


def load_node(val):
    n = int.from_bytes(val[:2], "big")
    out = []
    for j in range(n):
        e = int.from_bytes(val[2 + 8 * j:10 + 8 * j], "big")
        if e not in out:  # DynamicSet::insert
            out.append(e)
    return out


def hnsw_state(enter_point, next_element_id, layer0=(0, 0), layers=()):
    def ls(v):
        return varint(1) + varint(v[0]) + varint(v[1])
    b = varint(1)
    b += b"\x00" if enter_point is None else b"\x01" + varint(enter_point)
    b += varint(next_element_id) + ls(layer0) + varint(len(layers))
    for v in layers:
        b += ls(v)
    return b


def storekey_unescape(b):
    out, i = bytearray(), 0
    while True:
        c = b[i]
        if c == 0:
            return bytes(out), i + 1
        if c == 1:
            out.append(b[i + 1])  # 01 00 -> 00, 01 01 -> 01
            i += 2
        else:
            out.append(c)
            i += 1
