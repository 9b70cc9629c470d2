"""Staging of the reference's persisted HNSW state into device memory.

What the Rust shim does is a plain KV range scan; every VALUE is handed over undecoded and the GPU does the byte
work (csrc/stage.cu):

  Hs  HnswState   idx/trees/hnsw/mod.rs:61-72   parse_hnsw_state()  (tiny, parsed on the host)
  He  element -> SerializedVector   idx/trees/vector.rs:32-56, hnsw/elements.rs:78-128
  Hn  (layer, node) -> neighbour list   idx/trees/graph.rs:104-126, hnsw/layer.rs:526-540

PARITY UNPINNED (one place: read_varint): the `revision 0.17.0` crate is not vendored; its unsigned varint
(< 251 one byte; 0xFB u16 LE; 0xFC u32 LE; 0xFD u64 LE) is recalled from upstream.  The reference's own byte-level
KATs (key/index/hv.rs:72-101) only exercise single-byte values.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def read_varint(buf, pos):
    b = buf[pos]
    if b < 251:
        return b, pos + 1
    nb = {251: 2, 252: 4, 253: 8}.get(b)
    if nb is None or pos + 1 + nb > len(buf):
        raise ValueError("unsupported or truncated varint")
    return int.from_bytes(buf[pos + 1:pos + 1 + nb], "little"), pos + 1 + nb


def _layer_state(buf, pos):
    rev, pos = read_varint(buf, pos)
    if rev != 1:
        raise ValueError(f"LayerState revision {rev} (expected 1)")
    version, pos = read_varint(buf, pos)
    chunks, pos = read_varint(buf, pos)
    return {"version": version, "chunks": chunks}, pos


def parse_hnsw_state(val):
    """HnswState (hnsw/mod.rs:61-72): {enter_point: Option<u64>, next_element_id, layer0, layers}."""
    buf = bytes(val)
    rev, pos = read_varint(buf, 0)
    if rev != 1:
        raise ValueError(f"HnswState revision {rev} (expected 1)")
    tag = buf[pos]
    pos += 1
    ep = None
    if tag == 1:
        ep, pos = read_varint(buf, pos)
    elif tag != 0:
        raise ValueError("bad Option tag")
    nxt, pos = read_varint(buf, pos)
    layer0, pos = _layer_state(buf, pos)
    n, pos = read_varint(buf, pos)
    layers = []
    for _ in range(n):
        ls, pos = _layer_state(buf, pos)
        layers.append(ls)
    if pos != len(buf):
        raise ValueError("trailing bytes after HnswState")
    return {"enter_point": ep, "next_element_id": nxt, "layer0": layer0, "layers": layers}


def pack_values(items):
    """[(id, value bytes)] in key order -> (blob u8, off u64[n+1], ids u64[n]) as the C ABI wants them."""
    ids = np.fromiter((int(i) for i, _ in items), np.uint64, len(items))
    off = np.zeros(len(items) + 1, np.uint64)
    if len(items):
        off[1:] = np.cumsum(np.fromiter((len(v) for _, v in items), np.uint64, len(items)))
    blob = np.frombuffer(b"".join(bytes(v) for _, v in items), np.uint8) if len(items) else np.zeros(1, np.uint8)
    return np.ascontiguousarray(blob), off, ids


def _vp(a):
    return C.c_void_p(a.ctypes.data)


def decode_vectors(ctx, items, dim, dev_ptr, n_rows, dtype="F32", dev_present=None):
    """He values -> rows of a DEVICE buffer (n_rows x dim of dtype).  Returns the number of rejected values."""
    blob, off, ids = pack_values(items)
    bad = C.c_uint64(0)
    L.check(L.lib().sdb_stage_decode_vectors(ctx.h, _vp(blob), _vp(off), _vp(ids), len(items), int(dim),
                                             L.DTYPE[dtype.upper()], int(n_rows), C.c_void_p(int(dev_ptr)),
                                             C.c_void_p(int(dev_present)) if dev_present else None, C.byref(bad)))
    return bad.value


def decode_nodes(ctx, items, n_elems):
    """Hn values of one layer -> (row_ptr u64[n_elems+1], col_idx u32[e], n_bad)."""
    blob, off, ids = pack_values(items)
    rp, ci, bad = C.c_void_p(), C.c_void_p(), C.c_uint64(0)
    L.check(L.lib().sdb_stage_decode_nodes(ctx.h, _vp(blob), _vp(off), _vp(ids), len(items), int(n_elems),
                                           C.byref(rp), C.byref(ci), C.byref(bad)))
    try:
        row_ptr = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_uint64)), (n_elems + 1,)).copy()
        e = int(row_ptr[-1])
        col_idx = np.ctypeslib.as_array(C.cast(ci, C.POINTER(C.c_uint32)), (max(e, 1),))[:e].copy()
    finally:
        L.lib().sdb_free(rp)
        L.lib().sdb_free(ci)
    return row_ptr, col_idx, bad.value
