"""Pins oracle/kvformats.py (and the product's Hs parser) to the reference's byte-level known answers."""
import numpy as np
import pytest

from oracle import kvformats as K

PREFIX = b"/*\x00\x00\x00\x01*\x00\x00\x00\x02*testtb\0+\0\0\0\x03!hv"
# key/index/hv.rs:72-101 -- the Hv key embeds the revisioned SerializedVector as an escaped byte slice
KATS = [
    ("I16", [1, 2, 3], b"\x01\x01\x04\x03\x01\x01\x01\0\x02\x01\0\x03\x01\0\0"),
    ("I32", [1, 2, 3], b"\x01\x01\x03\x03\x01\x01\x01\0\x01\0\x01\0\x02\x01\0\x01\0\x01\0\x03\x01\0\x01\0\x01\0\0"),
    ("I64", [1, 2, 3], b"\x01\x01\x02\x03\x01\x01\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\x02\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\x03\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\0"),
    ("F32", [1.0, 2.0, 3.0], b"\x01\x01\x01\x01\x03\x01\0\x01\0\x80\x3F\x01\0\x01\0\x01\0\x40\x01\0\x01\0\x40\x40\0"),
    ("F64", [1.0, 2.0, 3.0], b"\x01\x01\x01\0\x03\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\xF0\x3F\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\x40\x01\0\x01\0\x01\0\x01\0\x01\0\x01\0\x08\x40\0"),
]


@pytest.mark.parametrize("variant,values,escaped", KATS)
def test_serialized_vector_matches_reference_key_kats(variant, values, escaped):
    raw, used = K.storekey_unescape(escaped)
    assert used == len(escaped)
    assert K.ser_vector(variant, values) == raw
    v, a = K.deser_vector(raw)
    assert v == variant and list(a) == values


def test_node_value_roundtrip_and_first_occurrence_wins():
    # graph.rs:104-126
    assert K.node_to_val([1, 2]) == b"\x00\x02" + (1).to_bytes(8, "big") + (2).to_bytes(8, "big")
    assert K.load_node(K.node_to_val([5, 3, 5, 9, 3])) == [5, 3, 9]
    assert K.load_node(K.node_to_val([])) == []


def test_product_state_parser_reads_what_the_format_writes():
    from surrealdb_b200.staging import parse_hnsw_state
    for ep, nxt, l0, ls in [(None, 0, (0, 0), ()), (7, 1000, (12, 0), ((3, 0), (4, 0))),
                            (70000, 10_000_000, (1 << 33, 2), ((300, 0),) * 6)]:
        st = parse_hnsw_state(K.hnsw_state(ep, nxt, l0, ls))
        assert st["enter_point"] == ep and st["next_element_id"] == nxt
        assert (st["layer0"]["version"], st["layer0"]["chunks"]) == l0
        assert [(l["version"], l["chunks"]) for l in st["layers"]] == list(ls)
    with pytest.raises(ValueError):
        parse_hnsw_state(K.hnsw_state(1, 2) + b"\x00")
