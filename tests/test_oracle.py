"""CPU tests of the checker itself: the plain-C restatement (oracle/pq_oracle.c) against
(a) the committed golden vectors produced by the reference and (b) the reference compiled
here (oracle/_ref).  This is what pins the oracle (prompt section 3)."""
import os

import numpy as np
import pytest

import oraclelib
from oraclelib import BYTE_ARRAY, Values

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_values(z, rg, c):
    p = f"rg{rg}_col{c}_"
    return Values(z[p + "is_null"], z[p + "vidx"], z[p + "fixed"], z[p + "str_off"], z[p + "chars"])


def test_oracle_matches_golden(oracle):
    z = np.load(os.path.join(GOLD, "mixed.npz"))
    h = oracle.open(os.path.join(GOLD, "mixed.parquet"))
    try:
        nrg, nc = (int(x) for x in z["shape"])
        assert oracle.num_rows(h) == int(z["num_rows"])
        assert oracle.num_row_groups(h) == nrg and oracle.num_columns(h) == nc
        assert np.array_equal(oracle.page_index(h), z["page_index"])
        for c in range(nc):
            ci = oracle.column_info(h, c)
            assert ci["name"].encode() == z[f"col{c}_name"].tobytes()
            assert [ci["type"], ci["column_index"], ci["max_def_level"], ci["max_rep_level"], ci["repetition"],
                    ci["converted"]] == list(z[f"col{c}_info"])
            for rg in range(nrg):
                assert oracle.read_column_by_idx(h, rg, c).diff(golden_values(z, rg, c)) is None
                p = oracle.read_pages(h, rg, c)
                assert np.array_equal(np.stack([p.page_num, p.page_type, p.num_values]), z[f"rg{rg}_col{c}_pages"])
            if ci["type"] == BYTE_ARRAY:
                pos, off, _ = oracle.string_iterator(h, ci["name"])
                assert np.array_equal(pos, z[f"col{c}_iter_pos"]) and np.array_equal(off, z[f"col{c}_iter_off"])
                for cs in (4096, 256):
                    t2c, n = oracle.chunk_index(h, ci["name"], cs)
                    assert n == int(z[f"col{c}_nchunks_{cs}"])
                    assert np.array_equal(t2c.astype(np.uint32), z[f"col{c}_t2c_{cs}"])
    finally:
        oracle.close(h)


def test_oracle_matches_reference_on_standard_files(oracle, ref, files):
    for name, path in files.items():
        hr, ho = ref.open(path), oracle.open(path)
        try:
            assert ref.num_rows(hr) == oracle.num_rows(ho)
            assert np.array_equal(ref.page_index(hr), oracle.page_index(ho))
            for c in range(ref.num_columns(hr)):
                ci = ref.column_info(hr, c)
                assert ci == oracle.column_info(ho, c)
                for rg in range(ref.num_row_groups(hr)):
                    d = ref.read_column_by_idx(hr, rg, c).diff(oracle.read_column_by_idx(ho, rg, c))
                    assert d is None, (name, c, rg, d)
                    pa, pb = ref.read_pages(hr, rg, c), oracle.read_pages(ho, rg, c)
                    assert np.array_equal(pa.page_num, pb.page_num) and np.array_equal(pa.page_type, pb.page_type)
                    assert np.array_equal(pa.num_values, pb.num_values) and np.array_equal(pa.first_value, pb.first_value)
                    assert pa.values.diff(pb.values) is None
                assert ref.read_column(hr, ci["name"]).diff(oracle.read_column(ho, ci["name"])) is None
                if ci["type"] == BYTE_ARRAY:
                    # chars of the reference iterator dangle for the last string of each page
                    # (parquet_reader.cpp:335-342: next() clears page_strings_ before returning),
                    # so positions and lengths are compared with the reference, bytes with read_column
                    x, y = ref.string_iterator(hr, ci["name"]), oracle.string_iterator(ho, ci["name"])
                    assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1])
                    full = oracle.read_column(ho, ci["name"])
                    assert np.array_equal(y[0], np.nonzero(full.is_null == 0)[0])
                    assert np.array_equal(y[2], full.chars)
                    a, b = ref.chunk_index(hr, ci["name"]), oracle.chunk_index(ho, ci["name"])
                    assert a[1] == b[1] and np.array_equal(a[0], b[0])
        finally:
            ref.close(hr)
            oracle.close(ho)


@pytest.mark.parametrize("bw", list(range(1, 33)))
def test_rle_codec_bit_widths(oracle, ref, bw):
    """RleBpEncoder -> RleDecoder round trip for every bit width 1..32 (reference headers),
    restatement against the reference decoder, with runs, literals and a ragged tail."""
    rng = np.random.default_rng(bw)
    hi = (1 << bw) - 1
    parts = [rng.integers(0, hi + 1, size=37, dtype=np.uint64), np.full(19, hi, dtype=np.uint64),
             rng.integers(0, min(hi, 3) + 1, size=64, dtype=np.uint64), np.zeros(5, dtype=np.uint64),
             rng.integers(0, hi + 1, size=3, dtype=np.uint64)]
    vals = np.concatenate(parts).astype(np.uint32)
    enc = ref.rle_encode(vals, bw)
    a = ref.rle_decode_i32(enc, bw, len(vals))
    b = oracle.rle_decode_i32(enc, bw, len(vals))
    assert np.array_equal(a, b)
    assert np.array_equal(a.view(np.uint32), vals)
    # asking for more values than the stream holds: zero fill (rle_decoder.hpp:21-24)
    a2 = ref.rle_decode_i32(enc, bw, len(vals) + 40)
    b2 = oracle.rle_decode_i32(enc, bw, len(vals) + 40)
    assert np.array_equal(a2, b2)


def test_page_chunk_index_spec(oracle):
    """a-20 frozen spec: greedy packing of a column's pages, chunk closes at >= 4096 bytes."""
    h = oracle.open(os.path.join(GOLD, "mixed.parquet"))
    try:
        idx = oracle.page_index(h)
        for c in range(oracle.num_columns(h)):
            pc, po, cf = oracle.page_chunk_index(h, c, 4096)
            sizes = idx[idx[:, 3] == oracle.column_info(h, c)["column_index"], 1]
            chunk, off, exp_c, exp_o, first = 0, 0, [], [], [0]
            for i, s in enumerate(sizes):
                if i and off >= 4096:
                    chunk += 1
                    off = 0
                    first.append(i)
                exp_c.append(chunk)
                exp_o.append(off)
                off += int(s)
            assert list(pc) == exp_c and list(po) == exp_o and list(cf) == first
    finally:
        oracle.close(h)


def test_oracle_regex_prune_is_pinned_to_re2(oracle, ref):
    """a-19: the oracle's page bitmaps == RE2 (pyarrow's match_substring_regex) applied to the values the
    REFERENCE's ColumnReader::read_pages returns for every page of the golden file"""
    pytest.importorskip("pyarrow")
    from oraclelib import BYTE_ARRAY, re2_page_bits
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mixed.parquet")
    ho = oracle.open(path)
    src, hs = (ref, ref.open(path)) if ref is not None else (oracle, ho)
    try:
        for c in range(oracle.num_columns(ho)):
            ci = oracle.column_info(ho, c)
            if ci["type"] != BYTE_ARRAY or ci["name"] == "wild":
                continue
            for pat in (r"^[a-z0-9._]+@[a-z0-9.]+\.com$", r"@mail7", r"Berlin|Dublin", r"^$", r"[^a-z]", r"\d{9}", r"example\.com!!$"):
                for neg in (False, True):
                    assert np.array_equal(oracle.regex_prune(ho, c, pat, neg), re2_page_bits(src, hs, c, pat, neg)), (ci["name"], pat, neg)
    finally:
        if src is not oracle:
            src.close(hs)
        oracle.close(ho)
