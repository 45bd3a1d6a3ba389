"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol
the headers declare, and the host reader's metadata / page index / raw page API agree with
the oracle -- no compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"PQG_API[^;(]*?\b(pq(?:g|r|gen)_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pq):
    L = ctypes.CDLL(pq.LIB_PATH)
    names = declared("pqg.h") + declared("pqg_reader.h") + declared("pqg_gen.h")
    assert len(names) > 60
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported by libpqg.so"
    assert sorted(pq.PQG_SYMBOLS) == declared("pqg.h")
    assert sorted(pq.PQR_SYMBOLS) == declared("pqg_reader.h")
    assert sorted(pq.PQGEN_SYMBOLS) == declared("pqg_gen.h")


def test_no_device_means_loud_failure(pq):
    if pq.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pq.PqgError, match="no CUDA device"):
        pq.Context(0)
    r = pq.Reader(os.path.join(GOLD, "mixed.parquet"))
    with pytest.raises(pq.PqgError, match="GPU decoder unavailable"):
        r.read_column_by_idx(0, 0)


def test_host_reader_metadata_matches_oracle(pq, oracle, files):
    for name, path in files.items():
        r = pq.Reader(path)
        h = oracle.open(path)
        try:
            assert r.num_rows == oracle.num_rows(h)
            assert r.num_row_groups == oracle.num_row_groups(h)
            assert r.num_columns == oracle.num_columns(h)
            assert r.num_pages == oracle.num_pages(h)
            assert np.array_equal(r.page_index(), oracle.page_index(h)), name
            for c in range(r.num_columns):
                assert r.column_info(c) == oracle.column_info(h, c)
                assert r.find_column(r.column_info(c)["name"]) == oracle.find_column(h, r.column_info(c)["name"])
            for rg in range(r.num_row_groups):
                assert r.row_group_num_rows(rg) == oracle.row_group_num_rows(h, rg)
            n = r.num_pages
            for pid in sorted({0, 1, n // 2, n - 1}):
                assert r.read_page_data(pid) == oracle.read_page_data(h, pid)
            assert r.read_pages_chunk(0, min(5, n - 1), 3000) == oracle.read_pages_chunk(h, 0, min(5, n - 1), 3000)
            assert r.read_pages_chunk(2, 2, 1 << 20) == oracle.read_pages_chunk(h, 2, 2, 1 << 20)
        finally:
            oracle.close(h)
            r.close()


def test_host_reader_error_text(pq, oracle):
    """same messages as the reference's std::runtime_error (SURVEY.md section 8 b)"""
    path = os.path.join(GOLD, "mixed.parquet")
    r = pq.Reader(path)
    n = r.num_pages
    for call, msg in [(lambda: r.read_page_data(n), f"Global page ID {n} out of range"),
                      (lambda: r.read_pages_chunk(n, n, 10), f"Start page ID {n} out of range"),
                      (lambda: r.read_pages_chunk(0, n, 10), f"End page ID {n} out of range"),
                      (lambda: r.read_pages_chunk(3, 2, 10), "Start page ID must be <= end page ID"),
                      (lambda: r.column_info(99), "Column index 99 out of range"),
                      (lambda: r.read_column("nope"), "Column not found: nope"),
                      (lambda: r.read_column_by_idx(7, 0), "Invalid row group index"),
                      (lambda: r.read_column_by_idx(0, 99), "Invalid column index"),
                      (lambda: r.string_iterator("id"), "Column 'id' is not BYTE_ARRAY (type: INT32)"),
                      (lambda: r.string_iterator("nope"), "Column not found: nope")]:
        with pytest.raises(pq.PqgError) as e:
            call()
        assert str(e.value) == msg
    h = oracle.open(path)
    with pytest.raises(RuntimeError, match=f"Global page ID {n} out of range"):
        oracle.read_page_data(h, n)
    oracle.close(h)
    with pytest.raises(pq.PqgError, match="cannot open file"):
        pq.Reader("/nonexistent/file.parquet")


def test_schema_string(pq, ref, files):
    for name, path in files.items():
        r = pq.Reader(path)
        h = ref.open(path)
        buf = ctypes.create_string_buffer(1 << 16)
        ref._fn("schema_string", ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64)(h, buf, len(buf))
        assert r.schema_string() == buf.value.decode()
        ref.close(h)
        r.close()


def test_descriptor_tables(pq, oracle):
    """the flat tables handed to the GPU: pages contiguous per chunk, row bases consistent"""
    path = os.path.join(GOLD, "mixed.parquet")
    r = pq.Reader(path)
    idx = r.page_index()
    for c in range(r.num_columns):
        chunks, nc, pages, npg, total = r.column_tables(c, -1)
        ci = r.column_info(c)
        sel = idx[idx[:, 3] == ci["column_index"]]
        assert npg == len(sel)
        assert [int(pages[i].payload_off) for i in range(npg)] == [int(x) for x in sel[:, 0]]
        assert [int(pages[i].payload_size) for i in range(npg)] == [int(x) for x in sel[:, 1]]
        row = 0
        for k in range(nc):
            ck = chunks[k]
            assert ck.out_row_base == row and ck.phys_type == ci["type"] and ck.max_def == ci["max_def_level"]
            for q in range(ck.first_page, ck.first_page + ck.n_pages):
                assert pages[q].chunk_idx == k and pages[q].out_row_base == row
                row += pages[q].num_values
            assert ck.num_values == row - ck.out_row_base
        assert total == row == r.num_rows
    r.close()


def test_multi_column_descriptor_tables(pq, files):
    """pqr_columns_tables: column k of the call owns the slots [k * S, (k + 1) * S); every chunk and
    page of the per-column tables shows up exactly once, pages contiguous per chunk, chunks with the
    largest dictionaries first; columns of different value widths are refused"""
    name = "fixed_dict" if "fixed_dict" in files else None
    if name is None:
        pytest.skip("needs the reference-written fixture files")
    r = pq.Reader(files[name])
    cols = [c for c in range(r.num_columns) if r.column_info(c)["type"] in (pq.INT64, pq.DOUBLE)]
    assert len(cols) >= 3
    chunks, nc, pages, npg, total = r.columns_tables(cols, -1)
    S = r.num_rows
    assert total == S * len(cols)
    singles = [r.column_tables(c, -1) for c in cols]
    assert nc == sum(t[1] for t in singles) and npg == sum(t[3] for t in singles)
    seen = set()
    pg = 0
    last_dict = None
    for i in range(nc):
        ck = chunks[i]
        k = cols.index(ck.column)
        base = k * S
        one = [singles[k][0][j] for j in range(singles[k][1]) if singles[k][0][j].row_group == ck.row_group][0]
        assert ck.out_row_base == base + one.out_row_base and ck.num_values == one.num_values
        assert (ck.dict_off, ck.dict_size, ck.dict_num_values, ck.has_dict, ck.n_pages) == \
               (one.dict_off, one.dict_size, one.dict_num_values, one.has_dict, one.n_pages)
        assert ck.first_page == pg
        row = ck.out_row_base
        for q in range(ck.n_pages):
            a, b = pages[pg + q], singles[k][2][one.first_page + q]
            assert a.chunk_idx == i and a.out_row_base == row
            assert (a.payload_off, a.payload_size, a.num_values, a.flags) == (b.payload_off, b.payload_size, b.num_values, b.flags)
            row += a.num_values
        pg += ck.n_pages
        d = ck.dict_num_values * 8 if ck.has_dict else 0
        assert last_dict is None or d <= last_dict  # heaviest gathers start first
        last_dict = d
        seen.add((ck.column, ck.row_group))
    assert len(seen) == nc
    c4 = [c for c in range(r.num_columns) if r.column_info(c)["type"] in (pq.INT32, pq.FLOAT)]
    if c4:
        with pytest.raises(pq.PqgError, match="value width"):
            r.columns_tables([cols[0], c4[0]], -1)
    with pytest.raises(pq.PqgError):
        r.columns_tables([], -1)
    r.close()
