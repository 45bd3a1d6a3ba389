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


def test_page_descriptors_carry_encoding_and_page_type(pq, tmp_path):
    """DataPageHeader.encoding travels in pqg_page_desc.flags bits 8..15 and DATA_PAGE_V2 pages are listed with
    PQG_PAGE_FLAG_V2, so that pqg_plan_create can refuse what the kernels do not decode (SURVEY 8 b)"""
    pa = pytest.importorskip("pyarrow")
    pqa = pytest.importorskip("pyarrow.parquet")
    rng = np.random.default_rng(1)
    n = 4000
    t = pa.table({"i": pa.array(rng.integers(0, 1 << 40, size=n), type=pa.int64()),
                  "d": pa.array(rng.integers(0, 10, size=n), type=pa.int64()),
                  "f": pa.array(rng.random(n), type=pa.float64())})
    base = dict(compression="NONE", data_page_version="1.0", write_statistics=False)

    def flags(path, col):
        r = pq.Reader(path)
        try:
            assert r.num_rows == n
            _, _, pages, npg, _ = r.column_tables(col, -1)
            return [pages[k].flags for k in range(npg)]
        finally:
            r.close()

    p = str(tmp_path / "v1.parquet")
    pqa.write_table(t, p, use_dictionary=["d"], **base)
    # PLAIN; the columns are nullable without nulls: the reader saw one level run per page (routing hints)
    assert set(flags(p, 0)) == {pq.PQG_PAGE_FLAG_LEVELS_SEEN | pq.PQG_PAGE_FLAG_NO_NULLS}
    tn = pa.table({"i": pa.array(rng.integers(0, 1 << 40, size=n), mask=rng.random(n) < 0.3, type=pa.int64())})
    pn = str(tmp_path / "nulls.parquet")
    pqa.write_table(tn, pn, use_dictionary=False, **base)
    assert set(flags(pn, 0)) == {pq.PQG_PAGE_FLAG_LEVELS_SEEN}       # nulls: seen, no NO_NULLS
    assert all(f & 1 and (f >> 8) in (2, 8) for f in flags(p, 1))    # dictionary
    p = str(tmp_path / "delta.parquet")
    pqa.write_table(t, p, use_dictionary=False, column_encoding={"i": "DELTA_BINARY_PACKED", "f": "BYTE_STREAM_SPLIT"}, **base)
    assert {f >> 8 for f in flags(p, 0)} == {5} and {f >> 8 for f in flags(p, 2)} == {9}
    p = str(tmp_path / "v2.parquet")
    pqa.write_table(t, p, use_dictionary=False, **dict(base, data_page_version="2.0"))
    assert all(f & pq.PQG_PAGE_FLAG_V2 for f in flags(p, 0))
    p = str(tmp_path / "snappy.parquet")
    pqa.write_table(t, p, **dict(base, compression="SNAPPY"))
    r = pq.Reader(p)
    with pytest.raises(pq.PqgError, match="Only uncompressed parquet files are supported"):
        r.column_tables(0, -1)
    r.close()


def _thrift_file(bool_list_len, declared_len=None):
    """a minimal file whose FileMetaData carries an unknown field (id 100) of type list<bool> in front of
    num_rows: bool elements are one byte each inside a list (compact protocol)"""
    def varint(x):
        out = bytearray()
        while x >= 0x80:
            out.append((x & 0x7F) | 0x80)
            x >>= 7
        out.append(x)
        return bytes(out)

    def zz(x):
        return varint((x << 1) ^ (x >> 63))
    md = bytearray()
    md += b"\x15" + zz(2)                      # 1: version i32
    md += b"\x19\x1c"                          # 2: schema list<struct>, 1 element
    md += b"\x48" + varint(6) + b"schema" + b"\x15" + zz(0) + b"\x00"   # name (4), num_children (5) = 0
    k = bool_list_len if declared_len is None else declared_len
    md += b"\x09" + zz(100)                    # id 100 (long form): list
    md += (bytes([(k << 4) | 1]) if k < 15 else b"\xf1" + varint(k)) + bytes([1, 2] * (bool_list_len // 2) + [1] * (bool_list_len % 2))
    md += b"\x06" + zz(3) + zz(123)            # 3: num_rows i64 (long form: the id goes backwards)
    md += b"\x19\x0c"                          # 4: row_groups, empty
    md += b"\x00"
    return b"PAR1" + bytes(md) + len(md).to_bytes(4, "little") + b"PAR1"


def test_thrift_skip_of_bool_lists_and_hostile_counts(pq):
    r = pq.Reader(data=np.frombuffer(_thrift_file(5), dtype=np.uint8))
    assert r.num_rows == 123 and r.num_row_groups == 0
    r.close()
    r = pq.Reader(data=np.frombuffer(_thrift_file(40), dtype=np.uint8))
    assert r.num_rows == 123
    r.close()
    # a count far beyond the buffer is a corrupt footer, not a 2^31-iteration loop
    with pytest.raises(pq.PqgError):
        pq.Reader(data=np.frombuffer(_thrift_file(4, declared_len=(1 << 31) - 1), dtype=np.uint8))


def test_extension_tables_are_host_side_only(pq, tmp_path):
    """the opt-in extensions (SNAPPY / DATA_PAGE_V2) are a reader-internal path: the table exports, which feed a plain
    pqg_plan_create, keep refusing such chunks -- with the reference's message by default, by name with extensions on"""
    pa = pytest.importorskip("pyarrow")
    pqa = pytest.importorskip("pyarrow.parquet")
    rng = np.random.default_rng(5)
    n = 3000
    t = pa.table({"i": pa.array(rng.integers(0, 100, size=n), type=pa.int64())})
    snappy, zstd, v2 = (str(tmp_path / f"{k}.parquet") for k in ("snappy", "zstd", "v2"))
    pqa.write_table(t, snappy, compression="SNAPPY", data_page_version="1.0")
    pqa.write_table(t, zstd, compression="ZSTD", data_page_version="1.0")
    pqa.write_table(t, v2, compression="NONE", data_page_version="2.0")
    for path, default_msg, ext_msg in ((snappy, "Only uncompressed parquet files are supported", "not exported"),
                                       (zstd, "Only uncompressed parquet files are supported", "SNAPPY-compressed parquet files are supported .*ZSTD"),
                                       (v2, None, "not exported")):
        r = pq.Reader(path)
        try:
            assert r.num_rows == n
            if default_msg:
                with pytest.raises(pq.PqgError, match=default_msg):
                    r.column_tables(0, -1)
            else:  # V2 pages are listed (flagged) so that plan creation can refuse them
                _, _, pages, npg, _ = r.column_tables(0, -1)
                assert npg >= 1 and all(pages[k].flags & pq.PQG_PAGE_FLAG_V2 for k in range(npg))
        finally:
            r.close()
        r = pq.Reader(path, extensions=True)
        try:
            with pytest.raises(pq.PqgError, match=ext_msg):
                r.column_tables(0, -1)
        finally:
            r.close()
