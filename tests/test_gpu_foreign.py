"""Foreign-writer coverage: files written by pyarrow -- uncompressed,
data-page v1, no statistics -- with what the reference's own writer never emits: pages of
64 KB .. 1 MB, bit-packed definition levels, multi-group literal runs, dictionary pages with
fallback to PLAIN inside a chunk.  The oracle (and the compiled reference, which reads such
files correctly: SURVEY.md section 4) give the expected values."""
import numpy as np
import pytest

from conftest import to_values

pa = pytest.importorskip("pyarrow")
pq_arrow = pytest.importorskip("pyarrow.parquet")

pytestmark = pytest.mark.gpu


def write(path, table, **kw):
    opts = dict(compression="NONE", data_page_version="1.0", write_statistics=False, use_dictionary=True)
    opts.update(kw)
    pq_arrow.write_table(table, path, **opts)
    return path


def tables(rng, n):
    nulls = rng.random(n) < 0.25
    i64 = rng.integers(-2**40, 2**40, size=n)
    small = rng.integers(0, 37, size=n)
    strs = np.array([f"value-{v:05d}-{'x' * (v % 9)}" for v in rng.integers(0, 2000, size=n)], dtype=object)
    runs = np.repeat(rng.integers(0, 5, size=n // 50 + 1), 50)[:n]
    return pa.table({
        "i64": pa.array(i64, type=pa.int64()),
        "i64n": pa.array(i64, mask=nulls, type=pa.int64()),
        "i32_small": pa.array(small.astype(np.int32), type=pa.int32()),
        "f64n": pa.array(rng.random(n), mask=nulls, type=pa.float64()),
        "f32": pa.array(rng.random(n).astype(np.float32), type=pa.float32()),
        "str": pa.array(strs, type=pa.string()),
        "strn": pa.array(strs, mask=nulls, type=pa.string()),
        "runs": pa.array(runs, mask=(rng.random(n) < 0.02), type=pa.int64()),
        "b": pa.array(rng.random(n) < 0.5, mask=nulls, type=pa.bool_()),
    })


@pytest.mark.parametrize("variant", ["dict_64k_pages", "plain_1m_pages", "small_pages", "dict_fallback"])
def test_pyarrow_files_match_oracle_and_reference(pq, oracle, tmp_path, variant):
    import oraclelib
    rng = np.random.default_rng(hash(variant) % 1000)
    n = 120_000
    t = tables(rng, n)
    kw = {"dict_64k_pages": dict(data_page_size=64 * 1024, row_group_size=50_000),
          "plain_1m_pages": dict(use_dictionary=False, data_page_size=1 << 20, row_group_size=n),
          "small_pages": dict(data_page_size=700, row_group_size=33_333),
          "dict_fallback": dict(dictionary_pagesize_limit=4096, data_page_size=16 * 1024, row_group_size=n)}[variant]
    path = write(str(tmp_path / f"{variant}.parquet"), t, **kw)
    ref = oraclelib.Ref() if oraclelib.Ref.available() else None
    r = pq.Reader(path)
    ho = oracle.open(path)
    hr = None
    if ref:
        try:
            hr = ref.open(path)  # the reference needs >= 256 bytes behind every page header
        except Exception:
            hr = None
    try:
        assert r.num_rows == n
        for c in range(r.num_columns):
            ci = r.column_info(c)
            for rg in range(r.num_row_groups):
                got = to_values(r.read_column_by_idx(rg, c))
                d = got.diff(oracle.read_column_by_idx(ho, rg, c))
                assert d is None, (variant, ci["name"], rg, d)
                if hr is not None:
                    d = got.diff(ref.read_column_by_idx(hr, rg, c))
                    assert d is None, ("vs reference", variant, ci["name"], rg, d)
            # and against pyarrow's own reading of the file
            col = t.column(ci["name"]).to_pylist()
            whole = r.read_column(ci["name"])
            isn = whole["is_null"].astype(bool)
            assert isn.tolist() == [v is None for v in col], (variant, ci["name"])
            if ci["type"] in (pq.INT32, pq.INT64):
                exp = np.array([0 if v is None else v for v in col], dtype=np.int64)
                if ci["type"] == pq.INT64:
                    gotv = whole["fixed"].astype(np.uint64).view(np.int64)
                else:
                    gotv = whole["fixed"].astype(np.uint32).view(np.int32).astype(np.int64)
                assert np.array_equal(gotv[~isn], exp[~isn]), (variant, ci["name"])
            if ci["type"] == pq.BYTE_ARRAY:
                off, chars = whole["str_off"], whole["chars"].tobytes()
                for i in (0, 1, n // 2, n - 1):
                    if col[i] is not None:
                        assert chars[int(off[i]):int(off[i + 1])].decode() == col[i]
                pat = r"^value-0[0-4]"
                bits, _ = r.regex_prune(c, pat)
                assert np.array_equal(bits, oracle.regex_prune(ho, c, pat, False)), (variant, ci["name"], "regex")
                t2c, nch = r.chunk_index(ci["name"], 4096)
                e2c, ench = oracle.chunk_index(ho, ci["name"], 4096)
                assert nch == ench and np.array_equal(t2c, e2c), (variant, ci["name"], "chunk index")
        # streaming path on the fixed-width columns
        cols = [c for c in range(r.num_columns) if r.column_info(c)["type"] in (pq.INT32, pq.INT64, pq.FLOAT, pq.DOUBLE)]
        vals = [np.zeros(n * 8, dtype=np.uint8) for _ in cols]
        masks = [np.zeros((n + 31) // 32 + 1, dtype=np.uint32) for _ in cols]
        st = r.read_columns_into(cols, [(v.ctypes.data, v.size, m.ctypes.data, m.size) for v, m in zip(vals, masks)])
        for c, v, m, s_ in zip(cols, vals, masks, st):
            exp = r.read_column(r.column_info(c)["name"])
            w = s_["width"]
            valid = ~exp["is_null"].astype(bool)
            pad = np.zeros((n, 8), dtype=np.uint8)
            pad[:, :w] = v[: n * w].reshape(n, w)
            assert np.array_equal(pad.view(np.uint64).reshape(n)[valid], exp["fixed"][valid]), (variant, c)
            if s_["has_validity"]:
                gv = ((m[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
                assert np.array_equal(gv, valid), (variant, c)
    finally:
        oracle.close(ho)
        if hr is not None:
            ref.close(hr)
        r.close()


def _compare_all(pq, oracle, path, expect_ref=True):
    """every column chunk of the file: GPU path vs oracle and (where it opens the file) the compiled reference"""
    import oraclelib
    ref = oraclelib.Ref() if oraclelib.Ref.available() else None
    r = pq.Reader(path)
    ho = oracle.open(path)
    hr = None
    if ref:
        try:
            hr = ref.open(path)
        except Exception:
            hr = None
    checked_ref = 0
    try:
        for c in range(r.num_columns):
            for rg in range(r.num_row_groups):
                got = to_values(r.read_column_by_idx(rg, c))
                d = got.diff(oracle.read_column_by_idx(ho, rg, c))
                assert d is None, (r.column_info(c)["name"], rg, d)
                if hr is not None:
                    d = got.diff(ref.read_column_by_idx(hr, rg, c))
                    assert d is None, ("vs reference", r.column_info(c)["name"], rg, d)
                    checked_ref += 1
        return r.num_columns, checked_ref
    finally:
        oracle.close(ho)
        if hr is not None:
            ref.close(hr)
        r.close()


def test_optional_struct_leaf_max_def_2_many_pages(pq, oracle, tmp_path):
    """optional struct -> optional leaf: max_def = 2, no repetition.  Deterministic in the reference
    (levels: src/reader/column_reader.cpp:146-170; schema walk: src/reader/parquet_reader.cpp:484-543):
    a slot is a value iff its level is 2.  Several hundred small pages per chunk: every page of such a
    chunk is host-listed for the general kernel (the hand-over list must hold them all)."""
    rng = np.random.default_rng(21)
    n = 150_000
    leaf_null = rng.random(n) < 0.2
    struct_null = rng.random(n) < 0.1
    i64 = rng.integers(-2**50, 2**50, size=n)
    f64 = rng.random(n)
    small = rng.integers(0, 50, size=n)
    strs = np.array([f"k{v:04d}" for v in rng.integers(0, 300, size=n)], dtype=object)
    st = pa.StructArray.from_arrays(
        [pa.array(i64, mask=leaf_null, type=pa.int64()), pa.array(f64, mask=leaf_null, type=pa.float64()),
         pa.array(small.astype(np.int32), mask=leaf_null, type=pa.int32()), pa.array(strs, mask=leaf_null, type=pa.string())],
        names=["x", "y", "z", "s"], mask=pa.array(struct_null))
    t = pa.table({"st": st, "plain": pa.array(i64, type=pa.int64())})
    for name, kw in (("small", dict(data_page_size=1500, row_group_size=75_000)),
                     ("small_plain", dict(data_page_size=1500, row_group_size=75_000, use_dictionary=False)),
                     ("big", dict(data_page_size=1 << 20, row_group_size=n))):
        path = write(str(tmp_path / f"nested_{name}.parquet"), t, **kw)
        r = pq.Reader(path)
        infos = [r.column_info(c) for c in range(r.num_columns)]
        r.close()
        assert [i["max_def_level"] for i in infos] == [2, 2, 2, 2, 1] and all(i["max_rep_level"] == 0 for i in infos)  # pyarrow fields are nullable
        ncols, _ = _compare_all(pq, oracle, path)
        assert ncols == 5
        # and pyarrow's reading: null iff struct or leaf is null
        r = pq.Reader(path)
        got = r.read_column("st.x") if r.find_column("st.x") >= 0 else r.read_column(infos[0]["name"])
        assert np.array_equal(got["is_null"].astype(bool), leaf_null | struct_null)
        valid = ~(leaf_null | struct_null)
        assert np.array_equal(got["fixed"].astype(np.uint64).view(np.int64)[valid], i64[valid])
        r.close()


def test_oversized_pages_with_nulls_beyond_the_big_page_kernel(pq, oracle, tmp_path):
    """one page of > 131072 slots with nulls: the one-CTA-per-page kernel hands it on to the general kernel"""
    rng = np.random.default_rng(22)
    n = 300_000
    nulls = rng.random(n) < 0.3
    t = pa.table({"a": pa.array(rng.integers(-2**40, 2**40, size=n), mask=nulls, type=pa.int64()),
                  "d": pa.array(rng.integers(0, 1000, size=n), mask=nulls, type=pa.int64())})
    path = write(str(tmp_path / "huge_pages.parquet"), t, data_page_size=8 << 20, row_group_size=n, dictionary_pagesize_limit=8 << 20)
    _compare_all(pq, oracle, path)


def test_int96_renders_like_the_reference(pq, oracle, tmp_path):
    """INT96 values come back as the STRING "INT96(high:low)" (reference src/reader/column_reader.cpp:257-264)"""
    import datetime
    rng = np.random.default_rng(23)
    n = 20_000
    base = datetime.datetime(2001, 1, 1)
    ts = [None if rng.random() < 0.1 else base + datetime.timedelta(seconds=int(s), microseconds=int(u))
          for s, u in zip(rng.integers(0, 10**9, size=n), rng.integers(0, 10**6, size=n))]
    t = pa.table({"ts": pa.array(ts, type=pa.timestamp("us")), "ts_req": pa.array([v or base for v in ts], type=pa.timestamp("ns"))})
    for name, kw in (("dict", {}), ("plain", dict(use_dictionary=False))):
        path = write(str(tmp_path / f"int96_{name}.parquet"), t, use_deprecated_int96_timestamps=True, data_page_size=2000, **kw)
        r = pq.Reader(path)
        assert r.column_info(0)["type"] == pq.INT96
        got = r.read_column("ts")
        r.close()
        assert int(got["vidx"][~got["is_null"].astype(bool)][0]) == 5  # the string alternative of Value
        s0 = got["chars"].tobytes()[int(got["str_off"][0]):int(got["str_off"][1])] if not got["is_null"][0] else b"INT96("
        assert s0.startswith(b"INT96(")
        _compare_all(pq, oracle, path)


def test_unsupported_files_fail_with_explicit_errors(pq, tmp_path):
    """compressed chunks: the reference's message (src/reader/column_reader.cpp:13-15); DATA_PAGE_V2 and
    DELTA_* / BYTE_STREAM_SPLIT pages: explicit errors instead of the reference's skipped / garbage decode
    (column_reader.cpp:66-67,173-222)"""
    rng = np.random.default_rng(24)
    n = 5000
    t = pa.table({"i": pa.array(rng.integers(0, 1 << 40, size=n), type=pa.int64()),
                  "f": pa.array(rng.random(n), type=pa.float64()),
                  "s": pa.array([f"s{v}" for v in rng.integers(0, 100, size=n)], type=pa.string())})
    cases = [
        ("snappy", dict(compression="SNAPPY"), "Only uncompressed parquet files are supported", ["i", "f", "s"]),
        ("zstd", dict(compression="ZSTD"), "Only uncompressed parquet files are supported", ["i", "f", "s"]),
        ("v2", dict(data_page_version="2.0"), "DATA_PAGE_V2", ["i", "f", "s"]),
        ("delta", dict(use_dictionary=False, column_encoding={"i": "DELTA_BINARY_PACKED"}), "DELTA_BINARY_PACKED", ["i"]),
        ("bss", dict(use_dictionary=False, column_encoding={"f": "BYTE_STREAM_SPLIT"}), "BYTE_STREAM_SPLIT", ["f"]),
        ("dlba", dict(use_dictionary=False, column_encoding={"s": "DELTA_LENGTH_BYTE_ARRAY"}), "DELTA_LENGTH_BYTE_ARRAY", ["s"]),
        ("dba", dict(use_dictionary=False, column_encoding={"s": "DELTA_BYTE_ARRAY"}), "DELTA_BYTE_ARRAY", ["s"]),
    ]
    for name, kw, msg, cols in cases:
        path = write(str(tmp_path / f"unsupported_{name}.parquet"), t, **kw)
        r = pq.Reader(path)
        try:
            assert r.num_rows == n  # the file opens: footer and page walk are fine
            for c in cols:
                with pytest.raises(pq.PqgError, match=msg):
                    r.read_column(c)
                with pytest.raises(pq.PqgError, match=msg):
                    r.read_columnar(r.find_column(c))
                if c == "s":
                    with pytest.raises(pq.PqgError, match=msg):
                        r.regex_prune(r.find_column(c), "s1")
                    with pytest.raises(pq.PqgError, match=msg):
                        r.string_iterator(c)
            # columns the option did not touch still decode
            for c in ("i", "f", "s"):
                if c not in cols:
                    got = r.read_column(c)
                    assert len(got["is_null"]) == n
        finally:
            r.close()


def _check_against_pyarrow(pq, r, t, n, label, skip=()):
    """every column of table t read through r.read_column vs pyarrow's own values (the oracle of the extension tests)"""
    for c in range(r.num_columns):
        ci = r.column_info(c)
        if ci["name"] in skip:
            continue
        col = t.column(ci["name"]).to_pylist()
        whole = r.read_column(ci["name"])
        isn = whole["is_null"].astype(bool)
        assert isn.tolist() == [v is None for v in col], (label, ci["name"], "nulls")
        ok = ~isn
        if ci["type"] in (pq.INT32, pq.INT64):
            exp = np.array([0 if v is None else v for v in col], dtype=np.int64)
            gotv = whole["fixed"].astype(np.uint64).view(np.int64) if ci["type"] == pq.INT64 else whole["fixed"].astype(np.uint32).view(np.int32).astype(np.int64)
            assert np.array_equal(gotv[ok], exp[ok]), (label, ci["name"])
        elif ci["type"] == pq.DOUBLE:
            exp = np.array([0.0 if v is None else v for v in col], dtype=np.float64)
            assert np.array_equal(whole["fixed"].astype(np.uint64).view(np.float64)[ok], exp[ok]), (label, ci["name"])
        elif ci["type"] == pq.FLOAT:
            exp = np.array([0.0 if v is None else v for v in col], dtype=np.float32)
            assert np.array_equal(whole["fixed"].astype(np.uint32).view(np.float32)[ok], exp[ok]), (label, ci["name"])
        elif ci["type"] == pq.BOOLEAN:
            exp = np.array([bool(v) for v in col], dtype=bool)
            assert np.array_equal(whole["fixed"].astype(bool)[ok], exp[ok]), (label, ci["name"])
        elif ci["type"] == pq.BYTE_ARRAY:
            off, chars = whole["str_off"], whole["chars"].tobytes()
            got = [None if isn[i] else chars[int(off[i]):int(off[i + 1])].decode() for i in range(n)]
            assert got == col, (label, ci["name"])
        # the columnar read takes the same plan
        cc = r.read_columnar(c)
        assert cc["num_slots"] == n, (label, ci["name"])


@pytest.mark.parametrize("variant", ["snappy_v1", "v2_plain", "v2_snappy", "snappy_small_pages", "v2_snappy_no_dict"])
def test_extensions_snappy_and_data_page_v2_against_pyarrow(pq, tmp_path, variant):
    """SURVEY 8 f-3, beyond the reference (which refuses compressed chunks and skips DATA_PAGE_V2): with extensions on, the
    pages are rewritten on the device into the DATA_PAGE layout (SNAPPY blocks decoded, a length word put in front of V2
    definition levels) and decoded by the usual kernels.  pyarrow is the oracle."""
    rng = np.random.default_rng(len(variant))
    n = 60_000
    t = tables(rng, n)
    # long repeats and text: SNAPPY emits copies with 1-, 2- and 4-byte offsets and overlapping runs
    t = t.append_column("rep", pa.array(["abcabcabc" * int(k) for k in rng.integers(0, 40, size=n)], type=pa.string()))
    t = t.append_column("zeros", pa.array(np.zeros(n, dtype=np.int64), mask=rng.random(n) < 0.5, type=pa.int64()))
    kw = {"snappy_v1": dict(compression="SNAPPY", data_page_size=64 * 1024, row_group_size=25_000),
          "v2_plain": dict(data_page_version="2.0", data_page_size=8 * 1024, row_group_size=n),
          "v2_snappy": dict(data_page_version="2.0", compression="SNAPPY", data_page_size=32 * 1024, row_group_size=40_000),
          "snappy_small_pages": dict(compression="SNAPPY", data_page_size=600, row_group_size=n),
          "v2_snappy_no_dict": dict(data_page_version="2.0", compression="SNAPPY", use_dictionary=False, data_page_size=1 << 20, row_group_size=n)}[variant]
    path = write(str(tmp_path / f"ext_{variant}.parquet"), t, **kw)
    # the default reader keeps the reference's refusals
    r0 = pq.Reader(path)
    try:
        with pytest.raises(pq.PqgError, match="Only uncompressed parquet files are supported|DATA_PAGE_V2"):
            r0.read_column("i64")
    finally:
        r0.close()
    r = pq.Reader(path, extensions=True)
    try:
        assert r.num_rows == n
        # (pyarrow writes BOOLEAN values of DATA_PAGE_V2 pages RLE-encoded: an encoding the decoder names and refuses)
        v2 = variant.startswith("v2")
        _check_against_pyarrow(pq, r, t, n, variant, skip=("b",) if v2 else ())
        if v2:
            with pytest.raises(pq.PqgError, match="encoding RLE is not supported"):
                r.read_column("b")
        # the streaming reads take the same plans (upload -> rewrite + decode -> download per column, no per-chunk overlap)
        cols = [c for c in range(r.num_columns) if r.column_info(c)["type"] in (pq.INT32, pq.INT64, pq.FLOAT, pq.DOUBLE)]
        vals = [np.zeros(n * 8, dtype=np.uint8) for _ in cols]
        masks = [np.zeros((n + 31) // 32 + 1, dtype=np.uint32) for _ in cols]
        st = r.read_columns_into(cols, [(v.ctypes.data, v.size, m.ctypes.data, m.size) for v, m in zip(vals, masks)])
        for c, v, m, s_ in zip(cols, vals, masks, st):
            exp = r.read_column(r.column_info(c)["name"])
            w = s_["width"]
            valid = ~exp["is_null"].astype(bool)
            pad = np.zeros((n, 8), dtype=np.uint8)
            pad[:, :w] = v[: n * w].reshape(n, w)
            assert np.array_equal(pad.view(np.uint64).reshape(n)[valid], exp["fixed"][valid]), (variant, c)
            if s_["has_validity"]:
                gv = ((m[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
                assert np.array_equal(gv, valid), (variant, c)
        # ... and so does the pipelined string read
        sc = r.find_column("strn")
        whole = r.read_column("strn")
        nrg = r.num_row_groups
        offs = np.zeros(n + 4 * nrg + 8, dtype=np.uint32)
        chars = np.zeros(len(whole["chars"]) + 16, dtype=np.uint8)
        bases = np.zeros(4 * nrg + 2, dtype=np.uint64)
        vmask = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
        st = r.read_strings_into(sc, 0, nrg, offs, chars, bases, vmask)
        assert st["num_slots"] == n and st["chars_size"] == len(whole["chars"])
        assert chars[:st["chars_size"]].tobytes() == whole["chars"].tobytes()
        # table exports keep refusing what a plain pqg_plan_create cannot decode
        with pytest.raises(pq.PqgError, match="not exported"):
            r.column_tables(0, -1)
    finally:
        r.close()


def test_extensions_refuse_other_codecs_and_report_corrupt_snappy(pq, tmp_path):
    rng = np.random.default_rng(3)
    n = 4000
    t = pa.table({"i": pa.array(rng.integers(0, 50, size=n), type=pa.int64())})
    path = write(str(tmp_path / "zstd.parquet"), t, compression="ZSTD")
    r = pq.Reader(path, extensions=True)
    try:
        with pytest.raises(pq.PqgError, match="SNAPPY-compressed parquet files are supported .*ZSTD"):
            r.read_column("i")
    finally:
        r.close()
    # a SNAPPY page with a flipped byte in its body: explicit error, no garbage
    path = write(str(tmp_path / "snappy.parquet"), t, compression="SNAPPY", use_dictionary=False)
    raw = bytearray(open(path, "rb").read())
    # find the data page payload through the page index (the reader walks headers only)
    r = pq.Reader(path)
    pi = r.page_index()
    r.close()
    off, size = int(pi[0][0]), int(pi[0][1])
    # the header sits in front of the payload: corrupt the SNAPPY preamble (uncompressed length) near the end of the page
    raw[off + size - 1] ^= 0xFF
    raw[off + size // 2] ^= 0x5A
    bad = str(tmp_path / "snappy_bad.parquet")
    open(bad, "wb").write(bytes(raw))
    r = pq.Reader(bad, extensions=True)
    try:
        try:
            got = r.read_column("i")
            # a flipped literal byte can still be a well-formed block: then the values differ but nothing crashes
            assert len(got["is_null"]) == n
        except pq.PqgError as e:
            assert "decompress" in str(e) or "ByteBuffer" in str(e) or "page" in str(e)
    finally:
        r.close()


def test_dictionary_form_read_of_big_foreign_pages_with_nulls(pq, tmp_path):
    """dictionary-form read (indices instead of strings) of a pyarrow dictionary column whose pages hold thousands of slots
    and nulls: those pages take the block decode (pqg_flat.cu), which for such plans emits the index itself"""
    rng = np.random.default_rng(31)
    n = 150_000
    words = np.array([f"w{v:05d}" for v in rng.integers(0, 3000, size=n)], dtype=object)
    t = pa.table({"s": pa.array(words, mask=rng.random(n) < 0.3, type=pa.string())})
    path = write(str(tmp_path / "dictform.parquet"), t, data_page_size=64 * 1024, row_group_size=60_000)
    r = pq.Reader(path)
    try:
        exp = r.read_column("s")
        idx, val, st = r.read_dictionary_indices(0)
        assert st["num_slots"] == n
        valid = ((val[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
        assert np.array_equal(~valid, exp["is_null"].astype(bool))
        assert not idx[~valid].any()
        exp_off, exp_chars = exp["str_off"].astype(np.int64), exp["chars"].tobytes()
        row = 0
        for rg in range(r.num_row_groups):
            nr = r.row_group_num_rows(rg)
            off, chars = r.chunk_dictionary(0, rg)
            got = [chars[off[k]:off[k + 1]] for k in idx[row:row + nr][valid[row:row + nr]]]
            want = [exp_chars[exp_off[i]:exp_off[i + 1]] for i in range(row, row + nr) if valid[i]]
            assert got == want, rg
            row += nr
    finally:
        r.close()
