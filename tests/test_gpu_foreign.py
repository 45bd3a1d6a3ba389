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
