"""Pipelined BYTE_ARRAY read (pqr_read_strings_into): one plan per row group on two alternating device contexts, results in
the caller's buffers -- compared with the whole-column columnar read (itself pinned to the oracle by test_gpu_parity.py)."""
import numpy as np
import pytest

pa = pytest.importorskip("pyarrow")
pq_arrow = pytest.importorskip("pyarrow.parquet")

pytestmark = pytest.mark.gpu


def _bits(words, n):
    i = np.arange(n)
    return ((words[i >> 5] >> (i & 31).astype(np.uint32)) & 1).astype(bool)


def _check(pq, r, col, rg0, rg1, slack=0):
    rows = sum(r.row_group_num_rows(rg) for rg in range(rg0, rg1))
    # expected: row group by row group through the plain read
    exp_strs, exp_null = [], []
    for rg in range(rg0, rg1):
        got = r.read_column_by_idx(rg, col)
        off, chars, isn = got["str_off"], got["chars"].tobytes(), got["is_null"].astype(bool)
        exp_null.append(isn)
        exp_strs += [None if isn[i] else chars[int(off[i]):int(off[i + 1])] for i in range(len(isn))]
    exp_null = np.concatenate(exp_null) if exp_null else np.zeros(0, dtype=bool)
    total_chars = sum(len(s) for s in exp_strs if s is not None)
    nrg = rg1 - rg0
    offsets = np.zeros(rows + 4 * nrg + 8, dtype=np.uint32)
    chars = np.zeros(total_chars + slack + 1, dtype=np.uint8)
    bases = np.zeros(4 * nrg + 2, dtype=np.uint64)
    validity = np.zeros((rows + 31) // 32 + 1, dtype=np.uint32)
    for attempt in range(2):  # the second call runs on the cached plans (no host sync between the passes)
        offsets[:] = 0xFFFFFFFF; chars[:] = 0xEE; validity[:] = 0xFFFFFFFF
        st = r.read_strings_into(col, rg0, rg1, offsets, (chars.ctypes.data, total_chars + slack), bases, validity)
        assert st["num_slots"] == rows and st["chars_size"] == total_chars, (attempt, st)
        nc = st["n_chunks"]
        assert int(bases[nc]) == total_chars
        # walk the chunks: chunk c owns offsets [row_base + c, row_base + n_c + c]
        got, row = [], 0
        valid = _bits(validity, rows) if st["has_validity"] else np.ones(rows, dtype=bool)
        raw = chars.tobytes()
        # the chunk row counts: one chunk per row group unless a dictionary switched inside it; recover them from the tables
        c = 0
        for rg in range(rg0, rg1):
            _, nck, _, _, _ = r.column_tables(col, rg)
            cks = r.column_tables(col, rg)[0]
            for k in range(nck):
                n = cks[k].num_values
                o = offsets[row + c: row + c + n + 1].astype(np.int64)
                b = int(bases[c])
                got += [None if not valid[row + i] else raw[b + o[i]: b + o[i + 1]] for i in range(n)]
                row += n
                c += 1
        assert c == nc and row == rows
        assert np.array_equal(~valid, exp_null), attempt
        assert got == exp_strs, attempt
        assert chars[total_chars:].tolist() == [0xEE] * (len(chars) - total_chars)  # nothing written past the strings
    return st


def test_pipelined_string_read_matches_the_row_group_reads(pq, tmp_path):
    rng = np.random.default_rng(12)
    n = 90_000
    nulls = rng.random(n) < 0.3
    words = np.array([f"city_{v:06d}_{'x' * (v % 7)}" for v in rng.integers(0, 5000, size=n)], dtype=object)
    emails = np.array([f"user{v:09d}@mail{v % 997:03d}.example.com" for v in rng.integers(0, 1 << 30, size=n)], dtype=object)
    t = pa.table({"dict_nulls": pa.array(words, mask=nulls, type=pa.string()),
                  "plain": pa.array(emails, type=pa.string()),
                  "plain_nulls": pa.array(emails, mask=nulls, type=pa.string()),
                  "i": pa.array(rng.integers(0, 9, size=n), type=pa.int64())})
    path = str(tmp_path / "strings.parquet")
    pq_arrow.write_table(t, path, compression="NONE", data_page_version="1.0", write_statistics=False, row_group_size=13_000,
                         data_page_size=8 * 1024, use_dictionary=["dict_nulls"])
    r = pq.Reader(path)
    try:
        nrg = r.num_row_groups
        assert nrg == 7
        for name in ("dict_nulls", "plain", "plain_nulls"):
            col = r.find_column(name)
            _check(pq, r, col, 0, nrg, slack=5)
            _check(pq, r, col, 2, 5)
            _check(pq, r, col, 3, 4)
        # an int column is refused, a chars buffer that is too small says how much is needed
        with pytest.raises(pq.PqgError, match="not BYTE_ARRAY"):
            r.read_strings_into(r.find_column("i"), 0, 1, np.zeros(20000, np.uint32), np.zeros(10, np.uint8), np.zeros(4, np.uint64))
        col = r.find_column("plain")
        with pytest.raises(pq.PqgError, match="chars buffer too small .* bytes needed"):
            r.read_strings_into(col, 0, nrg, np.zeros(n + 64, np.uint32), np.zeros(1000, np.uint8), np.zeros(nrg + 2, np.uint64))
        # ... and the reader still works afterwards
        _check(pq, r, col, 0, nrg)
    finally:
        r.close()


def test_pipelined_string_read_on_the_writer_format(pq, tmp_path):
    """the reference writer's own layout (1 KB pages, RLE levels): dictionary column with nulls + PLAIN column, 5 row groups"""
    rows = 200_000
    a = pq.synth_strings(pq.PQGEN_CITY64K, rows, 3, null_permille=300)
    b = pq.synth_strings(pq.PQGEN_EMAILS, rows, 4)
    g = pq.generate([("city", pq.BYTE_ARRAY, 1, 0), ("email", pq.BYTE_ARRAY, 0, 0)], [a, b], [40_000] * 5)  # (name, type, OPTIONAL / REQUIRED, UTF8)
    img = g.to_numpy()
    g.free()
    r = pq.Reader(data=img)
    try:
        for col in (0, 1):
            _check(pq, r, col, 0, r.num_row_groups)
    finally:
        r.close()
