"""GPU parity tests of the regex page-pruning scan and the 4 KB chunk
indexes (a-18 tuple level: reference src/main.cpp:21-32; a-20 page level), through the C-ABI,
against the oracle (oracle/regex_oracle.c backtracking matcher over the reference-order
values; the loop of src/main.cpp restated in oracle/pq_oracle.c)."""
import numpy as np
import pytest

from oraclelib import BYTE_ARRAY, re2_page_bits

pytestmark = pytest.mark.gpu

PATTERNS = [
    r"^[a-z0-9._]+@[a-z0-9.]+\.com$",      # BASELINE config 4
    r"@mail7",                               # unanchored literal
    r"^user1[0-9]*@",
    r"example\.com!!$",
    r"^$",                                   # matches only empty strings
    r"Berlin|Dublin",
    r"city_0000[0-4]",
    r"(ab|cd)+x?y{2,3}",
    r"[^a-z]",
    r".",
    r"\d{9}",
]


def string_columns(r):
    return [c for c in range(r.num_columns) if r.column_info(c)["type"] == BYTE_ARRAY]


def test_regex_prune_matches_oracle(pq, oracle, files):
    for name, path in files.items():
        r = pq.Reader(path)
        ho = oracle.open(path)
        try:
            for c in string_columns(r):
                for pat in PATTERNS:
                    for neg in (False, True):
                        got, ms = r.regex_prune(c, pat, neg)
                        exp = oracle.regex_prune(ho, c, pat, neg)
                        assert got.shape == exp.shape, (name, c, pat)
                        assert np.array_equal(got, exp), (name, r.column_info(c)["name"], pat, neg, int(np.nonzero(got != exp)[0][0]))
        finally:
            oracle.close(ho)
            r.close()


def test_regex_prune_matches_re2(pq, oracle, files):
    """page-match sets identical to RE2's PartialMatch over the reference reader's values (ASCII / valid UTF-8 columns)"""
    pytest.importorskip("pyarrow")
    import oraclelib
    ref = oraclelib.Ref() if oraclelib.Ref.available() else None
    path = files["golden_mixed"]
    r = pq.Reader(path)
    src = ref or oracle
    h = src.open(path)
    try:
        for c in string_columns(r):
            if r.column_info(c)["name"] == "wild":
                continue  # arbitrary bytes: not valid UTF-8
            for pat in PATTERNS:
                for neg in (False, True):
                    got, _ = r.regex_prune(c, pat, neg)
                    exp = re2_page_bits(src, h, c, pat, neg)
                    assert np.array_equal(got, exp), (r.column_info(c)["name"], pat, neg)
    finally:
        src.close(h)
        r.close()


def test_regex_rejects_unsupported_and_non_string(pq, files):
    r = pq.Reader(files["golden_mixed"])
    sc = string_columns(r)[0]
    for bad in (r"(a)\1", r"(?=a)b", r"a{2,1}", r"[z-a]", r"("):
        with pytest.raises(pq.PqgError):
            r.regex_prune(sc, bad)
    other = [c for c in range(r.num_columns) if c not in string_columns(r)][0]
    with pytest.raises(pq.PqgError):
        r.regex_prune(other, "a")
    r.close()


def test_chunk_index_matches_oracle(pq, oracle, files):
    for name, path in files.items():
        r = pq.Reader(path)
        ho = oracle.open(path)
        try:
            for c in string_columns(r):
                cname = r.column_info(c)["name"]
                for cs in (4096, 1, 100, 1 << 20):
                    got, n = r.chunk_index(cname, cs)
                    exp, ne = oracle.chunk_index(ho, cname, cs)
                    assert n == ne, (name, cname, cs, n, ne)
                    assert np.array_equal(got, exp), (name, cname, cs, int(np.nonzero(got != exp)[0][0]))
        finally:
            oracle.close(ho)
            r.close()


def test_chunk_index_errors(pq, files):
    r = pq.Reader(files["golden_mixed"])
    with pytest.raises(pq.PqgError, match="Column not found"):
        r.chunk_index("nope")
    other = [c for c in range(r.num_columns) if c not in string_columns(r)][0]
    with pytest.raises(pq.PqgError, match="is not BYTE_ARRAY"):
        r.chunk_index(r.column_info(other)["name"])
    r.close()


def test_page_chunk_index_matches_oracle(pq, oracle, files):
    for name, path in files.items():
        r = pq.Reader(path)
        ho = oracle.open(path)
        try:
            for c in range(r.num_columns):
                for cs in (4096, 1000, 1 << 16):
                    a = r.page_chunk_index(c, cs)
                    b = oracle.page_chunk_index(ho, c, cs)
                    for x, y, what in zip(a, b, ("page_chunk", "page_off", "chunk_first_page")):
                        assert np.array_equal(x, y), (name, c, cs, what)
        finally:
            oracle.close(ho)
            r.close()


def test_chunk_index_shards_stitch_like_one_run(pq, oracle, files):
    """multi-GPU recipe on one GPU: decode row groups as separate shards, chain them with
    carry_in / carry_out, compare with the single-shard answer (host gather of offsets)."""
    import ctypes as C
    path = files.get("strings", files["golden_mixed"])
    r = pq.Reader(path)
    L = pq.lib()
    ctx = pq.Context(0)
    img = np.fromfile(path, dtype=np.uint8)
    buf = ctx.upload(img.ctypes.data, img.size)
    try:
        for c in string_columns(r):
            cname = r.column_info(c)["name"]
            whole, n_whole = r.chunk_index(cname, 4096)
            carry, base, pieces = 0, 0, []
            for rg in range(r.num_row_groups):
                plan = ctx.plan(buf, r.column_tables(c, rg))
                plan.run()
                plan.finish()
                n = plan.num_slots
                ids = np.zeros(n + 1, dtype=np.uint32)
                nch, cout, ms = C.c_uint64(0), C.c_uint64(0), C.c_float(0)
                rc = L.pqg_chunk_index(ctx.h, plan.h, 4096, carry, base, ids.ctypes.data, C.byref(nch), C.byref(cout), C.byref(ms))
                assert rc == 0, ctx.err()
                pieces.append(ids[:n].astype(np.uint64))
                base += nch.value - 1
                carry = cout.value
                plan.destroy()
            got = np.concatenate(pieces)
            assert base + 1 == n_whole, (cname, base + 1, n_whole)
            assert np.array_equal(got, whole), cname
    finally:
        ctx.buf_free(buf)
        ctx.close()
        r.close()


def test_cli_tools(pq, oracle, files):
    """bin/parser (--regex-column / --chunk-index) and bin/index_test print the oracle's numbers"""
    import os
    import re
    import subprocess
    bindir = os.path.join(pq.PKG_DIR, "bin")
    path = files["golden_mixed"]
    ho = oracle.open(path)
    try:
        col = oracle.find_column(ho, "email")
        out = subprocess.run([os.path.join(bindir, "parser"), path, "--regex-column", "email", "--regex",
                              r"^[a-z0-9._]+@[a-z0-9.]+\.com$", "--neg-regex"], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        bits = oracle.regex_prune(ho, col, r"^[a-z0-9._]+@[a-z0-9.]+\.com$", True)
        assert int(re.search(r"Pages scanned: (\d+)", out.stdout).group(1)) == len(bits)
        assert int(re.search(r"Pages prunable: (\d+)", out.stdout).group(1)) == int((bits == 0).sum())
        out = subprocess.run([os.path.join(bindir, "parser"), path, "--chunk-index", "email"], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        _, n = oracle.chunk_index(ho, "email", 4096)
        assert f"Total tuples: {oracle.num_rows(ho)}" in out.stdout and f"Total chunks: {n}" in out.stdout
        out = subprocess.run([os.path.join(bindir, "index_test"), path, "email"], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr + out.stdout
        _, _, cf = oracle.page_chunk_index(ho, col, 4096)
        assert f"Total chunks: {len(cf)}" in out.stdout and "Lookup self-check: ok" in out.stdout
        out = subprocess.run([os.path.join(bindir, "parser"), path, "--regex-column", "email", "--regex", r"(a)\1"],
                             capture_output=True, text=True, timeout=120)
        assert out.returncode != 0 and "regex" in out.stderr
        out = subprocess.run([os.path.join(bindir, "parser"), path], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and "email" in out.stdout and "data pages" in out.stdout
    finally:
        oracle.close(ho)


def test_device_side_predicate_on_decoded_columns(pq):
    """pqg_plan_filter (device-resident consumer): value <op> constant over the decoded column on the device,
    nulls never match; against numpy on the same column"""
    rng = np.random.default_rng(8)
    n = 300_001
    isn = (rng.random(n) < 0.2).astype(np.uint8)
    i64 = rng.integers(-50, 50, size=n, dtype=np.int64)
    f64 = rng.random(n) * 100 - 50  # (distinct values: a PLAIN column -- the generator refuses NaN in a dictionary)
    f64[::50] = 1.25
    f64[::97] = np.nan
    i32 = rng.integers(-5, 5, size=n, dtype=np.int32)
    g = pq.generate([("a", pq.INT64, 1, -1), ("b", pq.DOUBLE, 1, -1), ("c", pq.INT32, 0, -1)],
                    [dict(fixed=i64, is_null=isn), dict(fixed=f64, is_null=isn), dict(fixed=i32)], [100_000, 100_000, 100_001])
    img = g.to_numpy()
    g.free()
    r = pq.Reader(data=img)
    ctx = pq.Context(0)
    buf = ctx.upload(img.ctypes.data, img.size)
    valid = isn == 0
    ops = {pq.PQG_CMP_EQ: np.equal, pq.PQG_CMP_NE: np.not_equal, pq.PQG_CMP_LT: np.less, pq.PQG_CMP_LE: np.less_equal,
           pq.PQG_CMP_GT: np.greater, pq.PQG_CMP_GE: np.greater_equal}
    try:
        for col, (vt, data, ok, c) in enumerate(((pq.INT64, i64, valid, 7), (pq.DOUBLE, f64, valid, 1.25), (pq.INT32, i32, np.ones(n, bool), -2))):
            plan = ctx.plan(buf, r.column_tables(col, -1))
            plan.run()
            plan.finish()
            for op, fn in ops.items():
                with np.errstate(invalid="ignore"):
                    exp = fn(data, c) & ok
                bits, cnt, ms = plan.filter(vt, op, c, n)
                got = ((bits[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
                assert cnt == int(exp.sum()), (col, op)
                assert np.array_equal(got, exp), (col, op)
                assert plan.filter(vt, op, c, n, want_bits=False)[1] == cnt
            with pytest.raises(pq.PqgError, match="value width"):
                plan.filter(pq.INT32 if vt != pq.INT32 else pq.INT64, pq.PQG_CMP_EQ, 1, n)
            plan.destroy()
    finally:
        ctx.buf_free(buf)
        ctx.close()
        r.close()
