"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every decode goes through the
C-ABI library (libpqg.so: host reader -> descriptor tables -> CUDA kernels) and is compared
slot by slot -- null flag, variant alternative, payload bits, string bytes -- with the
oracle (oracle/liboracle.so) and, where present, with the unmodified reference itself
(oracle/_ref/libpqref.so) on files written by the reference's own writer."""
import os

import numpy as np
import pytest

from conftest import to_values
from oraclelib import BYTE_ARRAY

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_gpu_present(pq):
    assert pq.device_count() > 0, "no CUDA device: the product has no CPU fallback"


def test_golden_vectors(pq):
    """committed reference outputs (tests/golden/make_golden.py) -- needs neither oracle nor _ref"""
    z = np.load(os.path.join(GOLD, "mixed.npz"))
    r = pq.Reader(os.path.join(GOLD, "mixed.parquet"))
    nrg, nc = (int(x) for x in z["shape"])
    assert np.array_equal(r.page_index(), z["page_index"])
    for c in range(nc):
        ci = r.column_info(c)
        for rg in range(nrg):
            got = r.read_column_by_idx(rg, c)
            p = f"rg{rg}_col{c}_"
            for k in ("is_null", "vidx", "fixed", "str_off", "chars"):
                assert np.array_equal(got[k], z[p + k]), (ci["name"], rg, k)
            pg = r.read_pages(rg, c)
            assert np.array_equal(np.stack([pg["page_num"], pg["page_type"], pg["num_values"]]), z[f"rg{rg}_col{c}_pages"])
        if ci["type"] == BYTE_ARRAY:
            pos, off, _ = r.string_iterator(ci["name"])
            assert np.array_equal(pos, z[f"col{c}_iter_pos"]) and np.array_equal(off, z[f"col{c}_iter_off"])
    r.close()


def test_read_column_matches_oracle_and_reference(pq, oracle, files):
    import oraclelib
    ref = oraclelib.Ref() if oraclelib.Ref.available() else None
    for name, path in files.items():
        r = pq.Reader(path)
        ho = oracle.open(path)
        hr = ref.open(path) if ref else None
        try:
            for c in range(r.num_columns):
                ci = r.column_info(c)
                for rg in range(r.num_row_groups):
                    got = to_values(r.read_column_by_idx(rg, c))
                    d = got.diff(oracle.read_column_by_idx(ho, rg, c))
                    assert d is None, (name, ci["name"], rg, d)
                    if hr:
                        d = got.diff(ref.read_column_by_idx(hr, rg, c))
                        assert d is None, ("vs reference", name, ci["name"], rg, d)
                whole = to_values(r.read_column(ci["name"]))
                assert whole.diff(oracle.read_column(ho, ci["name"])) is None, (name, ci["name"])
                one = to_values(r.read_column(ci["name"], rg=r.num_row_groups - 1))
                assert one.diff(oracle.read_column_by_idx(ho, r.num_row_groups - 1, c)) is None
        finally:
            oracle.close(ho)
            if hr:
                ref.close(hr)
            r.close()


def test_read_pages_matches_oracle(pq, oracle, files):
    for name, path in files.items():
        r = pq.Reader(path)
        ho = oracle.open(path)
        try:
            for c in range(r.num_columns):
                for rg in range(r.num_row_groups):
                    a, b = r.read_pages(rg, c), oracle.read_pages(ho, rg, c)
                    assert np.array_equal(a["page_num"], b.page_num) and np.array_equal(a["page_type"], b.page_type)
                    assert np.array_equal(a["num_values"], b.num_values) and np.array_equal(a["first_value"], b.first_value)
                    assert to_values(a["values"]).diff(b.values) is None, (name, c, rg)
        finally:
            oracle.close(ho)
            r.close()


def test_string_iterator_matches_oracle(pq, oracle, files):
    for name, path in files.items():
        r = pq.Reader(path)
        ho = oracle.open(path)
        try:
            for c in range(r.num_columns):
                ci = r.column_info(c)
                if ci["type"] != BYTE_ARRAY:
                    continue
                a, b = r.string_iterator(ci["name"]), oracle.string_iterator(ho, ci["name"])
                for x, y in zip(a, b):
                    assert np.array_equal(x, y), (name, ci["name"])
        finally:
            oracle.close(ho)
            r.close()


def test_columnar_api_is_consistent_with_values(pq, files):
    """the Value-free columnar result carries the same information"""
    path = files["golden_mixed"]
    r = pq.Reader(path)
    for c in range(r.num_columns):
        ci = r.column_info(c)
        col = r.read_columnar(c, -1)
        vals = r.read_column(ci["name"])
        n = col["num_slots"]
        assert n == len(vals["is_null"])
        valid = np.ones(n, dtype=bool)
        if col["has_validity"]:
            valid = ((col["validity"][np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
        assert np.array_equal(~valid, vals["is_null"].astype(bool))
        if ci["type"] == BYTE_ARRAY:
            lens = np.diff(vals["str_off"].astype(np.int64))
            got = []
            for k in range(col["n_chunks"]):
                base = int(col["chunk_row_base"][k])
                end = int(col["chunk_row_base"][k + 1]) if k + 1 < col["n_chunks"] else n
                off = col["offsets"][base + k: end + k + 1].astype(np.int64)
                got.append(np.diff(off))
                chars = col["chars"][int(col["char_bases"][k]) + int(off[0]): int(col["char_bases"][k]) + int(off[-1])]
                exp = vals["chars"][int(vals["str_off"][base]): int(vals["str_off"][end])]
                assert np.array_equal(chars, exp)
            assert np.array_equal(np.concatenate(got), lens)
        else:
            w = col["width"]
            raw = col["values"].reshape(n, w)
            if w <= 8:
                pad = np.zeros((n, 8), dtype=np.uint8)
                pad[:, :w] = raw
                bits = pad.view(np.uint64).reshape(n)
                assert np.array_equal(bits[valid], vals["fixed"][valid])
                assert not bits[~valid].any()
    r.close()


def test_read_columns_into_pipelined_matches_read_column(pq, files):
    """streaming path (cached plans, per-row-group H2D -> decode -> D2H) == the Value path,
    twice in a row (the second call reuses the cached plans and device buffers)"""
    for name in ("fixed_plain", "fixed_dict", "golden_mixed"):
        if name not in files:
            continue
        r = pq.Reader(files[name])
        cols = [c for c in range(r.num_columns) if r.column_info(c)["type"] in (pq.INT32, pq.INT64, pq.FLOAT, pq.DOUBLE)]
        n = r.num_rows
        for rep in range(2):
            vals = [np.full(n * 8 + 16, 0xAB, dtype=np.uint8) for _ in cols]
            vmask = [np.zeros((n + 31) // 32 + 1, dtype=np.uint32) for _ in cols]
            st = r.read_columns_into(cols, [(v.ctypes.data, v.size, m.ctypes.data, m.size) for v, m in zip(vals, vmask)])
            for c, v, m, s_ in zip(cols, vals, vmask, st):
                exp = r.read_column(r.column_info(c)["name"])
                w = s_["width"]
                assert s_["num_slots"] == n
                valid = ~exp["is_null"].astype(bool)
                if s_["has_validity"]:
                    got_valid = ((m[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
                    assert np.array_equal(got_valid, valid), (name, c, rep)
                else:
                    assert valid.all()
                raw = v[: n * w].reshape(n, w)
                pad = np.zeros((n, 8), dtype=np.uint8)
                pad[:, :w] = raw
                bits = pad.view(np.uint64).reshape(n)
                assert np.array_equal(bits[valid], exp["fixed"][valid]), (name, c, rep)
                assert not bits[~valid].any()
                assert (v[n * w:] == 0xAB).all()  # nothing written past the column
        # one row group only
        rg = r.num_row_groups - 1
        nr = r.row_group_num_rows(rg)
        c = cols[0]
        v = np.zeros(nr * 8, dtype=np.uint8)
        st = r.read_columns_into([c], [(v.ctypes.data, v.size, None, 0)], rg=rg)
        exp = r.read_column(r.column_info(c)["name"], rg=rg)
        w = st[0]["width"]
        pad = np.zeros((nr, 8), dtype=np.uint8)
        pad[:, :w] = v[: nr * w].reshape(nr, w)
        assert np.array_equal(pad.view(np.uint64).reshape(nr)[~exp["is_null"].astype(bool)], exp["fixed"][~exp["is_null"].astype(bool)])
        r.close()


def test_config1_full_size_against_the_reference(pq, oracle, tmp_path):
    """BASELINE.json configs[0] at full size: 1 M rows, INT32 REQUIRED PLAIN `id` + OPTIONAL
    BYTE_ARRAY `city` (8 names, 30 % nulls).  File from the workload generator (byte-identical to
    the reference writer, tests/test_gen_cpu.py); expected values from the UNMODIFIED reference
    where oracle/_ref is present, else from the oracle."""
    import fixtures
    import oraclelib
    from oraclelib import BYTE_ARRAY as BA, INT32, OPTIONAL, REQUIRED, UTF8
    n = 1_000_000
    rng = np.random.default_rng(42)
    idx = rng.integers(0, len(fixtures.CITIES), size=n)
    lens = np.array([len(c) for c in fixtures.CITIES], dtype=np.uint64)[idx]
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    chars = np.frombuffer(b"".join(fixtures.CITIES[i] for i in idx), dtype=np.uint8)
    isn = (rng.integers(0, 10, size=n) < 3).astype(np.uint8)
    specs = [("id", INT32, REQUIRED, -1), ("city", BA, OPTIONAL, UTF8)]
    g = pq.generate(specs, [dict(fixed=np.arange(n, dtype=np.int32)), dict(str_off=off, chars=chars, is_null=isn)], [n])
    path = g.write(str(tmp_path / "cfg1.parquet"))
    g.free()
    r = pq.Reader(path)
    src = oraclelib.Ref() if oraclelib.Ref.available() else oracle
    h = src.open(path)
    try:
        # page geometry of SURVEY.md 8(d) "Config 1": 3907 PLAIN pages of 256 values, 977 dictionary data pages of 1024
        assert r.num_pages == 3907 + 977
        for name in ("id", "city"):
            got = to_values(r.read_column(name))
            d = got.diff(src.read_column(h, name))
            assert d is None, (name, d)
        t2c, nch = r.chunk_index("city", 4096)
        e2c, ench = src.chunk_index(h, "city", 4096)
        assert nch == ench and np.array_equal(t2c, e2c)
    finally:
        src.close(h)
        r.close()


def test_dictionary_form_read_reconstructs_the_strings(pq, oracle, files):
    """late materialisation: uint32 dictionary indices per slot + the chunk dictionaries give
    back exactly the strings (and nulls) of read_column; columns with PLAIN pages are refused"""
    seen_dict = seen_plain = 0
    for name, path in files.items():
        r = pq.Reader(path)
        try:
            for c in range(r.num_columns):
                ci = r.column_info(c)
                if ci["type"] != BYTE_ARRAY:
                    continue
                try:
                    idx, val, st = r.read_dictionary_indices(c)
                except pq.PqgError as e:
                    assert "not dictionary-encoded throughout" in str(e), str(e)
                    seen_plain += 1
                    continue
                seen_dict += 1
                exp = r.read_column(ci["name"])
                n = len(exp["is_null"])
                assert st["num_slots"] == n and st["width"] == 4
                valid = np.ones(n, dtype=bool)
                if val is not None:
                    valid = ((val[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
                assert np.array_equal(~valid, exp["is_null"].astype(bool)), (name, ci["name"])
                assert not idx[~valid].any()
                row = 0
                exp_off, exp_chars = exp["str_off"].astype(np.int64), exp["chars"].tobytes()
                for rg in range(r.num_row_groups):
                    nr = r.row_group_num_rows(rg)
                    off, chars = r.chunk_dictionary(c, rg)
                    for i in list(range(row, min(row + 200, row + nr))) + [row + nr - 1]:
                        if valid[i]:
                            k = int(idx[i])
                            assert chars[off[k]:off[k + 1]] == exp_chars[exp_off[i]:exp_off[i + 1]], (name, ci["name"], i)
                    row += nr
        finally:
            r.close()
    assert seen_dict >= 3 and seen_plain >= 2


def test_large_string_dictionary(pq, oracle, tmp_path):
    """a 70 K-entry BYTE_ARRAY dictionary (1.2 MB dictionary page: the segmented speculative walk of
    k_dict_prepare with 1024 segments), OPTIONAL with 30 % nulls, two row groups"""
    from oraclelib import BYTE_ARRAY as BA, OPTIONAL, UTF8
    rng = np.random.default_rng(17)
    n, nkeys = 600_000, 70_000  # distinct <= non_null / 5 keeps the chunk dictionary-encoded
    keys = np.frombuffer(b"".join(b"city_%06d_x" % i for i in range(nkeys)), dtype=np.uint8).reshape(nkeys, 13)
    idx = rng.integers(0, nkeys, size=2 * n)
    col = dict(str_off=np.arange(2 * n + 1, dtype=np.uint64) * 13, chars=keys[idx].reshape(-1),
               is_null=(rng.random(2 * n) < 0.3).astype(np.uint8))
    g = pq.generate([("s", BA, OPTIONAL, UTF8)], [col], [n, n])
    path = g.write(str(tmp_path / "bigdict.parquet"))
    g.free()
    r = pq.Reader(path)
    h = oracle.open(path)
    try:
        chunks, nc, _, _, _ = r.column_tables(0, -1)
        assert nc == 2 and all(chunks[i].has_dict and chunks[i].dict_num_values > 60_000 for i in range(nc))
        got = to_values(r.read_column("s"))
        assert got.diff(oracle.read_column(h, "s")) is None
        bits, _ = r.regex_prune(0, r"^city_0000[0-9]{2}_x$")
        assert np.array_equal(bits, oracle.regex_prune(h, 0, r"^city_0000[0-9]{2}_x$", False))
        ix, val, _ = r.read_dictionary_indices(0)
        off, chars = r.chunk_dictionary(0, 1)
        exp = oracle.read_column_by_idx(h, 1, 0)
        valid = ((val[np.arange(2 * n) >> 5] >> (np.arange(2 * n) & 31).astype(np.uint32)) & 1).astype(bool)[n:]
        assert np.array_equal(~valid, exp.is_null.astype(bool))
        for i in range(0, n, 997):
            if valid[i]:
                k = int(ix[n + i])
                assert chars[off[k]:off[k + 1]] == exp.string(i)
    finally:
        oracle.close(h)
        r.close()


def test_multi_column_plan_matches_the_per_column_plans(pq, files):
    """pqr_columns_tables: several columns of one value width in ONE plan (column k owns the
    slots [k * S, (k + 1) * S)), chunks listed largest dictionary first -- the decoded bits and
    the validity must equal what one plan per column produces"""
    for name in ("fixed_plain", "fixed_dict"):
        if name not in files:
            continue
        path = files[name]
        img = np.fromfile(path, dtype=np.uint8)
        r = pq.Reader(path)
        for types, dt in (((pq.INT64, pq.DOUBLE), np.uint64), ((pq.INT32, pq.FLOAT), np.uint32)):
            cols = [c for c in range(r.num_columns) if r.column_info(c)["type"] in types]
            if len(cols) < 2:
                continue
            ctx = pq.Context(0)
            buf = ctx.upload(img.ctypes.data, img.size)
            t = r.columns_tables(cols, -1)
            S = t[4] // len(cols)
            plan = ctx.plan(buf, t)
            plan.run()
            plan.finish()
            vals = np.zeros(S * len(cols), dtype=dt)
            valid = np.zeros((S * len(cols) + 31) // 32 + 1, dtype=np.uint32)
            has_validity = plan.validity_ptr is not None and plan.validity_ptr != 0
            plan.download(values=vals.ctypes.data, validity=valid.ctypes.data if has_validity else None)
            ctx.sync()
            slot = np.arange(S * len(cols))
            vbits = ((valid[slot >> 5] >> (slot & 31).astype(np.uint32)) & 1).astype(bool) if has_validity else np.ones(S * len(cols), dtype=bool)
            for k, c in enumerate(cols):
                one = ctx.plan(buf, r.column_tables(c, -1))
                one.run()
                one.finish()
                v1 = np.zeros(S, dtype=dt)
                m1 = np.zeros((S + 31) // 32 + 1, dtype=np.uint32)
                hv = one.validity_ptr is not None and one.validity_ptr != 0
                one.download(values=v1.ctypes.data, validity=m1.ctypes.data if hv else None)
                ctx.sync()
                b1 = ((m1[np.arange(S) >> 5] >> (np.arange(S) & 31).astype(np.uint32)) & 1).astype(bool) if hv else np.ones(S, dtype=bool)
                assert np.array_equal(vals[k * S:(k + 1) * S], v1), (name, r.column_info(c)["name"])
                assert np.array_equal(vbits[k * S:(k + 1) * S], b1), (name, r.column_info(c)["name"])
                one.destroy()
            plan.destroy()
            ctx.buf_free(buf)
            ctx.close()
        # widths cannot be mixed
        c4 = [c for c in range(r.num_columns) if r.column_info(c)["type"] == pq.INT32]
        c8 = [c for c in range(r.num_columns) if r.column_info(c)["type"] == pq.INT64]
        if c4 and c8:
            with pytest.raises(pq.PqgError, match="value width"):
                r.columns_tables([c4[0], c8[0]], -1)
        r.close()
