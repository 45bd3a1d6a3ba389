"""The workload generator (include/pqg_gen.h) must emit files BYTE-IDENTICAL to the reference's
ParquetWriter for the same values -- that is what lets bench.py claim its synthetic inputs are
"files written by the repo's own writer" at sizes the real writer cannot produce in time.
Needs oracle/_ref (the compiled reference); skipped where it is absent."""
import os

import numpy as np
import pytest

import fixtures
from oraclelib import BYTE_ARRAY, INT64, REQUIRED, fixed_col


def _concat(cols):
    """columns of several row groups -> one whole-file column"""
    if "fixed" in cols[0]:
        fx = np.concatenate([c["fixed"] for c in cols])
        out = dict(fixed=fx)
    else:
        offs, base = [np.zeros(1, dtype=np.uint64)], 0
        for c in cols:
            offs.append(c["str_off"][1:] + np.uint64(base))
            base += int(c["str_off"][-1])
        out = dict(str_off=np.concatenate(offs), chars=np.concatenate([c["chars"] for c in cols]))
    if any(c.get("is_null") is not None for c in cols):
        out["is_null"] = np.concatenate([
            c["is_null"] if c.get("is_null") is not None else
            np.zeros(len(c["fixed"]) if "fixed" in c else len(c["str_off"]) - 1, dtype=np.uint8) for c in cols])
    return out


def _nrows(col):
    return len(col["fixed"]) if "fixed" in col else len(col["str_off"]) - 1


@pytest.mark.parametrize("name", sorted(fixtures.standard_files()))
def test_generator_is_byte_identical_to_reference_writer(pq, ref, tmp_path, name):
    specs, fn = fixtures.standard_files()[name]
    rgs = fn(np.random.default_rng(7))
    path = str(tmp_path / "ref.parquet")
    ref.write_file(path, specs, rgs)
    want = np.fromfile(path, dtype=np.uint8)
    cols = [_concat([rg[c] for rg in rgs]) for c in range(len(specs))]
    g = pq.generate(specs, cols, [_nrows(rg[0]) for rg in rgs], threads=3)
    got = g.to_numpy()
    assert got.size == want.size, (name, got.size, want.size)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (name, "first differing byte", int(bad[0]))
    p2 = g.write(str(tmp_path / "gen.parquet"))
    assert np.array_equal(np.fromfile(p2, dtype=np.uint8), want)


def test_generator_edge_shapes(pq, ref, tmp_path):
    """short RLE tails, partial bit-packed groups, all-null pages, empty row group, wide varints"""
    rng = np.random.default_rng(3)
    n = 40000
    runs = np.repeat(rng.integers(0, 50, size=n // 3 + 1), rng.integers(1, 7, size=n // 3 + 1))[:n]
    isn = np.zeros(n, dtype=np.uint8)
    isn[5000:9000] = 1  # several all-null pages
    isn[rng.random(n) < 0.02] = 1
    specs = [("runs", INT64, 1, -1), ("two", INT64, REQUIRED, -1), ("s", BYTE_ARRAY, 1, 0),
             ("big", BYTE_ARRAY, REQUIRED, -1)]
    strs = [b"k%d" % (v % 11) for v in runs]
    big = [bytes([65 + (i % 26)]) * (1 + (i * 37) % 2500) for i in range(300)]
    big = (big * (n // 300 + 1))[:n]
    from oraclelib import strings_to_col
    rg1 = [fixed_col(runs.astype(np.int64) * 3, isn), fixed_col((np.arange(n) // 777 % 2).astype(np.int64)),
           strings_to_col(strs, isn), strings_to_col(big)]
    m = 11
    rg2 = [fixed_col(np.arange(m, dtype=np.int64) % 2, np.zeros(m, dtype=np.uint8)), fixed_col(np.zeros(m, dtype=np.int64)),
           strings_to_col([b"x"] * m, np.ones(m, dtype=np.uint8)), strings_to_col([b""] * m)]
    path = str(tmp_path / "edge.parquet")
    ref.write_file(path, specs, [rg1, rg2])
    want = np.fromfile(path, dtype=np.uint8)
    cols = [_concat([rg1[c], rg2[c]]) for c in range(4)]
    got = pq.generate(specs, cols, [n, m]).to_numpy()
    assert got.size == want.size and np.array_equal(got, want)


def test_generator_rejects_bad_input(pq):
    with pytest.raises(pq.PqgError):
        pq.generate([("d", pq.DOUBLE, 0, -1)], [dict(fixed=np.array([1.0, np.nan] * 20))], [40])
    with pytest.raises(pq.PqgError):
        pq.generate([("r", pq.INT32, 0, -1)], [dict(fixed=np.arange(10, dtype=np.int32), is_null=np.ones(10, dtype=np.uint8))], [10])
