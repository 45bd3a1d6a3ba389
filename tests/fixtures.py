"""Seeded synthetic column shapes + files written by the REFERENCE's ParquetWriter (through
oracle/_ref) for parity tests.  TEST INFRASTRUCTURE.

Fixture rule: the reference reader fetches every page header with
a fixed 256-byte read and never clears the stream state, so the last page header of the
file must start >= 256 bytes before EOF.  `write_ref_file` appends a PLAIN INT64 padding
column and asserts the rule on the written file.
"""
import os

import numpy as np

from oraclelib import (BOOLEAN, BYTE_ARRAY, DOUBLE, FLOAT, INT32, INT64, OPTIONAL, REQUIRED, UTF8,
                       fixed_col, strings_to_col)

CITIES = [b"Amsterdam", b"Berlin", b"Copenhagen", b"Dublin", b"Edinburgh", b"Florence", b"Geneva", b"Helsinki"]


def nulls(rng, n, frac):
    if frac <= 0:
        return None
    return (rng.random(n) < frac).astype(np.uint8)


def col_int32_plain(rng, n, null_frac=0.0):
    return fixed_col(rng.permutation(n).astype(np.int32) * 7 - 3, nulls(rng, n, null_frac))


def col_int64_plain(rng, n, null_frac=0.0):
    return fixed_col(rng.integers(-2**62, 2**62, size=n, dtype=np.int64), nulls(rng, n, null_frac))


def col_double_plain(rng, n, null_frac=0.0):
    a = rng.random(n)
    if n > 8:
        a[3] = np.nan
        a[5] = -0.0
        a[7] = np.inf
    return fixed_col(a, nulls(rng, n, null_frac))


def col_float_plain(rng, n, null_frac=0.0):
    return fixed_col(rng.random(n).astype(np.float32), nulls(rng, n, null_frac))


def col_int64_dict(rng, n, nkeys, null_frac=0.0, runs=False):
    k = rng.integers(0, nkeys, size=n, dtype=np.int64)
    if runs:  # long equal runs -> RLE runs inside the index stream
        k = np.repeat(k[: max(1, n // 9)], 9)[:n]
        if len(k) < n:
            k = np.concatenate([k, np.zeros(n - len(k), dtype=np.int64)])
    return fixed_col(k * 2654435761, nulls(rng, n, null_frac))


def col_double_dict(rng, n, nkeys, null_frac=0.0):
    k = rng.integers(0, nkeys, size=n, dtype=np.int64)
    return fixed_col(k.astype(np.float64) * 0.37, nulls(rng, n, null_frac))


def col_int32_dict(rng, n, nkeys, null_frac=0.0):
    k = rng.integers(0, nkeys, size=n, dtype=np.int64)
    return fixed_col((k * 977 - 5).astype(np.int32), nulls(rng, n, null_frac))


def col_float_dict(rng, n, nkeys, null_frac=0.0):
    k = rng.integers(0, nkeys, size=n, dtype=np.int64)
    return fixed_col((k * 0.25).astype(np.float32), nulls(rng, n, null_frac))


def col_bool(rng, n, null_frac=0.0):
    return fixed_col(rng.integers(0, 2, size=n, dtype=np.int64).astype(np.uint64), nulls(rng, n, null_frac))


def col_city(rng, n, null_frac=0.3):
    idx = rng.integers(0, len(CITIES), size=n)
    return strings_to_col([CITIES[i] for i in idx], nulls(rng, n, null_frac))


def col_str_dict(rng, n, nkeys, null_frac=0.3):
    idx = rng.integers(0, nkeys, size=n)
    return strings_to_col([b"city_%06d_x" % i for i in idx], nulls(rng, n, null_frac))


SHORTS = [b"", b"a", b"bb", b"\x00\x00\x00", b"ccc", b"abcdabcdabcd", b"\x03", b"zz\x00"]


def col_str_short_dict(rng, n, null_frac=0.2):
    """dictionary of empty / 1-3 byte / NUL-bearing strings: defeats the length-prefix
    speculation (false candidates) -> the sequential fallbacks must produce the same result"""
    idx = rng.integers(0, len(SHORTS), size=n)
    return strings_to_col([SHORTS[i] for i in idx], nulls(rng, n, null_frac))


def col_email(rng, n, null_frac=0.0, noise_frac=0.2):
    out = []
    u = rng.integers(0, 10**9, size=n)
    m = rng.integers(0, 1000, size=n)
    z = rng.random(n)
    for i in range(n):
        if z[i] < noise_frac / 2:
            out.append(b"user%d.mail%d.example.com" % (u[i], m[i]))  # no '@'
        elif z[i] < noise_frac:
            out.append(b"user%d@mail%d.example.com!!" % (u[i], m[i]))  # trailing junk
        else:
            out.append(b"user%d@mail%d.example.com" % (u[i], m[i]))
    return strings_to_col(out, nulls(rng, n, null_frac))


def col_str_varlen(rng, n, null_frac=0.1, maxlen=300):
    """PLAIN strings with wild lengths: empty strings, > 255 bytes, NUL bytes, non-ASCII."""
    out = []
    lens = rng.integers(0, maxlen, size=n)
    for i in range(n):
        L = int(lens[i]) if i % 7 else 0
        b = rng.integers(0, 256, size=L, dtype=np.uint8).tobytes() + (b"#%d" % i)
        out.append(b if i % 7 else b"")
    return strings_to_col(out, nulls(rng, n, null_frac))


def col_str_tiny(rng, n, null_frac=0.0):
    """PLAIN strings of 0..6 bytes over {NUL, 0x01, 'a', 'b'}: runs of empty strings, strings that
    start with NUL, prefixes that look like each other's shadows -- every heuristic of the parallel
    length-prefix discovery must fail safely into the sequential walk.  Keep n small: the writer
    stays PLAIN only while distinct > non_null / 5."""
    alphabet = np.frombuffer(b"\x00\x01ab", dtype=np.uint8)
    lens = rng.integers(0, 7, size=n)
    lens[rng.random(n) < 0.25] = 0
    out = [alphabet[rng.integers(0, 4, size=int(L))].tobytes() for L in lens]
    return strings_to_col(out, nulls(rng, n, null_frac))


def pad_col(n):
    return fixed_col(np.arange(n, dtype=np.int64) * 1000003 + 17)


def write_ref_file(ref, path, specs, row_groups, pad=True):
    """specs: [(name, type, repetition, converted)]; row_groups: [[col, ...], ...]."""
    specs = list(specs)
    rgs = [list(rg) for rg in row_groups]
    if pad:
        specs.append(("zz_pad", INT64, REQUIRED, -1))
        for rg in rgs:
            n = len(rg[0]["fixed"]) if "fixed" in rg[0] else len(rg[0]["str_off"]) - 1
            rg.append(pad_col(max(n, 0)))
    ref.write_file(path, specs, rgs)
    return path


def check_eof_rule(oracle, path):
    h = oracle.open(path)
    try:
        idx = oracle.page_index(h)
    finally:
        oracle.close(h)
    size = os.path.getsize(path)
    # header start is unknown from the index; payload start - 64 is a safe lower bound
    if len(idx):
        last = int(idx[:, 0].max())
        assert last - 64 + 256 <= size, f"{path}: last page too close to EOF for the reference reader"


# name -> (specs, builder(rng) -> row_groups).  Row counts are chosen so every page shape of
# SURVEY.md section 8 (a) shows up: full pages, ragged last pages, several row groups.
def standard_files():
    files = {}

    def add(name, specs, fn):
        files[name] = (specs, fn)

    add("cfg1_small",
        [("id", INT32, REQUIRED, -1), ("city", BYTE_ARRAY, OPTIONAL, UTF8)],
        lambda rng: [[fixed_col(np.arange(20000, dtype=np.int32)), col_city(rng, 20000)]])
    add("fixed_plain",
        [("i32", INT32, REQUIRED, -1), ("i64", INT64, REQUIRED, -1), ("f32", FLOAT, REQUIRED, -1),
         ("f64", DOUBLE, REQUIRED, -1), ("i32n", INT32, OPTIONAL, -1), ("i64n", INT64, OPTIONAL, -1),
         ("f64n", DOUBLE, OPTIONAL, -1), ("f32n", FLOAT, OPTIONAL, -1)],
        lambda rng: [[col_int32_plain(rng, n), col_int64_plain(rng, n), col_float_plain(rng, n),
                      col_double_plain(rng, n), col_int32_plain(rng, n, 0.3), col_int64_plain(rng, n, 0.5),
                      col_double_plain(rng, n, 0.05), col_float_plain(rng, n, 0.9)] for n in (5000, 777)])
    add("fixed_dict",
        [("d8", INT64, REQUIRED, -1), ("d12", INT64, REQUIRED, -1), ("d16", INT64, REQUIRED, -1),
         ("dd", DOUBLE, REQUIRED, -1), ("d3n", INT64, OPTIONAL, -1), ("druns", INT64, OPTIONAL, -1),
         ("di32", INT32, OPTIONAL, -1), ("df32", FLOAT, REQUIRED, -1), ("d1", INT64, REQUIRED, -1)],
        lambda rng: [[col_int64_dict(rng, n, 256), col_int64_dict(rng, n, 4096), col_int64_dict(rng, n, 40000),
                      col_double_dict(rng, n, 1000), col_int64_dict(rng, n, 7, 0.3),
                      col_int64_dict(rng, n, 300, 0.2, runs=True), col_int32_dict(rng, n, 100, 0.4),
                      col_float_dict(rng, n, 33), col_int64_dict(rng, n, 2)] for n in (250000, 12345)])
    add("strings",
        [("city", BYTE_ARRAY, OPTIONAL, UTF8), ("s64k", BYTE_ARRAY, OPTIONAL, UTF8),
         ("email", BYTE_ARRAY, REQUIRED, UTF8), ("emailn", BYTE_ARRAY, OPTIONAL, UTF8),
         ("wild", BYTE_ARRAY, OPTIONAL, -1), ("allnull", BYTE_ARRAY, OPTIONAL, UTF8), ("shorts", BYTE_ARRAY, OPTIONAL, -1)],
        lambda rng: [[col_city(rng, n), col_str_dict(rng, n, 3000), col_email(rng, n), col_email(rng, n, 0.3),
                      col_str_varlen(rng, n), strings_to_col([b""] * n, np.ones(n, dtype=np.uint8)), col_str_short_dict(rng, n)]
                     for n in (30000, 4321)])
    add("tiny_plain",
        [("tiny", BYTE_ARRAY, REQUIRED, -1), ("tinyn", BYTE_ARRAY, OPTIONAL, -1)],
        lambda rng: [[col_str_tiny(rng, n), col_str_tiny(rng, n, 0.3)] for n in (1500, 700)])
    add("bools",
        [("b", BOOLEAN, REQUIRED, -1), ("bn", BOOLEAN, OPTIONAL, -1)],
        lambda rng: [[col_bool(rng, n), col_bool(rng, n, 0.3)] for n in (5000, 100)])
    return files


def make_file(ref, oracle, name, path, seed=42):
    specs, fn = standard_files()[name]
    rng = np.random.default_rng(seed)
    write_ref_file(ref, path, specs, fn(rng))
    check_eof_rule(oracle, path)
    return path
