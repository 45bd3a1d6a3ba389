"""Generates the committed golden fixtures (run in the container that has /root/reference):

    python tests/golden/make_golden.py

For every fixture: <name>.parquet is written by the REFERENCE's ParquetWriter
(src/writer/parquet_writer.cpp, through oracle/_ref) and <name>.npz holds what the REFERENCE's
reader returns for it: read_column_by_idx for every (row group, column) as value dumps
(oracle/valdump.h), ColumnReader::read_pages page tables, the page index, the
StringColumnIterator positions / lengths and the chunk-index prototype's output
(src/main.cpp:21-32 with chunk sizes 4096 and 256).  Small on purpose (a few hundred KB).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fixtures as fx  # noqa: E402
from oraclelib import (BOOLEAN, BYTE_ARRAY, DOUBLE, FLOAT, INT32, INT64, OPTIONAL, REQUIRED, UTF8, Oracle, Ref,  # noqa: E402
                       fixed_col, strings_to_col)


def golden_specs():
    g = {}
    g["mixed"] = (
        [("id", INT32, REQUIRED, -1), ("city", BYTE_ARRAY, OPTIONAL, UTF8), ("i64", INT64, REQUIRED, -1),
         ("f64n", DOUBLE, OPTIONAL, -1), ("d8", INT64, REQUIRED, -1), ("d12n", INT64, OPTIONAL, -1),
         ("dd", DOUBLE, REQUIRED, -1), ("f32", FLOAT, OPTIONAL, -1), ("email", BYTE_ARRAY, REQUIRED, UTF8),
         ("wild", BYTE_ARRAY, OPTIONAL, -1), ("b", BOOLEAN, OPTIONAL, -1), ("druns", INT64, OPTIONAL, -1)],
        lambda rng: [[fixed_col(np.arange(n, dtype=np.int32) + 1000 * k), fx.col_city(rng, n), fx.col_int64_plain(rng, n),
                      fx.col_double_plain(rng, n, 0.2), fx.col_int64_dict(rng, n, 200), fx.col_int64_dict(rng, n, 1500, 0.3),
                      fx.col_double_dict(rng, n, 50), fx.col_float_plain(rng, n, 0.5), fx.col_email(rng, n),
                      fx.col_str_varlen(rng, n, 0.1, 60), fx.col_bool(rng, n, 0.3),
                      fx.col_int64_dict(rng, n, 40, 0.25, runs=True)]
                     for k, n in enumerate((3000, 517))])
    return g


def dump_vals(prefix, v, out):
    out[prefix + "is_null"] = v.is_null
    out[prefix + "vidx"] = v.vidx
    out[prefix + "fixed"] = v.fixed
    out[prefix + "str_off"] = v.str_off
    out[prefix + "chars"] = v.chars


def main():
    ref, orc = Ref(), Oracle()
    for name, (specs, fn) in golden_specs().items():
        path = os.path.join(HERE, name + ".parquet")
        rng = np.random.default_rng(7)
        fx.write_ref_file(ref, path, specs, fn(rng))
        fx.check_eof_rule(orc, path)
        h = ref.open(path)
        out = {"page_index": ref.page_index(h), "num_rows": np.int64(ref.num_rows(h))}
        nrg, nc = ref.num_row_groups(h), ref.num_columns(h)
        out["shape"] = np.array([nrg, nc], dtype=np.int64)
        for c in range(nc):
            ci = ref.column_info(h, c)
            out[f"col{c}_info"] = np.array([ci["type"], ci["column_index"], ci["max_def_level"], ci["max_rep_level"],
                                            ci["repetition"], ci["converted"]], dtype=np.int64)
            out[f"col{c}_name"] = np.frombuffer(ci["name"].encode(), dtype=np.uint8)
            for rg in range(nrg):
                dump_vals(f"rg{rg}_col{c}_", ref.read_column_by_idx(h, rg, c), out)
                p = ref.read_pages(h, rg, c)
                out[f"rg{rg}_col{c}_pages"] = np.stack([p.page_num, p.page_type, p.num_values]).astype(np.int64)
            if ci["type"] == BYTE_ARRAY:
                pos, off, _ = ref.string_iterator(h, ci["name"])
                out[f"col{c}_iter_pos"] = pos
                out[f"col{c}_iter_off"] = off
                for cs in (4096, 256):
                    t2c, n = ref.chunk_index(h, ci["name"], cs)
                    out[f"col{c}_t2c_{cs}"] = t2c.astype(np.uint32)
                    out[f"col{c}_nchunks_{cs}"] = np.int64(n)
        ref.close(h)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, os.path.getsize(path), "bytes;", os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes npz")


if __name__ == "__main__":
    main()
