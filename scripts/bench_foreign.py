"""Decode throughput on pyarrow-written files (uncompressed, data page v1): INT64 PLAIN and
dictionary columns with 64 KB and 1 MB pages -- the pages exceed the 8 KB tiles, so they run
through the general kernel (one warp per page).  usage: python scripts/bench_foreign.py [rows] [page_bytes] [column]
(page_bytes / column restrict the run to one page size / one of plain, dict, plain_nulls, dict_nulls: ncu captures)"""
import json
import os
import sys
import tempfile

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pqa

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pqb200 as pq

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
only_page = int(sys.argv[2]) if len(sys.argv) > 2 else 0
only_col = sys.argv[3] if len(sys.argv) > 3 else None
rng = np.random.default_rng(2)
nulls = rng.random(rows) < 0.25
t = pa.table({"plain": pa.array(rng.integers(-2**60, 2**60, size=rows), type=pa.int64()),
              "dict": pa.array(rng.integers(0, 4096, size=rows) * 977, type=pa.int64()),
              "plain_nulls": pa.array(rng.integers(-2**60, 2**60, size=rows), mask=nulls, type=pa.int64()),
              "dict_nulls": pa.array(rng.integers(0, 4096, size=rows) * 977, mask=nulls, type=pa.int64())})
out = []
for page in (8 * 1024, 64 * 1024, 1 << 20):
    if only_page and page != only_page:
        continue
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "f.parquet")
        pqa.write_table(t, p, compression="NONE", data_page_version="1.0", write_statistics=False, data_page_size=page,
                        row_group_size=rows // 4, use_dictionary=["dict", "dict_nulls"])
        img = np.fromfile(p, dtype=np.uint8)
    r = pq.Reader(data=img)
    ctx = pq.Context(0)
    buf = ctx.upload(img.ctypes.data, img.size)
    ctx.set_profiling(True)
    for c in range(4):
        if only_col and t.column_names[c] != only_col:
            continue
        plan = ctx.plan(buf, r.column_tables(c, -1))
        for _ in range(4):
            plan.run()
            plan.finish()
        tm = plan.timings_avg(3)
        vals = np.zeros(rows, dtype=np.int64)
        valid = np.zeros((rows + 31) // 32 + 1, dtype=np.uint32)
        plan.download(values=vals.ctypes.data, validity=valid.ctypes.data)
        ctx.sync()
        exp = t.column(c).fill_null(0).to_numpy()
        if c >= 2:
            v = ((valid[np.arange(rows) >> 5] >> (np.arange(rows) & 31).astype(np.uint32)) & 1).astype(bool)
            assert np.array_equal(v, ~nulls), (page, c)
            assert np.array_equal(vals[v], exp[v].astype(np.int64)) and not vals[~v].any(), (page, c)
        else:
            assert np.array_equal(vals, exp), (page, c)
        out.append({"page_bytes": page, "column": t.column_names[c], "pages": r.column_tables(c, -1)[3], "ms": tm["total_ms"],
                    "tiles_ms": tm["fixed_ms"], "general_ms": tm["general_ms"],
                    "in_plus_out_GBps": (plan.bytes_in + plan.bytes_out) / tm["total_ms"] / 1e6})
        plan.destroy()
    ctx.buf_free(buf)
    ctx.close()
    r.close()
print(json.dumps({"rows": rows, "results": out}, indent=1))
