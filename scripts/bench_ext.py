"""Extension path: pyarrow files with SNAPPY pages and / or DATA_PAGE_V2 framing through
Reader(extensions=True).read_columnar -- the rewrite kernel (k_xform: SNAPPY decode, V2 level framing) + the usual decode.
usage: python scripts/bench_ext.py [rows]"""
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
rng = np.random.default_rng(5)
nulls = rng.random(rows) < 0.25
t = pa.table({"plain": pa.array(rng.integers(0, 1 << 20, size=rows), type=pa.int64()),          # compressible: 3 of 8 bytes
              "dict_nulls": pa.array(rng.integers(0, 4096, size=rows) * 977, mask=nulls, type=pa.int64()),
              "str": pa.array([f"user{v:07d}@mail{v % 97}.example.com" for v in rng.integers(0, 1 << 22, size=rows)], type=pa.string())})
out = []
for label, kw in (("snappy_v1", dict(compression="SNAPPY", data_page_version="1.0")),
                  ("v2_none", dict(compression="NONE", data_page_version="2.0")),
                  ("v2_snappy", dict(compression="SNAPPY", data_page_version="2.0"))):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "f.parquet")
        pqa.write_table(t, p, write_statistics=False, data_page_size=64 * 1024, row_group_size=rows // 4,
                        use_dictionary=["dict_nulls"], **kw)
        img = np.fromfile(p, dtype=np.uint8)
    r = pq.Reader(data=img, extensions=True)
    for c in range(r.num_columns):
        best = None
        for _ in range(3):
            cc = r.read_columnar(c)
            if best is None or cc["kernel_ms"] < best["kernel_ms"]:
                best = cc
        out.append({"file": label, "column": r.column_info(c)["name"], "stored_bytes": int(sum(1 for _ in ())) or None,
                    "decoded_in_bytes": int(best["bytes_in"]), "out_bytes": int(best["bytes_out"]), "kernel_ms": best["kernel_ms"],
                    "in_plus_out_GBps": (best["bytes_in"] + best["bytes_out"]) / best["kernel_ms"] / 1e6})
    r.close()
print(json.dumps({"rows": rows, "note": "kernel_ms = the whole run on the device (rewrite + decode), first-run sizing included for strings; "
                  "bytes_in = UNCOMPRESSED page bytes", "results": out}, indent=1))
