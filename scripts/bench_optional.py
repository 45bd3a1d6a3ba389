"""OPTIONAL fixed-width columns on device-resident pages: INT64 PLAIN and dictionary (bw 8 / 16),
30 % nulls (definition levels as RLE runs, what the reference writer emits).
usage: python scripts/bench_optional.py [rows]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_scans as bench
import pqb200 as pq

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(11)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
isn = (rng.random(rows) < 0.3).astype(np.uint8)
cols = {"i64n_plain": rng.integers(-2**62, 2**62, size=rows, dtype=np.int64),
        "i64n_d8": rng.integers(0, 256, size=rows, dtype=np.int64) * 2654435761,
        "i64n_d16": rng.integers(0, 65536, size=rows, dtype=np.int64) * 2654435761}
specs = [(k, 2, 1, -1) for k in cols]
g = pq.generate(specs, [dict(fixed=v, is_null=isn) for v in cols.values()], bench.rg_split(rows, 5_000_000))
img = g.to_numpy()
g.free()
r = pq.Reader(data=img)
ctx = pq.Context(0)
buf = ctx.upload(img.ctypes.data, img.size)
ctx.set_profiling(True)
out = []
for c, name in enumerate(cols):
    plan = ctx.plan(buf, r.column_tables(c, -1))
    for _ in range(5):
        plan.run()
        plan.finish()
    tm = plan.timings_avg(4)
    # parity against the generator's input
    vals = np.zeros(rows, dtype=np.int64)
    valid = np.zeros((rows + 31) // 32 + 1, dtype=np.uint32)
    plan.download(values=vals.ctypes.data, validity=valid.ctypes.data)
    ctx.sync()
    v = ((valid[np.arange(rows) >> 5] >> (np.arange(rows) & 31).astype(np.uint32)) & 1).astype(bool)
    assert np.array_equal(v, isn == 0), name
    assert np.array_equal(vals[v], cols[name][v]) and not vals[~v].any(), name
    ms = tm["total_ms"]
    out.append({"column": name, "ms": ms, "tiles_ms": tm["fixed_ms"], "general_ms": tm["general_ms"], "dict_ms": tm["dict_ms"],
                "in_plus_out_GBps": (plan.bytes_in + plan.bytes_out) / ms / 1e6, "frac": (plan.bytes_in + plan.bytes_out) / ms / 1e6 / peak})
    plan.destroy()
print(json.dumps({"rows": rows, "results": out}, indent=1))
