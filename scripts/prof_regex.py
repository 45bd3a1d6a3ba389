"""The regex page-pruning scan on a device-resident PLAIN e-mail column (cfg4 shape), for ncu captures and quick numbers.
usage: python scripts/prof_regex.py [rows]"""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_scans as bs
import pqb200 as pq

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
col = pq.synth_strings(pq.PQGEN_EMAILS, rows, 99)
g = pq.generate([("email", 6, 0, 0)], [col], bs.rg_split(rows, 1_250_000))
img = g.to_numpy()
g.free()
r = pq.Reader(data=img)
ctx = pq.Context(0)
buf = ctx.upload(img.ctypes.data, img.size)
t = r.column_tables(0, -1)
plan = ctx.plan(buf, t)
L = pq.lib()
out = {}
for neg in (0, 1):
    dfa = pq.regex_compile(bs.EMAIL_PATTERN)
    bits = np.zeros((t[3] + 31) // 32 + 1, dtype=np.uint32)
    ms = ctypes.c_float(0)
    for _ in range(4):
        assert L.pqg_regex_scan(ctx.h, plan.h, dfa, neg, bits.ctypes.data, ctypes.byref(ms)) == 0, ctx.err()
    L.pqg_dfa_free(dfa)
    out["neg" if neg else "pos"] = {"ms": ms.value, "Gpages_per_s": t[3] / ms.value / 1e6, "payload_GBps": plan.bytes_in / ms.value / 1e6}
for ts in (0, 1):  # A/B: tile barrier off / on
    plan.set_option(2, ts)
    dfa = pq.regex_compile(bs.EMAIL_PATTERN)
    bits = np.zeros((t[3] + 31) // 32 + 1, dtype=np.uint32)
    ms = ctypes.c_float(0)
    best = 1e9
    for _ in range(6):
        assert L.pqg_regex_scan(ctx.h, plan.h, dfa, 0, bits.ctypes.data, ctypes.byref(ms)) == 0, ctx.err()
        best = min(best, ms.value)
    L.pqg_dfa_free(dfa)
    out["tile_barrier_%d_ms" % ts] = best
print(json.dumps({"rows": rows, "pages": t[3], "pattern": bs.EMAIL_PATTERN, **out}))
