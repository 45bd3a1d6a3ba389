"""probe: regex scan timing on the bench's cfg4-shaped column (GPU)"""
import ctypes, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, pqb200 as pq
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
col = bench.cfg4_email_column(rows, 99)
g = pq.generate([("email", 6, 0, 0)], [col], bench.rg_split(rows, 1_250_000))
img = g.to_numpy()
r = pq.Reader(data=img)
ctx = pq.Context(0)
buf = ctx.upload(img.ctypes.data, img.size)
t = r.column_tables(0, -1)
plan = ctx.plan(buf, t)
L = pq.lib()
for pat in (bench.EMAIL_PATTERN, "zzz", "^user"):
    dfa = pq.regex_compile(pat)
    bits = np.zeros((t[3] + 31) // 32 + 1, dtype=np.uint32)
    ms = ctypes.c_float(0)
    for _ in range(3):
        L.pqg_regex_scan(ctx.h, plan.h, dfa, 0, bits.ctypes.data, ctypes.byref(ms))
    print(pat, "states", L.pqg_dfa_num_states(dfa), "ms", ms.value, "pages", t[3], "Mpages/s", t[3] / ms.value / 1e3)
# decode timing of the same column (string kernels)
ctx.set_profiling(True)
for _ in range(3):
    plan.run(); plan.finish()
print("decode", plan.timings(), "bytes_in", plan.bytes_in, "bytes_out", plan.bytes_out)
