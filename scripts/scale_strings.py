"""Scale check of the string paths through the reader API: ~3 GB PLAIN BYTE_ARRAY column
(80 M email-like values, 64 row groups): regex pruning, tuple-level chunk index, page-level chunk
index, decode -- sanity numbers + self-consistency (no oracle at this size).
usage: python scripts/scale_strings.py [rows]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pqb200 as pq

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 80_000_000
t0 = time.time()
col = bench.cfg4_email_column(rows, 3)
g = pq.generate([("email", 6, 0, 0)], [col], bench.rg_split(rows, 1_250_000))
img = g.to_numpy()
g.free()
print(f"file {img.size / 1e9:.2f} GB built in {time.time() - t0:.1f}s", flush=True)
r = pq.Reader(data=img)
print("pages", r.num_pages, "row groups", r.num_row_groups, "open/page scan s", round(r.page_scan_seconds, 3))
t0 = time.time()
bits, ms = r.regex_prune(0, bench.EMAIL_PATTERN)
print(f"regex: {len(bits)} pages, {int((bits == 0).sum())} prunable, kernel {ms:.2f} ms, wall {time.time() - t0:.2f} s (includes the upload)")
nbits, ms2 = r.regex_prune(0, bench.EMAIL_PATTERN, neg=True)
# every page is hit by the pattern or by its negation (pages are non-empty)
assert np.all((bits | nbits) == 1)
# noise rows come in blocks of 2000 every 10000: pages inside a noise block must be prunable for the positive pattern
frac = (bits == 0).mean()
assert 0.15 < frac < 0.25, frac
t0 = time.time()
t2c, nch = r.chunk_index("email", 4096)
print(f"chunk index: {nch} chunks, wall {time.time() - t0:.2f} s")
w = 2 + 33  # to_string(33) + 33 bytes per value
per_chunk = -(-4096 // w)
assert nch == -(-rows // per_chunk), (nch, rows, per_chunk)
assert t2c[0] == 0 and t2c[-1] == nch - 1 and np.all(np.diff(t2c.astype(np.int64)) >= 0)
assert np.array_equal(t2c[: per_chunk * 1000: per_chunk], np.arange(1000, dtype=np.uint64))
t0 = time.time()
pc, po, cf = r.page_chunk_index(0, 4096)
print(f"page chunk index: {len(cf)} chunks over {len(pc)} pages, wall {time.time() - t0:.2f} s")
assert pc[0] == 0 and pc[-1] == len(cf) - 1 and np.all(np.diff(pc.astype(np.int64)) >= 0) and np.all(po[cf] == 0)
print("scale check ok")
