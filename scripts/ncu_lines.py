"""Per-source-line instruction and stall-sample shares of one kernel in an .ncu-rep.

ncu's CSV export of the source page carries metrics only for the SASS view; the line table comes
from nvdisasm over the cubin inside libpqg.so (same build as the capture, or the instruction
counts will not line up and the script says so).
usage: python scripts/ncu_lines.py <report.ncu-rep> <mangled-kernel-substring> <cubin-stem e.g. pqg_scan> [units] [launch index in the report]
`units` divides the executed-instruction counts (e.g. pages in the launch) to print per-unit costs."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern, stem = sys.argv[1], sys.argv[2], sys.argv[3]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
skip = sys.argv[5] if len(sys.argv) > 5 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "duckdb-parquet-parser_b200", "libpqg.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", stem, so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if ".section" in l and ".text" in l and kern in l][0]
cur, seq = None, []
for l in sass[start + 1:]:
    if ".section" in l and ".text" in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# one section per profiled launch: a "Kernel Name" row, a header row, then one row per instruction
heads = [i for i, r in enumerate(rows) if "Instructions Executed" in r]
pick = int(skip) if skip is not None else 0
h = heads[pick]
end = heads[pick + 1] - 1 if pick + 1 < len(heads) else len(rows)
hdr, data = rows[h], [r for r in rows[h + 1:end] if len(r) == len(rows[h])]
if len(data) != len(seq):
    sys.exit(f"instruction count mismatch: report {len(data)} vs cubin {len(seq)} (different build?)")
ie, iss = hdr.index("Instructions Executed"), hdr.index("# Samples")
ex, sm = collections.Counter(), collections.Counter()
for (c, _), r in zip(seq, data):
    ex[c] += int(r[ie])
    sm[c] += int(r[iss])
te, ts = sum(ex.values()), sum(sm.values())
print(f"total executed warp instructions {te} ({te / units:.1f} per unit), samples {ts}")
for k in sorted(ex, key=lambda k: (k is None, k)):
    if ex[k] > te * 0.002 or sm[k] > ts * 0.005:
        print(f"{str(k):38s} {ex[k] / units:9.1f}  {100 * ex[k] / te:5.2f}% exec  {100 * sm[k] / max(ts, 1):5.2f}% samples")
