"""String-path measurements on device-resident pages (not part of bench.py's headline line):
  cfg1 'city'   OPTIONAL BYTE_ARRAY, 8-entry dictionary, 30 % nulls       (BASELINE configs[0])
  cfg3 shape    OPTIONAL BYTE_ARRAY, 64 K-entry dictionary, 30 % nulls     (configs[2], scaled)
  cfg4 shape    PLAIN BYTE_ARRAY email-like                                (configs[3], scaled)
decode (size pass + copy pass), tuple-level 4 KB chunk index, page-level chunk index.
usage: python scripts/bench_strings.py [rows] > profiles/...json"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_scans as bench
import pqb200 as pq

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
only = sys.argv[2] if len(sys.argv) > 2 else ""  # cfg1 | cfg3 | cfg4: one workload only (ncu captures)
rg_rows = 2_500_000
rng = np.random.default_rng(7)
L = pq.lib()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]


def dict_strings(n, nkeys, fmt, width):
    keys = np.frombuffer(b"".join(fmt % i for i in range(nkeys)), dtype=np.uint8).reshape(nkeys, width)
    idx = rng.integers(0, nkeys, size=n)
    chars = keys[idx].reshape(-1)
    off = np.arange(n + 1, dtype=np.uint64) * width
    isn = (rng.random(n) < 0.3).astype(np.uint8)
    return dict(str_off=off, chars=chars, is_null=isn)


def run(name, col, rep):
    g = pq.generate([("s", 6, rep, 0)], [col], bench.rg_split(rows, rg_rows))
    img = g.to_numpy()
    g.free()
    r = pq.Reader(data=img)
    ctx = pq.Context(0)
    buf = ctx.upload(img.ctypes.data, img.size)
    t = r.column_tables(0, -1)
    plan = ctx.plan(buf, t)
    ctx.set_profiling(True)
    for _ in range(4):
        plan.run()
        plan.finish()
    tm = plan.timings_avg(3)
    n = plan.num_slots
    ids = np.zeros(n + 1, dtype=np.uint32)
    nch, cout, ms = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_float(0)
    cms = []
    for _ in range(3):
        rc = L.pqg_chunk_index(ctx.h, plan.h, 4096, 0, 0, ids.ctypes.data, ctypes.byref(nch), ctypes.byref(cout), ctypes.byref(ms))
        assert rc == 0, ctx.err()
        cms.append(ms.value)
    # regex page pruning on the same device-resident column
    rx = {}
    for pat in ("^[A-Za-z]+_?0*[0-9]+_x$", "@mail7[0-9]{2}\\."):
        dfa = pq.regex_compile(pat)
        bits = np.zeros((t[3] + 31) // 32 + 1, dtype=np.uint32)
        for _ in range(3):
            rc = L.pqg_regex_scan(ctx.h, plan.h, dfa, 0, bits.ctypes.data, ctypes.byref(ms))
            assert rc == 0, ctx.err()
        L.pqg_dfa_free(dfa)
        rx[pat] = {"ms": ms.value, "Mpages_per_s": t[3] / ms.value / 1e3, "pages_hit": int(sum(bin(int(x)).count("1") for x in bits))}
    t0 = time.perf_counter()
    pc, po, cf = r.page_chunk_index(0, 4096)
    page_s = time.perf_counter() - t0
    # dictionary-form decode (late materialisation): uint32 index per slot instead of chars + offsets
    dict_form = None
    try:
        import ctypes as C
        h = C.c_void_p()
        chunks, nc, pages, npg, _ = t
        if L.pqg_plan_create_dict_indices(ctx.h, buf, chunks, nc, pages, npg, C.byref(h)) == 0:
            dplan = pq.Plan(ctx, h)
            for _ in range(4):
                dplan.run()
                dplan.finish()
            dtm = dplan.timings_avg(3)
            dict_form = {"ms": dtm["total_ms"], "tiles_ms": dtm["fixed_ms"], "general_ms": dtm["general_ms"],
                         "bytes_out": dplan.bytes_out, "in_plus_out_GBps": (dplan.bytes_in + dplan.bytes_out) / dtm["total_ms"] / 1e6,
                         "decoded_page_GBps": dplan.bytes_in / dtm["total_ms"] / 1e6}
            dplan.destroy()
    except Exception as e:  # PLAIN columns: refused by design
        dict_form = {"error": str(e)}
    bi, bo = plan.bytes_in, plan.bytes_out
    dec_ms = tm["total_ms"]
    out = {"workload": name, "rows": rows, "pages": t[3], "bytes_in": bi, "bytes_out": bo,
           "decode_ms": dec_ms, "size_pass_ms": tm["str_size_ms"], "copy_pass_ms": tm["str_copy_ms"], "dict_prepare_ms": tm["dict_ms"],
           "decoded_page_GBps": bi / dec_ms / 1e6, "in_plus_out_GBps": (bi + bo) / dec_ms / 1e6, "frac_of_hbm_peak": (bi + bo) / dec_ms / 1e6 / peak,
           "chunk_index_ms": min(cms), "chunks": int(nch.value), "page_chunk_index_wall_ms": page_s * 1e3, "page_chunks": int(len(cf)), "regex": rx, "dictionary_form": dict_form}
    plan.destroy()
    ctx.buf_free(buf)
    ctx.close()
    r.close()
    return out


def cfg1_city():
    names = [b"Berlin", b"Dublin", b"Paris", b"Rome", b"Vienna", b"Amsterdam", b"Lisbon", b"Copenhagen"]
    idx = rng.integers(0, 8, size=rows)
    lens = np.array([len(x) for x in names], dtype=np.uint64)[idx]
    off = np.zeros(rows + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    chars = np.frombuffer(b"".join(names[i] for i in idx[:0]), dtype=np.uint8)
    table = np.zeros((8, 10), dtype=np.uint8)
    for i, x in enumerate(names):
        table[i, :len(x)] = np.frombuffer(x, dtype=np.uint8)
    mask = np.arange(10)[None, :] < np.array([len(x) for x in names])[idx][:, None]
    chars = table[idx][mask]
    return dict(str_off=off, chars=chars, is_null=(rng.random(rows) < 0.3).astype(np.uint8))


work = {"cfg1": lambda: run("cfg1 city: 8 city names (variable length), 30% nulls", cfg1_city(), 1),
        "cfg3": lambda: run("cfg3 shape: 64K-entry dictionary 'city_%06u_x', 30% nulls", pq.synth_strings(pq.PQGEN_CITY64K, rows, 333, null_permille=300), 1),
        "cfg4": lambda: run("cfg4 shape: PLAIN email-like strings", pq.synth_strings(pq.PQGEN_EMAILS, rows, 99), 0)}
res = [f() for k, f in work.items() if not only or k == only]
print(json.dumps({"hbm_peak_GBps": peak, "results": res}, indent=1))
