#!/usr/bin/env python
"""bench.py -- decoded input-page GB/s of the B200 Parquet page decoder (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU ColumnReader

Headline workload (BASELINE.json configs[1], SURVEY.md section 8 d "Config 2"): a 100 M-row file as
10 row groups x 10 M rows with 7 columns -- INT64 PLAIN, DOUBLE PLAIN, INT64 dictionary with
2^8 / 2^12 / 2^16 / 2^20 distinct keys (index bit widths 8 / 12 / 16 / 20) and DOUBLE
dictionary with 2^16 keys -- produced by the workload generator, which is byte-identical to the
reference's ParquetWriter (tests/test_gen_cpu.py).  One step = one decode of all 7 columns
through ONE plan (pqr_columns_tables; PQG_BENCH_PER_COLUMN=1 runs one plan per column instead).

  value      Sigma page payload bytes (data + dictionary pages) / device time, image resident in
             HBM, K steps timed with CUDA events on the decoder's stream, max over ranks.  Before the
             timed steps the plan's device output is compared with the generator's input columns on
             the device, all rows of every column (roofline.parity).
  e2e        same metric through the reference-facing reader API (pqr_* C-ABI) with HOST buffers:
             every step uploads the column chunks from pinned host memory and reads the
             decoded columns back to the host.  e2e.platform_ceiling: the same bytes as bare copies
             (what the box allows); e2e.cold: a fresh reader + its first read.
  roofline   dominant kernel (k_fixed_tiles): (bytes_in + bytes_out) of its launches / their
             CUDA-event durations inside the timed steps, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference
             the UNMODIFIED reference (oracle/_ref/libpqref.so, compiled from /root/reference
             sources) reading a bounded sample (the first row groups of the same data, written
             as its own file) with one ParquetReader per host thread.
  strings / regex / chunk_index  (bench_scans.py; --no-scans skips them)
             BASELINE.json configs[2..4] next to the headline, at every N, each with its own roofline,
             cpu_baseline (N = 1) and parity statement: the cfg3 dictionary-string shape (100 M rows per
             GPU), the regex page-pruning scan over ONE ~10 GB file whose row groups are split over the
             ranks (bitmaps gathered on the host), and the 4 KB chunk index over a mixed PLAIN /
             dictionary column of 6.4 GB per GPU with the carry chain across the ranks.

N > 1 (torchrun): every rank decodes its own 100 M-row shard (row groups are independent; no
data-path collective), "scaling": "weak"; value = total bytes / max-over-ranks time.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "decoded_page_GBps"
UNIT = "GB/s"
WORKLOAD = "cfg2: 100M-row INT64/DOUBLE PLAIN + dictionary bw 8/12/16/20, 10 row groups x 10M rows"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line: the JSON result.  Libraries (NCCL prints its version banner
# on stdout, ...) are redirected to stderr by pointing fd 1 at fd 2 for the whole run.
_REAL_STDOUT = None


def claim_stdout():
    """called first thing in main() (not at import: scripts/ reuse this module's workloads)"""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ── workload ────────────────────────────────────────────────────────────────────────────
def cfg2_specs():
    I64, F64, REQ = 2, 5, 0
    return [("i64_plain", I64, REQ, -1), ("f64_plain", F64, REQ, -1), ("i64_d8", I64, REQ, -1),
            ("i64_d12", I64, REQ, -1), ("i64_d16", I64, REQ, -1), ("i64_d20", I64, REQ, -1),
            ("f64_d16", F64, REQ, -1)]


def cfg2_columns(rows, seed):
    """SURVEY.md 8(d) Config 2: uniform 64-bit ints, uniform [0,1) doubles, dictionary keys
    k * 2654435761 (int64) / k * 0.37 (double) with k uniform in [0, 2^b)."""
    rng = np.random.default_rng(seed)
    cols = [dict(fixed=rng.integers(-2**63, 2**63 - 1, size=rows, dtype=np.int64)),
            dict(fixed=rng.random(rows))]
    for b in (8, 12, 16, 20):
        k = rng.integers(0, 1 << b, size=rows, dtype=np.int64)
        cols.append(dict(fixed=k * 2654435761))
    k = rng.integers(0, 1 << 16, size=rows, dtype=np.int64)
    cols.append(dict(fixed=k.astype(np.float64) * 0.37))
    return cols


def scratch_dir(need_bytes):
    """RAM-backed scratch for the CPU arm's sample file when it has room (page cache warm either way)"""
    import shutil
    import tempfile
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            if os.path.isdir(d) and shutil.disk_usage(d).free > need_bytes:
                return d
        except OSError:
            pass
    return tempfile.gettempdir()


def rg_split(rows, rg_rows):
    out = [rg_rows] * (rows // rg_rows)
    if rows % rg_rows:
        out.append(rows % rg_rows)
    return out


def build_file(pq, rows, rg_rows, seed):
    t0 = time.time()
    cols = cfg2_columns(rows, seed)
    t1 = time.time()
    g = pq.generate(cfg2_specs(), cols, rg_split(rows, rg_rows))
    t2 = time.time()
    log(f"[bench] synthetic columns {t1 - t0:.1f}s, writer-identical encode {t2 - t1:.1f}s, file {g.size / 1e9:.3f} GB")
    return g, cols


# ── clocks ──────────────────────────────────────────────────────────────────────────────
class ClockSampler:
    """SM clock + throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.thread = [], set(), False, None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report that, never fake numbers
            self.nv = None
            self.err = str(e)

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join()
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": "nvml unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ── CPU arm: the unmodified reference ───────────────────────────────────────────────────
def ref_sample_file(pq, cols, rows, rg_rows, n_rgs, path):
    """first n_rgs row groups of the same data, as their own (writer-identical) file"""
    take = min(rows, rg_rows * n_rgs)
    sub = [dict(fixed=c["fixed"][:take]) for c in cols]
    g = pq.generate(cfg2_specs(), sub, rg_split(take, rg_rows))
    g.write(path)
    size = g.size
    g.free()
    return take, size


def payload_bytes_of(pq, path):
    r = pq.Reader(path)
    idx = r.page_index()
    total = int(idx[:, 1].sum())
    # dictionary pages are not in the page index: add them from the descriptor tables
    for c in range(r.num_columns):
        chunks, nc, _, _, _ = r.column_tables(c, -1)
        total += sum(int(chunks[i].dict_size) for i in range(nc) if chunks[i].has_dict)
    r.close()
    return total


def cpu_reference_run(pq, cols, rows, rg_rows, sample_rgs, steps, warmup, threads):
    import oraclelib
    if not oraclelib.Ref.available():
        raise RuntimeError("oracle/_ref/libpqref.so is missing (built by __graft_entry__.build() where /root/reference exists)")
    ref = oraclelib.Ref()
    path = os.path.join(scratch_dir(3 << 30), f"pqg_bench_ref_{os.getpid()}.parquet")
    take, _ = ref_sample_file(pq, cols, rows, rg_rows, sample_rgs, path)
    try:
        payload = payload_bytes_of(pq, path)
        n_rg = len(rg_split(take, rg_rows))
        items = [(rg, c) for rg in range(n_rg) for c in range(len(cols))]
        rgs = [i[0] for i in items]
        cs = [i[1] for i in items]
        times = []
        for s in range(warmup + steps):
            t, nv = ref.time_read_chunks(path, rgs, cs, threads)
            assert nv == take * len(cols), (nv, take)
            if s >= warmup:
                times.append(t)
    finally:
        os.unlink(path)
    mean = sum(times) / len(times)
    return dict(value=payload / mean / 1e9, seconds=mean, payload=payload, rows=take, values=take * len(cols),
                threads=threads, sample=f"first {n_rg} row group(s) = {take} rows x {len(cols)} columns of the same data "
                                        f"({payload / 1e6:.0f} MB payload), ParquetReader::read_column_by_idx per chunk, "
                                        f"one reader per thread, page cache warm")


class _DevView:
    """a raw device pointer as a CUDA-array-interface object (torch.as_tensor wraps it without a copy)"""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_parity_check(plans, cols, rows, fused):
    import torch
    checked = 0
    for k, col in enumerate(cols):
        exp = torch.from_numpy(np.ascontiguousarray(col["fixed"]).view(np.int64)).cuda()
        p = plans[0] if fused else plans[k]
        base = p.values_ptr + (k * rows * 8 if fused else 0)
        got = torch.as_tensor(_DevView(base, rows), device="cuda")
        if not torch.equal(got, exp):
            bad = int((got != exp).nonzero()[0].item())
            raise AssertionError(f"bench parity (timed arm, device): column {k} differs first at row {bad}")
        checked += rows
        del exp, got
    torch.cuda.empty_cache()
    return {"checked_values": checked, "how": "device output of the timed plan == the generator's input columns (torch.equal on the device, "
            "all rows, every column), before the timed steps; page errors checked after them"}


# ── GPU arm ─────────────────────────────────────────────────────────────────────────────
def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--rg-rows", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample-rgs", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scans", action="store_true", help="skip the strings / regex / chunk-index measurements (BASELINE configs[2..4])")
    ap.add_argument("--strings-rows", type=int, default=100_000_000, help="cfg3 shape: rows per GPU (0 = skip)")
    ap.add_argument("--regex-rows", type=int, default=270_000_000, help="cfg4: values of the ONE ~10 GB file (37 bytes of payload each; 0 = skip)")
    ap.add_argument("--chunk-rgs", type=int, default=64, help="cfg5 shape: alternating PLAIN / dictionary row groups of 5 M rows per GPU (64 = 6.4 GB; 0 = skip)")
    ap.add_argument("--chunk-rg-rows", type=int, default=5_000_000)
    ap.add_argument("--scan-steps", type=int, default=5)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    config = {"workload": WORKLOAD, "rows_per_gpu": a.rows, "row_group_rows": a.rg_rows, "columns": [s[0] for s in cfg2_specs()],
              "l2": "inputs (2.6 GB) and outputs (5.6 GB) per step exceed the 126 MB L2; no flush needed",
              "sharding": "one process per GPU, each decodes its own row-group shard; no collective on the data path"}
    import pqb200 as pq
    pq.lib()
    cores = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return 0
        # the sample only needs the first row groups: do not generate the whole file
        take = min(a.rows, a.rg_rows * a.cpu_sample_rgs)
        cols = cfg2_columns(take, 1234)
        r = cpu_reference_run(pq, cols, take, a.rg_rows, a.cpu_sample_rgs, a.steps, a.warmup, cores)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64/f64 bit moves", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; this framework has no CPU fallback")
    torch.cuda.set_device(local)
    affinity = None
    if world > 1 and hasattr(os, "sched_setaffinity"):
        # every rank keeps to its own cores (page-header scan, staging copies): the ranks share one host
        try:
            cores_all = sorted(os.sched_getaffinity(0))
            mine = cores_all[local::world] if len(cores_all) >= world else cores_all
            os.sched_setaffinity(0, mine)
            affinity = {"cores": len(mine), "of": len(cores_all)}
        except OSError:
            affinity = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gen, cols = build_file(pq, a.rows, a.rg_rows, 1234 + rank)
    size = gen.size
    # pinned host image of the file (what an application would have read / mapped)
    host = torch.empty(size + 64, dtype=torch.uint8, pin_memory=True)
    gen.emit(host.data_ptr(), size)
    gen.free()

    # host side of the boundary: footer + page headers -> flat descriptor tables (timed apart)
    t0 = time.time()
    reader = pq.Reader.from_pointer(host.data_ptr(), size, device=local)
    open_s = time.time() - t0
    ncols = reader.num_columns
    n_pages = reader.num_pages
    tables = [reader.column_tables(c, -1) for c in range(ncols)]

    # ---- kernel arm: image resident in HBM --------------------------------------------
    stream = torch.cuda.Stream()
    ctx = pq.Context(local, stream.cuda_stream)
    t0 = time.time()
    dev_img = torch.empty(size + 64, dtype=torch.uint8, device="cuda")
    h2d0, h2d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        h2d0.record()
        dev_img[:size].copy_(host[:size], non_blocking=True)
        dev_img[size:].zero_()
        h2d1.record()
    stream.synchronize()
    h2d_ms = h2d0.elapsed_time(h2d1)
    image = ctx.wrap_device(dev_img.data_ptr(), size)
    # One plan over all columns (they share the 8-byte value width): one dictionary-preparation
    # launch, tile launches grouped by dictionary footprint, one sweep of the general kernel.
    # PQG_BENCH_PER_COLUMN=1 times one plan per column instead (the layout of earlier rounds).
    fused = os.environ.get("PQG_BENCH_PER_COLUMN", "0") != "1"
    col_plans = [ctx.plan(image, t) for t in tables]
    plans = [ctx.plan(image, reader.columns_tables(list(range(ncols)), -1))] if fused else col_plans
    bytes_in = sum(p.bytes_in for p in plans)
    ctx.set_profiling(True)

    def step():
        for p in plans:
            p.run()

    # per-column detail for the roofline object (outside the timed region)
    for _ in range(4):
        for p in col_plans:
            p.run()
    for p in col_plans:
        p.finish()
    col_tm = [p.timings_avg(3) for p in col_plans]
    col_bytes = [(p.bytes_in, p.bytes_out) for p in col_plans]
    # A/B: the partitioned-dictionary mode (dictionaries of 32 KB .. 512 KB split over sibling CTAs' shared memories; opt-in,
    # measured slower) against the default L2 gather, on the columns it applies to
    col_l2 = {}
    for i, (spec, p) in enumerate(zip(cfg2_specs(), col_plans)):
        if spec[0].endswith("_d16") and os.environ.get("PQG_BENCH_NO_AB", "0") != "1":  # (profiling runs skip the A/B: fewer launches to step over)
            p.set_option(pq.PQG_OPT_PARTITIONED_DICT, 1)
            for _ in range(4):
                p.run()
            p.finish()
            col_l2[i] = p.timings_avg(3)["fixed_ms"]
            p.set_option(pq.PQG_OPT_PARTITIONED_DICT, 0)
    if fused:
        for p in col_plans:
            p.destroy()

    for _ in range(warmup):
        step()
    for p in plans:
        p.finish()
    bytes_out = sum(p.bytes_out for p in plans)
    # parity of the TIMED arm at full size, before timing: the plan's device output against the generator's
    # input columns (the generator is byte-identical to the reference's writer, tests/test_gen_cpu.py, so these
    # ARE the values the reference's reader returns for this file), compared on the device, column by column
    value_parity = device_parity_check(plans, cols, a.rows, fused)
    launches0 = ctx.launches
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    with torch.cuda.stream(stream):
        ev0.record()
        for _ in range(a.steps):
            step()
        ev1.record()
    stream.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launches - launches0
    for p in plans:
        p.finish()  # raises on any page error
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / a.steps
    value = bytes_in * world / (ms_per_step * 1e-3) / 1e9

    # dominant kernel: k_fixed_tiles, one launch per column; CUDA events around every launch
    # of the timed steps (the last <= 8 runs of each plan are kept)
    tm = [p.timings_avg(min(a.steps, 8)) for p in plans]
    k_ms = sum(t["fixed_ms"] for t in tm)
    tile_launches = sum(t["tile_launches"] for t in tm)
    k_bytes = bytes_in + bytes_out
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_fixed_tiles<8> (TMA-staged page tiles)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_step": k_bytes, "kernel_ms_per_step": k_ms, "launches_per_step": tile_launches,
                "plan": "one plan over all columns" if fused else "one plan per column", "parity": value_parity,
                "kernel_share_of_step": k_ms / ms_per_step if ms_per_step else None,
                "general_kernel_ms_per_step": sum(t["general_ms"] for t in tm), "dict_prepare_ms_per_step": sum(t["dict_ms"] for t in tm),
                "per_column": [dict({"column": s[0], "ms": t["fixed_ms"], "dict_ms": t["dict_ms"], "general_ms": t["general_ms"],
                                     "GBps_in_plus_out": (b[0] + b[1]) / (t["fixed_ms"] * 1e-3) / 1e9 if t["fixed_ms"] > 0 else None},
                                    **({"ms_with_partitioned_smem_dictionary_(opt-in,_rejected)": col_l2[i]} if i in col_l2 else {}))
                               for i, (s, t, b) in enumerate(zip(cfg2_specs(), col_tm, col_bytes))],
                "per_column_note": "one plan per column, timed apart from the headline step"}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        roofline["traffic"] = tr.get("dram_bytes_per_step")
        roofline["traffic_source"] = tr.get("source")
    except Exception:
        pass

    for p in plans:
        p.destroy()
    ctx.buf_free(image)
    del dev_img
    torch.cuda.empty_cache()

    # ---- e2e arm: host buffers in, host columns out, through the reader C-ABI ---------
    # pqr_read_columns_into: per row group H2D -> decode -> D2H on three streams; the decoded
    # columns land in caller-owned pinned host buffers.  Every step moves the file bytes up and
    # the decoded columns down again; only descriptor tables / device buffers are kept.
    e2e_steps = max(1, min(a.e2e_steps, a.steps))
    outs = [torch.empty(a.rows, dtype=torch.int64, pin_memory=True) for _ in range(ncols)]
    dsts = [(o.data_ptr(), o.numel() * 8, None, 0) for o in outs]
    stats = None

    def e2e_step():
        nonlocal stats
        stats = reader.read_columns_into(list(range(ncols)), dsts, -1)

    for _ in range(2):
        e2e_step()  # warm-up (builds and caches the plans)
    for c in range(ncols):  # bit-exact check against the generator's input columns
        got = outs[c].numpy().view(np.uint64)
        exp = np.ascontiguousarray(cols[c]["fixed"]).view(np.uint64)
        if not np.array_equal(got, exp):
            raise AssertionError(f"bench parity check failed on column {c}")
        outs[c].zero_()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    for c in range(ncols):
        if not np.array_equal(outs[c].numpy().view(np.uint64)[-4096:], np.ascontiguousarray(cols[c]["fixed"]).view(np.uint64)[-4096:]):
            raise AssertionError(f"bench parity check (timed step) failed on column {c}")
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d_bytes = sum(s["h2d_bytes"] for s in stats)
    d2h_bytes = sum(s["d2h_bytes"] for s in stats)
    e2e = {"value": bytes_in * world / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "api": "pqr_open_memory + pqr_read_columns_into (pinned host file image in, pinned host columnar buffers out; "
                  "H2D/decode/D2H pipelined per row group)",
           "parity": "decoded columns bit-identical to the generator's input arrays (full check after warm-up, tail check after the timed steps)",
           "pcie_GBps": {"h2d": h2d_bytes / e2e_s / 1e9, "d2h": d2h_bytes / e2e_s / 1e9}}
    # ---- what the platform gives: the same bytes as plain copies, nothing else (pinned host <-> device, both directions at
    #      once, every rank at the same time) -- the ceiling of ANY host-to-host path on this box
    probe = None
    try:
        d_in = torch.empty(int(h2d_bytes), dtype=torch.uint8, device="cuda")
        d_out = torch.empty(int(d2h_bytes), dtype=torch.uint8, device="cuda")
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        flat_out = [o.view(torch.uint8) for o in outs]
        times = []
        for it in range(3):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_in):
                d_in.copy_(host[:int(h2d_bytes)], non_blocking=True)
            with torch.cuda.stream(s_out):
                off = 0
                for fo in flat_out:
                    fo.copy_(d_out[off:off + fo.numel()], non_blocking=True)
                    off += fo.numel()
            s_in.synchronize()
            s_out.synchronize()
            times.append(time.perf_counter() - t0)
        probe_s = min(times[1:])
        if world > 1:
            t = torch.tensor([probe_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            probe_s = float(t.item())
        probe = {"value": bytes_in * world / probe_s / 1e9, "unit": UNIT, "ms": probe_s * 1e3,
                 "h2d_GBps": h2d_bytes * world / probe_s / 1e9, "d2h_GBps": d2h_bytes * world / probe_s / 1e9,
                 "what": "the step's H2D and D2H bytes as bare cudaMemcpyAsync from / to pinned memory on two streams, all ranks at once, no kernels: "
                         "the same metric if decoding were free"}
        del d_in, d_out
    except Exception as ex:  # never lose the headline to the probe
        probe = {"error": str(ex)}
    e2e["platform_ceiling"] = probe
    if probe and probe.get("value"):
        e2e["frac_of_platform_ceiling"] = e2e["value"] / probe["value"]
    # ---- cold: a fresh reader (footer + page-header scan on the host threads, descriptor tables, plan and device buffers
    #      built) and its first read -- what the reference's arm pays inside every read_column_by_idx
    try:
        for o in outs:
            o.zero_()
        barrier()
        t0 = time.perf_counter()
        r2 = pq.Reader.from_pointer(host.data_ptr(), size, device=local)
        r2.read_columns_into(list(range(ncols)), dsts, -1)
        torch.cuda.synchronize()
        cold_s = time.perf_counter() - t0
        open_s2 = r2.page_scan_seconds if hasattr(r2, "page_scan_seconds") else None
        r2.close()
        for c in range(ncols):
            if not np.array_equal(outs[c].numpy().view(np.uint64)[::4099], np.ascontiguousarray(cols[c]["fixed"]).view(np.uint64)[::4099]):
                raise AssertionError(f"bench parity check (cold read) failed on column {c}")
        if world > 1:
            t = torch.tensor([cold_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cold_s = float(t.item())
        e2e["cold"] = {"value": bytes_in * world / cold_s / 1e9, "unit": UNIT, "ms": cold_s * 1e3, "page_scan_ms": None if open_s2 is None else open_s2 * 1e3,
                       "includes": "pqr_open_memory (footer parse + page-header scan of every column chunk on the host threads) + descriptor tables + "
                                   "plan / device-buffer creation + the first pipelined read of all columns"}
    except AssertionError:
        raise
    except Exception as ex:
        e2e["cold"] = {"error": str(ex)}
    if affinity:
        e2e["rank_affinity"] = affinity
    reader.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64/f64 bit moves", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline,
            "bytes": {"payload_in_per_step": int(bytes_in), "decoded_out_per_step": int(bytes_out), "pages": int(n_pages),
                      "values_per_step": a.rows * ncols},
            "ingest": {"host_open_page_scan_s": open_s, "h2d_ms": h2d_ms, "h2d_GBps": size / (h2d_ms * 1e-3) / 1e9 if h2d_ms else None},
            "frac_of_hbm_peak_in_plus_out": (bytes_in + bytes_out) * world / (ms_per_step * 1e-3) / 1e9 / (peak * world)}
    # BASELINE configs[2..4] next to the headline, at every N (bench_scans.py)
    del outs, host
    torch.cuda.empty_cache()
    if not a.no_scans:
        import bench_scans as bs
        dd = bs.Dist(rank, world)
        psrc = roofline["peak_source"]
        want_cpu = world == 1 and not a.no_cpu_baseline
        scan_steps = max(1, min(a.scan_steps, a.steps))
        if a.strings_rows > 0:
            t0 = time.time()
            line_strings = bs.strings_bench(pq, local, stream, dd, scan_steps, warmup, peak, psrc, a.strings_rows, min(a.rg_rows, a.strings_rows), want_cpu)
            line_strings["wall_s"] = time.time() - t0
            line["strings"] = line_strings
        if a.regex_rows > 0:
            t0 = time.time()
            line_regex = bs.regex_bench(pq, local, stream, dd, scan_steps, warmup, peak, psrc, a.regex_rows, min(1_250_000, a.regex_rows), want_cpu)
            line_regex["wall_s"] = time.time() - t0
            line["regex"] = line_regex
        if a.chunk_rgs > 0:
            t0 = time.time()
            line_ci = bs.chunk_index_bench(pq, local, stream, dd, min(scan_steps, 3), 1, peak, psrc, a.chunk_rgs, a.chunk_rg_rows, want_cpu)
            line_ci["wall_s"] = time.time() - t0
            line["chunk_index"] = line_ci
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            r = cpu_reference_run(pq, cols, a.rows, a.rg_rows, a.cpu_sample_rgs, 1, 0, cores)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                                    "sample": r["sample"], "seconds": r["seconds"], "Mvalues_per_s": r["values"] / r["seconds"] / 1e6}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"unavailable: {e}"}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
