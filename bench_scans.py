"""bench_scans.py -- the BASELINE.json configurations next to the cfg2 headline of bench.py, measured at every N:

  strings      configs[2]: OPTIONAL BYTE_ARRAY, 65 536-entry dictionary ("city_%06u_x"), 30 % nulls, RLE definition
               levels; 100 M rows per GPU (weak scaling: the 1 B-row column is 10 such shards), decoded to
               offsets + chars + validity
  regex        configs[3]: --regex-column page pruning over ONE ~10 GB PLAIN BYTE_ARRAY file (305 M e-mail-like values,
               anchored pattern, once plain and once --neg-regex).  N > 1: the file's row groups are split over the ranks
               by pqr_shard_row_groups (strong scaling); every rank scans its shard, the page bitmaps are gathered on the
               host (multi_gpu.gather_page_bits) and compared with a whole-file scan on rank 0
  chunk_index  configs[4]: the 4 KB chunk index over a mixed PLAIN / dictionary string column, row groups alternating
               between the two, 6.25 GB per GPU (8 GPUs = the 50 GB of the configuration); the shards of the ranks form ONE
               logical column: the carry chain crosses the ranks (multi_gpu.chain_carry)

Every object carries its own `roofline` (algorithmic bytes of SURVEY.md 8 d over the CUDA-event kernel time, against
MEASURED_PEAKS.json), a `cpu_baseline` on a bounded sample, and a `parity` statement: the full-size device output against
the generator's inputs where the domain allows it, plus a sample row group against the oracle (tests/oraclelib: the only
use of oracle/ here is as the checker and as the CPU baseline)."""
import ctypes as C
import os
import threading
import time

import numpy as np

EMAIL_PATTERN = r"^[a-z0-9._]+@[a-z0-9.]+\.com$"
BYTE_ARRAY, OPTIONAL, REQUIRED, UTF8 = 6, 1, 0, 0


def log(*a):
    import sys
    print(*a, file=sys.stderr, flush=True)


class Dist:
    """torch.distributed when the process group exists (N > 1), no-ops otherwise"""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(self, x):
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x):
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())


class _DevView:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def dev_tensor(ptr, n, typestr):
    import torch
    return torch.as_tensor(_DevView(ptr, n, typestr), device="cuda")


def scratch_dir(need_bytes):
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


def roofline(kernel, bytes_in, bytes_out, ms, peak, peak_source):
    ach = (bytes_in + bytes_out) / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
            "traffic": None, "algorithmic_bytes": int(bytes_in + bytes_out), "kernel_ms": ms, "peak_source": peak_source}


def host_image(gen):
    """the generated file as pageable host bytes (numpy); the job is freed"""
    buf = np.empty(gen.size + 64, dtype=np.uint8)
    gen.emit(buf.ctypes.data, gen.size)
    size = gen.size
    gen.free()
    return buf, size


def upload_range(pq, ctx, buf, host, lo, hi):
    """file bytes [lo, hi) -> the same offsets of the device image"""
    lo &= ~15
    pq.lib().pqg_buf_write(ctx.h, buf, lo, host.ctypes.data + lo, hi - lo)


# ── configs[2]: dictionary strings ─────────────────────────────────────────────────────────────────────────────
def strings_bench(pq, device, stream, dist, steps, warmup, peak, peak_source, rows, rg_rows, cpu_baseline=True):
    import torch
    t0 = time.time()
    seed = 333
    first_row = dist.rank * rows  # the ranks' shards are consecutive pieces of one column
    col = pq.synth_strings(pq.PQGEN_CITY64K, rows, seed, first_row=first_row, null_permille=300)
    specs = [("city", BYTE_ARRAY, OPTIONAL, UTF8)]
    gen = pq.generate(specs, [col], rg_split(rows, rg_rows))
    host, size = host_image(gen)
    log(f"[bench] strings workload: {rows} rows, file {size / 1e9:.3f} GB, built in {time.time() - t0:.1f}s")
    reader = pq.Reader.from_pointer(host.ctypes.data, size, device=device)
    ctx = pq.Context(device, stream.cuda_stream)
    image = ctx.upload(host.ctypes.data, size)
    tables = reader.column_tables(0, -1)
    plan = ctx.plan(image, tables)
    n_pages = tables[3]
    ctx.set_profiling(True)
    for _ in range(max(warmup, 2)):
        plan.run()
        plan.finish()
    bytes_in, bytes_out = plan.bytes_in, plan.bytes_out
    # full-size parity on the device: chars == the input strings of the non-null rows, in order; offsets == 13 * rank
    L = 13
    isn = torch.from_numpy(col["is_null"][:rows]).cuda().bool()
    exp_chars = torch.from_numpy(col["chars"][:rows * L]).cuda().view(rows, L)[~isn].reshape(-1)
    got_chars = dev_tensor(plan.chars_ptr, plan.chars_size, "|u1")
    n_chunks = tables[1]
    if got_chars.numel() != exp_chars.numel() or not torch.equal(got_chars, exp_chars):
        raise AssertionError("strings bench parity: decoded chars differ from the generator's input strings")
    # offsets: chunk c owns [row_base_c + c, row_base_c + c + n_c]; relative to the chunk's chars
    offs = dev_tensor(plan.offsets_ptr, rows + n_chunks, "<u4").to(torch.int64)
    valid = (~isn).to(torch.int64)
    r0 = 0
    for c in range(n_chunks):
        n_c = int(tables[0][c].num_values)
        exp = torch.zeros(n_c + 1, dtype=torch.int64, device="cuda")
        torch.cumsum(valid[r0:r0 + n_c] * L, 0, out=exp[1:])
        if not torch.equal(offs[r0 + c:r0 + c + n_c + 1], exp):
            raise AssertionError(f"strings bench parity: offsets of chunk {c} differ")
        r0 += n_c
    vwords = dev_tensor(plan.validity_ptr, (rows + 31) // 32, "<u4").to(torch.int64)
    bitpos = torch.arange(32, device="cuda", dtype=torch.int64)
    gotv = ((vwords.view(-1, 1) >> bitpos) & 1).reshape(-1)[:rows].bool()
    if not torch.equal(gotv, ~isn):
        raise AssertionError("strings bench parity: validity differs")
    del exp_chars, got_chars, offs, valid, vwords, gotv, isn
    torch.cuda.empty_cache()
    # timed steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    with torch.cuda.stream(stream):
        ev0.record()
        for _ in range(steps):
            plan.run()
        ev1.record()
    stream.synchronize()
    dist.barrier()
    plan.finish()
    ms = dist.allmax(ev0.elapsed_time(ev1) / steps)
    tm = plan.timings_avg(min(steps, 8))
    k_ms = tm["dict_ms"] + tm["str_size_ms"] + tm["str_copy_ms"]
    res = {"metric": "decoded_page_GBps", "unit": "GB/s", "value": bytes_in * dist.world / (ms * 1e-3) / 1e9, "ms_per_step": ms, "scaling": "weak",
           "workload": f"cfg3 shape: OPTIONAL BYTE_ARRAY, 65536-entry dictionary, 30 % nulls, RLE def levels, {rows} rows x {dist.world} GPU(s) "
                       f"({len(rg_split(rows, rg_rows))} row groups of {rg_rows} per GPU), device resident",
           "rows_per_gpu": rows, "pages_per_gpu": int(n_pages), "payload_in_per_gpu": int(bytes_in), "decoded_out_per_gpu": int(bytes_out),
           "in_plus_out_GBps": (bytes_in + bytes_out) * dist.world / (ms * 1e-3) / 1e9,
           "kernel_ms": {"dict_prepare": tm["dict_ms"], "size_pass": tm["str_size_ms"], "copy_pass": tm["str_copy_ms"]},
           "roofline": roofline("k_dict_prepare<0> + k_str_pages<sizes> + k_str_pages<copy>", bytes_in, bytes_out, k_ms, peak, peak_source),
           "parity": "full size, on the device: chars == the input strings of the non-null rows, offsets == 13 x rank, validity == !is_null"}
    # end to end through the reader: pinned host file image in, pinned host offsets / chars / validity out, row group by row
    # group on two alternating contexts (pqr_read_strings_into) -- against the same column uploaded, decoded and downloaded
    # as ONE plan (pqr_read_columnar's way: nothing overlaps)
    chars_size = int(plan.chars_size)
    pin_img = torch.empty(size + 64, dtype=torch.uint8, pin_memory=True)
    pin_img.numpy()[:size] = host[:size]
    r2 = pq.Reader.from_pointer(pin_img.data_ptr(), size, device=device)
    n_rg = r2.num_row_groups
    o_off = torch.empty(rows + n_chunks + 8, dtype=torch.int32, pin_memory=True)
    o_chars = torch.empty(chars_size + 64, dtype=torch.uint8, pin_memory=True)
    o_val = torch.empty((rows + 31) // 32 + 1, dtype=torch.int32, pin_memory=True)
    o_base = np.zeros(n_chunks + 2, dtype=np.uint64)
    dsts = ((o_off.data_ptr(), o_off.numel()), (o_chars.data_ptr(), o_chars.numel()), o_base, (o_val.data_ptr(), o_val.numel()))
    e2e_steps = max(2, min(steps, 5))
    st = None
    for _ in range(2):
        st = r2.read_strings_into(0, 0, n_rg, dsts[0], dsts[1], dsts[2], dsts[3])
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        st = r2.read_strings_into(0, 0, n_rg, dsts[0], dsts[1], dsts[2], dsts[3])
    e2e_ms = dist.allmax((time.perf_counter() - t0) * 1e3 / e2e_steps)
    if st["chars_size"] != chars_size or not np.array_equal(o_chars.numpy()[:1 << 20], dev_tensor(plan.chars_ptr, 1 << 20, "|u1").cpu().numpy()):
        raise AssertionError("strings bench e2e: the pipelined read differs from the plan's output")
    # one plan, nothing overlapped: upload, run, finish, download, sync
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pq.lib().pqg_buf_write(ctx.h, image, 0, pin_img.data_ptr(), size)
        plan.run()
        plan.finish()
        plan.download(offsets=o_off.data_ptr(), chars=o_chars.data_ptr(), validity=o_val.data_ptr())
        ctx.sync()
    serial_ms = dist.allmax((time.perf_counter() - t0) * 1e3 / e2e_steps)
    res["e2e"] = {"value": bytes_in * dist.world / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e2e_ms,
                  "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]),
                  "api": "pqr_read_strings_into: pinned file image in, pinned offsets / chars / validity out, one cached plan per row group on two "
                         "alternating contexts (H2D + size pass of row group k + 1 under the copy pass + D2H of row group k)",
                  "one_plan_no_overlap_ms": serial_ms, "speedup_vs_one_plan": serial_ms / e2e_ms if e2e_ms else None}
    r2.close()
    del pin_img, o_off, o_chars, o_val
    # sample row group against the oracle (+ CPU baseline: the reference's ColumnReader on a bounded sample)
    if dist.rank == 0:
        res["parity"] += "; " + _strings_sample_check(pq, specs, col, plan, tables, rg_rows)
        if cpu_baseline:
            res["cpu_baseline"] = _strings_cpu_baseline(pq, specs, seed)
    plan.destroy()
    ctx.buf_free(image)
    reader.close()
    ctx.close()
    return res


def _strings_sample_check(pq, specs, col, plan, tables, rg_rows):
    """row group 0 of the timed plan against the oracle's read_column of the same row group (written as its own file)"""
    import oraclelib
    oraclelib.build_oracle()
    orc = oraclelib.Oracle()
    take = min(rg_rows, 2_000_000)  # the chunk of a row group depends on its rows only: a shorter row group is its own chunk
    sub = dict(str_off=col["str_off"][:take + 1], chars=col["chars"][:take * 13], is_null=col["is_null"][:take])
    path = os.path.join(scratch_dir(1 << 30), f"pqg_bench_str_{os.getpid()}.parquet")
    g = pq.generate(specs, [sub], [take])
    g.write(path)
    g.free()
    try:
        h = orc.open(path)
        exp = orc.read_column_by_idx(h, 0, 0)
        orc.close(h)
        r = pq.Reader(path)
        got = r.read_column_by_idx(0, 0)
        r.close()
    finally:
        os.unlink(path)
    g_ = oraclelib.Values(got["is_null"], got["vidx"], got["fixed"], got["str_off"], got["chars"])
    d = g_.diff(exp)
    if d is not None:
        raise AssertionError(f"strings bench: GPU read_column differs from the oracle on the sample: {d}")
    return f"the first {take} rows as their own file: pqr_read_column_by_idx bit-exact with the oracle"


def _strings_cpu_baseline(pq, specs, seed, n_rgs=None, rg_rows=500_000):
    import oraclelib
    cores = os.cpu_count() or 1
    if not oraclelib.Ref.available():
        return {"value": None, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": "unavailable: oracle/_ref/libpqref.so missing"}
    ref = oraclelib.Ref()
    n_rgs = n_rgs or max(2 * cores, 8)
    rows = n_rgs * rg_rows
    col = pq.synth_strings(pq.PQGEN_CITY64K, rows, seed, null_permille=300)
    path = os.path.join(scratch_dir(2 << 30), f"pqg_bench_strref_{os.getpid()}.parquet")
    g = pq.generate(specs, [col], [rg_rows] * n_rgs)
    g.write(path)
    g.free()
    try:
        r = pq.Reader(path)
        payload = int(r.page_index()[:, 1].sum())
        chunks, nc, _, _, _ = r.column_tables(0, -1)
        payload += sum(int(chunks[i].dict_size) for i in range(nc))
        r.close()
        t, nv = ref.time_read_chunks(path, list(range(n_rgs)), [0] * n_rgs, cores)
    finally:
        os.unlink(path)
    return {"value": payload / t / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference", "seconds": t, "Mvalues_per_s": nv / t / 1e6,
            "sample": f"{n_rgs} row groups x {rg_rows} rows of the same generator ({payload / 1e6:.0f} MB payload), "
                      f"ParquetReader::read_column_by_idx per chunk, one reader per thread, page cache warm"}


# ── configs[3]: regex page pruning ─────────────────────────────────────────────────────────────────────────────
def regex_bench(pq, device, stream, dist, steps, warmup, peak, peak_source, rows, rg_rows, cpu_baseline=True):
    import torch
    mg = __import__("importlib").import_module("duckdb-parquet-parser_b200.multi_gpu")
    specs = [("email", BYTE_ARRAY, REQUIRED, UTF8)]
    seed = 99
    path = None
    t0 = time.time()
    if dist.world == 1:
        col = pq.synth_strings(pq.PQGEN_EMAILS, rows, seed)
        gen = pq.generate(specs, [col], rg_split(rows, rg_rows))
        del col
        host, size = host_image(gen)
        reader = pq.Reader.from_pointer(host.ctypes.data, size, device=device)
    else:
        # ONE file for all ranks: rank 0 writes it, everybody maps it
        path = os.path.join(scratch_dir(12 << 30), "pqg_bench_regex_shared.parquet")
        if dist.rank == 0:
            col = pq.synth_strings(pq.PQGEN_EMAILS, rows, seed)
            gen = pq.generate(specs, [col], rg_split(rows, rg_rows))
            del col
            gen.write(path)
            gen.free()
        dist.barrier()
        host = np.memmap(path, dtype=np.uint8, mode="r")
        size = host.size
        reader = pq.Reader(path, device=device)
    log(f"[bench] regex workload: {rows} rows, file {size / 1e9:.3f} GB, ready in {time.time() - t0:.1f}s")
    ctx = pq.Context(device, stream.cuda_stream)
    L = pq.lib()
    b = reader.shard_row_groups(0, dist.world)
    rg0, rg1 = b[dist.rank], b[dist.rank + 1]
    tables = reader.column_tables_rgs(0, rg0, rg1)
    n_local = tables[3]
    buf = C.c_void_p()
    ctx.check(L.pqg_buf_alloc(ctx.h, size, C.byref(buf)))
    whole = dist.world > 1 and dist.rank == 0  # rank 0 also keeps the whole file for the identity check
    if whole or dist.world == 1:
        ctx.check(L.pqg_buf_write(ctx.h, buf, 0, host.ctypes.data, size))
    elif n_local:
        lo = min(int(tables[2][0].payload_off), int(tables[0][0].dict_off) if tables[0][0].has_dict else 1 << 62) & ~15
        last = tables[2][n_local - 1]
        hi = int(last.payload_off) + int(last.payload_size)
        ctx.check(L.pqg_buf_write(ctx.h, buf, lo, host.ctypes.data + lo, hi - lo))
    ctx.sync()
    plan = ctx.plan(buf, tables) if n_local else None
    bytes_local = plan.bytes_in if plan else 0
    out = {}
    for neg in (0, 1):
        dfa = pq.regex_compile(EMAIL_PATTERN)
        bits = np.zeros((n_local + 31) // 32 + 1, dtype=np.uint32)
        ms = C.c_float(0)
        times = []
        dist.barrier()
        for it in range(warmup + steps):
            if plan:
                ctx.check(L.pqg_regex_scan(ctx.h, plan.h, dfa, neg, bits.ctypes.data, C.byref(ms)))
            if it >= warmup:
                times.append(ms.value)
        L.pqg_dfa_free(dfa)
        idx = np.arange(n_local)
        unpacked = ((bits[idx >> 5] >> (idx & 31).astype(np.uint32)) & 1).astype(np.uint8)
        t_g = time.perf_counter()
        allbits, per_rank = mg.gather_page_bits(unpacked, dist.world)
        gather_s = time.perf_counter() - t_g
        out[neg] = dict(ms=dist.allmax(sum(times) / max(len(times), 1)), bits=allbits, per_rank=per_rank, gather_ms=dist.allmax(gather_s * 1e3))
    n_pages = len(out[0]["bits"])
    bytes_in = int(dist.allsum(bytes_local))
    parity = []
    if dist.rank == 0:
        if whole:
            # the single-GPU answer on rank 0: one plan over the whole file
            wt = reader.column_tables(0, -1)
            wp = ctx.plan(buf, wt)
            for neg in (0, 1):
                dfa = pq.regex_compile(EMAIL_PATTERN)
                bits = np.zeros((wt[3] + 31) // 32 + 1, dtype=np.uint32)
                ms = C.c_float(0)
                ctx.check(L.pqg_regex_scan(ctx.h, wp.h, dfa, neg, bits.ctypes.data, C.byref(ms)))
                L.pqg_dfa_free(dfa)
                idx = np.arange(wt[3])
                single = ((bits[idx >> 5] >> (idx & 31).astype(np.uint32)) & 1).astype(np.uint8)
                if not np.array_equal(single, out[neg]["bits"]):
                    raise AssertionError("regex bench: the gathered bitmap differs from the single-GPU scan of the whole file")
            wp.destroy()
            parity.append(f"bitmaps gathered from {dist.world} shards identical to a single-GPU scan of the whole file (both polarities)")
        parity.append(_regex_sample_check(pq, specs, seed, rg_rows, out))
    ms = out[0]["ms"]
    res = {"metric": "regex_pruned_pages_per_s", "unit": "pages/s", "value": n_pages / (ms * 1e-3), "pattern": EMAIL_PATTERN, "scaling": "strong",
           "workload": f"cfg4: {rows} PLAIN BYTE_ARRAY e-mail-like values, {n_pages} pages, {bytes_in / 1e9:.2f} GB payload in ONE file "
                       f"({len(rg_split(rows, rg_rows))} row groups), device resident"
                       + (f", row groups split over {dist.world} GPUs by pqr_shard_row_groups" if dist.world > 1 else ""),
           "kernel_ms": ms, "payload_GBps": bytes_in / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_in / (ms * 1e-3) / 1e9 / (peak * dist.world),
           "pages": int(n_pages), "pages_per_rank": [int(x) for x in out[0]["per_rank"]], "pages_pruned": int((out[0]["bits"] == 0).sum()),
           "host_gather_ms": out[0]["gather_ms"],
           "neg_regex": {"value": n_pages / (out[1]["ms"] * 1e-3), "kernel_ms": out[1]["ms"], "pages_pruned": int((out[1]["bits"] == 0).sum()),
                         "host_gather_ms": out[1]["gather_ms"]},
           "roofline": roofline("k_regex_tiles (per GPU: the slowest shard)", bytes_in / dist.world, (n_pages / dist.world + 7) // 8, ms, peak, peak_source),
           "parity": "; ".join(parity) if parity else None}
    if dist.rank == 0 and cpu_baseline:
        res["cpu_baseline"] = _regex_cpu_baseline(pq, specs, seed, rg_rows)
    if plan:
        plan.destroy()
    ctx.buf_free(buf)
    reader.close()
    ctx.close()
    del host
    dist.barrier()
    if path and dist.rank == 0:
        os.unlink(path)
    return res


def _sample_file(pq, kind, specs, seed, first_row, rows, tag):
    col = pq.synth_strings(kind, rows, seed, first_row=first_row)
    path = os.path.join(scratch_dir(1 << 30), f"pqg_bench_{tag}_{os.getpid()}_{first_row}.parquet")
    g = pq.generate(specs, [col], [rows])
    g.write(path)
    g.free()
    return path


def _regex_sample_check(pq, specs, seed, rg_rows, out):
    """row group 3 (it holds a noise block boundary) through the oracle: reference-order decode + backtracking matcher"""
    import oraclelib
    oraclelib.build_oracle()
    orc = oraclelib.Oracle()
    rg = 3 if len(out[0]["bits"]) > 0 else 0
    path = _sample_file(pq, pq.PQGEN_EMAILS, specs, seed, rg * rg_rows, rg_rows, "rx")
    try:
        h = orc.open(path)
        exp = orc.regex_prune(h, 0, EMAIL_PATTERN, False)
        expn = orc.regex_prune(h, 0, EMAIL_PATTERN, True)
        orc.close(h)
    finally:
        os.unlink(path)
    # pages of row group rg in the column's global page order: every row group has the same page count here
    # (fixed-length values), except that the last one may be shorter
    k = len(exp)
    a = rg * k
    if not (np.array_equal(exp, out[0]["bits"][a:a + k]) and np.array_equal(expn, out[1]["bits"][a:a + k])):
        raise AssertionError("regex bench: page bitmap differs from the oracle on the sample row group")
    return f"row group {rg} ({rg_rows} values, {k} pages): page bitmaps of both polarities identical to the oracle"


def _regex_cpu_baseline(pq, specs, seed, rg_rows):
    """the oracle's restatement (the reference's regex mode has no source: kind "port"), one row group per thread"""
    import oraclelib
    oraclelib.build_oracle()
    threads = max(1, min(os.cpu_count() or 1, 16))
    paths = [_sample_file(pq, pq.PQGEN_EMAILS, specs, seed, t * rg_rows, rg_rows, "rxcpu") for t in range(threads)]
    pages = [0] * threads

    def work(t):
        orc = oraclelib.Oracle()
        h = orc.open(paths[t])
        pages[t] = len(orc.regex_prune(h, 0, EMAIL_PATTERN, False))
        orc.close(h)
    try:
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        dt = time.perf_counter() - t0
    finally:
        for p in paths:
            os.unlink(p)
    return {"value": sum(pages) / dt, "unit": "pages/s", "cores": threads, "kind": "port", "seconds": dt,
            "sample": f"{threads} row groups x {rg_rows} values ({sum(pages)} pages), one per thread: oracle/regex_oracle.c over the reference-order "
                      f"decode (file open + page walk included)"}


# ── configs[4]: 4 KB chunk index over a mixed PLAIN / dictionary column ────────────────────────────────────────────
def mixed_column(pq, first_rg, n_rgs, rg_rows, seed):
    """row groups alternate: even = PLAIN (near-unique e-mails), odd = dictionary (65 536 city names)"""
    lens = [33 if (first_rg + g) % 2 == 0 else 13 for g in range(n_rgs)]
    chars = np.empty(sum(lens) * rg_rows, dtype=np.uint8)
    off = np.empty(n_rgs * rg_rows + 1, dtype=np.uint64)
    cpos = 0
    for g in range(n_rgs):
        kind = pq.PQGEN_EMAILS if lens[g] == 33 else pq.PQGEN_CITY64K
        pq.synth_strings(kind, rg_rows, seed, first_row=(first_rg + g) * rg_rows, off_base=cpos,
                         chars=chars[cpos:cpos + lens[g] * rg_rows], str_off=off[g * rg_rows:(g + 1) * rg_rows + 1])
        cpos += lens[g] * rg_rows
    return dict(str_off=off, chars=chars), lens


def chunk_index_bench(pq, device, stream, dist, steps, warmup, peak, peak_source, n_rgs, rg_rows, cpu_baseline=True):
    import torch
    mg = __import__("importlib").import_module("duckdb-parquet-parser_b200.multi_gpu")
    specs = [("comment", BYTE_ARRAY, REQUIRED, UTF8)]
    seed = 555
    t0 = time.time()
    first_rg = dist.rank * n_rgs
    col, lens = mixed_column(pq, first_rg, n_rgs, rg_rows, seed)
    gen = pq.generate(specs, [col], [rg_rows] * n_rgs)
    rows = n_rgs * rg_rows
    weight_total_exp = sum((L + 2) * rg_rows for L in lens)  # decimal digits of 33 / 13 = 2
    del col
    host, size = host_image(gen)
    log(f"[bench] chunk-index workload: {rows} rows in {n_rgs} alternating PLAIN / dictionary row groups, file {size / 1e9:.3f} GB, built in {time.time() - t0:.1f}s")
    reader = pq.Reader.from_pointer(host.ctypes.data, size, device=device)
    ctx = pq.Context(device, stream.cuda_stream)
    image = ctx.upload(host.ctypes.data, size)
    tables = reader.column_tables(0, -1)
    plan = ctx.plan(image, tables)
    n_pages = tables[3]
    ctx.set_profiling(True)
    L_ = pq.lib()
    S = 4096
    rec = []
    job = C.c_void_p()
    n_chunks_local = base = 0
    for it in range(warmup + steps):
        if job:
            L_.pqg_chunk_job_free(ctx.h, job)
            job = C.c_void_p()
        dist.barrier()
        w0 = time.perf_counter()
        plan.run()
        plan.finish()
        dec = plan.timings()
        pms, ems = C.c_float(0), C.c_float(0)
        ctx.check(L_.pqg_chunk_index_prepare(ctx.h, plan.h, S, C.byref(job), C.byref(pms)))
        w1 = time.perf_counter()

        def stitch(carry):
            n, c = C.c_uint64(0), C.c_uint64(0)
            ctx.check(L_.pqg_chunk_index_stitch(ctx.h, job, carry, C.byref(n), C.byref(c)))
            return n.value, c.value
        base, n_chunks_local = mg.chain_carry(stitch, dist.rank, dist.world)
        w2 = time.perf_counter()
        ctx.check(L_.pqg_chunk_index_emit(ctx.h, job, base, None, C.byref(ems)))
        w3 = time.perf_counter()
        if it >= warmup:
            rec.append(dict(decode_ms=dec["total_ms"], prepare_ms=pms.value, emit_ms=ems.value, chain_wall_ms=(w2 - w1) * 1e3,
                            parallel_wall_ms=(w1 - w0 + w3 - w2) * 1e3, wall_ms=(w3 - w0) * 1e3))
    mean = {k: sum(r[k] for r in rec) / len(rec) for k in rec[0]}
    kernel_ms = dist.allmax(mean["decode_ms"] + mean["prepare_ms"] + mean["emit_ms"])
    wall_ms = dist.allmax(mean["wall_ms"])
    chain_ms = dist.allmax(mean["chain_wall_ms"])
    total_chunks = int(dist.allmax(base + n_chunks_local))
    # full-size checks on the device: ids are monotone, start at this rank's base (or base + 1 when the carried chunk was
    # full), end at base + n - 1; the shard's weight is what the generator put in
    ids = dev_tensor(L_.pqg_chunk_job_ids(job), rows, "<u4").to(torch.int64)
    if int(L_.pqg_chunk_job_total_weight(job)) != weight_total_exp:
        raise AssertionError("chunk-index bench: the shard's total weight differs from the generator's")
    if not bool((ids[1:] >= ids[:-1]).all()) or int(ids[-1]) != base + n_chunks_local - 1 or int(ids[0]) not in (base, base + 1):
        raise AssertionError("chunk-index bench: chunk ids are not a monotone chain over the shard")
    step_ = (ids[1:] - ids[:-1])
    if int(step_.max()) > 1:
        raise AssertionError("chunk-index bench: a chunk id was skipped")
    # every closed chunk holds >= 4096 bytes and closing one value earlier would not have: weights per id
    w = torch.empty(rows, dtype=torch.int64, device="cuda")
    r0 = 0
    for g in range(n_rgs):
        w[r0:r0 + rg_rows] = lens[g] + 2
        r0 += rg_rows
    per = torch.zeros(n_chunks_local + 1, dtype=torch.int64, device="cuda")
    per.index_add_(0, ids - base, w)
    lastw = torch.zeros(n_chunks_local + 1, dtype=torch.int64, device="cuda")
    last_idx = torch.nonzero(step_ > 0).reshape(-1)
    lastw[ids[last_idx] - base] = w[last_idx]
    closed = per[1:n_chunks_local - 1] if n_chunks_local > 2 else per[:0]  # chunks that start and end inside this shard
    closed_last = lastw[1:n_chunks_local - 1] if n_chunks_local > 2 else lastw[:0]
    if closed.numel() and (int(closed.min()) < S or int((closed - closed_last).max()) >= S):
        raise AssertionError("chunk-index bench: a chunk closed too early or too late")
    del ids, w, per, lastw, step_
    torch.cuda.empty_cache()
    bytes_in = plan.bytes_in
    bytes_out = plan.bytes_out + 4 * rows + 8 * n_pages
    res = {"metric": "chunk_index_GBps", "unit": "GB/s", "value": bytes_in * dist.world / (wall_ms * 1e-3) / 1e9, "scaling": "weak",
           "workload": f"cfg5 shape: 4 KB chunk index over a mixed PLAIN / dictionary BYTE_ARRAY column, {n_rgs} alternating row groups of {rg_rows} rows "
                       f"per GPU = {bytes_in / 1e9:.2f} GB payload per GPU x {dist.world} GPU(s) ({bytes_in * dist.world / 1e9:.1f} GB in all); the ranks' shards are "
                       f"consecutive pieces of ONE column, the carry chain crosses the ranks",
           "rows_per_gpu": rows, "pages_per_gpu": int(n_pages), "total_chunks": total_chunks, "ms_per_step_wall": wall_ms,
           "kernel_ms": {"decode": mean["decode_ms"], "prepare(weights+scan+walks)": mean["prepare_ms"], "emit(cuts+ids)": mean["emit_ms"]},
           "serial_chain_ms": chain_ms, "rows_per_s": rows * dist.world / (wall_ms * 1e-3),
           "kernel_GBps": bytes_in * dist.world / (kernel_ms * 1e-3) / 1e9,
           "roofline": roofline("k_str_pages + scan + chain kernels", bytes_in, bytes_out, kernel_ms, peak, peak_source),
           "parity": "full size, on the device: shard weight == the generator's, ids monotone without gaps from id_base to id_base + n - 1, "
                     "every chunk closed inside the shard holds >= 4096 bytes and < 4096 without its last value"}
    if dist.rank == 0:
        res["parity"] += "; " + _chunk_sample_check(pq, specs, seed, rg_rows, L_, job, ctx)
        if cpu_baseline:
            res["cpu_baseline"] = _chunk_cpu_baseline(pq, specs, seed, rg_rows)
    L_.pqg_chunk_job_free(ctx.h, job)
    plan.destroy()
    ctx.buf_free(image)
    reader.close()
    ctx.close()
    return res


def _chunk_sample_file(pq, specs, seed, rg_rows, n_rgs):
    col, _ = mixed_column(pq, 0, n_rgs, rg_rows, seed)
    path = os.path.join(scratch_dir(2 << 30), f"pqg_bench_ci_{os.getpid()}.parquet")
    g = pq.generate(specs, [col], [rg_rows] * n_rgs)
    g.write(path)
    g.free()
    return path


def _chunk_sample_check(pq, specs, seed, rg_rows, L_, job, ctx):
    """rank 0's first two row groups (one PLAIN, one dictionary): the chain from row 0 depends on nothing behind them"""
    import oraclelib
    import torch
    oraclelib.build_oracle()
    orc = oraclelib.Oracle()
    n = 2 * rg_rows
    path = _chunk_sample_file(pq, specs, seed, rg_rows, 2)
    try:
        h = orc.open(path)
        exp, _ = orc.chunk_index(h, "comment", 4096)
        orc.close(h)
    finally:
        os.unlink(path)
    got = dev_tensor(L_.pqg_chunk_job_ids(job), n, "<u4").cpu().numpy().astype(np.uint64)
    if not np.array_equal(got, np.asarray(exp, dtype=np.uint64)):
        raise AssertionError("chunk-index bench: tuple_to_chunk differs from the oracle on the first two row groups")
    return f"the first 2 row groups ({n} rows): tuple_to_chunk identical to the oracle's restatement of src/main.cpp:21-32"


def _chunk_cpu_baseline(pq, specs, seed, rg_rows):
    """the unmodified reference: StringColumnIterator + the loop of src/main.cpp (single-threaded by construction)"""
    import oraclelib
    if not oraclelib.Ref.available():
        return {"value": None, "unit": "GB/s", "cores": 1, "kind": "reference", "sample": "unavailable: oracle/_ref/libpqref.so missing"}
    ref = oraclelib.Ref()
    path = _chunk_sample_file(pq, specs, seed, rg_rows, 2)
    try:
        r = pq.Reader(path)
        payload = int(r.page_index()[:, 1].sum())
        chunks, nc, _, _, _ = r.column_tables(0, -1)
        payload += sum(int(chunks[i].dict_size) for i in range(nc))
        r.close()
        h = ref.open(path)
        t0 = time.perf_counter()
        _, n = ref.chunk_index(h, "comment", 4096)
        dt = time.perf_counter() - t0
        ref.close(h)
    finally:
        os.unlink(path)
    return {"value": payload / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "reference", "seconds": dt, "rows_per_s": 2 * rg_rows / dt,
            "sample": f"2 row groups x {rg_rows} rows ({payload / 1e6:.0f} MB payload, {n} chunks): StringColumnIterator + the loop of src/main.cpp:21-32"}
