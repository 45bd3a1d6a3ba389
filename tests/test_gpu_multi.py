"""The row-group split on REAL devices (SURVEY.md section 8 e): world_size 2, one process per GPU, every rank decodes /
scans / indexes its own contiguous run of row groups of ONE file (pqr_shard_row_groups) with the CUDA kernels, the
host gathers page bitmaps and chains the chunk index (multi_gpu.py) -- and the gathered results must be identical to
the single-GPU run of the whole file.  Skipped on boxes with one GPU (the gloo world_size-2 test in
tests/test_multi_cpu.py covers the exchange there)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
PATTERN = r"^[a-z0-9._]+@[a-z0-9.]+\.com$"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, outdir):
    for p in (HERE, os.path.dirname(HERE)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    import torch
    import torch.distributed as dist
    import pqb200 as pq
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    mg = importlib.import_module("duckdb-parquet-parser_b200.multi_gpu")
    r = pq.Reader(path, device=rank)
    res = {}
    for c in range(r.num_columns):
        ci = r.column_info(c)
        name = ci["name"]
        if ci["type"] == pq.BYTE_ARRAY:
            bits, per = mg.regex_prune_sharded(r, c, PATTERN, False, rank, world)
            nbits, _ = mg.regex_prune_sharded(r, c, "Berlin|user1", True, rank, world)
            t2c, total = mg.chunk_index_sharded(r, name, 512, rank, world, col=c)
            res[name] = ("str", bits, nbits, t2c, total, per)
        elif ci["type"] in (pq.INT32, pq.INT64, pq.FLOAT, pq.DOUBLE):
            b = mg.shard_bounds(r, c, world)
            rows = sum(r.row_group_num_rows(g) for g in range(b[rank], b[rank + 1]))
            w = 4 if ci["type"] in (pq.INT32, pq.FLOAT) else 8
            vals = np.zeros(max(rows * w, 8), dtype=np.uint8)
            mask = np.zeros((rows + 31) // 32 + 1, dtype=np.uint32)
            st = r.read_columns_into_rgs([c], [(vals.ctypes.data, vals.size, mask.ctypes.data, mask.size)], b[rank], b[rank + 1])[0]
            valid = ((mask[np.arange(rows) >> 5] >> (np.arange(rows) & 31).astype(np.uint32)) & 1).astype(bool) if st["has_validity"] else np.ones(rows, dtype=bool)
            parts = [None] * world
            dist.all_gather_object(parts, (vals[: rows * w].copy(), valid))
            res[name] = ("fixed", np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), w)
    r.close()
    np.save(os.path.join(outdir, f"r{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_row_group_split_matches_the_single_gpu_run(pq, files, tmp_path):
    if pq.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    for fname in [k for k in ("strings", "fixed_dict", "golden_mixed") if k in files]:
        path = files[fname]
        out = tmp_path / fname
        out.mkdir()
        mp.spawn(_worker, args=(world, _free_port(), path, str(out)), nprocs=world, join=True)
        got = [np.load(str(out / f"r{r}.npy"), allow_pickle=True)[0] for r in range(world)]
        r = pq.Reader(path, device=0)
        try:
            assert r.num_row_groups >= 2
            for c in range(r.num_columns):
                ci = r.column_info(c)
                name = ci["name"]
                if name not in got[0]:
                    continue
                if ci["type"] == pq.BYTE_ARRAY:
                    bits, _ = r.regex_prune(c, PATTERN, False)
                    nbits, _ = r.regex_prune(c, "Berlin|user1", True)
                    t2c, total = r.chunk_index(name, 512)
                    for rank in range(world):
                        _, gb, gn, gt, gtot, per = got[rank][name]
                        assert np.array_equal(gb, bits) and np.array_equal(gn, nbits), (fname, name, rank)
                        assert gtot == total and np.array_equal(gt, t2c), (fname, name, rank, gtot, total)
                        assert sum(per) == len(bits)
                else:
                    whole = r.read_column(name)
                    valid = ~whole["is_null"].astype(bool)
                    for rank in range(world):
                        _, gv, gvalid, w = got[rank][name]
                        n = len(valid)
                        pad = np.zeros((n, 8), dtype=np.uint8)
                        pad[:, :w] = gv.reshape(n, w)
                        assert np.array_equal(gvalid, valid), (fname, name, rank)
                        assert np.array_equal(pad.view(np.uint64).reshape(n)[valid], whole["fixed"][valid]), (fname, name, rank)
        finally:
            r.close()
