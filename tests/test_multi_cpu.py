"""N > 1 host logic under gloo, world_size 2, on CPU: shard boundaries, the gather of per-shard
page bitmaps and the chunk-index carry chain (duckdb-parquet-parser_b200/multi_gpu.py).  The
per-rank operations (which are GPU kernels in the product) are stood in for by an
oracle-backed shim with the same signatures -- this test is about the exchange, not the
kernels; the kernels' shard semantics (carry_in / id_base) are covered on the GPU in
tests/test_gpu_scan.py::test_chunk_index_shards_stitch_like_one_run."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "mixed.parquet")


class OracleOps:
    """per-rank ops with the signatures of pqb200.Reader, computed by the oracle"""

    def __init__(self, path):
        import oraclelib
        import pqb200
        oraclelib.build_oracle()
        self.o = oraclelib.Oracle()
        self.h = self.o.open(path)
        self.path = path
        self.rd = None
        self.pq = pqb200

    def _reader_meta(self):
        # footer-only metadata through the product's host parser (no GPU involved)
        if self.rd is None:
            self.rd = self.pq.Reader(self.path)
        return self.rd

    def find_column(self, name):
        return self.o.find_column(self.h, name)

    def shard_row_groups(self, col, n):
        return self._reader_meta().shard_row_groups(col, n)

    def _pages_per_rg(self, col):
        out = []
        for rg in range(self.o.num_row_groups(self.h)):
            pg = self.o.read_pages(self.h, rg, col)
            out.append(int((pg.page_type == 0).sum()))
        return out

    def regex_prune_rgs(self, col, rg0, rg1, pattern, neg=False):
        bits = self.o.regex_prune(self.h, col, pattern, neg)
        per = self._pages_per_rg(col)
        a, b = sum(per[:rg0]), sum(per[:rg1])
        return bits[a:b], 0.0

    # the phased chunk index (multi_gpu.chunk_index_sharded): prepare / stitch / emit
    def chunk_index_prepare_rgs(self, name, rg0, rg1, chunk_size=4096):
        pos, off, _ = self.o.string_iterator(self.h, name)
        rows = [self.o.row_group_num_rows(self.h, rg) for rg in range(self.o.num_row_groups(self.h))]
        r0, r1 = sum(rows[:rg0]), sum(rows[:rg1])
        lens = np.diff(off.astype(np.int64))
        keep = (pos >= r0) & (pos < r1)
        return dict(pos=pos[keep] - r0, lens=lens[keep], rows=r1 - r0, chunk_size=chunk_size)

    def chunk_index_stitch(self, job, carry_in):
        local = np.zeros(len(job["pos"]), dtype=np.uint32)
        bytes_, cid = carry_in, 0
        for k, ln in enumerate(job["lens"]):  # the loop of reference src/main.cpp:21-32 on the shard's values
            if bytes_ >= job["chunk_size"]:
                bytes_, cid = 0, cid + 1
            bytes_ += len(str(int(ln))) + int(ln)
            local[k] = cid
        job["local"] = local
        return cid + 1, bytes_

    def chunk_index_emit(self, job, id_base):
        ids = np.zeros(job["rows"], dtype=np.uint32)
        ids[job["pos"]] = id_base + job["local"]
        return ids


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
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mg = importlib.import_module("duckdb-parquet-parser_b200.multi_gpu")
    ops = OracleOps(path)
    res = {}
    for name in ("city", "email"):
        col = ops.find_column(name)
        bits, per = mg.regex_prune_sharded(ops, col, r"^[a-z0-9._]+@[a-z0-9.]+\.com$", False, rank, world)
        nbits, _ = mg.regex_prune_sharded(ops, col, "Berlin|user1", True, rank, world)
        t2c, total = mg.chunk_index_sharded(ops, name, 512, rank, world, col=col)
        res[name] = (bits, nbits, t2c, total, per)
    res["max"] = mg.gather_max(float(rank + 1), world)
    np.save(os.path.join(outdir, f"r{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gather_matches_single_rank(tmp_path, pq, oracle):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, GOLD, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(str(tmp_path / f"r{r}.npy"), allow_pickle=True)[0] for r in range(world)]
    h = oracle.open(GOLD)
    try:
        for name in ("city", "email"):
            col = oracle.find_column(h, name)
            exp_bits = oracle.regex_prune(h, col, r"^[a-z0-9._]+@[a-z0-9.]+\.com$", False)
            exp_nbits = oracle.regex_prune(h, col, "Berlin|user1", True)
            exp_t2c, exp_total = oracle.chunk_index(h, name, 512)
            for r in range(world):
                bits, nbits, t2c, total, per = got[r][name]
                assert np.array_equal(bits, exp_bits) and np.array_equal(nbits, exp_nbits), (name, r)
                assert total == exp_total and np.array_equal(t2c, exp_t2c), (name, r, total, exp_total)
                assert sum(per) == len(exp_bits) and len(per) == world
        assert got[0]["max"] == got[1]["max"] == 2.0
    finally:
        oracle.close(h)


def test_shard_bounds_cover_all_row_groups(pq):
    r = pq.Reader(GOLD)
    nrg = r.num_row_groups
    for n in (1, 2, 3, 8):
        for col in (-1, 0, 1):
            b = r.shard_row_groups(col, n)
            assert len(b) == n + 1 and b[0] == 0 and b[-1] == nrg
            assert all(b[i] <= b[i + 1] for i in range(n))
    with pytest.raises(pq.PqgError):
        r.shard_row_groups(0, 0)
    r.close()
