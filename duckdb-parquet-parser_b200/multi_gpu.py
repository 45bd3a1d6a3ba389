"""Multi-GPU host logic: one process per GPU (torch.distributed is only the plumbing), every
rank owns a contiguous run of row groups (SURVEY.md section 8 e).  Row groups share nothing,
so there is NO collective on the data path -- only a host gather of
  * per-shard page bitmaps (concatenated in rank order = global page order of the column), and
  * the chunk-index chain: rank r needs the bytes left in the open chunk by rank r-1
    (`carry`), and the number of chunks closed before it (`base`).
`ops` is the per-rank reader (pqb200.Reader: GPU kernels through the C-ABI).  The CPU tests run
the same functions under gloo with an oracle-backed `ops` to cover the N > 1 logic.
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def shard_bounds(ops, col, world):
    """n+1 row-group boundaries; identical on every rank (derived from the footer only)"""
    return ops.shard_row_groups(col, world)


def gather_page_bits(local_bits, world):
    """host gather of the per-shard page bitmaps (one byte per page): concatenation in rank order = global page order of
    the column.  -> (bits of the whole column on every rank, pages per rank)"""
    local = np.ascontiguousarray(local_bits, dtype=np.uint8)
    if world == 1:
        return local, [len(local)]
    parts = [None] * world
    _dist().all_gather_object(parts, (np.packbits(local), len(local)))  # one bit per page on the wire
    return np.concatenate([np.unpackbits(b)[:n] for b, n in parts]), [n for _, n in parts]


def regex_prune_sharded(ops, col, pattern, neg=False, rank=0, world=1):
    """-> (page bits of the whole column on every rank, pages per rank)"""
    b = shard_bounds(ops, col, world)
    local, _ = ops.regex_prune_rgs(col, b[rank], b[rank + 1], pattern, neg)
    return gather_page_bits(local, world)


def chain_carry(stitch, rank=0, world=1):
    """The only ordered step of the sharded chunk index: rank r waits for (carry, id_base) of rank r-1, stitches its own
    prepared shard (`stitch(carry_in) -> (n_chunks, carry_out)`: a host loop, microseconds) and passes the pair on.
    -> (id_base of this rank, n_chunks of this rank)"""
    dist = _dist() if world > 1 else None
    carry, base = 0, 0
    if world > 1 and rank > 0:
        msg = [None]
        dist.recv_object_list(msg, src=rank - 1)
        carry, base = msg[0]
    n, carry_out = stitch(carry)
    if world > 1 and rank + 1 < world:
        dist.send_object_list([(int(carry_out), int(base + n - 1))], dst=rank + 1)
    return int(base), int(n)


def chunk_index_sharded(ops, name, chunk_size=4096, rank=0, world=1, col=None, gather=True):
    """-> (tuple_to_chunk of the whole column on every rank, total chunks).
    Three phases (include/pqg.h: pqg_chunk_index_prepare / _stitch / _emit): every rank uploads, decodes and prepares its
    shard at the same time; the carry then travels rank 0 -> 1 -> ... as two integers, each rank spending only a host
    stitch on it; the ids are materialised on all ranks at once again and gathered on the host."""
    if col is None:
        col = ops.find_column(name)
    b = shard_bounds(ops, col, world)
    job = ops.chunk_index_prepare_rgs(name, b[rank], b[rank + 1], chunk_size)
    base, n = chain_carry(lambda carry: ops.chunk_index_stitch(job, carry), rank, world)
    ids = ops.chunk_index_emit(job, base)
    out = ids.astype(np.uint64)  # id_base already applied to the non-null rows on the device
    if world == 1:
        return out, int(n)
    parts = [None] * world
    _dist().all_gather_object(parts, (out if gather else None, int(base + n)))
    total = parts[-1][1]
    if not gather:
        return out, int(total)
    return np.concatenate([p[0] for p in parts]), int(total)


def gather_max(value, world):
    """max over ranks of a host float (timings are taken on the device per rank)"""
    if world == 1:
        return value
    vals = [None] * world
    _dist().all_gather_object(vals, float(value))
    return max(vals)
