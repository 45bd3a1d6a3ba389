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


def regex_prune_sharded(ops, col, pattern, neg=False, rank=0, world=1):
    """-> (page bits of the whole column on every rank, pages per rank)"""
    b = shard_bounds(ops, col, world)
    local, _ = ops.regex_prune_rgs(col, b[rank], b[rank + 1], pattern, neg)
    local = np.ascontiguousarray(local, dtype=np.uint8)
    if world == 1:
        return local, [len(local)]
    parts = [None] * world
    _dist().all_gather_object(parts, local)
    return np.concatenate(parts), [len(p) for p in parts]


def chunk_index_sharded(ops, name, chunk_size=4096, rank=0, world=1, col=None):
    """-> (tuple_to_chunk of the whole column on every rank, total chunks).
    Every rank decodes its shard first (that is the heavy, parallel part and happens inside
    chunk_index_rgs before the chain needs the carry only in its last step); the carry then
    travels rank 0 -> 1 -> ... as two integers."""
    dist = _dist() if world > 1 else None
    if col is None:
        col = ops.find_column(name)
    b = shard_bounds(ops, col, world)
    carry, base = 0, 0
    if world > 1 and rank > 0:
        msg = [None]
        dist.recv_object_list(msg, src=rank - 1)
        carry, base = msg[0]
    ids, n, carry_out = ops.chunk_index_rgs(name, b[rank], b[rank + 1], chunk_size, carry, base)
    if world > 1 and rank + 1 < world:
        dist.send_object_list([(int(carry_out), int(base + n - 1))], dst=rank + 1)
    out = ids.astype(np.uint64)  # id_base already applied to the non-null rows on the device
    if world == 1:
        return out, int(n)
    parts = [None] * world
    dist.all_gather_object(parts, (out, int(base + n)))
    total = parts[-1][1]
    return np.concatenate([p[0] for p in parts]), int(total)


def gather_max(value, world):
    """max over ranks of a host float (timings are taken on the device per rank)"""
    if world == 1:
        return value
    vals = [None] * world
    _dist().all_gather_object(vals, float(value))
    return max(vals)
