"""Multi-GPU plumbing: pairs are independent, so rank i of N aligns its own shard of the batch on its own GPU
(no collective on the data path).  Shards are contiguous for equal-length batches and COST-BALANCED for mixed lengths
(cost = band cells, + qlen + tlen with a CIGAR; largest first, dealt in serpentine order so every GPU also gets
the same mix of lengths).  The only optional exchange is an all-gather of the fixed-size 64-byte result
records so that every rank ends with the whole batch's results (BASELINE north_star: "an all-gather only to
collate results").  `align_fn` is the per-rank aligner -- ksw2_b200.Context.align_packed on a GPU box; the CPU
tests inject a checker so the sharding / collation logic is covered under gloo without a GPU."""
import numpy as np

from . import RESULT_DTYPE


def shard_bounds(n, world):
    return [n * i // world for i in range(world + 1)]


def pair_cost(qoff, toff, w, cigar):
    """work estimate per pair: cells of the band (SURVEY 8d geometry, closed form) + the traceback's share"""
    ql = np.diff(np.asarray(qoff, dtype=np.int64)); tl = np.diff(np.asarray(toff, dtype=np.int64))
    short = np.minimum(ql, tl)
    if w is None:
        band = short
    else:                                                                             # scalar band or one band per pair; < 0: none
        wv = np.broadcast_to(np.asarray(w, dtype=np.int64), short.shape)
        band = np.where(wv < 0, short, np.minimum(short, 2 * wv + 1))                 # lanes per diagonal once the band is full
    return band * (ql + tl) // 2 + 1 + (ql + tl if cigar else 0)


def balanced_shards(qoff, toff, w, world, cigar=False):
    """index arrays, one per rank: pairs sorted by cost (descending) and dealt 0,1,..,N-1,N-1,..,1,0,0,1,.. (serpentine)"""
    cost = pair_cost(qoff, toff, w, cigar)
    order = np.argsort(-cost, kind="stable")
    pos = np.arange(len(order)) % (2 * world)
    owner = np.where(pos < world, pos, 2 * world - 1 - pos)
    return [np.sort(order[owner == r]) for r in range(world)]


def gather_pairs(qcat, qoff, tcat, toff, idx, jcat=None):
    """the sub-batch made of pairs idx (any order): (qcat, qoff, tcat, toff, jcat) with fresh offsets"""
    idx = np.asarray(idx, dtype=np.int64)
    def take(cat, off):
        lens = (off[idx + 1] - off[idx]).astype(np.int64)
        new = np.zeros(len(idx) + 1, dtype=np.int64); np.cumsum(lens, out=new[1:])
        if new[-1] == 0:
            return np.zeros(1, np.uint8), new
        src = np.repeat(off[idx] - new[:-1], lens) + np.arange(new[-1], dtype=np.int64)
        return np.ascontiguousarray(cat[src]), new
    qs, qo = take(qcat, np.asarray(qoff, dtype=np.int64))
    ts, to = take(tcat, np.asarray(toff, dtype=np.int64))
    js = None if jcat is None else take(jcat, np.asarray(toff, dtype=np.int64))[0]
    return qs, qo, ts, to, js


def shard_slice(qcat, qoff, tcat, toff, lo, hi, jcat=None):
    q0, q1, t0, t1 = int(qoff[lo]), int(qoff[hi]), int(toff[lo]), int(toff[hi])
    qs = qcat[q0:q1] if q1 > q0 else np.zeros(1, np.uint8)
    ts = tcat[t0:t1] if t1 > t0 else np.zeros(1, np.uint8)
    js = None if jcat is None else (jcat[t0:t1] if t1 > t0 else np.zeros(1, np.uint8))
    return (np.ascontiguousarray(qs), np.ascontiguousarray(qoff[lo:hi + 1] - q0), np.ascontiguousarray(ts),
            np.ascontiguousarray(toff[lo:hi + 1] - t0), js)


def align_balanced(align_fn, P, qcat, qoff, tcat, toff, rank, world, w=-1, cigar=False, gather=True, device=None, jcat=None):
    """Cost-balanced variant of align_sharded for mixed-length batches: rank r aligns balanced_shards(...)[r]; with gather=True every
    rank returns all n records in the caller's order.  Returns (records, local_cigars, local_indices)."""
    n = len(qoff) - 1
    shards = balanced_shards(qoff, toff, w, world, cigar)
    idx = shards[rank]
    qs, qo, ts, to, js = gather_pairs(qcat, qoff, tcat, toff, idx, jcat)
    res, cigs = align_fn(P, qs, qo, ts, to, js) if len(idx) else (np.zeros(0, RESULT_DTYPE), [])
    if not gather or world == 1:
        out = np.zeros(n, dtype=RESULT_DTYPE) if world == 1 else res
        if world == 1:
            out[idx] = res
        return out, cigs, idx
    import torch
    import torch.distributed as dist
    per = max(1, max(len(x) for x in shards))
    buf = np.zeros((per, RESULT_DTYPE.itemsize // 4), dtype=np.int32)
    buf[: len(idx)] = res.view(np.int32).reshape(len(idx), RESULT_DTYPE.itemsize // 4)    # (an empty shard must still reach the all_gather)
    mine = torch.from_numpy(buf)
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.zeros(n, dtype=RESULT_DTYPE)
    for i, p in enumerate(parts):
        k = len(shards[i])
        out[shards[i]] = np.ascontiguousarray(p.cpu().numpy()[:k]).view(RESULT_DTYPE).reshape(k)
    return out, cigs, idx


def align_sharded(align_fn, P, qcat, qoff, tcat, toff, rank, world, gather=True, device=None, jcat=None):
    """Each rank aligns pairs [bounds[rank], bounds[rank+1]); with gather=True every rank returns all n records
    (CIGARs stay local to the rank that produced them: (records, local_cigars, (lo, hi)))."""
    n = len(qoff) - 1
    b = shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    qs, qo, ts, to, js = shard_slice(qcat, qoff, tcat, toff, lo, hi, jcat)
    res, cigs = align_fn(P, qs, qo, ts, to, js) if hi > lo else (np.zeros(0, RESULT_DTYPE), [])
    if not gather or world == 1:
        return res, cigs, (lo, hi)
    import torch
    import torch.distributed as dist
    per = max(1, max(b[i + 1] - b[i] for i in range(world)))
    buf = np.zeros((per, RESULT_DTYPE.itemsize // 4), dtype=np.int32)
    buf[: hi - lo] = res.view(np.int32).reshape(hi - lo, RESULT_DTYPE.itemsize // 4)
    mine = torch.from_numpy(buf)
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.zeros(n, dtype=RESULT_DTYPE)
    for i, p in enumerate(parts):
        k = b[i + 1] - b[i]
        out[b[i]: b[i + 1]] = np.ascontiguousarray(p.cpu().numpy()[:k]).view(RESULT_DTYPE).reshape(k)
    return out, cigs, (lo, hi)
