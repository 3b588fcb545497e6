"""Multi-GPU plumbing: pairs are independent, so rank i of N aligns a contiguous shard of the batch on its own GPU
(no collective on the data path).  The only optional exchange is an all-gather of the fixed-size 64-byte result
records so that every rank ends with the whole batch's results (BASELINE north_star: "an all-gather only to
collate results").  `align_fn` is the per-rank aligner -- ksw2_b200.Context.align_packed on a GPU box; the CPU
tests inject a checker so the sharding / collation logic is covered under gloo without a GPU."""
import numpy as np

from . import RESULT_DTYPE


def shard_bounds(n, world):
    return [n * i // world for i in range(world + 1)]


def shard_slice(qcat, qoff, tcat, toff, lo, hi, jcat=None):
    q0, q1, t0, t1 = int(qoff[lo]), int(qoff[hi]), int(toff[lo]), int(toff[hi])
    qs = qcat[q0:q1] if q1 > q0 else np.zeros(1, np.uint8)
    ts = tcat[t0:t1] if t1 > t0 else np.zeros(1, np.uint8)
    js = None if jcat is None else (jcat[t0:t1] if t1 > t0 else np.zeros(1, np.uint8))
    return (np.ascontiguousarray(qs), np.ascontiguousarray(qoff[lo:hi + 1] - q0), np.ascontiguousarray(ts),
            np.ascontiguousarray(toff[lo:hi + 1] - t0), js)


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
    per = max(b[i + 1] - b[i] for i in range(world))
    buf = np.zeros((per, RESULT_DTYPE.itemsize // 4), dtype=np.int32)
    buf[: hi - lo] = res.view(np.int32).reshape(hi - lo, -1)
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
