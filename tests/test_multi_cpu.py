"""N>1 host logic on CPU: world_size-2 gloo run of ksw2_b200.multi.align_sharded (sharding + all-gather collation of the
fixed-size result records).  The per-rank aligner is replaced by the oracle (test-only injection): what is under test
here is the plumbing, the GPU aligner itself is covered by test_gpu_parity.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import harness as H


def _oracle_align(P, qcat, qoff, tcat, toff, jcat=None):
    from ksw2_b200 import RESULT_DTYPE
    hp = H.make_params(P["kind"], H.simple_mat(5, 2, 4), **P["par"])
    res, cig, _ = H.run_cpu("oracle", hp, None, None, packed=(qcat, qoff, tcat, toff))
    out = np.zeros(len(qoff) - 1, dtype=RESULT_DTYPE)
    for name in ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end", "n_cigar"):
        out[name] = res[:, H.FIELDS.index(name)]
    return out, cig


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ksw2_b200.multi import align_sharded
    rng = np.random.default_rng(123)                       # same data on every rank
    qs, ts = [], []
    for i in range(37):
        L = int(rng.integers(20, 200)); t = rng.integers(0, 4, L).astype(np.uint8); qq = t.copy(); qq[rng.random(L) < 0.1] = 1
        qs.append(qq[: L - int(rng.integers(0, 5))]); ts.append(t)
    qcat, qoff = H.pack(qs); tcat, toff = H.pack(ts)
    P = dict(kind="extd2", par=dict(q=4, e=2, q2=24, e2=1, w=40, zdrop=100, flag=0))
    allres, cigs, (lo, hi) = align_sharded(_oracle_align, P, qcat, qoff, tcat, toff, rank, world)
    full, fcig = _oracle_align(P, qcat, qoff, tcat, toff)
    ok = bool(np.array_equal(allres, full)) and all(np.array_equal(a, b) for a, b in zip(cigs, fcig[lo:hi])) and (hi - lo) in (18, 19)
    # cost-balanced shards (mixed lengths): every rank ends with all records in the caller's order, its CIGARs stay local
    from ksw2_b200.multi import align_balanced
    bres, bcigs, idx = align_balanced(_oracle_align, P, qcat, qoff, tcat, toff, rank, world, w=40, cigar=True)
    ok = ok and bool(np.array_equal(bres, full)) and all(np.array_equal(a, fcig[i]) for a, i in zip(bcigs, idx)) and abs(len(idx) - 18.5) < 1
    q.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_alignment_with_gloo_allgather():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert got == [(0, True), (1, True)]


def test_shard_bounds_cover_everything():
    from ksw2_b200.multi import shard_bounds
    for n in (0, 1, 7, 1000):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))


def test_balanced_shards_are_balanced_and_complete():
    from ksw2_b200.multi import balanced_shards, pair_cost, gather_pairs
    rng = np.random.default_rng(4)
    L = np.exp(rng.uniform(np.log(150), np.log(20000), 5000)).astype(np.int64)          # BASELINE config 5: log-uniform 150 bp .. 20 kb
    qoff = np.concatenate([[0], np.cumsum(L)]); toff = np.concatenate([[0], np.cumsum(L + rng.integers(-10, 10, len(L)))])
    cost = pair_cost(qoff, toff, 500, True)
    for world in (1, 2, 3, 8):
        sh = balanced_shards(qoff, toff, 500, world, True)
        assert sorted(np.concatenate(sh).tolist()) == list(range(len(L)))
        loads = np.array([cost[x].sum() for x in sh], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.01
    cat = rng.integers(0, 4, int(qoff[-1])).astype(np.uint8)
    qs, qo, ts, to, _ = gather_pairs(cat, qoff, cat, qoff, np.array([5, 0, 17]))
    assert np.array_equal(qs[qo[1]:qo[2]], cat[qoff[0]:qoff[1]]) and np.array_equal(ts[:to[1]], cat[qoff[5]:qoff[6]]) and qo[-1] == L[[5, 0, 17]].sum()


def _worker_small(rank, world, port, q):
    """fewer pairs than ranks (and no pairs at all): every rank must still reach the all-gather"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ksw2_b200.multi import align_sharded, align_balanced
    t = np.array([0, 1, 2, 3, 0, 1, 2, 3], np.uint8)
    P = dict(kind="extz2", par=dict(q=4, e=2, w=-1, zdrop=-1, flag=0))
    ok = True
    for n in (1, 0):
        qcat, qoff = H.pack([t] * n); tcat, toff = H.pack([t] * n)
        a, _, _ = align_sharded(_oracle_align, P, qcat, qoff, tcat, toff, rank, world)
        b, _, _ = align_balanced(_oracle_align, P, qcat, qoff, tcat, toff, rank, world, w=-1, cigar=True)
        ok = ok and len(a) == n and len(b) == n and (n == 0 or (int(a["score"][0]) == 16 and int(b["score"][0]) == 16))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_fewer_pairs_than_ranks():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_small, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert got == [(0, True), (1, True)]


def test_pair_cost_with_a_band_per_pair():
    from ksw2_b200.multi import pair_cost, balanced_shards
    qoff = np.array([0, 100, 1100, 21100], np.int64)
    a = pair_cost(qoff, qoff, np.array([10, -1, 500], np.int32), False)
    assert a[0] == 21 * 100 + 1 and a[1] == 1000 * 1000 + 1 and a[2] == 1001 * 20000 + 1
    assert sorted(np.concatenate(balanced_shards(qoff, qoff, np.array([10, -1, 500]), 2)).tolist()) == [0, 1, 2]
