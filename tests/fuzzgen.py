"""Deterministic fuzz-case generators for the differential tests (SURVEY.md Appendix A.8 domain)."""
import numpy as np

LENS = [1, 5, 16, 17, 33, 64, 100, 150, 257]
AB = [(2, 4), (1, 2), (1, 4), (2, 8)]
QE = [(4, 2), (2, 1), (6, 1), (12, 2)]
DUAL = [(4, 2, 24, 1), (4, 2, 13, 1), (2, 1, 32, 0), (6, 1, 6, 1), (24, 1, 4, 2), (12, 2, 40, 1), (5, 3, 30, 1)]
WS = [-1, 3, 7, 10, 15, 16, 17, 31, 40, 100]
ZD = [-1, 10, 30, 100, 400]
EB = [0, 5, 20]
FLAGS = [0, 1, 2, 4, 5, 6, 8, 0x18, 0x19, 0x40, 0x41, 0x42, 0x44, 0x80, 0x82, 0xc0, 0xc6, 0x09, 0x48, 0x58]
SPL = [(2, 1, 32, 4), (2, 1, 32, 9), (4, 2, 24, 5), (2, 1, 10, 0), (3, 1, 3, 4), (2, 2, 17, 7)]
SFLAGS = [0, 0x100, 0x200, 0x300, 0x500, 0x180, 0x580, 0x101, 0x102, 0x104, 0x140, 0x1c0, 0x118, 0x119, 0x282, 0x400, 0x108]


def rand_pair(rng, tl=None, div=None):
    tl = int(rng.choice(LENS)) if tl is None else tl
    div = rng.uniform(0.02, 0.30) if div is None else div
    t = rng.integers(0, 4, tl).astype(np.uint8)
    q = []
    for b in t:
        x = rng.random()
        if x < div * 0.25:
            continue
        if x < div * 0.5:
            q.extend(rng.integers(0, 4, int(rng.integers(1, 4))).tolist())
        if x > 1 - div * 0.5:
            q.append((int(b) + int(rng.integers(1, 4))) & 3)
        else:
            q.append(int(b))
    q = np.asarray(q, dtype=np.uint8)
    mode = rng.integers(0, 6)
    if mode == 0 and len(q) > 4:      # truncate tail
        q = q[: max(1, len(q) - int(rng.integers(1, max(2, len(q) // 2))))]
    elif mode == 1:                   # random tail (forces z-drop)
        k = int(rng.integers(1, max(2, len(q) // 2 + 1)))
        q = np.concatenate([q[: max(0, len(q) - k)], rng.integers(0, 4, k).astype(np.uint8)])
    elif mode == 2:                   # extra tail
        q = np.concatenate([q, rng.integers(0, 4, int(rng.integers(1, 40))).astype(np.uint8)])
    if len(q) == 0:
        q = rng.integers(0, 4, 1).astype(np.uint8)
    if rng.random() < 0.4:            # N runs
        for s in (q, t):
            if len(s) > 2 and rng.random() < 0.7:
                p = int(rng.integers(0, len(s)))
                s[p: p + int(rng.integers(1, 6))] = 4
    return q, t


def splice_pair(rng):
    ql = int(rng.integers(5, 151))
    q = rng.integers(0, 4, ql).astype(np.uint8)
    ins_len = int(rng.choice([0, 8, 20, 45, 90]))
    t = q.copy()
    if ins_len:
        pos = int(rng.integers(1, max(2, ql - 1)))
        kind = int(rng.integers(0, 4))
        body = rng.integers(0, 4, ins_len).astype(np.uint8)
        if kind == 0:
            body[:2] = [2, 3]; body[-2:] = [0, 2]                 # GT..AG
        elif kind == 1 and ins_len >= 6:
            body[:3] = [2, 3, 0]; body[-3:] = [1, 0, 2]           # GTA..CAG
        elif kind == 2:
            body[:2] = [1, 3]; body[-2:] = [0, 1]                 # CT..AC
        t = np.concatenate([q[:pos], body, q[pos:]])
    div = rng.uniform(0, 0.10)
    mask = rng.random(len(t)) < div
    t = t.copy()
    t[mask] = (t[mask] + rng.integers(1, 4, int(mask.sum()))) & 3
    if rng.random() < 0.3 and len(t) > 3:
        p = int(rng.integers(0, len(t)))
        t[p: p + int(rng.integers(1, 4))] = 4
    if rng.random() < 0.2:
        q = q.copy(); p = int(rng.integers(0, len(q))); q[p: p + 2] = 4
    junc = None
    if rng.random() < 0.5:
        junc = ((np.arange(len(t)) * 7 + int(rng.integers(0, 16))) % 23 < 3).astype(np.uint8) * np.uint8(rng.integers(1, 16))
    return q, t, junc


RFLAGS = [0, 1, 2, 0x40, 0x41, 0x42, 0x80, 0x82, 0xc0, 0xc2]


def rows_band(rng, prs, flag):
    """Band for the row-wise entry points (ksw_extz / ksw_extd) inside the reference's DEFINED domain: with a CIGAR the
    traceback start (tlen-1, qlen-1) must lie in the band (else ksw_backtrack reads unwritten heap, ksw2.h:143);
    score-only only needs every row's band to start at or before the query end (else ksw2_extz.c:113 writes past eh[])."""
    w = int(rng.choice(WS))
    if w < 0:
        return w
    if flag & 1:
        need = max(max(0, len(t) - 1 - len(q)) for q, t in prs)
    else:
        need = max(abs(len(t) - len(q)) for q, t in prs)
    return max(w, need)


def rows_batches(seed, n_iter, npairs=4):
    """(kind, mat, params dict, queries, targets) for ksw_extz / ksw_extd inside the reference's defined domain"""
    import harness as H
    rng = np.random.default_rng(seed)
    for it in range(n_iter):
        kind = ["extz", "extd"][it % 2]
        a, b = AB[rng.integers(len(AB))]
        prs = [rand_pair(rng) for _ in range(npairs)]
        mat = H.simple_mat(5, a, b, 0 if rng.random() < 0.7 else -1)
        if rng.random() < 0.2:                       # the row-wise kernels read the whole matrix
            mat = rng.integers(-6, 4, 25).astype(np.int8)
        fl = int(rng.choice(RFLAGS))
        kw = dict(w=rows_band(rng, prs, fl), zdrop=int(rng.choice(ZD)), flag=fl)
        if kind == "extz":
            kw["q"], kw["e"] = QE[rng.integers(len(QE))]
        else:
            kw["q"], kw["e"], kw["q2"], kw["e2"] = DUAL[rng.integers(len(DUAL))]
        yield kind, mat, kw, [p[0] for p in prs], [p[1] for p in prs]


def extf2_batches(seed, n_iter, npairs=4):
    """(params dict, queries, targets) for ksw_extf2_sse: q = mch, q2 = mis (either sign, the reference takes -|mis|), zdrop = xdrop"""
    rng = np.random.default_rng(seed)
    for it in range(n_iter):
        prs = [rand_pair(rng) for _ in range(npairs)]
        kw = dict(q=int(rng.choice([1, 2, 3])), q2=int(rng.choice([-1, -2, -4, 3, 6])), e=int(rng.choice([1, 2, 3, 5])),
                  w=int(rng.choice(WS)), zdrop=int(rng.choice([-1, 5, 20, 50, 200])), flag=1)
        yield kw, [p[0] for p in prs], [p[1] for p in prs]


def gg_batches(seed, n_iter, kinds=("gg",), npairs=4):
    """(kind, mat, params, queries, targets) for the global-alignment entry points; flag 1 = no CIGAR pointers.  The band always
    contains the end cell (the reference's traceback reads unwritten memory otherwise)."""
    import harness as H
    rng = np.random.default_rng(seed)
    for it in range(n_iter):
        kind = kinds[it % len(kinds)]
        prs = [rand_pair(rng) for _ in range(npairs)]
        a, b = AB[rng.integers(len(AB))]
        q, e = QE[rng.integers(len(QE))]
        mat = H.simple_mat(5, a, b, 0 if rng.random() < 0.7 else -1)
        w = int(rng.choice(WS))
        if w >= 0:
            w = max(w, max(abs(len(t) - len(qq)) for qq, t in prs))
        yield kind, mat, dict(q=q, e=e, w=w, flag=int(rng.choice([0, 0, 1]))), [p[0] for p in prs], [p[1] for p in prs]
