#!/usr/bin/env python
"""bench.py -- GCUPS of the ksw2 hot path on B200 (driver contract: see the task statement / DESIGN.md section "Measurement").

  python bench.py [--gpus N --steps K --warmup W] [--impl ours|reference] [--workload c1..c5] [--configs all|none|c1,c3,..] [--pairs P]

Headline (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE.json configs[1] -- 1 M x 150 bp pairs per GPU, ksw_extz2 extension
with Z-drop, w=100, score only; weak scaling (every rank aligns its own batch).  A "step" is one pass of the hot path over the
batch.  `value` = in-band DP cells of all ranks / device time (CUDA events, max over ranks) with the inputs resident in HBM;
`e2e` = the same metric through the C-ABI batch call ksw2b_align() with pinned HOST buffers (H2D + kernels + D2H inside the timed
region); `e2e.pageable` the same from plain malloc'ed buffers, `e2e.batch_api` through the array-of-pointers call ksw2b_extz2_batch.

`configs`: every other BASELINE.json configuration, each with its own value / e2e / roofline / parity flag:
  c1  test/MT-human.fa x test/MT-orang.fa (16.5 kb), extz2 global with CIGAR         one pair, rank 0 only
  c3  100 k x 5 kb ONT-like pairs per GPU, extd2 dual gap, w=500, zdrop=400, CIGAR      weak scaling
  c4  10 k x 50 kb pairs IN ALL, extz2 global, no band, score only                    strong scaling (contiguous shards)
  c5  200 k pairs IN ALL, 150 bp - 20 kb log-uniform, half extz2 / half extd2, band per pair min(500, ceil(0.2 len) + 50),
      zdrop=400, KSW_EZ_RIGHT, CIGAR                                                   strong scaling (cost-balanced shards)
For these, one pass through ksw2b_align_ex() with host buffers gives both numbers: `e2e` from the host clock around the call and
`value` from the device span of its kernels (CUDA events inside the library, first kernel to last kernel).
With N > 1, rank 0 additionally drives all N GPUs from ONE process through ksw2b_multi_align() (`c_api_multi`).

`--impl reference` times the CPU implementation (oracle/_ref = the unmodified reference if it was built, else the oracle port) on the
host cores, on bounded samples of the same workloads.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "GCUPS"
FLAG_RIGHT = 0x02
WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "c1": dict(name="MT-human x MT-orang (16.5 kb) extz2 global, no band, CIGAR", kinds=["extz2"], pairs=1, scaling="replica",
               par=dict(q=4, e=2, w=-1, zdrop=-1, end_bonus=0, flag=0), steps=3, warm=1),
    # BASELINE.json configs[1]
    "c2": dict(name="1M x 150bp extz2 extension, Z-drop, w=100, score-only", kinds=["extz2"], L=150, pairs=1_000_000, scaling="weak",
               par=dict(q=4, e=2, w=100, zdrop=100, end_bonus=0, flag=0x41), seed=20260925, steps=5, warm=3),
    # BASELINE.json configs[2]
    "c3": dict(name="100k x 5kb ONT-like extd2 dual-gap, w=500, zdrop=400, CIGAR", kinds=["extd2"], L=5000, pairs=100_000, scaling="weak",
               model=3, par=dict(q=4, e=2, q2=24, e2=1, w=500, zdrop=400, end_bonus=0, flag=0), seed=20260926, steps=2, warm=1),
    # BASELINE.json configs[3]: 10k pairs in all, sharded over the GPUs
    "c4": dict(name="10k x 50kb extz2 global, no band, score-only (one warp per pair)", kinds=["extz2"], L=50000, pairs=10_000, scaling="strong",
               model=4, par=dict(q=4, e=2, w=-1, zdrop=-1, end_bonus=0, flag=0x01), seed=20260927, steps=1, warm=0),
    # BASELINE.json configs[4]: 200k pairs in all, cost-balanced over the GPUs; even pairs extz2, odd pairs extd2
    "c5": dict(name="200k mixed 150bp-20kb, extz2+extd2, band per pair, zdrop=400, right-aligned CIGAR", kinds=["extz2", "extd2"], L=0, pairs=200_000,
               scaling="strong", model=5, par=dict(q=4, e=2, q2=24, e2=1, w=500, zdrop=400, end_bonus=0, flag=FLAG_RIGHT), seed=20260928, steps=2, warm=1),
}
NAMES = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end"]


# ----------------------------------------------------------------------------------------------------------
# synthetic data: C2 vectorised numpy (unchanged since round 1), the others tools/ksw2b_gen.c
# ----------------------------------------------------------------------------------------------------------
def _gen_c2_chunk(n, L, seed):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 4, (n, L), dtype=np.uint8)
    ev = rng.integers(0, 65536, (n, L), dtype=np.uint16)
    dele, ins = ev < 131, (ev >= 131) & (ev < 262)                       # 0.2 % each
    sub = (ev >= 262) & (ev < 917)                                       # 1 %
    shift = np.cumsum(dele.view(np.int8) - ins.view(np.int8), axis=1, dtype=np.int16)
    src = shift + np.arange(L, dtype=np.int16)[None, :]
    rnd = rng.integers(0, 4, (n, L), dtype=np.uint8)
    q = np.take_along_axis(t, np.clip(src, 0, L - 1).astype(np.intp), axis=1)
    oob = (src < 0) | (src >= L) | ins
    q[oob] = rnd[oob]
    q[sub] = (q[sub] + 1 + (rnd[sub] % 3)) & 3
    junk = rng.random(n) < 0.10
    k = rng.integers(40, 76, n)
    jm = junk[:, None] & (np.arange(L)[None, :] >= (L - k)[:, None])
    q[jm] = rnd[jm]
    q[(ev >= 917) & (ev < 1572)] = 4                                     # 1 % N
    return q.reshape(-1), t.reshape(-1)


def gen_c2(n, L, seed):
    """targets uniform ACGT; query = target with 1% sub, 0.2% ins, 0.2% del (re-padded to L); 10% of pairs get
    their last 40-75 query bases randomised (forces Z-drop); 1% of query bases -> N.  Chunked over threads."""
    from concurrent.futures import ThreadPoolExecutor
    nch = max(1, min(32, n // 20000))
    bounds = [n * i // nch for i in range(nch + 1)]
    with ThreadPoolExecutor(max_workers=min(nch, os.cpu_count() or 1)) as ex:
        parts = list(ex.map(lambda i: _gen_c2_chunk(bounds[i + 1] - bounds[i], L, seed * 1000 + i), range(nch)))
    q = np.concatenate([p[0] for p in parts]); t = np.concatenate([p[1] for p in parts])
    off = np.arange(n + 1, dtype=np.int64) * L
    return np.ascontiguousarray(q), off, np.ascontiguousarray(t), off.copy()


GEN_SRC = os.path.join(ROOT, "tools", "ksw2b_gen.c")
GEN_LIB = os.path.join(ROOT, "tools", "libksw2b_gen.so")
_gen = None


def build_gen(force=False):
    if force or not os.path.exists(GEN_LIB) or os.path.getmtime(GEN_LIB) < os.path.getmtime(GEN_SRC):
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-o", GEN_LIB, GEN_SRC, "-lpthread", "-lm"])
    return GEN_LIB


def genlib():
    global _gen
    if _gen is None:
        G = C.CDLL(build_gen())
        G.ksg_lengths.restype = None; G.ksg_lengths.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        G.ksg_sizes.restype = None; G.ksg_sizes.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        G.ksg_generate.restype = C.c_int64
        G.ksg_generate.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        G.ksg_cells.restype = None
        G.ksg_cells.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _gen = G
    return _gen


def gen_model(model, seed, L, idx=None, n=None, nthreads=None):
    """pairs idx (int64 array of global pair numbers; None: 0..n-1) of the model-3/4/5 workloads: (qcat, qoff, tcat, toff)"""
    G = genlib()
    if idx is not None:
        idx = np.ascontiguousarray(idx, dtype=np.int64); m = len(idx); ip = idx.ctypes.data
    else:
        m = int(n); ip = None
    tb, qb = C.c_int64(0), C.c_int64(0)
    G.ksg_sizes(model, seed, L, m, ip, C.byref(tb), C.byref(qb))
    tcat = np.empty(max(1, tb.value), np.uint8); qcat = np.empty(max(1, qb.value), np.uint8)
    toff = np.zeros(m + 1, np.int64); qoff = np.zeros(m + 1, np.int64)
    tot = G.ksg_generate(model, seed, L, m, ip, nthreads or (os.cpu_count() or 1), tcat.ctypes.data, toff.ctypes.data, qcat.ctypes.data, qoff.ctypes.data)
    if tot < 0:
        raise RuntimeError(f"ksg_generate failed ({tot})")
    return qcat[: max(1, tot)], qoff, tcat, toff


def gen_c3(n, L, seed):
    """n pairs of the config-3 model (ONT-like copies of L-base targets)"""
    return gen_model(3, seed, L, n=n)


def gen_c4(n, L, seed):
    """n pairs of the config-4 model (~90 %-identity copies of L-base targets)"""
    return gen_model(4, seed, L, n=n)


def model_lengths(model, seed, L, n):
    tl = np.zeros(n, np.int32)
    genlib().ksg_lengths(model, seed, L, n, None, tl.ctypes.data)
    return tl


def cells_lanes(qoff, toff, w, n_diag=None):
    """(in-band cells, padded direction bytes incl. 8 B of off/off_end per diagonal) per pair: SURVEY 8(d) conventions.
    w: scalar band or int32 array; n_diag bounds the diagonals (the reference stops at the Z-drop diagonal)."""
    ql = np.ascontiguousarray(np.diff(qoff), dtype=np.int32); tl = np.ascontiguousarray(np.diff(toff), dtype=np.int32)
    n = len(ql)
    wv = np.ascontiguousarray(np.broadcast_to(np.asarray(w, dtype=np.int32), (n,)))
    nd = np.ascontiguousarray(n_diag, dtype=np.int32) if n_diag is not None else None
    cells = np.zeros(n, np.int64); lanes = np.zeros(n, np.int64)
    genlib().ksg_cells(n, ql.ctypes.data, tl.ctypes.data, wv.ctypes.data, nd.ctypes.data if nd is not None else None, cells.ctypes.data, lanes.ctypes.data,
                       min(32, os.cpu_count() or 1))
    return cells, lanes


def band_of(tlen):
    """config 5: band per pair = min(500, ceil(0.2 len) + 50)"""
    return np.minimum(500, (np.asarray(tlen, dtype=np.int64) + 4) // 5 + 50).astype(np.int32)


def golden_c1():
    z = np.load(os.path.join(ROOT, "tests", "golden", "seqs.npz"))
    exp = [c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "expected.json"))) if c["name"] == "mt_extz2"][0]
    return np.ascontiguousarray(z["mt_q"]), np.ascontiguousarray(z["mt_t"]), exp


class Batch:
    """one ksw2b_align_ex call: pairs that share a parameter block"""
    def __init__(self, kind, par, qcat, qoff, tcat, toff, w=None, gidx=None):
        self.kind, self.par, self.qcat, self.qoff, self.tcat, self.toff, self.w, self.gidx = kind, par, qcat, qoff, tcat, toff, w, gidx
        self.n = len(qoff) - 1

    def band(self):
        return self.w if self.w is not None else self.par.get("w", -1)

    def subset(self, sel):
        """the pairs sel (local indices, ascending) as a new batch"""
        import ksw2_b200.multi as M
        qs, qo, ts, to, _ = M.gather_pairs(self.qcat, self.qoff, self.tcat, self.toff, sel)
        return Batch(self.kind, self.par, qs, qo, ts, to, None if self.w is None else np.ascontiguousarray(self.w[sel]), None)


def build_batches(wl, rank, world, pairs=0, sample=0):
    """the batches of workload wl that THIS rank aligns (sample > 0: a bounded sample of the whole workload instead, for the CPU arm)"""
    W = WORKLOADS[wl]
    n = pairs or W["pairs"]
    if wl == "c1":
        q, t, _ = golden_c1()
        return [Batch("extz2", W["par"], q, np.array([0, len(q)], np.int64), t, np.array([0, len(t)], np.int64))]
    if wl == "c2":
        m = sample or n
        qcat, qoff, tcat, toff = gen_c2(m, W["L"], W["seed"] + 1000 * (0 if sample else rank))
        return [Batch("extz2", W["par"], qcat, qoff, tcat, toff)]
    if wl == "c3":                                            # weak: rank r aligns pairs [r n, (r+1) n) of an endless stream
        m = sample or n
        idx = np.arange(m, dtype=np.int64) + (0 if sample else rank * n)
        qcat, qoff, tcat, toff = gen_model(3, W["seed"], W["L"], idx=idx)
        return [Batch("extd2", W["par"], qcat, qoff, tcat, toff, gidx=idx)]
    if wl == "c4":                                            # strong: contiguous shard of the n pairs
        if sample:
            idx = np.arange(sample, dtype=np.int64) * (n // sample)
        else:
            idx = np.arange(n * rank // world, n * (rank + 1) // world, dtype=np.int64)
        qcat, qoff, tcat, toff = gen_model(4, W["seed"], W["L"], idx=idx)
        return [Batch("extz2", W["par"], qcat, qoff, tcat, toff, gidx=idx)]
    if wl == "c5":                                            # strong: cost-balanced shard, per kind
        import ksw2_b200.multi as M
        tl = model_lengths(5, W["seed"], 0, n).astype(np.int64)
        wv = band_of(tl)
        out = []
        for k, kind in enumerate(W["kinds"]):
            gi = np.arange(k, n, 2, dtype=np.int64)           # even pairs extz2, odd pairs extd2
            if sample:
                gi = gi[:: max(1, len(gi) // max(1, sample // 2))][: max(1, sample // 2)]
            else:
                off = np.zeros(len(gi) + 1, np.int64); np.cumsum(tl[gi], out=off[1:])
                shards = M.balanced_shards(off, off, wv[gi], world, cigar=True)
                gi = gi[shards[rank]]
            qcat, qoff, tcat, toff = gen_model(5, W["seed"], 0, idx=gi)
            par = dict(W["par"]); par["w"] = -1
            if kind == "extz2":
                par = {k_: v for k_, v in par.items() if k_ not in ("q2", "e2")}
            out.append(Batch(kind, par, qcat, qoff, tcat, toff, w=np.ascontiguousarray(wv[gi]), gidx=gi))
        return out
    raise ValueError(wl)


# ----------------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.p:
            return None
        time.sleep(0.15)
        self.p.terminate()
        rows = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.15] or [r for _, r in self.rows[-3:]]
        if not rows:
            return None
        try:
            sm = sorted(float(r[0]) for r in rows)
            reasons = [n for i, n in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]) if any(r[3 + i].lower().startswith("active") for r in rows)]
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}
        except Exception:
            return None


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_run(batch, nthreads, want_cigar=False, which=None):
    """the CPU implementation on a batch: (seconds, fields, cigars, kind_str)"""
    import harness as H
    if which is None:
        which = "ref" if H.have_ref() else "oracle"
    P = H.make_params(batch.kind, H.simple_mat(5, 2, 4), **batch.par)
    res, cigs, secs = H.run_cpu(which, P, None, None, nthreads=nthreads, want_cigar=want_cigar, packed=(batch.qcat, batch.qoff, batch.tcat, batch.toff), w=batch.w)
    return secs, res, cigs, ("reference" if which == "ref" else "port")


def cpu_cells(batch, fields_res=None):
    """cells the CPU semantics executes on a batch: needs the diagonal where each pair stopped, which the oracle port reports"""
    import harness as H
    if batch.par.get("zdrop", -1) < 0:                      # no Z-drop: every diagonal is executed (the band may still end the sweep early)
        return int(cells_lanes(batch.qoff, batch.toff, batch.band())[0].sum())
    cl = np.zeros(batch.n, dtype=np.int64)
    P = H.make_params(batch.kind, H.simple_mat(5, 2, 4), **batch.par)
    H.run_cpu("oracle", P, None, None, nthreads=os.cpu_count() or 1, want_cigar=False, packed=(batch.qcat, batch.qoff, batch.tcat, batch.toff), cells_out=cl, w=batch.w)
    return int(cl.sum())


def cigar_text(c):
    return "".join(f"{int(x) >> 4}{'MIDN___=X'[int(x) & 0xf]}" for x in c)


def algorithmic_bytes(batch, res, score_only):
    """inputs at 1 B/base + one 56-B ksw_extz_t per pair; with a CIGAR also the reference's padded direction bytes,
    off/off_end (8 B per diagonal), one byte per traceback step and 4 B per CIGAR word"""
    b = float(batch.qoff[-1] + batch.toff[-1] + 56 * batch.n)
    if not score_only:
        _, lanes = cells_lanes(batch.qoff, batch.toff, batch.band())
        b += float(lanes.sum()) + float(res["n_cigar"].sum()) * 4 + float(batch.qoff[-1] + batch.toff[-1])
    return b


# ----------------------------------------------------------------------------------------------------------
# one configuration through the C-ABI batch call with host buffers (sub-configs: value from the device span inside the call)
# ----------------------------------------------------------------------------------------------------------
def run_config(wl, K, local, rank, world, dist_, gloo, steps=0, warm=-1, pairs=0, no_cpu=False, pinned=True):
    import torch
    import harness as H
    W = WORKLOADS[wl]
    steps = steps or W["steps"]; warm = W["warm"] if warm < 0 else warm
    ncores = os.cpu_count() or 1
    active = not (W["scaling"] == "replica" and rank != 0)
    t_gen = time.time()
    batches = build_batches(wl, rank, world, pairs) if active else []
    t_gen = time.time() - t_gen
    mat = H.simple_mat(5, 2, 4)
    ctx = K.Context(local)
    ctx.set_timing(True)
    keep = []
    for b in batches:                                           # the caller's buffers: pinned host memory
        if pinned:
            hq = torch.empty(len(b.qcat), dtype=torch.uint8, pin_memory=True); hq.numpy()[:] = b.qcat
            ht = torch.empty(len(b.tcat), dtype=torch.uint8, pin_memory=True); ht.numpy()[:] = b.tcat
            keep.append((hq, ht)); b.hq, b.ht = hq.numpy(), ht.numpy()
        else:
            b.hq, b.ht = b.qcat, b.tcat
        b.P = K.make_params(b.kind, mat, **b.par)

    def one_pass(collect=False):
        out, span, fill, nfill, launches, h2d, d2h = [], 0.0, 0.0, 0, 0, 0, 0
        t0 = time.perf_counter()
        for b in batches:
            r, cg = ctx.align_packed(b.P, b.hq, b.qoff, b.ht, b.toff, None, b.w, want_cigars=collect)
            f, nf, sp, nl = ctx.last_timing()
            x, y = ctx.last_transfer_bytes()
            span += sp; fill += f; nfill += nf; launches += nl; h2d += x; d2h += y
            out.append((r, cg))
        return time.perf_counter() - t0, span, fill, nfill, launches, h2d, d2h, out

    def barrier():
        if world > 1:
            dist_.barrier(group=gloo)

    if active and warm == 0 and batches:                         # at least load the kernels: a tiny pass
        tiny = [b.subset(np.arange(min(b.n, 64))) for b in batches]
        for tb in tiny:
            ctx.align_packed(K.make_params(tb.kind, mat, **tb.par), tb.qcat, tb.qoff, tb.tcat, tb.toff, None, tb.w, want_cigars=False)
    for _ in range(warm if active else 0):
        one_pass()
    barrier()
    wall = span = fill = 0.0
    nfill = launches = h2d = d2h = 0
    last = None
    for s in range(steps if active else 0):
        w_, sp, f, nf, nl, x, y, out = one_pass(collect=(s == steps - 1))
        wall += w_; span += sp; fill += f; nfill += nf; launches += nl; h2d, d2h = x, y
        last = out
    # executed cells (reference semantics: up to the diagonal where the pair stopped)
    cells = 0
    for b, (r, _) in zip(batches, last or []):
        c, _ = cells_lanes(b.qoff, b.toff, b.band(), r["n_diag"])
        cells += int(c.sum())
    if world > 1:
        tt = torch.tensor([wall, span * 1e-3, fill * 1e-3], device="cuda", dtype=torch.float64); dist_.all_reduce(tt, op=dist_.ReduceOp.MAX)
        wall, span, fill = float(tt[0]), float(tt[1]) * 1e3, float(tt[2]) * 1e3
        ct = torch.tensor([cells, sum(b.n for b in batches), h2d, d2h, launches], device="cuda", dtype=torch.int64); dist_.all_reduce(ct)
        cells_all, pairs_all, h2d, d2h, launches = (int(v) for v in ct.tolist())
    else:
        cells_all, pairs_all = cells, sum(b.n for b in batches)
    out = None
    parity, cpu = None, None
    # parity of the timed results against the CPU reference on a sample + the CPU baseline (rank 0)
    if rank == 0 and active and not no_cpu:
        ok, ccells, csecs, kind, ns_tot = True, 0, 0.0, "port", 0
        for b, (r, cg) in zip(batches, last):
            if wl == "c1":
                _, _, exp = golden_c1()
                ok = ok and all(int(r[k][0]) == exp["fields"][k] for k in NAMES + ["n_cigar"])
                ok = ok and hashlib.md5((cigar_text(cg[0]) + "\n").encode("latin1")).hexdigest() == exp["cigar_md5"]
                sel = np.arange(1)
            else:
                ns = {"c3": 192, "c4": ncores, "c5": 384}.get(wl, 1000)          # (c4: one 50 kb pair per host thread, ~2.5 s)
                ns = min(b.n, ns // len(batches))
                sel = np.unique(np.linspace(0, b.n - 1, ns).astype(np.int64))
            sb = b.subset(sel)
            with_cig = not (b.par["flag"] & 1)
            secs, cres, ccig, kind = cpu_run(sb, ncores, want_cigar=with_cig)
            ok = ok and all(np.array_equal(r[nm][sel], cres[:, H.FIELDS.index(nm)]) for nm in NAMES)
            if with_cig:
                ok = ok and np.array_equal(r["n_cigar"][sel], cres[:, H.FIELDS.index("n_cigar")])
                ok = ok and all(np.array_equal(cg[int(i)], c) for i, c in zip(sel, ccig))
            c, _ = cells_lanes(sb.qoff, sb.toff, sb.band(), r["n_diag"][sel])
            ccells += int(c.sum()); csecs += secs; ns_tot += len(sel)
        parity = bool(ok)
        cpu = {"value": ccells / max(csecs, 1e-9) / 1e9, "unit": "GCUPS", "cores": ncores, "kind": kind,
               "sample": f"{ns_tot} pairs spread over rank 0's share, one pass, {ncores} threads; all fields" + (" + every CIGAR word" if not (W['par']['flag'] & 1) else "") + f" bit-equal to GPU: {parity}"}
    if rank == 0:
        peak, how = measured_peak()
        score_only = bool(W["par"]["flag"] & 1)
        alg = sum(algorithmic_bytes(b, r, score_only) for b, (r, _) in zip(batches, last or []))
        alg_all = alg * (pairs_all / max(1, sum(b.n for b in batches)))          # ranks hold equal shares of the same distribution
        kern = "ks_fill_warp_kernel" if wl in ("c1", "c4") else "ks_fill_kernel"
        out = {"workload": W["name"], "pairs": pairs_all, "scaling": W["scaling"], "steps": steps, "warmup": warm,
               "value": cells_all * steps / max(span, 1e-9) / 1e6, "unit": "GCUPS",
               "value_from": "device span of the kernels inside the timed ksw2b_align_ex calls (CUDA events in the library, max over ranks)",
               "ms_per_step": span / max(1, steps),
               "e2e": {"value": cells_all * steps / max(wall, 1e-9) / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "host_buffers": "pinned" if pinned else "pageable"},
               "roofline": {"bound": "hbm", "achieved": alg_all * steps / max(fill, 1e-9) / 1e6, "peak": peak, "unit": "GB/s",
                            "frac": alg_all * steps / max(fill, 1e-9) / 1e6 / peak, "traffic": ncu_traffic(wl, pairs_all), "kernel": f"{kern}<{'+'.join(W['kinds'])}>",
                            "algorithmic_bytes_per_step": alg_all, "fill_ms_per_step": fill / max(1, steps), "fill_launches_per_step": nfill / max(1, steps)},
               "cells_per_step": cells_all, "gpu_launches": launches, "parity": parity, "cpu_baseline": cpu, "gen_s": round(t_gen, 2)}
    ctx.close()
    del keep
    return out


def ncu_traffic(wl, pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fill kernel from the committed ncu --set full capture of this round
    (profiles/r2_traffic.json: bytes per pair of the profiled launch), scaled to this launch; None when no capture exists"""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name))).get(wl)
            if tj:
                return float(tj["bytes_per_pair"]) * pairs
        except Exception:
            pass
    return None


# ----------------------------------------------------------------------------------------------------------
def reference_arm(a, W, wl, rank):
    """--impl reference: the CPU implementation on bounded samples (rank 0 only)"""
    if rank != 0:
        return 0
    ncores = os.cpu_count() or 1

    def sample_run(name, ns, steps, warm):
        bs = build_batches(name, 0, 1, 0, sample=ns) if name != "c1" else build_batches("c1", 0, 1)
        cells = sum(cpu_cells(b) for b in bs)
        times, kind = [], "port"
        for s in range(warm + steps):
            secs = 0.0
            for b in bs:
                t, _, _, kind = cpu_run(b, ncores, want_cigar=not (b.par["flag"] & 1))
                secs += t
            if s >= warm:
                times.append(secs)
        T = sum(times)
        return cells * len(times) / T / 1e9, 1e3 * T / max(1, len(times)), kind, sum(b.n for b in bs)

    ns = a.cpu_sample or {"c1": 1, "c2": 200_000, "c3": 256, "c4": ncores, "c5": 1024}[wl]
    val, ms, kind, npairs = sample_run(wl, ns, a.steps, a.warmup)
    cfg = {"workload": W["name"], "pairs_per_step": npairs, "kinds": W["kinds"], **W["par"], "scoring": "a=2 b=4 N=0"}
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GCUPS", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
           "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": ncores, "kind": kind, "sample": f"{npairs} pairs of the workload per step, all host threads"},
           "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    if a.configs != "none" and not a.workload:
        cf = {}
        for name in (["c1", "c3", "c4", "c5"] if a.configs == "all" else a.configs.split(",")):
            n2 = {"c1": 1, "c3": 256, "c4": ncores, "c5": 1024}[name]
            v, m_, k, npn = sample_run(name, n2, 1, 0)
            cf[name] = {"workload": WORKLOADS[name]["name"], "value": v, "unit": "GCUPS", "ms_per_step": m_, "cores": ncores, "kind": k, "sample": f"{npn} pairs, one pass, all host threads"}
        out["configs"] = cf
    print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="", choices=[""] + list(WORKLOADS), help="make this workload the headline (default: c2 + all configs)")
    ap.add_argument("--configs", default="all", help="all | none | comma list of c1,c3,c4,c5: the extra configurations reported next to the headline")
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample")
    ap.add_argument("--panel", type=int, default=0); ap.add_argument("--threads", type=int, default=0); ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--approx", action="store_true", help="add KSW_EZ_APPROX_MAX (0x08) to the headline workload's flag: the reference's fast mode (README -g)")
    a = ap.parse_args()
    if a.approx:
        for w_ in WORKLOADS.values():
            w_["par"] = dict(w_["par"], flag=w_["par"]["flag"] | 0x08); w_["name"] += " + KSW_EZ_APPROX_MAX"
        a.configs = "none" 
    wl = a.workload or "c2"
    W = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return reference_arm(a, W, wl, rank)
    warm = max(a.warmup, 3)
    import harness as H
    import torch
    import ksw2_b200 as K
    L = K.lib()                               # raises if the CUDA extension is missing: no fallback
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device")
    torch.cuda.set_device(local)
    dist = gloo = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")
    mat = H.simple_mat(5, 2, 4)
    ncores = os.cpu_count() or 1

    if wl != "c2":                            # an explicit non-default headline: that configuration alone, same JSON shape
        c = run_config(wl, K, local, rank, world, dist, gloo, steps=a.steps if a.steps != 5 else 0, pairs=a.pairs, no_cpu=a.no_cpu)
        if rank == 0:
            out = {"metric": METRIC, "value": c["value"], "unit": "GCUPS", "n_gpus": world, "steps": c["steps"], "warmup": c["warmup"], "ms_per_step": c["ms_per_step"],
                   "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "int8", "data": "synthetic",
                   "config": {"workload": c["workload"], "pairs": c["pairs"], **W["par"], "scoring": "a=2 b=4 N=0"},
                   "roofline": c["roofline"], "cpu_baseline": c["cpu_baseline"], "e2e": c["e2e"], "gpu_launches": c["gpu_launches"], "parity_sample_ok": c["parity"],
                   "cells_per_step": c["cells_per_step"], "value_from": c["value_from"]}
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ headline: C2, weak scaling
    n = a.pairs or W["pairs"]
    cfg = {"workload": W["name"], "pairs_per_gpu": n, "kind": "extz2", **W["par"], "scoring": "a=2 b=4 N=0",
           "l2_policy": "inputs larger than L2 (no flush)" if n * W["L"] * 2 > 130e6 else "inputs fit L2 (small run)"}
    b = build_batches("c2", rank, world, n)[0]
    qcat, qoff, tcat, toff = b.qcat, b.qoff, b.tcat, b.toff
    # pinned host staging (the caller's buffers of the e2e path) + resident device copies (the `value` path)
    hq = torch.empty(len(qcat), dtype=torch.uint8, pin_memory=True); hq.numpy()[:] = qcat
    ht = torch.empty(len(tcat), dtype=torch.uint8, pin_memory=True); ht.numpy()[:] = tcat
    dq = hq.cuda(non_blocking=True); dt = ht.cuda(non_blocking=True)
    ctx = K.Context(local)
    if a.panel or a.threads or a.ctas:
        ctx.set_tuning(a.panel, a.threads, a.ctas)
    P = K.make_params("extz2", mat, **W["par"])
    plan = L.ksw2b_plan_create(ctx.h, C.byref(P), n, qoff.ctypes.data, toff.ctypes.data)
    if not plan:
        raise RuntimeError("plan_create: " + L.ksw2b_last_error().decode())
    cells = int(L.ksw2b_plan_cells(plan))
    L.ksw2b_plan_set_timing(plan, 1)          # CUDA events around every DP-fill launch, on the launching stream
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def run_dev():
        rc = L.ksw2b_plan_run(plan, dq.data_ptr(), dt.data_ptr(), None, sp)
        if rc:
            raise RuntimeError("plan_run: " + L.ksw2b_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        run_dev()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    fill_ms, fill_launches = 0.0, 0
    for _ in range(a.steps):
        run_dev()
        nl = C.c_int(0)
        fill_ms += float(L.ksw2b_plan_fill_ms(plan, C.byref(nl))); fill_launches += nl.value      # waits for this step's fill launches
    e1.record(stream)
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = int(L.ksw2b_plan_launches(plan)) * a.steps
    if world > 1:
        tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
    res = np.zeros(n, dtype=K.RESULT_DTYPE)
    cigp = C.POINTER(C.c_uint32)()
    rc = L.ksw2b_plan_fetch(plan, res.ctypes.data, C.byref(cigp), sp)
    if rc:
        raise RuntimeError("plan_fetch: " + L.ksw2b_last_error().decode())
    # cells the reference semantics executes (stops at the Z-drop diagonal): SURVEY.md 8(d) cell convention
    cells_exec = int(cells_lanes(qoff, toff, W["par"]["w"], res["n_diag"])[0].sum())
    if world > 1:
        ct = torch.tensor([cells_exec], device="cuda", dtype=torch.int64); dist.all_reduce(ct); cells_all = int(ct.item())
    else:
        cells_all = cells_exec
    value = cells_all * a.steps / (ms * 1e-3) / 1e9
    cpu = None
    parity = None
    if rank == 0 and not a.no_cpu:
        ns = min(n, a.cpu_sample or 400_000)
        sb = Batch("extz2", W["par"], qcat[: qoff[ns]], qoff[: ns + 1], tcat[: toff[ns]], toff[: ns + 1])
        secs, cres, _, kind = cpu_run(sb, ncores)
        parity = bool(all(np.array_equal(res[nm][:ns], cres[:, H.FIELDS.index(nm)]) for nm in NAMES))
        ccells = int(cells_lanes(sb.qoff, sb.toff, W["par"]["w"], res["n_diag"][:ns])[0].sum())
        cpu = {"value": ccells / secs / 1e9, "unit": "GCUPS", "cores": ncores, "kind": kind,
               "sample": f"first {ns} pairs of rank 0's batch, one pass, {ncores} threads; fields bit-equal to GPU: {parity}"}
    L.ksw2b_plan_destroy(plan)

    # end-to-end through the C-ABI batch call with host buffers: pinned, pageable, and the array-of-pointers flavour
    res2 = np.zeros(n, dtype=K.RESULT_DTYPE)                 # pageable result buffer (the pageable / array-of-pointers legs)
    hres = torch.empty(n * K.RESULT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)       # pinned result buffer of the pinned leg: no staging copy
    res2p = hres.numpy().view(K.RESULT_DTYPE)
    hqn, htn = hq.numpy(), ht.numpy()

    def run_e2e(qbuf, tbuf, rbuf=None):
        cg = C.POINTER(C.c_uint32)()
        rbuf = res2 if rbuf is None else rbuf
        rc = L.ksw2b_align(ctx.h, C.byref(P), n, qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data, toff.ctypes.data, None, rbuf.ctypes.data, C.byref(cg))
        if rc:
            raise RuntimeError("ksw2b_align: " + L.ksw2b_last_error().decode())

    def timed(fn, reps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([te], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX); te = float(tt.item())
        return cells_all * reps / te / 1e9

    esteps = max(10, a.steps)
    e2e_val = timed(lambda: run_e2e(hqn, htn, res2p), esteps)
    res2[:] = res2p
    h2d, d2h = ctx.last_transfer_bytes()
    xfer = (h2d, d2h) if h2d and d2h else (len(qcat) + len(tcat), 64 * n)
    same = bool(np.array_equal(res2["score"], res["score"]) and np.array_equal(res2["max"], res["max"]))
    e2e_page = timed(lambda: run_e2e(qcat, tcat), max(3, esteps // 2))
    # array-of-pointers flavour (the reference's argument lists, one ksw_extz_t per pair), pageable memory
    qp = (qcat.ctypes.data + qoff[:-1]).astype(np.uint64); tp = (tcat.ctypes.data + toff[:-1]).astype(np.uint64)
    qlen = np.diff(qoff).astype(np.int32); tlen = np.diff(toff).astype(np.int32)
    ez = np.zeros(n * 7, dtype=np.uint64)                                        # n x ksw_extz_t (56 bytes), zero-initialised like cli.c:208
    L.ksw2b_extz2_batch.restype = C.c_int
    L.ksw2b_extz2_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int8, C.c_void_p, C.c_int8, C.c_int8,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]

    def run_ptrs():
        rc = L.ksw2b_extz2_batch(ctx.h, None, n, qlen.ctypes.data, qp.ctypes.data, tlen.ctypes.data, tp.ctypes.data, 5, mat.ctypes.data, W["par"]["q"], W["par"]["e"],
                                 W["par"]["w"], W["par"]["zdrop"], W["par"]["end_bonus"], W["par"]["flag"], ez.ctypes.data)
        if rc:
            raise RuntimeError("ksw2b_extz2_batch: " + L.ksw2b_last_error().decode())

    e2e_ptrs = timed(run_ptrs, 3)
    clocks = sampler.stop(t_wall0, time.time()) if sampler else None       # sampled from the device-timed steps to the end of the end-to-end legs (all under load)
    same_ptrs = bool(np.array_equal(ez.view(np.int32).reshape(n, 14)[:, 7], res["score"]))
    ctx.close()
    del dq, dt, hq, ht, hres, res2p
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ the other BASELINE configurations
    configs = {}
    todo = [] if a.configs == "none" else (["c1", "c3", "c4", "c5"] if a.configs == "all" else [c for c in a.configs.split(",") if c])
    for name in todo:
        c = run_config(name, K, local, rank, world, dist, gloo, no_cpu=a.no_cpu)
        if rank == 0:
            configs[name] = c
    # ------------------------------------------------------------------ all N GPUs from ONE process through the C API (rank 0)
    multi = None
    if world > 1 and todo:
        if rank == 0:
            multi = {}
            try:
                mc = K.MultiContext(world)
                for name in [x for x in ("c4", "c5") if x in todo]:
                    Wm = WORKLOADS[name]
                    bs = build_batches(name, 0, 1)
                    for bb in bs:                                   # contexts, kernels and staging of every device: a small untimed call
                        tb = bb.subset(np.arange(min(bb.n, 64 * world)))
                        mc.align_packed(K.make_params(tb.kind, mat, **tb.par), tb.qcat, tb.qoff, tb.tcat, tb.toff, None, tb.w, want_cigars=False)
                    t0 = time.perf_counter(); cells_m, ok = 0, True
                    outs = []
                    for bb in bs:
                        r, _ = mc.align_packed(K.make_params(bb.kind, mat, **bb.par), bb.qcat, bb.qoff, bb.tcat, bb.toff, None, bb.w, want_cigars=False)
                        outs.append(r)
                    te = time.perf_counter() - t0
                    for bb, r in zip(bs, outs):
                        cells_m += int(cells_lanes(bb.qoff, bb.toff, bb.band(), r["n_diag"])[0].sum())
                    pr, sp_ = mc.last()
                    multi[name] = {"value": cells_m / te / 1e9, "unit": "GCUPS", "pairs": sum(bb.n for bb in bs), "devices": world, "seconds": te,
                                   "what": "ONE process, ksw2b_multi_align over all devices, pageable host buffers in/out, one timed pass after a 64-pair-per-device warm-up call",
                                   "pairs_per_device_last_call": pr.tolist(), "device_span_ms_last_call": [round(float(x), 2) for x in sp_]}
                mc.close()
            except Exception as ex:                            # the per-rank numbers above stand on their own
                multi = {"error": str(ex)[:300]}
        dist.barrier(group=gloo)

    # ------------------------------------------------------------------ collation of the results over NCCL (north_star: "an all-gather only to collate")
    collate = None
    if world > 1 and todo:
        try:
            import ksw2_b200.multi as M
            m = 4000
            tlm = model_lengths(5, WORKLOADS["c5"]["seed"], 0, m).astype(np.int64)
            gq, gqo, gt, gto = gen_model(5, WORKLOADS["c5"]["seed"], 0, n=m)                 # the same 4000 mixed-length pairs on every rank
            Pc = K.make_params("extd2", mat, q=4, e=2, q2=24, e2=1, w=300, zdrop=400, flag=FLAG_RIGHT)
            cx = K.Context(local)
            t0 = time.perf_counter()
            allres, mycigs, myidx = M.align_balanced(lambda P_, a_, b_, c_, d_, e_: cx.align_packed(P_, a_, b_, c_, d_, e_), Pc, gq, gqo, gt, gto, rank, world,
                                                     w=300, cigar=True, gather=True, device=torch.device("cuda", local))
            dt_ = time.perf_counter() - t0
            cx.close()
            # every rank must hold the same n records in the caller's order: compare a checksum of the scores across ranks, and rank 0 checks a sample
            chk = torch.tensor([int(allres["score"].astype(np.int64).sum()), int(allres["max_t"].astype(np.int64).sum()), int(allres["n_cigar"].astype(np.int64).sum())],
                               device="cuda", dtype=torch.int64)
            lo_, hi_ = chk.clone(), chk.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            same_everywhere = bool(torch.equal(lo_, hi_))
            ok_sample = None
            if rank == 0:
                sel = np.arange(0, m, 40)
                sb = Batch("extd2", dict(q=4, e=2, q2=24, e2=1, w=300, zdrop=400, end_bonus=0, flag=FLAG_RIGHT), gq, gqo, gt, gto).subset(sel)
                _, cres, _, _ = cpu_run(sb, ncores)
                ok_sample = bool(all(np.array_equal(allres[nm][sel], cres[:, H.FIELDS.index(nm)]) for nm in NAMES))
                collate = {"pairs": m, "ranks": world, "backend": "nccl", "records_identical_on_all_ranks": same_everywhere, "sample_equals_reference": ok_sample,
                           "seconds": dt_, "what": "ksw2_b200.multi.align_balanced: cost-balanced shards aligned per rank, 64-byte result records all-gathered on the device"}
        except Exception as ex:
            if rank == 0:
                collate = {"error": str(ex)[:300]}
        dist.barrier(group=gloo)

    if rank == 0:
        peak, how = measured_peak()
        alg_bytes = float(qoff[-1] + toff[-1] + 56 * n)            # inputs at 1 B/base + one 56-B ksw_extz_t per pair
        # the dominant kernel = the DP fill; its own launch durations (CUDA events inside the library, same timed region)
        ach = alg_bytes * a.steps / (max(fill_ms, 1e-9) * 1e-3) / 1e9
        out = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": ms / a.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic", "config": cfg,
               "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic("c2", n),
                            "peak_source": how, "kernel": "ks_fill_kernel<extz2>", "algorithmic_bytes_per_step": alg_bytes,
                            "fill_launches_per_step": fill_launches / max(1, a.steps), "fill_ms_per_step": fill_ms / max(1, a.steps),
                            "fill_share_of_step": fill_ms / ms,
                            "note": "integer-ALU bound path: algorithmic traffic is tiny next to HBM peak (see DESIGN.md)"},
               "cpu_baseline": cpu,
               "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": int(xfer[0]), "d2h_bytes_per_step": int(xfer[1]),
                       "steps": esteps, "same_results_as_device_path": same, "host_buffers": "pinned",
                       "pageable": {"value": e2e_page, "unit": "GCUPS", "what": "the same ksw2b_align call from plain (pageable) numpy buffers"},
                       "batch_api": {"value": e2e_ptrs, "unit": "GCUPS", "same_scores": same_ptrs,
                                     "what": "ksw2b_extz2_batch: arrays of the reference's arguments (pointers, lengths) in pageable memory, one ksw_extz_t per pair"}},
               "gpu_launches": launches, "clocks": clocks, "cells_per_step": cells_all, "cells_full_band_rank0": cells, "parity_sample_ok": parity,
               "configs": configs}
        if multi is not None:
            out["c_api_multi"] = multi
        if collate is not None:
            out["allgather_collate"] = collate
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
