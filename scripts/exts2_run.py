#!/usr/bin/env python
"""One batch through the splice-aware kernel (ksw_exts2_sse semantics): 40 000 x 1 kb cDNA-like pairs (target = query with two GT..AG introns of
200-600 bp), score only, checked against the CPU checker on a sample; prints GCUPS.  Used for the ncu capture of the exts2 fill kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
import ksw2_b200 as K
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
rng = np.random.default_rng(11)
qs, ts = [], []
for i in range(n):
    q = rng.integers(0, 4, 600).astype(np.uint8)
    parts, last = [], 0
    for cut in sorted(rng.choice(np.arange(100, 500), 2, replace=False)):
        parts.append(q[last:cut]); last = cut
        intron = rng.integers(0, 4, int(rng.integers(200, 600))).astype(np.uint8)
        intron[0], intron[1], intron[-2], intron[-1] = 2, 3, 0, 2            # GT ... AG
        parts.append(intron)
    parts.append(q[last:])
    t = np.concatenate(parts)
    qq = q.copy(); m = rng.random(len(qq)) < 0.02; qq[m] = (qq[m] + 1) & 3
    qs.append(qq); ts.append(t)
mat = H.simple_mat(5, 1, 2)
kw = dict(q=2, e=1, q2=32, noncan=9, zdrop=-1, flag=0x101)                 # KSW_EZ_SPLICE_FOR + score only
qcat, qoff = K.pack(qs); tcat, toff = K.pack(ts)
ctx = K.Context(0); ctx.set_timing(True)
P = K.make_params("exts2", mat, **kw)
for _ in range(2):
    res, _ = ctx.align_packed(P, qcat, qoff, tcat, toff)
f, nf, span, nl = ctx.last_timing()
cells = int(bench.cells_lanes(qoff, toff, -1, res["n_diag"])[0].sum())
ns = 200
exp, _, _ = H.run_cpu("ref" if H.have_ref() else "oracle", H.make_params("exts2", mat, **kw), qs[:ns], ts[:ns], nthreads=os.cpu_count(), want_cigar=False)
ok = all(np.array_equal(res[nm][:ns], exp[:, H.FIELDS.index(nm)]) for nm in ("max", "max_q", "max_t", "mqe", "mte", "score", "zdropped"))
print(f"exts2: {n} pairs, {cells / span / 1e6:.1f} GCUPS (device span {span:.1f} ms, fill {f:.1f} ms), parity on {ns} pairs: {ok}")
