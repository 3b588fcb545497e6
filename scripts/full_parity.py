#!/usr/bin/env python
"""One-off FULL parity runs (not samples): every pair of BASELINE config 5 (200 000 mixed pairs, two kernels, band per pair, right-aligned CIGAR),
the first 20 000 pairs of config 3 and all 1 000 000 pairs of config 2 on the GPU against the unmodified reference on the host cores --
all ksw_extz_t fields and every CIGAR word.  Prints one JSON line (kept under profiles/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
import ksw2_b200 as K
import bench

mat = H.simple_mat(5, 2, 4)
ncores = os.cpu_count() or 1
which = "ref" if H.have_ref() else "oracle"
out = {"checker": "unmodified reference (oracle/_ref)" if which == "ref" else "oracle port", "cores": ncores}
ctx = K.Context(0)


def check(name, batches):
    tot = bad = cig_bad = 0
    t_gpu = t_cpu = 0.0
    for b in batches:
        t0 = time.time()
        res, cigs = ctx.align_packed(K.make_params(b.kind, mat, **b.par), b.qcat, b.qoff, b.tcat, b.toff, None, b.w)
        t_gpu += time.time() - t0
        P = H.make_params(b.kind, mat, **b.par)
        with_cig = not (b.par["flag"] & 1)
        exp, ecig, secs = H.run_cpu(which, P, None, None, nthreads=ncores, want_cigar=with_cig, packed=(b.qcat, b.qoff, b.tcat, b.toff), w=b.w)
        t_cpu += secs
        ok = np.ones(b.n, dtype=bool)
        for nm in bench.NAMES + (["n_cigar"] if with_cig else []):
            ok &= res[nm] == exp[:, H.FIELDS.index(nm)]
        bad += int((~ok).sum()); tot += b.n
        if with_cig:
            cig_bad += sum(0 if np.array_equal(a, c) else 1 for a, c in zip(cigs, ecig))
    out[name] = {"pairs": tot, "pairs_with_a_field_mismatch": bad, "pairs_with_a_cigar_mismatch": cig_bad, "gpu_seconds_incl_python": round(t_gpu, 2), "cpu_seconds": round(t_cpu, 2)}


check("c5_all_200k_pairs", bench.build_batches("c5", 0, 1))
check("c3_first_20k_pairs", bench.build_batches("c3", 0, 1, 20000))
check("c2_all_1M_pairs", bench.build_batches("c2", 0, 1))
print(json.dumps(out))
