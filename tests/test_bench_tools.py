"""CPU tests of the bench tooling (tools/ksw2b_gen.c through bench.py): the workload generators are pure functions of
(model, seed, pair index), the cell accounting follows SURVEY 8(d), and the CPU checkers honour a band per pair."""
import os
import sys

import numpy as np

import harness as H

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_generators_are_functions_of_the_pair_index():
    for model, L in ((3, 500), (4, 2000), (5, 0)):
        q, qo, t, to = bench.gen_model(model, 77, L, n=200, nthreads=3)
        idx = np.array([3, 50, 199], np.int64)
        q2, qo2, t2, to2 = bench.gen_model(model, 77, L, idx=idx, nthreads=1)
        for k, i in enumerate(idx):
            assert np.array_equal(q2[qo2[k]:qo2[k + 1]], q[qo[i]:qo[i + 1]]) and np.array_equal(t2[to2[k]:to2[k + 1]], t[to[i]:to[i + 1]])
        assert int(q.max()) <= 3 and int(t.max()) <= 3
        tl = bench.model_lengths(model, 77, L, 200)
        assert np.array_equal(tl, np.diff(to))
        if model == 5:
            assert tl.min() >= 150 and tl.max() <= 20000 and len(np.unique(tl)) > 100
        else:
            assert (tl == L).all() and abs(np.diff(qo).mean() - L) < 0.05 * L
    q3, _, _, _ = bench.gen_model(3, 78, 500, n=10)
    q4, _, _, _ = bench.gen_model(3, 77, 500, n=10)
    assert not np.array_equal(q3[:400], q4[:400])


def test_cell_accounting_matches_the_harness_formula():
    rng = np.random.default_rng(1)
    ql = rng.integers(1, 400, 60); tl = rng.integers(1, 400, 60); w = rng.choice([-1, 0, 3, 50, 1000], 60)
    qoff = np.concatenate([[0], np.cumsum(ql)]); toff = np.concatenate([[0], np.cumsum(tl)])
    cells, lanes = bench.cells_lanes(qoff, toff, w)
    for i in range(60):
        assert cells[i] == H.band_cells(int(ql[i]), int(tl[i]), int(w[i])), i
    nd = np.minimum(ql + tl - 1, rng.integers(1, 300, 60))
    c2, _ = bench.cells_lanes(qoff, toff, w, nd)
    for i in range(60):
        assert c2[i] == H.band_cells(int(ql[i]), int(tl[i]), int(w[i]), r_stop=int(nd[i])), i
    assert (lanes >= cells).all()
    assert np.array_equal(bench.band_of([150, 2249, 2250, 20000]), [80, 500, 500, 500])


def test_cpu_checkers_take_a_band_per_pair():
    """ksd_run_w: per-pair band == single calls with that band, for the restatement and for the reference build"""
    rng = np.random.default_rng(9)
    qs, ts, ws = [], [], []
    for i in range(40):
        L = int(rng.integers(20, 300)); t = rng.integers(0, 4, L).astype(np.uint8)
        q = H.mutate(rng, t, sub=0.05, ins=0.02, dele=0.02)
        qs.append(q if len(q) else t[:1].copy()); ts.append(t); ws.append(int(rng.choice([-1, 5, 20, 64])))
    for which in (["oracle", "ref"] if H.have_ref() else ["oracle"]):
        for kind in ("extz2", "extd2"):
            P = H.make_params(kind, H.simple_mat(5, 2, 4), w=-1, zdrop=100, flag=0)
            res, cig, _ = H.run_cpu(which, P, qs, ts, nthreads=3, w=np.asarray(ws, np.int32))
            for i in range(40):
                P1 = H.make_params(kind, H.simple_mat(5, 2, 4), w=ws[i], zdrop=100, flag=0)
                r1, c1, _ = H.run_cpu(which, P1, [qs[i]], [ts[i]])
                assert np.array_equal(res[i, :11], r1[0, :11]) and np.array_equal(cig[i], c1[0]), (which, kind, i)
