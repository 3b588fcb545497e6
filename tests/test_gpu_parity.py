"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI library
ksw2_b200/libksw2_b200.so and is compared bit-for-bit (all ksw_extz_t fields + CIGAR) with the oracle on
the same inputs; nothing here reads /root/reference."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import fuzzgen as F
import harness as H
from test_oracle import CASES, SEQS, cli_text, fuzz_batches

pytestmark = pytest.mark.gpu
CMP = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end", "n_cigar"]


@pytest.fixture(scope="module")
def K():
    import ksw2_b200
    return ksw2_b200


@pytest.fixture(scope="module")
def ctx(K):
    c = K.Context(0)
    yield c
    c.close()


def to_k(K, P):
    mat = np.ctypeslib.as_array(P.mat, shape=(P.m * P.m,)).copy()
    return K.make_params(P.kind, mat, m=P.m, q=P.q, e=P.e, q2=P.q2, e2=P.e2, w=P.w, zdrop=P.zdrop, end_bonus=P.end_bonus,
                         flag=P.flag, noncan=P.noncan, junc_bonus=P.junc_bonus)


def check(K, ctx, P, qs, ts, js=None, nthreads=4):
    exp, ecig, _ = H.run_cpu("oracle", P, qs, ts, js, nthreads=nthreads)
    res, cigs = ctx.align(to_k(K, P), qs, ts, js)
    for name in CMP:
        got, want = res[name], exp[:, H.FIELDS.index(name)]
        if not np.array_equal(got, want):
            i = int(np.nonzero(got != want)[0][0])
            raise AssertionError(f"{name} differs at pair {i}: got {got[i]} want {want[i]} (kind {P.kind} flag {hex(P.flag)} w {P.w} zdrop {P.zdrop} "
                                 f"qlen {len(qs[i])} tlen {len(ts[i])}); {int((got != want).sum())}/{len(qs)} pairs differ")
    if not (P.flag & 1):
        for i, (a, b) in enumerate(zip(cigs, ecig)):
            assert np.array_equal(a, b), f"CIGAR differs at pair {i}: {H.cigar_str(a)[:60]} vs {H.cigar_str(b)[:60]}"


def test_fuzz_vs_oracle(K, ctx):
    n = 0
    for P, qs, ts, js in fuzz_batches(4242, 360):
        check(K, ctx, P, qs, ts, js, nthreads=1)        # incl. flags with KSW_EZ_APPROX_MAX (0x08): the tile engine's tracker
        n += 1
    assert n == 360


def test_eqx_vs_oracle(K, ctx):
    """KSW_EZ_EQX (extd2): '=' / 'X' ops, intended ksw_cigar2eqx semantics (reference post-pass is broken: parity unpinned)"""
    from test_sim_engine import eqx_batches
    for P, qs, ts in eqx_batches(32, 40):
        check(K, ctx, P, qs, ts, None, nthreads=1)


def test_fuzz_tunings(K):
    """panel heights / CTA shapes must not change results"""
    for panel, thr, cps in [(1, 32, 1), (3, 64, 2), (8, 128, 2), (32, 128, 1), (40, 64, 1)]:
        c = K.Context(0)
        c.set_tuning(panel, thr, cps)
        for P, qs, ts, js in fuzz_batches(100 + panel, 40):
            check(K, c, P, qs, ts, js, nthreads=1)
        c.close()


def test_warp_mode_fuzz_and_long(K):
    """one warp per alignment (ks_fill_warp_kernel): same results, incl. multi-wave (> 32 blocks) pairs and odd panel heights"""
    rng = np.random.default_rng(8)
    for wp in (1, 5, 32, 128):
        c = K.Context(0)
        c.set_mode(2, wp)
        for P, qs, ts, js in fuzz_batches(500 + wp, 45):
            check(K, c, P, qs, ts, js, nthreads=1)
        qs, ts = [], []
        for i in range(12):
            L = int(rng.integers(400, 3000)); t = rng.integers(0, 4, L).astype(np.uint8); q = t.copy(); m = rng.random(L) < 0.08; q[m] = (q[m] + 1) & 3
            qs.append(np.ascontiguousarray(q[int(rng.integers(0, 20)):])); ts.append(t)
        for kind, fl, w in (("extz2", 0, -1), ("extd2", 2, 300), ("extd2", 0x41, 100), ("extz2", 0x80, 700)):
            check(K, c, H.make_params(kind, H.simple_mat(5, 2, 4), w=w, zdrop=300, flag=fl), qs, ts, nthreads=8)
        c.close()


@pytest.mark.parametrize("name", ["mt_extz2", "mt_extd2_r", "mt_exts2", "p50_extz2_w500_z400", "p50_extd2_w64"])
def test_golden_warp_mode(K, name):
    c = {c["name"]: c for c in CASES}[name]
    cx = K.Context(0); cx.set_mode(2, 0)
    P = K.make_params(c["kind"], H.simple_mat(5, *c["mat"]), **c["params"])
    res, cig = cx.align(P, [SEQS[c["q"]]], [SEQS[c["t"]]])
    for k in CMP:
        assert int(res[k][0]) == c["fields"][k], (name, k)
    if c["cigar_md5"] is not None:
        assert hashlib.md5((cli_text(cig[0]) + "\n").encode("latin1")).hexdigest() == c["cigar_md5"]
    cx.close()


GOLD = ["mt_extd2_42241_w751_z400_approx", "t1_0_extz2", "t1_1_extd2", "t1_2_extz2", "t1_2_extd2", "t1_3_extz2", "t1_4_extd2", "t5_regression_extz2", "readme_extz2",
        "mt_extz2", "mt_extz2_r", "mt_extd2", "mt_extd2_r", "mt_exts2", "p50_extz2_w500_z400", "p50_extd2_w64", "p50_extz2_w500_z50",
        "p50_extd2_w500_z50", "mt_extz2_w20", "p50_extz2_w10", "p50_extd2_w10", "p50_extz2_w30", "p50_extd2_w30", "p50_extz2_w64", "p50_extz2_w100",
        "p50_extd2_w100",
        "t1_0_extz", "t1_1_extd", "t1_2_extz", "t1_2_extd", "t1_3_extd", "t1_4_extz", "readme_extz", "mt_extz_w100_z200",
        "mt_extd_w751_z400_x", "p50_extz_w500_s", "p50_extd_w500",
        "t1_0_gg", "t1_1_gg2", "t1_2_gg", "t1_2_gg2", "t1_2_gg2_sse", "t1_3_gg2_sse", "t1_2_extf2", "t1_4_extf2", "mt_gg_w200", "mt_gg2_w200",
        "mt_extf2_w300_x100", "p50_extf2_w500"]


@pytest.mark.parametrize("name", GOLD)
def test_golden(K, ctx, name):
    c = {c["name"]: c for c in CASES}[name]
    P = K.make_params(c["kind"], H.simple_mat(5, *c["mat"]), **c["params"])
    res, cig = ctx.align(P, [SEQS[c["q"]]], [SEQS[c["t"]]])
    for k in CMP:
        assert int(res[k][0]) == c["fields"][k], (name, k)
    if c["cigar_md5"] is not None:
        assert hashlib.md5((cli_text(cig[0]) + "\n").encode("latin1")).hexdigest() == c["cigar_md5"]


def test_single_pair_entry_points(K):
    """the unchanged ksw2.h entry points, incl. re-use and doubling of ez->cigar (ksw2.h:113-123)"""
    L = K.lib()
    rng = np.random.default_rng(3)
    mat = H.simple_mat(5, 2, 4)
    ez = K.ExtzT()
    for it in range(12):
        tl = int(rng.integers(20, 400))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = t.copy(); q[rng.random(tl) < 0.1] = 3; q = np.ascontiguousarray(q[: max(5, tl - int(rng.integers(0, 9)))])
        kind = it % 2
        P = H.make_params("extz2" if kind == 0 else "extd2", mat, w=-1 if it % 3 else 40, zdrop=-1 if it % 4 else 50, flag=[0, 2, 0x40, 0x80][it % 4])
        exp, ecig, _ = H.run_cpu("oracle", P, [q], [t])
        m_before = ez.m_cigar
        if kind == 0:
            L.ksw_extz2_sse(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data, 4, 2, P.w, P.zdrop, 0, P.flag, C.byref(ez))
        else:
            L.ksw_extd2_sse(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data, 4, 2, 24, 1, P.w, P.zdrop, 0, P.flag, C.byref(ez))
        got = [ez.max_zd & 0x7fffffff, ez.max_zd >> 31, ez.max_q, ez.max_t, ez.mqe, ez.mqe_t, ez.mte, ez.mte_q, ez.score, ez.n_cigar, ez.reach_end]
        assert got == [int(x) for x in exp[0][:11]], (it, got, exp[0])
        assert [ez.cigar[i] for i in range(ez.n_cigar)] == [int(x) for x in ecig[0]]
        m = m_before
        while m < ez.n_cigar:
            m = m << 1 if m else 4
        assert ez.m_cigar == m
    assert C.sizeof(K.ExtzT) == 56


def test_rows_fuzz_vs_oracle(K, ctx):
    """ksw_extz / ksw_extd semantics (row-wise kernels) through ksw2b_align, all fields + CIGAR"""
    n = 0
    for kind, mat, kw, qs, ts in F.rows_batches(31337, 100, npairs=40):
        check(K, ctx, H.make_params(kind, mat, **kw), qs, ts, nthreads=4)
        n += len(qs)
    assert n == 4000


def test_extf2_fuzz_and_entry_point(K, ctx):
    """ksw_extf2_sse semantics through ksw2b_align and through the exported single-pair symbol"""
    mat = H.simple_mat(5, 2, 4)
    n = 0
    for kw, qs, ts in F.extf2_batches(99, 100, npairs=40):
        check(K, ctx, H.make_params("extf2", mat, **kw), qs, ts, nthreads=4)
        n += len(qs)
    assert n == 4000
    L = K.lib()
    rng = np.random.default_rng(8)
    ez = K.ExtzT()
    for it in range(6):
        tl = int(rng.integers(20, 300))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = t.copy(); q[rng.random(tl) < 0.1] = 3; q = np.ascontiguousarray(q[: max(5, tl - int(rng.integers(0, 9)))])
        P = H.make_params("extf2", mat, q=2, q2=-4, e=2, w=[-1, 20, 40][it % 3], zdrop=[-1, 30][it % 2], flag=1)
        exp, _, _ = H.run_cpu("oracle", P, [q], [t])
        L.ksw_extf2_sse(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 2, -4, 2, P.w, P.zdrop, C.byref(ez))
        got = [ez.max_zd & 0x7fffffff, ez.max_zd >> 31, ez.max_q, ez.max_t, ez.mqe, ez.mqe_t, ez.mte, ez.mte_q, ez.score]
        assert got == [int(x) for x in exp[0][:9]], (it, got, exp[0])


def test_gg_fuzz_and_entry_point(K, ctx):
    """ksw_gg: batches through ksw2b_align, then the exported symbol with the reference's pointer-triple CIGAR interface"""
    n = 0
    for kind, mat, kw, qs, ts in F.gg_batches(555, 90, kinds=("gg", "gg2", "gg2_sse"), npairs=30):
        check(K, ctx, H.make_params(kind, mat, **kw), qs, ts, nthreads=4)
        n += len(qs)
    assert n == 2700
    L = K.lib()
    rng = np.random.default_rng(9)
    mat = H.simple_mat(5, 2, 4)
    m_cig, n_cig, cig = C.c_int(0), C.c_int(0), C.POINTER(C.c_uint32)()
    for it in range(6):
        tl = int(rng.integers(20, 300))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = t.copy(); q[rng.random(tl) < 0.1] = 3; q = np.ascontiguousarray(q[: max(5, tl - int(rng.integers(0, 9)))])
        P = H.make_params("gg", mat, w=[-1, 40][it % 2], flag=0)
        exp, ecig, _ = H.run_cpu("oracle", P, [q], [t])
        sc = L.ksw_gg(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data, 4, 2, P.w, C.byref(m_cig), C.byref(n_cig), C.byref(cig))
        assert sc == int(exp[0][8]) and n_cig.value == len(ecig[0])
        assert [cig[i] for i in range(n_cig.value)] == [int(x) for x in ecig[0]]
        sc2 = L.ksw_gg(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data, 4, 2, P.w, None, None, None)
        assert sc2 == sc
        for name in ("gg2", "gg2_sse"):
            P2 = H.make_params(name, mat, w=P.w, flag=0)
            exp2, ecig2, _ = H.run_cpu("oracle", P2, [q], [t])
            sc3 = getattr(L, "ksw_" + name)(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data, 4, 2, P.w,
                                            C.byref(m_cig), C.byref(n_cig), C.byref(cig))
            assert sc3 == int(exp2[0][8]) and [cig[i] for i in range(n_cig.value)] == [int(x) for x in ecig2[0]], name


def test_rows_single_pair_entry_points(K):
    """ksw_extz / ksw_extd exported with the reference prototypes (ksw2.h:61-62,67-68)"""
    L = K.lib()
    rng = np.random.default_rng(5)
    mat = H.simple_mat(5, 2, 4)
    ez = K.ExtzT()
    for it in range(10):
        tl = int(rng.integers(20, 300))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = t.copy(); q[rng.random(tl) < 0.1] = 3; q = np.ascontiguousarray(q[: max(5, tl - int(rng.integers(0, 9)))])
        dual = it % 2
        P = H.make_params("extd" if dual else "extz", mat, w=-1 if it % 3 else 40, zdrop=-1 if it % 4 else 50, flag=[0, 2, 0x40, 0x80][it % 4])
        exp, ecig, _ = H.run_cpu("oracle", P, [q], [t])
        if dual:
            L.ksw_extd(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data, 4, 2, 24, 1, P.w, P.zdrop, P.flag, C.byref(ez))
        else:
            L.ksw_extz(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data, 4, 2, P.w, P.zdrop, P.flag, C.byref(ez))
        got = [ez.max_zd & 0x7fffffff, ez.max_zd >> 31, ez.max_q, ez.max_t, ez.mqe, ez.mqe_t, ez.mte, ez.mte_q, ez.score, ez.n_cigar, ez.reach_end]
        assert got == [int(x) for x in exp[0][:11]], (it, got, exp[0])
        assert [ez.cigar[i] for i in range(ez.n_cigar)] == [int(x) for x in ecig[0]]


def test_concurrent_single_pair_calls_are_combined(K):
    """many host threads calling the unchanged one-pair-per-call API at once; the library combines the calls
    that are in flight into GPU batches (group commit) and every caller still gets exactly its own result"""
    import threading
    L = K.lib()
    L.ksw2b_combine_stats.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    mat = H.simple_mat(5, 2, 4)
    nthr, per = 16, 40
    rng = np.random.default_rng(77)
    work = []
    for t in range(nthr):
        items = []
        for i in range(per):
            tl = int(rng.integers(30, 400))
            tt = rng.integers(0, 4, tl).astype(np.uint8)
            q = tt.copy(); q[rng.random(tl) < 0.08] = 2; q = np.ascontiguousarray(q[: max(5, tl - int(rng.integers(0, 12)))])
            items.append((q, tt))
        work.append(items)
    # three parameter sets in flight at the same time (threads 0-7: extz2 extension, 8-11: extd2 global CIGAR, 12-15: ksw_extz rows)
    def params(t):
        if t < 8:
            return H.make_params("extz2", mat, w=50, zdrop=100, flag=0x40)
        if t < 12:
            return H.make_params("extd2", mat, w=-1, zdrop=-1, flag=0)
        return H.make_params("extz", mat, w=60, zdrop=200, flag=0)
    got = [[None] * per for _ in range(nthr)]
    c0, b0 = C.c_ulonglong(0), C.c_ulonglong(0)
    L.ksw2b_combine_stats(C.byref(c0), C.byref(b0))

    def run(t):
        P = params(t)
        ez = K.ExtzT()
        for i, (q, tt) in enumerate(work[t]):
            if t < 8:
                L.ksw_extz2_sse(None, len(q), q.ctypes.data, len(tt), tt.ctypes.data, 5, mat.ctypes.data, 4, 2, P.w, P.zdrop, 0, P.flag, C.byref(ez))
            elif t < 12:
                L.ksw_extd2_sse(None, len(q), q.ctypes.data, len(tt), tt.ctypes.data, 5, mat.ctypes.data, 4, 2, 24, 1, P.w, P.zdrop, 0, P.flag, C.byref(ez))
            else:
                L.ksw_extz(None, len(q), q.ctypes.data, len(tt), tt.ctypes.data, 5, mat.ctypes.data, 4, 2, P.w, P.zdrop, P.flag, C.byref(ez))
            got[t][i] = ([ez.max_zd & 0x7fffffff, ez.max_zd >> 31, ez.max_q, ez.max_t, ez.mqe, ez.mqe_t, ez.mte, ez.mte_q, ez.score, ez.n_cigar, ez.reach_end],
                         [ez.cigar[k] for k in range(ez.n_cigar)])
    th = [threading.Thread(target=run, args=(t,)) for t in range(nthr)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    for t in range(nthr):
        exp, ecig, _ = H.run_cpu("oracle", params(t), [w[0] for w in work[t]], [w[1] for w in work[t]])
        for i in range(per):
            assert got[t][i][0] == [int(x) for x in exp[i][:11]], (t, i)
            assert got[t][i][1] == [int(x) for x in ecig[i]], (t, i)
    c1, b1 = C.c_ulonglong(0), C.c_ulonglong(0)
    L.ksw2b_combine_stats(C.byref(c1), C.byref(b1))
    assert c1.value - c0.value == nthr * per
    assert b1.value - b0.value <= nthr * per          # usually far fewer launches than calls


def test_invalid_and_empty_inputs(K, ctx):
    mat = H.simple_mat(5, 2, 4)
    q = np.array([0, 1, 2, 3], np.uint8)
    # mismatch too negative -> early return with reset ez (ksw2_extz2_sse.c:82)
    P = K.make_params("extz2", H.simple_mat(5, 2, 40), q=4, e=2)
    res, cig = ctx.align(P, [q], [q])
    assert res["score"][0] == -0x40000000 and res["max"][0] == 0 and res["max_t"][0] == -1 and res["n_cigar"][0] == 0
    # empty query / target inside a batch
    P = K.make_params("extd2", mat)
    res, cig = ctx.align(P, [q, np.zeros(0, np.uint8), q], [q, q, np.zeros(0, np.uint8)])
    assert res["score"][1] == -0x40000000 and res["score"][2] == -0x40000000 and res["score"][0] == 8
    # exts2 with q2 <= q+e -> early return
    P = K.make_params("exts2", mat, q=3, e=1, q2=3, noncan=4, flag=0x100)
    res, _ = ctx.align(P, [q], [q])
    assert res["score"][0] == -0x40000000


def synth_reads(rng, n, L, sub=0.01, indel=0.002, junk_frac=0.1):
    ts = rng.integers(0, 4, (n, L)).astype(np.uint8)
    qs = []
    for i in range(n):
        t = ts[i]
        q = t.copy()
        m = rng.random(L) < sub
        q[m] = (q[m] + rng.integers(1, 4, int(m.sum()))) & 3
        for _ in range(rng.poisson(indel * L)):
            p = int(rng.integers(1, L - 1))
            if rng.random() < 0.5:
                q = np.concatenate([q[:p], q[p + 1:], rng.integers(0, 4, 1).astype(np.uint8)])
            else:
                q = np.concatenate([q[:p], rng.integers(0, 4, 1).astype(np.uint8), q[p:-1]])
        if rng.random() < junk_frac:
            k = int(rng.integers(40, 76))
            q[-k:] = rng.integers(0, 4, k)
        q[rng.random(L) < 0.01] = 4
        qs.append(np.ascontiguousarray(q))
    return qs, [np.ascontiguousarray(t) for t in ts]


def test_c2_sample_150bp_extension(K, ctx):
    """BASELINE config 2 geometry on a 20k-pair sample: 150 bp, w=100, Z-drop, score only (flag 0x41)"""
    rng = np.random.default_rng(20260925)
    qs, ts = synth_reads(rng, 20000, 150)
    P = H.make_params("extz2", H.simple_mat(5, 2, 4), q=4, e=2, w=100, zdrop=100, flag=0x41)
    check(K, ctx, P, qs, ts, nthreads=8)


def test_c3_sample_5kb_dual_gap_cigar(K, ctx):
    """BASELINE config 3 geometry on a small sample: 5 kb ONT-like, extd2, w=500, zdrop=400, CIGAR"""
    rng = np.random.default_rng(20260926)
    qs, ts = [], []
    for i in range(48):
        t = rng.integers(0, 4, 5000).astype(np.uint8)
        q = []
        for b in t:
            x = rng.random()
            if x < 0.035:
                continue
            if x < 0.07:
                q.extend(rng.integers(0, 4, int(rng.geometric(0.7))).tolist())
            q.append(int((b + rng.integers(1, 4)) & 3) if x > 0.97 else int(b))
        qs.append(np.asarray(q, np.uint8)); ts.append(t)
    P = H.make_params("extd2", H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=500, zdrop=400, flag=0)
    check(K, ctx, P, qs, ts, nthreads=8)
    P = H.make_params("extd2", H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=500, zdrop=400, flag=2)
    check(K, ctx, P, qs[:16], ts[:16], nthreads=8)


def test_mixed_lengths_right_aligned(K, ctx):
    """BASELINE config 5 flavour: mixed 150 bp .. 6 kb, extz2 + extd2, right-aligned CIGAR"""
    rng = np.random.default_rng(9)
    qs, ts = [], []
    for i in range(96):
        L = int(np.exp(rng.uniform(np.log(150), np.log(6000))))
        t = rng.integers(0, 4, L).astype(np.uint8)
        q = t.copy(); m = rng.random(L) < rng.uniform(0.01, 0.12); q[m] = (q[m] + 1) & 3
        cut = int(rng.integers(0, 20))
        qs.append(np.ascontiguousarray(q[cut:])); ts.append(t)
    for kind in ("extz2", "extd2"):
        P = H.make_params(kind, H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=300, zdrop=400, flag=2)
        check(K, ctx, P, qs, ts, nthreads=8)


def _run_plan(K, ctx, P, qcat, qoff, tcat, toff):
    """device-resident path (ksw2b_plan_*): returns the result records"""
    import torch
    L = K.lib()
    dq = torch.from_numpy(qcat).cuda(); dt = torch.from_numpy(tcat).cuda()
    n = len(qoff) - 1
    plan = L.ksw2b_plan_create(ctx.h, C.byref(P), n, qoff.ctypes.data, toff.ctypes.data)
    assert plan, L.ksw2b_last_error()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.ksw2b_plan_run(plan, dq.data_ptr(), dt.data_ptr(), None, sp) == 0, L.ksw2b_last_error()
    res = np.zeros(n, dtype=K.RESULT_DTYPE)
    cg = C.POINTER(C.c_uint32)()
    assert L.ksw2b_plan_fetch(plan, res.ctypes.data, C.byref(cg), sp) == 0, L.ksw2b_last_error()
    L.ksw2b_plan_destroy(plan)
    return res


def test_full_size_c2_properties(K):
    """BASELINE config 2 at its FULL size (1 M pairs x 150 bp, flag 0x41) through size-independent properties:
    results do not depend on the launch shape (panel / CTA tuning), on the entry point (device-resident plan vs host-buffer
    ksw2b_align), nor on the order of the pairs (a permuted batch gives the permuted results); a 20 k-pair sample spread
    over the batch is bit-equal to the oracle; and the executed-cell accounting is consistent (n_diag within bounds)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    n = 1_000_000
    qcat, qoff, tcat, toff = bench.gen_c2(n, 150, 20260925)
    mat = H.simple_mat(5, 2, 4)
    par = dict(q=4, e=2, w=100, zdrop=100, end_bonus=0, flag=0x41)
    P = K.make_params("extz2", mat, **par)
    c1 = K.Context(0)
    r1 = _run_plan(K, c1, P, qcat, qoff, tcat, toff)
    c2 = K.Context(0); c2.set_tuning(7, 64, 3)
    r2 = _run_plan(K, c2, P, qcat, qoff, tcat, toff)
    names = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end", "n_cigar", "n_diag"]
    for nm in names:
        assert np.array_equal(r1[nm], r2[nm]), nm
    # host-buffer entry point
    r3 = np.zeros(n, dtype=K.RESULT_DTYPE)
    cg = C.POINTER(C.c_uint32)()
    L = K.lib()
    assert L.ksw2b_align(c1.h, C.byref(P), n, qcat.ctypes.data, qoff.ctypes.data, tcat.ctypes.data, toff.ctypes.data, None, r3.ctypes.data, C.byref(cg)) == 0
    for nm in names:
        assert np.array_equal(r1[nm], r3[nm]), nm
    # permutation invariance on a 200 k slice
    m = 200_000
    perm = np.random.default_rng(1).permutation(m)
    q2 = qcat[: m * 150].reshape(m, 150)[perm].reshape(-1).copy(); t2 = tcat[: m * 150].reshape(m, 150)[perm].reshape(-1).copy()
    r4 = _run_plan(K, c1, P, q2, qoff[: m + 1], t2, toff[: m + 1])
    for nm in names:
        assert np.array_equal(r4[nm], r1[nm][:m][perm]), nm
    # oracle on a strided sample
    idx = np.arange(0, n, 50)
    sq = qcat.reshape(n, 150)[idx]; st_ = tcat.reshape(n, 150)[idx]
    exp, _, _ = H.run_cpu("oracle", H.make_params("extz2", mat, **par), list(sq), list(st_), nthreads=8, want_cigar=False)
    for nm in names[:10]:
        assert np.array_equal(r1[nm][idx], exp[:, H.FIELDS.index(nm)]), nm
    assert r1["n_diag"].min() >= 1 and r1["n_diag"].max() <= 299
    assert int(r1["zdropped"].sum()) > 0                # some junk-tailed pairs do Z-drop (max - H > 100 needs a long bad tail)
    c1.close(); c2.close()


def test_mid_size_c3_modes_agree(K):
    """BASELINE config 3 geometry (5 kb, extd2, w=500, CIGAR) on 600 pairs: the thread-per-pair and the warp-per-pair kernels,
    and a chunked run (direction arena forced small via many pairs per call is not needed: two half batches) agree bit for bit
    incl. every CIGAR word; a sample is checked against the oracle."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    n = 600
    qcat, qoff, tcat, toff = bench.gen_c3(n, 5000, 20260926)
    qs = [qcat[qoff[i]:qoff[i + 1]] for i in range(n)]; ts = [tcat[toff[i]:toff[i + 1]] for i in range(n)]
    P = K.make_params("extd2", H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=500, zdrop=400, flag=0)
    ca = K.Context(0); ca.set_mode(1, 0)
    cb = K.Context(0); cb.set_mode(2, 64)
    ra, ga = ca.align(P, qs, ts)
    rb, gb = cb.align(P, qs, ts)
    for nm in CMP + ["n_diag"]:
        assert np.array_equal(ra[nm], rb[nm]), nm
    for x, y in zip(ga, gb):
        assert np.array_equal(x, y)
    # halves == whole
    rh, gh = ca.align(P, qs[: n // 2], ts[: n // 2])
    for nm in CMP:
        assert np.array_equal(ra[nm][: n // 2], rh[nm]), nm
    for x, y in zip(ga[: n // 2], gh):
        assert np.array_equal(x, y)
    # CIGAR consistency: every CIGAR consumes exactly the aligned prefixes
    for i in range(n):
        if ra["n_cigar"][i] == 0:
            continue
        ops = ga[i] & 15; lens = ga[i] >> 4
        tl = int(lens[(ops == 0) | (ops == 2) | (ops == 3)].sum()); ql = int(lens[(ops == 0) | (ops == 1)].sum())
        if ra["zdropped"][i]:
            assert tl == ra["max_t"][i] + 1 and ql == ra["max_q"][i] + 1
        else:
            assert tl == 5000 and ql == len(qs[i])
    exp, ecig, _ = H.run_cpu("oracle", H.make_params("extd2", H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=500, zdrop=400, flag=0), qs[:40], ts[:40], nthreads=8)
    for nm in CMP:
        assert np.array_equal(ra[nm][:40], exp[:, H.FIELDS.index(nm)]), nm
    for x, y in zip(ga[:40], ecig):
        assert np.array_equal(x, y)
    ca.close(); cb.close()


# ---------------------------------------------------------------------------------------------------------
# round 2: band per pair, BASELINE config 5 geometry, thread-mode goldens on the long pairs, several GPUs from one caller,
# the one-live-plan rule, the caller's kalloc arenas
# ---------------------------------------------------------------------------------------------------------
def _check_w(K, ctx, P, qs, ts, w, nthreads=8, which="oracle"):
    exp, ecig, _ = H.run_cpu(which, P, qs, ts, None, nthreads=nthreads, w=w)
    res, cigs = ctx.align(to_k(K, P), qs, ts, None, w=w)
    for name in CMP:
        got, want = res[name], exp[:, H.FIELDS.index(name)]
        assert np.array_equal(got, want), (name, int(np.nonzero(got != want)[0][0]))
    if not (P.flag & 1):
        for i, (a, b) in enumerate(zip(cigs, ecig)):
            assert np.array_equal(a, b), f"CIGAR differs at pair {i}"
    return res, cigs


def test_band_per_pair(K, ctx):
    """ksw2b_align_ex: one band per pair (what a batch collected from minimap2-style calls carries) == n single calls with that band"""
    rng = np.random.default_rng(123)
    qs, ts, ws = [], [], []
    for i in range(300):
        L = int(np.exp(rng.uniform(np.log(30), np.log(1500))))
        t = rng.integers(0, 4, L).astype(np.uint8)
        q = H.mutate(rng, t, sub=0.05, ins=0.01, dele=0.01)
        qs.append(q if len(q) else t[:1].copy()); ts.append(t)
        ws.append(int(rng.choice([-1, 0, 1, 7, 16, 33, 100, 257, 5000, min(500, (L + 4) // 5 + 50)])))
    w = np.asarray(ws, np.int32)
    for kind, fl in (("extz2", 0x41), ("extz2", 2), ("extd2", 0), ("extd2", 0x42)):
        _check_w(K, ctx, H.make_params(kind, H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=-1, zdrop=200, flag=fl), qs, ts, w)
    # a uniform-length score-only batch with differing bands must not take the device-generated job table
    t = [rng.integers(0, 4, 150).astype(np.uint8) for _ in range(400)]
    _check_w(K, ctx, H.make_params("extz2", H.simple_mat(5, 2, 4), q=4, e=2, w=100, zdrop=100, flag=0x41), [x.copy() for x in t], t,
             np.asarray(rng.choice([10, 50, 100], 400), np.int32))


def test_c5_geometry_thread_mode(K):
    """BASELINE config 5 at its stated extremes, in THREAD mode (the kernel bench.py times): 20 kb pairs, w=500, extz2 and extd2,
    KSW_EZ_RIGHT + CIGAR, mixed with short pairs and their per-pair bands; every field and CIGAR word against the CPU checker"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    c = K.Context(0); c.set_mode(1, 0)
    tl = bench.model_lengths(5, 20260928, 0, 4000).astype(np.int64)
    order = np.argsort(-tl)
    idx = np.sort(np.concatenate([order[:6], order[1000:1010], order[-10:]]))        # the longest (~20 kb), some mid, the shortest (150 bp)
    qcat, qoff, tcat, toff = bench.gen_model(5, 20260928, 0, idx=idx)
    qs = [qcat[qoff[i]:qoff[i + 1]] for i in range(len(idx))]; ts = [tcat[toff[i]:toff[i + 1]] for i in range(len(idx))]
    w = bench.band_of(tl[idx])
    assert max(len(t) for t in ts) > 19000 and int(w.max()) == 500
    which = "ref" if H.have_ref() else "oracle"
    for kind in ("extz2", "extd2"):
        P = H.make_params(kind, H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=-1, zdrop=400, flag=2)
        res, _ = _check_w(K, c, P, qs, ts, w, which=which)
        assert int(res["n_cigar"].min()) > 0
    c.close()


@pytest.mark.parametrize("name", ["mt_extz2", "mt_extz2_r", "mt_extd2", "p50_extz2_w500_z400", "p50_extd2_w64", "p50_extd2_w500_z50", "p50_extz2_w100"])
def test_golden_thread_mode(K, name):
    """the long golden pairs (MT 16.5 kb, phage 50 kb) through the thread-per-pair kernel (a lone pair would auto-select the warp kernel)"""
    c = {c["name"]: c for c in CASES}[name]
    cx = K.Context(0); cx.set_mode(1, 0)
    P = K.make_params(c["kind"], H.simple_mat(5, *c["mat"]), **c["params"])
    res, cig = cx.align(P, [SEQS[c["q"]]], [SEQS[c["t"]]])
    for k in CMP:
        assert int(res[k][0]) == c["fields"][k], (name, k)
    if c["cigar_md5"] is not None:
        assert hashlib.md5((cli_text(cig[0]) + "\n").encode("latin1")).hexdigest() == c["cigar_md5"]
    cx.close()


def test_multi_device_set_one_caller(K, ctx):
    """ksw2b_multi_align: the device set shards one batch (contiguous for equal lengths, cost-balanced otherwise), results
    and CIGARs come back in the caller's order.  Uses every visible GPU, and two contexts on device 0 when there is only one."""
    import torch
    nd = torch.cuda.device_count()
    devs = list(range(nd)) if nd > 1 else [0, 0]
    mc = K.MultiContext(devs)
    rng = np.random.default_rng(5150)
    # mixed lengths + CIGAR + band per pair
    qs, ts = [], []
    for i in range(500):
        L = int(np.exp(rng.uniform(np.log(40), np.log(3000))))
        t = rng.integers(0, 4, L).astype(np.uint8)
        q = H.mutate(rng, t, sub=0.04, ins=0.01, dele=0.01)
        qs.append(q if len(q) else t[:1].copy()); ts.append(t)
    w = np.minimum(500, (np.asarray([len(t) for t in ts]) + 4) // 5 + 50).astype(np.int32)
    qcat, qoff = K.pack(qs); tcat, toff = K.pack(ts)
    for kind, fl in (("extd2", 2), ("extz2", 0x41)):
        P = K.make_params(kind, H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=-1, zdrop=400, flag=fl)
        r1, c1 = ctx.align_packed(P, qcat, qoff, tcat, toff, None, w)
        r2, c2 = mc.align_packed(P, qcat, qoff, tcat, toff, None, w)
        for nm in CMP + ["n_diag"]:
            assert np.array_equal(r1[nm], r2[nm]), nm
        for a, b in zip(c1, c2):
            assert np.array_equal(a, b)
        pairs, _ = mc.last()
        assert int(pairs.sum()) == 500 and int(pairs.min()) > 0
    # equal lengths: contiguous shards
    t = rng.integers(0, 4, (3001, 150)).astype(np.uint8)
    q = t.copy(); q[rng.random(q.shape) < 0.03] = 1
    off = np.arange(3002, dtype=np.int64) * 150
    P = K.make_params("extz2", H.simple_mat(5, 2, 4), q=4, e=2, w=100, zdrop=100, flag=0x41)
    r1, _ = ctx.align_packed(P, q.reshape(-1), off, t.reshape(-1), off)
    r2, _ = mc.align_packed(P, q.reshape(-1), off, t.reshape(-1), off)
    for nm in CMP + ["n_diag"]:
        assert np.array_equal(r1[nm], r2[nm]), nm
    # fewer pairs than devices, and an empty batch
    r3, _ = mc.align_packed(P, q[:1].reshape(-1), off[:2], t[:1].reshape(-1), off[:2])
    assert int(r3["score"][0]) == int(r1["score"][0]) or int(r3["max"][0]) == int(r1["max"][0])
    r4, _ = mc.align_packed(P, np.zeros(1, np.uint8), np.zeros(1, np.int64), np.zeros(1, np.uint8), np.zeros(1, np.int64))
    assert len(r4) == 0
    mc.close()


def test_one_live_plan_per_context(K, ctx):
    """a plan borrows its context's buffers: a second plan, or ksw2b_align, on the same context is refused (-4) while one is alive"""
    L = K.lib()
    P = K.make_params("extz2", H.simple_mat(5, 2, 4), w=50, zdrop=100, flag=0x41)
    off = np.arange(11, dtype=np.int64) * 100
    p1 = L.ksw2b_plan_create(ctx.h, C.byref(P), 10, off.ctypes.data, off.ctypes.data)
    assert p1
    p2 = L.ksw2b_plan_create(ctx.h, C.byref(P), 10, off.ctypes.data, off.ctypes.data)
    assert not p2 and b"one live plan" in L.ksw2b_last_error()
    q = np.zeros(1000, np.uint8); res = np.zeros(10, dtype=K.RESULT_DTYPE); cg = C.POINTER(C.c_uint32)()
    assert L.ksw2b_align(ctx.h, C.byref(P), 10, q.ctypes.data, off.ctypes.data, q.ctypes.data, off.ctypes.data, None, res.ctypes.data, C.byref(cg)) != 0
    L.ksw2b_plan_destroy(p1)
    assert L.ksw2b_align(ctx.h, C.byref(P), 10, q.ctypes.data, off.ctypes.data, q.ctypes.data, off.ctypes.data, None, res.ctypes.data, C.byref(cg)) == 0
    assert int(res["max"][0]) == 200


def test_caller_kalloc_arenas(K):
    """ez->cigar lives in the CALLER's kalloc arena (reference ksw2.h:103-119, kalloc.c:136): a program that links the reference's kalloc.c,
    calls the drop-in symbols with km = km_init() from several threads and frees with kfree(km, ez.cigar) (oracle/kalloc_interop.c,
    prebuilt into oracle/_ref/ where the reference exists); it also compares every call with the reference kernels it links"""
    import os, subprocess
    exe = os.path.join(H.ORACLE_DIR, "_ref", "kalloc_interop")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/kalloc_interop was not built (no reference on this box)")
    out = subprocess.run([exe, K.LIB_PATH, "6", "40"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatches" in out.stdout


def test_ring_schedule_banded_pairs(K):
    """one warp per alignment on the ring schedule (ks_fill_ring_kernel: banded pairs, effective band <= 512; what long CIGAR pairs run on):
    the fuzz domain (pairs outside the schedule's domain keep the automatic choice), 5 kb dual-gap pairs at w = 500 / 512 with every CIGAR
    word, truncated pairs whose band runs empty, approximate-max"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    c = K.Context(0); c.set_mode(4, 0)
    for P, qs, ts, js in fuzz_batches(909, 120):
        check(K, c, P, qs, ts, js, nthreads=1)
    q, qo, t, to = bench.gen_model(3, 5, 5000, n=24)
    qs = [q[qo[i]:qo[i + 1]] for i in range(24)]; ts = [t[to[i]:to[i + 1]] for i in range(24)]
    qs[3] = qs[3][:3000]; ts[5] = ts[5][:3500]; qs[7] = qs[7][:150]
    for kind, fl, w, zd in (("extd2", 0, 500, 400), ("extd2", 2, 512, 400), ("extz2", 0x41, 500, 100), ("extz2", 2, 511, -1), ("extd2", 8, 500, 400), ("extd2", 0x18, 300, 200)):
        check(K, c, H.make_params(kind, H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=w, zdrop=zd, flag=fl), qs, ts, nthreads=8)
    c.close()


def test_cta_per_pair_mode(K):
    """one CTA (256 lanes) per alignment (ks_fill_cta_kernel: a handful of very long pairs): fuzz domain, the MT pair with its golden CIGAR"""
    c = K.Context(0); c.set_mode(3, 0)
    for P, qs, ts, js in fuzz_batches(910, 60):
        check(K, c, P, qs, ts, js, nthreads=1)
    g = {x["name"]: x for x in CASES}["mt_extz2"]
    res, cig = c.align(K.make_params(g["kind"], H.simple_mat(5, *g["mat"]), **g["params"]), [SEQS[g["q"]]], [SEQS[g["t"]]])
    for k in CMP:
        assert int(res[k][0]) == g["fields"][k], k
    assert hashlib.md5((cli_text(cig[0]) + "\n").encode("latin1")).hexdigest() == g["cigar_md5"]
    c.close()
    # the automatic choice for a lone 16.5 kb pair is this kernel too: same answer from a default context
    c2 = K.Context(0)
    res2, cig2 = c2.align(K.make_params(g["kind"], H.simple_mat(5, *g["mat"]), **g["params"]), [SEQS[g["q"]]], [SEQS[g["t"]]])
    assert np.array_equal(cig2[0], cig[0]) and int(res2["score"][0]) == g["fields"]["score"]
    c2.close()


def test_cigar_runs_in_several_chunks(K, monkeypatch):
    """CIGAR batches larger than the direction arena run chunk by chunk, each chunk's CIGAR words fetched with one chunk of lag from a
    double-buffered staging area: forced here with a 24 MB arena (KSW2B_ARENA_MB) on 300 x ~2 kb pairs; results equal the one-chunk run"""
    rng = np.random.default_rng(2024)
    qs, ts = [], []
    for i in range(300):
        L = int(rng.integers(800, 2400)); t = rng.integers(0, 4, L).astype(np.uint8)
        q = H.mutate(rng, t, sub=0.05, ins=0.02, dele=0.02)
        qs.append(q if len(q) else t[:1].copy()); ts.append(t)
    P = K.make_params("extd2", H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=300, zdrop=400, flag=0)
    c1 = K.Context(0)
    r1, g1 = c1.align(P, qs, ts)
    monkeypatch.setenv("KSW2B_ARENA_MB", "24")
    c2 = K.Context(0)
    r2, g2 = c2.align(P, qs, ts)
    r3, g3 = c2.align(P, qs, ts)                       # and again on the same context (buffers re-used)
    monkeypatch.delenv("KSW2B_ARENA_MB")
    for nm in CMP + ["n_diag"]:
        assert np.array_equal(r1[nm], r2[nm]) and np.array_equal(r1[nm], r3[nm]), nm
    for a, b, c_ in zip(g1, g2, g3):
        assert np.array_equal(a, b) and np.array_equal(a, c_)
    exp, ecig, _ = H.run_cpu("oracle", H.make_params("extd2", H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=300, zdrop=400, flag=0), qs[:60], ts[:60], nthreads=8)
    for a, b in zip(g2[:60], ecig):
        assert np.array_equal(a, b)
    c1.close(); c2.close()


def test_mixed_length_chunk_is_split_between_ring_and_threads(K):
    """a rank's share of a mixed-length batch (no arena pressure): the long head of the length-sorted chunk goes to the ring schedule, the rest runs
    one thread per pair (plan_build's tail split); results equal the all-threads run and the CPU checker"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    n = 2400
    tl = bench.model_lengths(5, 20260928, 0, n).astype(np.int64)
    qcat, qoff, tcat, toff = bench.gen_model(5, 20260928, 0, n=n)
    w = bench.band_of(tl)
    ca = K.Context(0)
    cb = K.Context(0); cb.set_mode(1, 0)
    for kind, fl in (("extd2", 2), ("extz2", 0x42)):
        P = K.make_params(kind, H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=-1, zdrop=400, flag=fl)
        ra, ga = ca.align_packed(P, qcat, qoff, tcat, toff, None, w)
        rb, gb = cb.align_packed(P, qcat, qoff, tcat, toff, None, w)
        for nm in CMP + ["n_diag"]:
            assert np.array_equal(ra[nm], rb[nm]), (kind, nm)
        for x, y in zip(ga, gb):
            assert np.array_equal(x, y)
        sel = np.argsort(-tl)[::40]
        qs = [qcat[qoff[i]:qoff[i + 1]] for i in sel]; ts = [tcat[toff[i]:toff[i + 1]] for i in sel]
        exp, ecig, _ = H.run_cpu("ref" if H.have_ref() else "oracle", H.make_params(kind, H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=-1, zdrop=400, flag=fl),
                                 qs, ts, nthreads=8, w=w[sel])
        for nm in CMP:
            assert np.array_equal(ra[nm][sel], exp[:, H.FIELDS.index(nm)]), (kind, nm)
        for i, c_ in zip(sel, ecig):
            assert np.array_equal(ga[int(i)], c_)
    ca.close(); cb.close()
