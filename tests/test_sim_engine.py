"""CPU tests of the DEVICE ENGINE SOURCE (ksw2_b200/csrc/ksw2_tile.cuh + ksw2_pair.cuh) built for the host
(tests/sim/ksw2_sim.cpp): the same per-thread code the GPU runs, executed sequentially, compared with the
oracle over the SURVEY.md A.8 fuzz domain with random panel heights and both score-row modes, and with the
golden vectors.  (The GPU parity tests proper are in test_gpu_parity.py; this keeps the engine honest on
machines without a GPU.  The product library never links the host build.)
"""
import hashlib

import numpy as np
import pytest

import fuzzgen as F
import harness as H
from test_oracle import CASES, SEQS, cli_text, fuzz_batches


def compare(P, qs, ts, js, panel, fs):
    a = H.run_cpu("oracle", P, qs, ts, js)
    b = H.run_sim(P, qs, ts, js, panel=panel, force_smode=fs)
    assert np.array_equal(a[0][:, :11], b[0][:, :11]), (P.kind, hex(P.flag), P.w, P.zdrop, panel, fs)
    for x, y in zip(a[1], b[1]):
        assert np.array_equal(x, y), (P.kind, hex(P.flag), P.w, P.zdrop, panel, fs)


def test_engine_fuzz_vs_oracle():
    rng = np.random.default_rng(5)
    n = 0
    for P, qs, ts, js in fuzz_batches(77, 450):
        compare(P, qs, ts, js, int(rng.choice([1, 2, 3, 5, 7, 16, 32, 64, 1000])), int(rng.integers(0, 4)))
        n += 1
    assert n == 450          # incl. KSW_EZ_APPROX_MAX cases (the tracker of the tile engine, ks_apx_step)


def test_engine_warp_driver_fuzz():
    """the warp-cooperative driver (ks_pair_fill_warp), simulated lane by lane: negative panel = warp mode with that panel height"""
    rng = np.random.default_rng(6)
    n = 0
    for P, qs, ts, js in fuzz_batches(78, 150):
        compare(P, qs, ts, js, -int(rng.choice([1, 2, 5, 16, 33, 128])), int(rng.integers(0, 4)))     # incl. KSW_EZ_APPROX_MAX: the tracker in ks_apx_step
        n += 1
    assert n == 150


GOLD_SIM = ["t1_0_extz2", "t1_1_extd2", "t1_2_extz2", "t1_2_extd2", "t1_3_extz2", "t1_4_extd2", "t5_regression_extz2", "readme_extz2",
            "mt_extz2", "mt_extd2_r", "mt_exts2", "p50_extz2_w500_z400", "p50_extd2_w64", "p50_extz2_w500_z50", "mt_extz2_w20",
            "p50_extz2_w10", "p50_extd2_w30", "p50_extz2_w100",
            "t1_2_extz", "t1_2_extd", "t1_4_extz", "readme_extz", "mt_extz_w100_z200", "mt_extd_w751_z400_x", "p50_extz_w500_s",
            "t1_2_gg", "t1_2_gg2", "t1_2_gg2_sse", "t1_2_extf2", "t1_4_extf2", "mt_gg_w200", "mt_gg2_w200", "mt_extf2_w300_x100", "p50_extf2_w500"]


@pytest.mark.parametrize("name", GOLD_SIM)
def test_engine_golden(name):
    c = {c["name"]: c for c in CASES}[name]
    P = H.make_params(c["kind"], H.simple_mat(5, *c["mat"]), **c["params"])
    res, cig = H.run_sim(P, [SEQS[c["q"]]], [SEQS[c["t"]]], panel=48)
    got = {k: int(v) for k, v in zip(H.FIELDS, res[0])}
    exp = dict(c["fields"])
    got.pop("m_cigar"); exp.pop("m_cigar")
    assert got == exp
    if c["cigar_md5"] is not None:
        assert hashlib.md5((cli_text(cig[0]) + "\n").encode("latin1")).hexdigest() == c["cigar_md5"]


def test_engine_edge_lengths():
    """qlen/tlen of 1, multiples of 16 +-1, band exactly |tlen-qlen|, band 0"""
    rng = np.random.default_rng(11)
    mat = H.simple_mat(5, 2, 4)
    for ql, tl in [(1, 1), (1, 40), (40, 1), (16, 16), (17, 15), (15, 17), (32, 33), (48, 33), (33, 48), (64, 1), (2, 130)]:
        q = rng.integers(0, 4, ql).astype(np.uint8); t = rng.integers(0, 4, tl).astype(np.uint8)
        for w in (-1, 0, 1, abs(tl - ql), abs(tl - ql) + 1, 15, 16):
            for kind in ("extz2", "extd2"):
                for fl in (0, 1, 2, 0x40, 0x42, 0x80):
                    P = H.make_params(kind, mat, w=w, zdrop=30, flag=fl)
                    compare(P, [q], [t], None, 3, 0)


def eqx_batches(seed, n_iter):
    """extd2 + KSW_EZ_EQX (0x800) over the fuzz domain; exact-max flags only keep the case count small"""
    rng = np.random.default_rng(seed)
    import fuzzgen as F
    for it in range(n_iter):
        prs = [F.rand_pair(rng) for _ in range(4)]
        a, b = F.AB[rng.integers(len(F.AB))]
        q, e, q2, e2 = F.DUAL[rng.integers(len(F.DUAL))]
        fl = int(rng.choice([0, 2, 0x40, 0x42, 0x80, 0x82, 0xc0, 0x08])) | 0x800
        P = H.make_params("extd2", H.simple_mat(5, a, b), q=q, e=e, q2=q2, e2=e2, w=int(rng.choice(F.WS)), zdrop=int(rng.choice(F.ZD)),
                          end_bonus=int(rng.choice(F.EB)), flag=fl)
        yield P, [p[0] for p in prs], [p[1] for p in prs]


def check_eqx_semantics(P, qs, ts, cigs):
    """KSW_EZ_EQX pinned by its definition (the reference's own post-pass corrupts the heap): collapsing =/X gives the
    plain CIGAR of the same call without the flag, and every '=' / 'X' base compares equal / different at the position counted
    from the first op (ksw2.h:163-182)."""
    P0 = H.KsdParams(P.kind, P.m, P.mat, P.q, P.e, P.q2, P.e2, P.w, P.zdrop, P.end_bonus, P.flag & ~0x800, P.noncan, P.junc_bonus)
    _, plain, _ = H.run_cpu("oracle", P0, qs, ts)
    for q, t, c, c0 in zip(qs, ts, cigs, plain):
        x = y = 0
        merged = []
        for w in c.tolist():
            op, ln = w & 15, w >> 4
            assert op != 0
            if op in (7, 8):
                eq = t[x:x + ln] == q[y:y + ln]
                assert eq.all() if op == 7 else (~eq).all()
                x += ln; y += ln; op = 0
            elif op in (2, 3):
                x += ln
            else:
                y += ln
            if merged and merged[-1][0] == op:
                merged[-1][1] += ln
            else:
                merged.append([op, ln])
        assert [(o, l) for o, l in merged] == [(w & 15, w >> 4) for w in c0.tolist()]


def test_engine_eqx():
    n = 0
    for P, qs, ts in eqx_batches(31, 60):
        a = H.run_cpu("oracle", P, qs, ts)
        check_eqx_semantics(P, qs, ts, a[1])
        b = H.run_sim(P, qs, ts, None, panel=7, force_smode=0)
        assert np.array_equal(a[0][:, :11], b[0][:, :11]), hex(P.flag)
        for x, y in zip(a[1], b[1]):
            assert np.array_equal(x, y), hex(P.flag)
        n += 1
    assert n == 60


def test_rows_engine_fuzz():
    """ksw_extz / ksw_extd device functions (ksw2_rows.cuh, host build) == oracle restatement"""
    n = 0
    for kind, mat, kw, qs, ts in F.rows_batches(909, 300):
        P = H.make_params(kind, mat, **kw)
        exp, ecig, _ = H.run_cpu("oracle", P, qs, ts)
        res, cig = H.run_sim(P, qs, ts)
        assert np.array_equal(res[:, :11], exp[:, :11]), (kind, kw)
        for a, b in zip(cig, ecig):
            assert np.array_equal(a, b), (kind, kw)
        n += len(qs)
    assert n == 1200


def test_extf2_engine_fuzz():
    """ksw_extf2_sse device function (ksw2_extf2.cuh, host build) == oracle restatement"""
    mat = H.simple_mat(5, 2, 4)
    n = 0
    for kw, qs, ts in F.extf2_batches(4711, 300):
        P = H.make_params("extf2", mat, **kw)
        exp, _, _ = H.run_cpu("oracle", P, qs, ts)
        res, _ = H.run_sim(P, qs, ts)
        assert np.array_equal(res[:, :9], exp[:, :9]), kw
        n += len(qs)
    assert n == 1200


def test_gg_engine_fuzz():
    """ksw_gg (row-wise device function), ksw_gg2 / ksw_gg2_sse (ksw2_gg2.cuh) == oracle restatements"""
    n = 0
    for kind, mat, kw, qs, ts in F.gg_batches(1234, 300, kinds=("gg", "gg2", "gg2_sse")):
        P = H.make_params(kind, mat, **kw)
        exp, ecig, _ = H.run_cpu("oracle", P, qs, ts)
        res, cig = H.run_sim(P, qs, ts)
        assert np.array_equal(res[:, :11], exp[:, :11]), (kind, kw)
        if not (P.flag & 1):
            for a, b in zip(cig, ecig):
                assert np.array_equal(a, b), (kind, kw)
        n += len(qs)
    assert n == 1200


def test_engine_warp_driver_wide_bands():
    """warp driver on bands many blocks wide: most waves are all-interior, i.e. the warp-uniform fast step (ks_pair_fill_warp) is what runs"""
    rng = np.random.default_rng(2)
    mat = H.simple_mat(5, 2, 4)
    t = rng.integers(0, 4, 1500).astype(np.uint8)
    q = H.mutate(rng, t, sub=0.05, ins=0.02, dele=0.02)
    for kind, kw in (("extz2", dict(w=-1, zdrop=-1, flag=0)), ("extz2", dict(w=700, zdrop=300, flag=0x41)), ("extd2", dict(w=-1, zdrop=-1, flag=2)),
                     ("exts2", dict(q=2, e=1, q2=32, noncan=4, zdrop=-1, flag=0x100))):
        P = H.make_params(kind, mat if kind != "exts2" else H.simple_mat(5, 1, 2), **kw)
        for panel in (-24, -128):
            compare(P, [q, t[:1100]], [t, q], None, panel, 0)


def _ring_ok(P, qs, ts):
    """the ring schedule serves pairs whose effective band is at most 512 (KS_RING_MAX_W); exts2 has no band"""
    if P.kind == 2:
        return all(max(len(q), len(t)) <= 512 for q, t in zip(qs, ts))
    return all((P.w if 0 <= P.w <= max(len(q), len(t)) else max(len(q), len(t))) <= 512 for q, t in zip(qs, ts))


def test_engine_ring_schedule_fuzz():
    """one warp per pair on the ring schedule (ks_pair_fill_ring: block k on lane k & 31 for its whole life), simulated lane by lane:
    panel -200000 selects it; incl. approximate-max and the coded-target reload"""
    rng = np.random.default_rng(16)
    n = 0
    for P, qs, ts, js in fuzz_batches(79, 220):
        if not _ring_ok(P, qs, ts):
            continue
        compare(P, qs, ts, js, -200000, int(rng.integers(0, 4)))
        n += 1
    assert n > 150


def test_engine_ring_schedule_long_banded_pairs():
    """ring schedule on pairs whose blocks wrap around the 32 lanes many times, at the schedule's band limit (2w < 1026)"""
    rng = np.random.default_rng(17)
    qs, ts = [], []
    for L in (5000, 3300, 2100):
        t = rng.integers(0, 4, L).astype(np.uint8)
        q = H.mutate(rng, t, sub=0.03, ins=0.03, dele=0.03)
        qs.append(q); ts.append(t)
    qs.append(qs[0][:2500]); ts.append(ts[0])                      # the band runs empty before the end
    for kind, fl, w, zd in (("extd2", 0, 500, 400), ("extd2", 2, 512, 400), ("extz2", 0x41, 500, 100), ("extz2", 2, 511, -1), ("extd2", 0x18, 300, 200)):
        P = H.make_params(kind, H.simple_mat(5, 2, 4), q=4, e=2, q2=24, e2=1, w=w, zdrop=zd, flag=fl)
        compare(P, qs, ts, None, -200000, 0)


def test_engine_cta_wide_wavefront_fuzz():
    """the wavefront over more than one warp (ks_pair_fill_warp with NL = 64 lanes in the simulation; 256 on the device): panel -(100000 + C)"""
    rng = np.random.default_rng(18)
    n = 0
    for P, qs, ts, js in fuzz_batches(80, 120):
        compare(P, qs, ts, js, -(100000 + int(rng.choice([1, 7, 40, 200, 3000]))), int(rng.integers(0, 4)))
        n += 1
    assert n == 120
