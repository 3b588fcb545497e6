"""CPU tests that pin the oracle (oracle/ksw2_oracle.c):

 1. against the golden vectors in tests/golden/ (produced by the unmodified reference, see
    scripts/make_golden.py) -- all 11 ksw_extz_t fields + CIGAR;
 2. against the anchors SURVEY.md Appendix B recorded from the reference's own CLI (hard-coded
    below, independent of our fixture generator);
 3. differentially against oracle/_ref/libksw2_ref.so (the reference compiled as-is) over the
    fuzz domain of SURVEY.md Appendix A.8 -- skipped where that build is absent.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import fuzzgen as F
import harness as H

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SEQS = np.load(os.path.join(GOLD, "seqs.npz"))
CASES = json.load(open(os.path.join(GOLD, "expected.json")))
SLOW = os.environ.get("KSW2_SLOW", "0") == "1"


def cli_text(c):
    return "".join(str(int(x) >> 4) + "MID\0"[int(x) & 0xf] for x in c)


def is_heavy(c):
    """un-banded 16.5 kb / 50 kb cases: seconds to minutes on the scalar oracle"""
    w = c["params"].get("w", -1)
    n = len(SEQS[c["t"]])
    if c["kind"] in ("extz", "extd"):           # the row-wise restatement is a plain int32 loop: a few seconds at most
        return False
    wide = c["kind"] == "exts2" or w < 0 or w > 1000
    return wide and (n > 20000 or (n > 10000 and not (c["params"].get("flag", 0) & 1) and c["name"] not in ("mt_extz2", "mt_exts2")))


def check_case(which, c):
    P = H.make_params(c["kind"], H.simple_mat(5, *c["mat"]), **c["params"])
    res, cig, _ = H.run_cpu(which, P, [SEQS[c["q"]]], [SEQS[c["t"]]])
    got = {k: int(v) for k, v in zip(H.FIELDS, res[0])}
    exp = dict(c["fields"])
    got.pop("m_cigar"); exp.pop("m_cigar")      # capacity depends on the allocator's history, not on the path
    assert got == exp, c["name"]
    if c["cigar_md5"] is not None:
        assert hashlib.md5((cli_text(cig[0]) + "\n").encode("latin1")).hexdigest() == c["cigar_md5"], c["name"]
    if "cigar" in c:
        assert H.cigar_str(cig[0]) == c["cigar"]


@pytest.mark.parametrize("c", [c for c in CASES if not is_heavy(c)], ids=lambda c: c["name"])
def test_oracle_golden(c):
    check_case("oracle", c)


@pytest.mark.slow
@pytest.mark.skipif(not SLOW, reason="set KSW2_SLOW=1")
@pytest.mark.parametrize("c", [c for c in CASES if is_heavy(c)], ids=lambda c: c["name"])
def test_oracle_golden_heavy(c):
    check_case("oracle", c)


# SURVEY.md Appendix B (captured from the reference's ksw2-test binary): score, max, max_t, max_q, md5
ANCHORS = {
    "mt_extz2": (16102, 17054, 16568, 16024, "ea0524d904ed8c1922b9d108bbc95724"),
    "mt_extz2_r": (16102, 17054, 16568, 16024, "db8b671f4dbfa2bd53cf755495898dfd"),
    "mt_extd2": (17127, 17614, 16568, 16024, "df0e77e43f48cfd00968cdac58da02d3"),
    "mt_extd2_r": (17127, 17614, 16568, 16024, "8e2c9cfb877ab8354ec1cbbbadb967da"),
    "mt_exts2": (8591, 9067, 16568, 16024, "76b16ab722b3e4ac48540fbe0a986269"),
    "p50_extz2": (69932, 70010, 49962, 49999, "6804b572b6415e349f18871eba2c0bc0"),
    "p50_extz2_r": (69932, 70010, 49962, 49999, "60eb8c9ac804ce9ba09147ac416fc470"),
    "p50_extd2": (70098, 70148, 49962, 49999, "f82abb282ef4195d7780990fc3198953"),
    "p50_extd2_r": (70098, 70148, 49962, 49999, "6c010e7765a00b070f3e019dbee76b12"),
    "p50_extz2_w500_z400": (69932, 70010, 49962, 49999, "6804b572b6415e349f18871eba2c0bc0"),
    "p50_extz2_w10": (-31264, 1554, 1102, 1101, None),
    "p50_extz2_w30": (-226, 7178, 35781, 35811, None),
    "p50_extz2_w64": (64170, 64248, 49962, 49999, None),
    "p50_extd2_w10": (-30826, 1555, 1102, 1101, None),
    "p50_extd2_w30": (-104, 7204, 35781, 35811, None),
    "p50_extd2_w64": (64314, 64364, 49962, 49999, None),
    "t5_regression_extz2": (-30, 5, 4, 4, None),
    "t1_4_extz2": (-46, 10, 4, 4, None),
    "t1_4_extd2": (-19, 10, 4, 4, None),
    "t1_2_extz2": (12, 48, 35, 33, None),
    "t1_2_extd2": (14, 48, 35, 33, None),
    "readme_extz2": (0, 2, 0, 0, None),
}


def test_fixture_matches_survey_anchors():
    by = {c["name"]: c for c in CASES}
    for name, (score, mx, mt, mq, md5) in ANCHORS.items():
        f = by[name]["fields"]
        assert (f["score"], f["max"], f["max_t"], f["max_q"]) == (score, mx, mt, mq), name
        if md5:
            assert by[name]["cigar_md5"] == md5, name
    r = by["readme_extz2"]
    assert r["cigar"] == "2D7M2D4M" and r["fields"]["mte_q"] == 9 and r["fields"]["mqe_t"] == 14   # rounded-en quirk
    assert by["t1_2_extz2"]["cigar"] == "5M2D27M6D7M2D4M3D3M3D2M2D6M"
    assert by["t1_2_extd2"]["cigar"] == "5M2D28M19D4M3I2M2I4M2D6M"
    assert by["p50_extz2_w500_z50"]["fields"]["score"] == H.C.c_int32(-0x40000000).value


def fuzz_batches(seed, n_iter):
    rng = np.random.default_rng(seed)
    for it in range(n_iter):
        kind = ["extz2", "extd2", "exts2"][it % 3]
        a, b = F.AB[rng.integers(len(F.AB))]
        if kind == "exts2":
            prs = [F.splice_pair(rng) for _ in range(4)]
            q, e, q2, nc = F.SPL[rng.integers(len(F.SPL))]
            mat = H.simple_mat(5, 1, 2) if rng.random() < 0.5 else H.simple_mat(5, a, b)
            P = H.make_params(kind, mat, q=q, e=e, q2=q2, noncan=nc, zdrop=int(rng.choice(F.ZD)),
                              junc_bonus=int(rng.choice([0, 3, 9])), flag=int(rng.choice(F.SFLAGS)))
            js = None
            if rng.random() < 0.5:
                js = [(p[2] if p[2] is not None else np.zeros(len(p[1]), np.uint8)) for p in prs]
            yield P, [p[0] for p in prs], [p[1] for p in prs], js
        else:
            prs = [F.rand_pair(rng) for _ in range(4)]
            mat = H.simple_mat(5, a, b, 0 if rng.random() < 0.7 else -1)
            fl = int(rng.choice(F.FLAGS))
            if kind == "extz2":
                q, e = F.QE[rng.integers(len(F.QE))]
                P = H.make_params(kind, mat, q=q, e=e, w=int(rng.choice(F.WS)), zdrop=int(rng.choice(F.ZD)),
                                  end_bonus=int(rng.choice(F.EB)), flag=fl)
            else:
                q, e, q2, e2 = F.DUAL[rng.integers(len(F.DUAL))]
                P = H.make_params(kind, mat, q=q, e=e, q2=q2, e2=e2, w=int(rng.choice(F.WS)), zdrop=int(rng.choice(F.ZD)),
                                  end_bonus=int(rng.choice(F.EB)), flag=fl)
            yield P, [p[0] for p in prs], [p[1] for p in prs], None


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_vs_reference_fuzz():
    n = 0
    for P, qs, ts, js in fuzz_batches(20260925, 600):
        a = H.run_cpu("ref", P, qs, ts, js)
        b = H.run_cpu("oracle", P, qs, ts, js)
        assert np.array_equal(a[0][:, :11], b[0][:, :11]), (P.kind, hex(P.flag), P.w, P.zdrop)
        for x, y in zip(a[1], b[1]):
            assert np.array_equal(x, y)
        n += len(qs)
    assert n == 2400


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_rows_oracle_vs_reference_fuzz():
    """kso_extz / kso_extd == the reference's ksw_extz / ksw_extd (ksw2_extz.c, ksw2_extd.c) inside the reference's defined domain"""
    n = 0
    for kind, mat, kw, qs, ts in F.rows_batches(20261017, 500):
        P = H.make_params(kind, mat, **kw)
        a = H.run_cpu("ref", P, qs, ts)
        b = H.run_cpu("oracle", P, qs, ts)
        assert np.array_equal(a[0][:, :11], b[0][:, :11]), (kind, kw)
        for x, y in zip(a[1], b[1]):
            assert np.array_equal(x, y), (kind, kw)
        n += len(qs)
    assert n == 2000


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_extf2_oracle_vs_reference_fuzz():
    """kso_extf2 == the reference's ksw_extf2_sse (ksw2_extf2_sse.c), SSE4.1 build"""
    n = 0
    mat = H.simple_mat(5, 2, 4)
    for kw, qs, ts in F.extf2_batches(20261018, 500):
        P = H.make_params("extf2", mat, **kw)
        a = H.run_cpu("ref", P, qs, ts)
        b = H.run_cpu("oracle", P, qs, ts)
        assert np.array_equal(a[0][:, :11], b[0][:, :11]), kw
        n += len(qs)
    assert n == 2000


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_gg_oracle_vs_reference_fuzz():
    """kso_gg == the reference's ksw_gg (ksw2_gg.c): score and CIGAR"""
    n = 0
    for kind, mat, kw, qs, ts in F.gg_batches(20261019, 400):
        P = H.make_params(kind, mat, **kw)
        a = H.run_cpu("ref", P, qs, ts)
        b = H.run_cpu("oracle", P, qs, ts)
        assert np.array_equal(a[0][:, :11], b[0][:, :11]), (kind, kw)
        for x, y in zip(a[1], b[1]):
            assert np.array_equal(x, y), (kind, kw)
        n += len(qs)
    assert n == 1600


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_gg2_oracle_vs_reference_fuzz():
    """kso_gg2 / kso_gg2_sse == the reference's ksw_gg2 / ksw_gg2_sse (ksw2_gg2.c, ksw2_gg2_sse.c).  ksw_gg2_sse needs its CIGAR
    pointers (it dereferences them unconditionally, :123) and reads UNINITIALISED heap when the traceback leaves the band on the
    right (kmalloc'ed matrix, off_end == NULL): such pairs are recognised by the reference disagreeing with ITSELF when it is
    run again with a different allocator history, and are outside the parity domain."""
    n = unstable = 0
    for kind, mat, kw, qs, ts in F.gg_batches(20261020, 400, kinds=("gg2", "gg2_sse")):
        if kind == "gg2_sse":
            kw["flag"] = 0
        P = H.make_params(kind, mat, **kw)
        a = H.run_cpu("ref", P, qs, ts)
        b = H.run_cpu("oracle", P, qs, ts)
        for i in range(len(qs)):
            n += 1
            if np.array_equal(a[0][i, :11], b[0][i, :11]) and np.array_equal(a[1][i], b[1][i]):
                continue
            assert kind == "gg2_sse" and a[0][i, 8] == b[0][i, 8], (kind, kw, i)          # the score never depends on it
            a2 = H.run_cpu("ref", P, [ts[i][::-1].copy(), qs[i]], [qs[i][::-1].copy(), ts[i]])
            a1 = H.run_cpu("ref", P, [qs[i]], [ts[i]])
            assert not (np.array_equal(a1[1][0], a[1][i]) and np.array_equal(a2[1][1], a[1][i])), (kind, kw, i)
            unstable += 1
    assert n == 1600 and unstable < 40


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")
def test_reference_golden_still_reproduces():
    """the fixture generator and the reference build agree today (guards against a stale fixture)"""
    for c in CASES:
        if len(SEQS[c["t"]]) < 1000:
            check_case("ref", c)
