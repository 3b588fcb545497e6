"""CPU tests of the DEVICE ENGINE SOURCE (ksw2_b200/csrc/ksw2_tile.cuh + ksw2_pair.cuh) built for the host
(tests/sim/ksw2_sim.cpp): the same per-thread code the GPU runs, executed sequentially, compared with the
oracle over the SURVEY.md A.8 fuzz domain with random panel heights and both score-row modes, and with the
golden vectors.  (The GPU parity tests proper are in test_gpu_parity.py; this keeps the engine honest on
machines without a GPU.  The product library never links the host build.)
"""
import hashlib

import numpy as np
import pytest

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
        compare(P, qs, ts, js, int(rng.choice([1, 2, 3, 5, 7, 16, 32, 64, 1000])), int(rng.integers(0, 2)))
        n += 1
    assert n == 450          # incl. KSW_EZ_APPROX_MAX cases (ksw2_scalar.cuh path)


GOLD_SIM = ["t1_0_extz2", "t1_1_extd2", "t1_2_extz2", "t1_2_extd2", "t1_3_extz2", "t1_4_extd2", "t5_regression_extz2", "readme_extz2",
            "mt_extz2", "mt_extd2_r", "mt_exts2", "p50_extz2_w500_z400", "p50_extd2_w64", "p50_extz2_w500_z50", "mt_extz2_w20",
            "p50_extz2_w10", "p50_extd2_w30", "p50_extz2_w100"]


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
