"""tools/ksw2b-test: the reference CLI's options and output format (cli.c:141-260) over ONE GPU batch.

Golden stdout in tests/golden/cli_expected.json comes from the reference's own CLI (oracle/_ref/ksw2-test = cli.c compiled as-is,
scripts/make_golden.py) on the t1/q1 fixture pairs, with the NUL byte it prints for N_SKIP replaced by 'N'.  CPU tests pin the
fixture (reference CLI and oracle-derived lines reproduce it); the GPU test runs the real binary."""
import json
import os
import subprocess

import numpy as np
import pytest

import harness as H
import ksw2_b200 as K

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = json.load(open(os.path.join(GOLD, "cli_expected.json")))
SEQS = np.load(os.path.join(GOLD, "seqs.npz"))
REF_CLI = os.path.join(H.ORACLE_DIR, "_ref", "ksw2-test")


def write_fixture(d):
    for fn, pre in (("t.fa", "t"), ("q.fa", "q")):
        with open(os.path.join(d, fn), "w") as f:
            for i in range(5):
                f.write(f">{pre}{i + 1} some description\n{''.join('ACGTN'[int(x)] for x in SEQS[f'{pre}1_{i}'])}\n")
    return os.path.join(d, "t.fa"), os.path.join(d, "q.fa")


def oracle_lines(args):
    """what ksw2b-test must print, computed with the CPU checker instead of the GPU (same option handling as tools/ksw2b_test.cpp)"""
    a, b, q, e, q2, e2, w, z, flag, pair, algo = 2, 4, 4, 2, 13, 1, -1, -1, 0, True, "extd"
    it = iter(args)
    for o in it:
        if o == "-t": algo = next(it)
        elif o == "-w": w = int(next(it))
        elif o == "-z": z = int(next(it))
        elif o == "-r": flag |= 2
        elif o == "-s": flag |= 1
        elif o == "-g": flag |= 0x18
        elif o == "-a": pair = False
        elif o == "-A": a = int(next(it))
        elif o == "-B": b = int(next(it))
        elif o == "-O":
            p = next(it).split(","); q = q2 = int(p[0]); q2 = int(p[1]) if len(p) > 1 else q2
        elif o == "-E":
            p = next(it).split(","); e = e2 = int(p[0]); e2 = int(p[1]) if len(p) > 1 else e2
    mat = H.simple_mat(5, a, b)
    kw = dict(q=q, e=e, q2=q2, e2=e2, w=w, zdrop=z, flag=flag)
    kind = {"gg": "gg", "gg2": "gg2", "gg2_sse": "gg2_sse", "extz": "extz", "extz2_sse": "extz2", "extd": "extd", "extd2_sse": "extd2",
            "extf2_sse": "extf2", "exts2_sse": "exts2", "test": "extd2"}[algo]
    if kind in ("gg", "gg2", "gg2_sse"):
        kw = dict(q=q, e=e, w=w, flag=(flag & 1) if kind != "gg2_sse" else 0)
    if kind == "extf2":
        kw = dict(q=int(mat[0]), q2=int(mat[1]), e=e, w=w, zdrop=z, flag=1)
    if kind == "exts2":
        mat = H.simple_mat(5, 1, 2); kw = dict(q=2, e=1, q2=32, noncan=4, zdrop=z, junc_bonus=0, flag=flag | 0x100)
    if algo == "test":
        kw = dict(q=4, e=2, q2=24, e2=1, w=751, zdrop=400, flag=8)
    P = H.make_params(kind, mat, **kw)
    prs = [(i, i) for i in range(5)] if pair else [(i, j) for j in range(5) for i in range(5)]
    res, cig, _ = H.run_cpu("oracle", P, [SEQS[f"q1_{j}"] for i, j in prs], [SEQS[f"t1_{i}"] for i, j in prs])
    out = ""
    for k, (i, j) in enumerate(prs):
        r = res[k]
        out += f"t{i + 1}\tq{j + 1}\t{r[8]}\t{r[0]}\t{r[3]}\t{r[2]}"
        if len(cig[k]):
            out += "\t" + "".join(f"{int(x) >> 4}{'MIDN___=X'[int(x) & 15]}" for x in cig[k])
        out += "\n"
    return out


def test_fixture_matches_the_oracle():
    for c in CASES:
        assert oracle_lines(c["args"]) == c["stdout"], c["args"]


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/ksw2-test not built (no /root/reference on this box)")
def test_reference_cli_reproduces_the_fixture(tmp_path):
    t, q = write_fixture(str(tmp_path))
    for c in CASES:
        out = subprocess.run([REF_CLI] + c["args"] + [t, q], capture_output=True).stdout.replace(b"\0", b"N").decode()
        assert out == c["stdout"], c["args"]


def test_cli_builds_and_prints_usage():
    exe = K.build_cli()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage: ksw2b-test" in r.stderr and "extz2_sse" in r.stderr


@pytest.mark.gpu
def test_cli_output_equals_reference_cli(tmp_path):
    exe = K.build_cli()
    t, q = write_fixture(str(tmp_path))
    for c in CASES[::3] + CASES[-8:]:          # (every process start pays ~1 s of CUDA context creation: a third of the cases + the option mixes)
        r = subprocess.run([exe] + c["args"] + [t, q], capture_output=True, text=True)
        assert r.returncode == 0, (c["args"], r.stderr)
        assert r.stdout == c["stdout"], c["args"]
    # gzip input, literal sequences (cli.c:216-218), -R
    subprocess.check_call(["gzip", "-k", t])
    r = subprocess.run([exe, "-t", "extz2_sse", "-R", "2", t + ".gz", q], capture_output=True, text=True)
    assert r.stdout == next(c for c in CASES if c["args"] == ["-t", "extz2_sse"])["stdout"]
    r = subprocess.run([exe, "-t", "extz2_sse", "ATAGCTAGCTAGCAT", "AGCTAcCGCAT"], capture_output=True, text=True)
    assert r.stdout.startswith("first\tsecond\t")
