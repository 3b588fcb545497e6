#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (run in the build container only).

Inputs : the reference's own test sequences (/root/reference/test/*.fa[.gz]; data, not source),
         nt4-encoded exactly as the reference CLI does (cli.c:17-34: A,C,G,T -> 0..3, else 4).
Outputs: tests/golden/seqs.npz        encoded sequences (so the tests never read /root/reference)
         tests/golden/expected.json   per case: parameters + all ksw_extz_t fields + CIGAR string + md5
The expected values come from oracle/_ref/libksw2_ref.so (reference compiled as-is, SSE4.1).
The md5 is over the CLI's text form (cut -f7 | md5sum), so it can be compared with the values
SURVEY.md Appendix B recorded from the reference's own ksw2-test binary.
"""
import gzip, hashlib, json, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402

REF_TEST = "/root/reference/test"
NT4 = np.full(256, 4, np.uint8)
for i, ch in enumerate("ACGT"):
    NT4[ord(ch)] = i
    NT4[ord(ch.lower())] = i


def read_fa(path):
    op = gzip.open if path.endswith(".gz") else open
    recs, name, buf = [], None, []
    with op(path, "rt") as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if name is not None:
                    recs.append((name, "".join(buf)))
                name, buf = line[1:].split()[0], []
            elif line:
                buf.append(line)
    if name is not None:
        recs.append((name, "".join(buf)))
    return recs


def enc(s):
    return NT4[np.frombuffer(s.encode(), dtype=np.uint8)]


def cli_cigar_text(c):
    # cli.c:141-142 prints "MID"[op]; op 3 (N) indexes the string terminator -> a NUL byte
    return "".join(str(int(x) >> 4) + "MID\0"[int(x) & 0xf] for x in c)


def main():
    H.build_oracle()
    assert H.have_ref(), "needs /root/reference"
    seqs = {}
    t1 = read_fa(f"{REF_TEST}/t1.fa"); q1 = read_fa(f"{REF_TEST}/q1.fa")
    for i, ((_, t), (_, q)) in enumerate(zip(t1, q1)):
        seqs[f"t1_{i}"] = enc(t); seqs[f"q1_{i}"] = enc(q)
    seqs["mt_t"] = enc(read_fa(f"{REF_TEST}/MT-human.fa")[0][1]); seqs["mt_q"] = enc(read_fa(f"{REF_TEST}/MT-orang.fa")[0][1])
    seqs["p50_t"] = enc(read_fa(f"{REF_TEST}/t2.fa.gz")[0][1]); seqs["p50_q"] = enc(read_fa(f"{REF_TEST}/q2.fa.gz")[0][1])
    seqs["readme_t"] = enc("ATAGCTAGCTAGCAT"); seqs["readme_q"] = enc("AGCTAcCGCAT")
    os.makedirs(f"{ROOT}/tests/golden", exist_ok=True)
    np.savez_compressed(f"{ROOT}/tests/golden/seqs.npz", **seqs)

    mat24 = H.simple_mat(5, 2, 4); mat12 = H.simple_mat(5, 1, 2)
    cases = []

    def add(name, kind, tkey, qkey, mat=(2, 4), store_cigar=True, **kw):
        m = mat24 if mat == (2, 4) else H.simple_mat(5, *mat)
        P = H.make_params(kind, m, **kw)
        res, cig, _ = H.run_cpu("ref", P, [seqs[qkey]], [seqs[tkey]])
        txt = cli_cigar_text(cig[0])
        c = dict(name=name, kind=kind, t=tkey, q=qkey, mat=list(mat), params=kw,
                 fields={k: int(v) for k, v in zip(H.FIELDS, res[0])},
                 cigar_md5=hashlib.md5((txt + "\n").encode("latin1")).hexdigest() if len(cig[0]) else None)
        if store_cigar and len(cig[0]) <= 4000:
            c["cigar"] = H.cigar_str(cig[0])
        cases.append(c)

    cli_z = dict(q=4, e=2, w=-1, zdrop=-1, end_bonus=0, flag=0)
    cli_d = dict(q=4, e=2, q2=13, e2=1, w=-1, zdrop=-1, end_bonus=0, flag=0)
    for i in range(5):
        add(f"t1_{i}_extz2", "extz2", f"t1_{i}", f"q1_{i}", **cli_z)
        add(f"t1_{i}_extd2", "extd2", f"t1_{i}", f"q1_{i}", **cli_d)
    add("t5_regression_extz2", "extz2", "t1_4", "q1_4", mat=(1, 9), q=16, e=1, w=10, zdrop=-1, end_bonus=0, flag=0)   # test/t1.fa:9
    add("readme_extz2", "extz2", "readme_t", "readme_q", **cli_z)
    for fl, tag in ((0, ""), (2, "_r")):
        add(f"mt_extz2{tag}", "extz2", "mt_t", "mt_q", **{**cli_z, "flag": fl})
        add(f"mt_extd2{tag}", "extd2", "mt_t", "mt_q", **{**cli_d, "flag": fl})
        add(f"p50_extz2{tag}", "extz2", "p50_t", "p50_q", **{**cli_z, "flag": fl})
        add(f"p50_extd2{tag}", "extd2", "p50_t", "p50_q", **{**cli_d, "flag": fl})
    add("mt_exts2", "exts2", "mt_t", "mt_q", mat=(1, 2), q=2, e=1, q2=32, noncan=4, zdrop=-1, junc_bonus=0, flag=0x100)  # cli.c:79-83
    add("p50_extz2_w500_z400", "extz2", "p50_t", "p50_q", **{**cli_z, "w": 500, "zdrop": 400})
    for w in (10, 30, 64, 100):
        add(f"p50_extz2_w{w}", "extz2", "p50_t", "p50_q", **{**cli_z, "w": w, "flag": 1})
        add(f"p50_extd2_w{w}", "extd2", "p50_t", "p50_q", **{**cli_d, "w": w, "flag": 1})
    add("p50_extz2_w500_z50", "extz2", "p50_t", "p50_q", **{**cli_z, "w": 500, "zdrop": 50})
    add("p50_extd2_w500_z50", "extd2", "p50_t", "p50_q", **{**cli_d, "w": 500, "zdrop": 50})
    add("mt_extz2_w20", "extz2", "mt_t", "mt_q", **{**cli_z, "w": 20})
    add("p50_extz2_sg", "extz2", "p50_t", "p50_q", **{**cli_z, "flag": 0x09})
    add("p50_extd2_sg", "extd2", "p50_t", "p50_q", **{**cli_d, "flag": 0x09})
    add("mt_extd2_42241_w751_z400_approx", "extd2", "mt_t", "mt_q", q=4, e=2, q2=24, e2=1, w=751, zdrop=400, end_bonus=0, flag=8)  # cli.c "test" algo
    # the row-wise entry points ksw_extz / ksw_extd (reference ksw2_extz.c, ksw2_extd.c; no end_bonus argument)
    row_z = dict(q=4, e=2, w=-1, zdrop=-1, flag=0)
    row_d = dict(q=4, e=2, q2=13, e2=1, w=-1, zdrop=-1, flag=0)
    for i in range(5):
        add(f"t1_{i}_extz", "extz", f"t1_{i}", f"q1_{i}", **row_z)
        add(f"t1_{i}_extd", "extd", f"t1_{i}", f"q1_{i}", **row_d)
    add("readme_extz", "extz", "readme_t", "readme_q", **row_z)
    add("mt_extz", "extz", "mt_t", "mt_q", **row_z)
    add("mt_extd_r", "extd", "mt_t", "mt_q", **{**row_d, "flag": 2})
    add("mt_extz_w100_z200", "extz", "mt_t", "mt_q", **{**row_z, "w": 100, "zdrop": 200})
    add("mt_extd_w751_z400_x", "extd", "mt_t", "mt_q", q=4, e=2, q2=24, e2=1, w=751, zdrop=400, flag=0x40)
    add("p50_extz_w500_s", "extz", "p50_t", "p50_q", **{**row_z, "w": 500, "flag": 1})
    add("p50_extd_w500", "extd", "p50_t", "p50_q", **{**row_d, "w": 500})
    # global-alignment entry points ksw_gg / ksw_gg2 / ksw_gg2_sse (score + CIGAR) and ksw_extf2_sse (q = mch, q2 = mis, zdrop = xdrop)
    for i in range(5):
        for kind in ("gg", "gg2", "gg2_sse"):
            add(f"t1_{i}_{kind}", kind, f"t1_{i}", f"q1_{i}", q=4, e=2, w=-1, flag=0)
        add(f"t1_{i}_extf2", "extf2", f"t1_{i}", f"q1_{i}", q=2, q2=-4, e=2, w=-1, zdrop=-1, flag=1)
    add("mt_gg_w200", "gg", "mt_t", "mt_q", q=4, e=2, w=200, flag=0)
    add("mt_gg2_w200", "gg2", "mt_t", "mt_q", q=4, e=2, w=200, flag=0)
    # (no mt_gg2_sse_w200: on that input the reference's traceback leaves the band and reads its kmalloc'ed, never-written matrix --
    #  the CIGAR changes with the allocator's history; see tests/test_oracle.py::test_gg2_oracle_vs_reference_fuzz)
    add("mt_extf2_w300_x100", "extf2", "mt_t", "mt_q", q=2, q2=-4, e=2, w=300, zdrop=100, flag=1)
    add("p50_extf2_w500", "extf2", "p50_t", "p50_q", q=1, q2=-2, e=1, w=500, zdrop=-1, flag=1)
    with open(f"{ROOT}/tests/golden/expected.json", "w") as f:
        json.dump(cases, f, indent=1)
    print(f"wrote {len(cases)} cases")


# (ksw_extz / ksw_extd with a band narrower than |tlen - qlen| read unwritten heap in the reference: -w 60 for them, not -w 10)
CLI_CASES = [["-t", a] + o for a in ("gg", "gg2", "gg2_sse", "extz", "extz2_sse", "extd", "extd2_sse", "extf2_sse", "exts2_sse", "test")
             for o in ([], ["-w", "60" if a in ("extz", "extd") else "10"], ["-s"])] + \
            [["-t", "extz2_sse", "-r"], ["-t", "extd2_sse", "-r", "-z", "30"], ["-t", "extz2_sse", "-z", "20", "-w", "20"], ["-t", "extd2_sse", "-a"],
             ["-t", "extz2_sse", "-g"], ["-t", "extd2_sse", "-O", "6,20", "-E", "3,1"], ["-t", "extz", "-A", "1", "-B", "3", "-a"], ["-t", "extd", "-z", "25"]]


def write_fasta(path, names, seqs):
    with open(path, "w") as f:
        for n, s in zip(names, seqs):
            f.write(f">{n}\n{''.join('ACGTN'[int(x)] for x in s)}\n")


def make_cli_golden():
    """expected stdout of the reference's own CLI (oracle/_ref/ksw2-test = cli.c compiled as-is) on the t1/q1 fixture pairs"""
    import subprocess, tempfile
    seqs = np.load(f"{ROOT}/tests/golden/seqs.npz")
    d = tempfile.mkdtemp()
    write_fasta(f"{d}/t.fa", [f"t{i + 1}" for i in range(5)], [seqs[f"t1_{i}"] for i in range(5)])
    write_fasta(f"{d}/q.fa", [f"q{i + 1}" for i in range(5)], [seqs[f"q1_{i}"] for i in range(5)])
    exe = f"{ROOT}/oracle/_ref/ksw2-test"
    out = []
    for args in CLI_CASES:
        if args[1] == "gg2_sse" and "-s" in args:
            continue
        # the reference prints "MID"[op]: a NUL byte for N_SKIP (cli.c:148); ksw2b-test prints 'N'
        txt = subprocess.run([exe] + args + [f"{d}/t.fa", f"{d}/q.fa"], capture_output=True).stdout.replace(b"\0", b"N").decode()
        out.append(dict(args=args, stdout=txt))
    with open(f"{ROOT}/tests/golden/cli_expected.json", "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {len(out)} CLI cases")


if __name__ == "__main__":
    main()
    make_cli_golden()
