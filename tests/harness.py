"""Test/bench harness (TEST INFRASTRUCTURE): ctypes access to the CPU checkers under oracle/.

  * oracle/libksw2_oracle.so   -- the in-repo C restatement (kso_ext{z,d,s}2)
  * oracle/_ref/libksw2_ref.so -- the unmodified reference compiled by oracle/Makefile (ksw_ext{z,d,s}2_sse)
  * oracle/libksw2_driver.so   -- multi-threaded batch runner over either (oracle/ref_driver.c)

Nothing here is imported by the product package (ksw2_b200/).
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_ORACLE = os.path.join(ORACLE_DIR, "libksw2_oracle.so")
LIB_REF = os.path.join(ORACLE_DIR, "_ref", "libksw2_ref.so")
LIB_DRIVER = os.path.join(ORACLE_DIR, "libksw2_driver.so")

FIELDS = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end", "m_cigar"]
NF = len(FIELDS)
KIND = {"extz2": 0, "extd2": 1, "exts2": 2, "extz": 3, "extd": 4, "extf2": 5, "gg": 6, "gg2": 7, "gg2_sse": 8}
SYM_REF = {0: b"ksw_extz2_sse", 1: b"ksw_extd2_sse", 2: b"ksw_exts2_sse", 3: b"ksw_extz", 4: b"ksw_extd", 5: b"ksw_extf2_sse", 6: b"ksw_gg", 7: b"ksw_gg2", 8: b"ksw_gg2_sse"}
SYM_ORACLE = {0: b"kso_extz2", 1: b"kso_extd2", 2: b"kso_exts2", 3: b"kso_extz", 4: b"kso_extd", 5: b"kso_extf2", 6: b"kso_gg", 7: b"kso_gg2", 8: b"kso_gg2_sse"}


class KsdParams(C.Structure):
    _fields_ = [("kind", C.c_int), ("m", C.c_int), ("mat", C.POINTER(C.c_int8)),
                ("q", C.c_int), ("e", C.c_int), ("q2", C.c_int), ("e2", C.c_int),
                ("w", C.c_int), ("zdrop", C.c_int), ("end_bonus", C.c_int), ("flag", C.c_int),
                ("noncan", C.c_int), ("junc_bonus", C.c_int)]


def build_oracle():
    """(Re)build the checkers; the reference build only happens where /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"], stdout=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(LIB_REF)


_driver = None


def driver():
    global _driver
    if _driver is None:
        if not os.path.exists(LIB_DRIVER) or not os.path.exists(LIB_ORACLE):
            build_oracle()
        _driver = C.CDLL(LIB_DRIVER)
        _driver.ksd_run.restype = C.c_int64
        _driver.ksd_run.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(KsdParams), C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_double), C.c_void_p]
        _driver.ksd_run_w.restype = C.c_int64
        _driver.ksd_run_w.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(KsdParams), C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.POINTER(C.c_double), C.c_void_p]
    return _driver


def simple_mat(m=5, a=2, b=4, sc_ambi=0):
    """m x m match/mismatch matrix with a wildcard last row/column (what cli.c:36-48 builds; our own code)."""
    mat = np.full((m, m), -abs(b), dtype=np.int8)
    np.fill_diagonal(mat, a)
    mat[m - 1, :] = sc_ambi
    mat[:, m - 1] = sc_ambi
    return np.ascontiguousarray(mat.reshape(-1))


def pack(seqs):
    """list of uint8 arrays -> (concatenated uint8 array, int64 offsets[n+1])"""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    cat = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
    if cat.size == 0:
        cat = np.zeros(1, np.uint8)
    return np.ascontiguousarray(cat), off


def make_params(kind, mat, m=5, q=4, e=2, q2=24, e2=1, w=-1, zdrop=-1, end_bonus=0, flag=0, noncan=0, junc_bonus=0):
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    P = KsdParams(KIND[kind] if isinstance(kind, str) else kind, m, mat.ctypes.data_as(C.POINTER(C.c_int8)),
                  q, e, q2, e2, w, zdrop, end_bonus, flag, noncan, junc_bonus)
    P._keep = mat
    return P


def run_cpu(which, P, queries, targets, juncs=None, nthreads=1, repeat=1, want_cigar=True, packed=None, cells_out=None, w=None):
    """Run a batch on a CPU checker. which: 'ref' | 'oracle'.
    w: optional band per pair (int32[n]; replaces P.w; pairs are then handed to the threads dynamically).
    Returns (fields int32[n,NF], cigars list[np.uint32 array], seconds)."""
    lib = LIB_REF if which == "ref" else LIB_ORACLE
    sym = (SYM_REF if which == "ref" else SYM_ORACLE)[P.kind]
    if packed is None:
        qcat, qoff = pack(queries)
        tcat, toff = pack(targets)
    else:
        qcat, qoff, tcat, toff = packed
    n = len(qoff) - 1
    jcat = None
    if juncs is not None:
        jcat, _ = pack(juncs)
    res = np.zeros((n, NF), dtype=np.int32)
    cig_off = np.zeros(n + 1, dtype=np.int64)
    secs = C.c_double(0)
    cap = 0
    buf = np.zeros(1, dtype=np.uint32)
    if want_cigar and not (P.flag & 1):
        cap = int((qoff[-1] + toff[-1]) + 2 * n + 16)
        buf = np.zeros(cap, dtype=np.uint32)
    if w is not None:
        w = np.ascontiguousarray(w, dtype=np.int32)
        assert len(w) == n
    rc = driver().ksd_run_w(lib.encode(), sym, C.byref(P), n, qcat.ctypes.data, qoff.ctypes.data, tcat.ctypes.data, toff.ctypes.data,
                          jcat.ctypes.data if jcat is not None else None, w.ctypes.data if w is not None else None, nthreads, repeat, res.ctypes.data,
                          cig_off.ctypes.data if cap else None, buf.ctypes.data if cap else None, cap, C.byref(secs),
                          cells_out.ctypes.data if cells_out is not None else None)
    if rc != 0:
        raise RuntimeError(f"ksd_run failed rc={rc}")
    cigs = [buf[cig_off[i]:cig_off[i + 1]].copy() for i in range(n)] if cap else [np.zeros(0, np.uint32)] * n
    return res, cigs, secs.value


def cigar_str(c):
    return "".join(f"{int(x) >> 4}{'MIDN___=X'[int(x) & 0xf]}" for x in c)


def band_cells(qlen, tlen, w, kind="extz2", r_stop=None):
    """In-band cell count of SURVEY.md 8(d): sum over executed diagonals of en0-st0+1 (numpy)."""
    if kind == "exts2" or w < 0:
        w = max(qlen, tlen)
    nd = qlen + tlen - 1 if r_stop is None else r_stop
    r = np.arange(nd, dtype=np.int64)
    st0 = np.maximum(np.maximum(0, r - qlen + 1), (r - w + 1) >> 1)
    en0 = np.minimum(np.minimum(tlen - 1, r), (r + w) >> 1)
    bad = np.nonzero(st0 > en0)[0]
    if len(bad):
        st0, en0 = st0[:bad[0]], en0[:bad[0]]
    return int((en0 - st0 + 1).sum())


# ---------------------------------------------------------------------------------------------
# synthetic data (shared by tests and bench.py; SURVEY.md 8(d) generators)
# ---------------------------------------------------------------------------------------------
def mutate(rng, t, sub=0.01, ins=0.002, dele=0.002, indel_geo=None):
    """Return a mutated copy of uint8 code array t (codes 0..3)."""
    out = []
    i = 0
    n = len(t)
    u = rng.random(n)
    for i in range(n):
        x = u[i]
        if x < dele:
            k = 1 if indel_geo is None else int(rng.geometric(indel_geo))
            continue  # (single-base deletions chained by repeated draws)
        if x < dele + ins:
            k = 1 if indel_geo is None else int(rng.geometric(indel_geo))
            out.extend(rng.integers(0, 4, k).tolist())
        if x > 1.0 - sub:
            out.append((int(t[i]) + int(rng.integers(1, 4))) & 3)
        else:
            out.append(int(t[i]))
    return np.asarray(out, dtype=np.uint8)


# ---------------------------------------------------------------------------------------------
# host simulation of the device engine (tests/sim/ksw2_sim.cpp) -- test infrastructure
# ---------------------------------------------------------------------------------------------
SIM_DEFS = os.environ.get("KS_SIM_DEFS", "").split()          # extra -D flags: fuzz a build variant of the device source (scripts/fuzz_campaign.py)
LIB_SIM = os.path.join(ROOT, "tests", "sim", "libksw2_sim" + ("_" + "".join(c for c in "".join(SIM_DEFS) if c.isalnum()) if SIM_DEFS else "") + ".so")
_sim = None


def build_sim():
    src = os.path.join(ROOT, "tests", "sim", "ksw2_sim.cpp")
    deps = [src] + [os.path.join(ROOT, "ksw2_b200", "csrc", f) for f in ("ksw2_prim.cuh", "ksw2_tile.cuh", "ksw2_pair.cuh", "ksw2_params.h", "ksw2_scalar.cuh", "ksw2_rows.cuh", "ksw2_extf2.cuh", "ksw2_gg2.cuh")]
    if os.path.exists(LIB_SIM) and all(os.path.getmtime(LIB_SIM) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas"] + SIM_DEFS + ["-o", LIB_SIM, src])


def sim():
    global _sim
    if _sim is None:
        build_sim()
        _sim = C.CDLL(LIB_SIM)
        _sim.kssim_run.restype = C.c_int64
        _sim.kssim_run.argtypes = [C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 10 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    return _sim


def run_sim(P, queries, targets, juncs=None, panel=32, force_smode=0):
    qcat, qoff = pack(queries)
    tcat, toff = pack(targets)
    n = len(queries)
    jcat = pack(juncs)[0] if juncs is not None else None
    res = np.zeros((n, NF), dtype=np.int32)
    cig_off = np.zeros(n + 1, dtype=np.int64)
    cap = int(qoff[-1] + toff[-1] + 2 * n + 16)
    buf = np.zeros(cap, dtype=np.uint32)
    rc = sim().kssim_run(P.kind, P.m, C.cast(P.mat, C.c_void_p), P.q, P.e, P.q2, P.e2, P.w, P.zdrop, P.end_bonus, P.flag, P.noncan,
                         P.junc_bonus, n, qcat.ctypes.data, qoff.ctypes.data, tcat.ctypes.data, toff.ctypes.data,
                         jcat.ctypes.data if jcat is not None else None, panel, force_smode, res.ctypes.data, cig_off.ctypes.data,
                         buf.ctypes.data, cap)
    if rc != 0:
        raise RuntimeError(f"kssim_run rc={rc}")
    return res, [buf[cig_off[i]:cig_off[i + 1]].copy() for i in range(n)]
