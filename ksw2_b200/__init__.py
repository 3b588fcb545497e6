"""ksw2_b200 -- B200-native (sm_100a) implementation of the lh3/ksw2 hot path.

The product is the C-ABI shared library `ksw2_b200/libksw2_b200.so` (CUDA kernels + host code in
ksw2_b200/csrc/, public headers in include/).  This Python module is only plumbing for tests and
bench.py: it builds the library with nvcc, loads it with ctypes and wraps the batch calls with numpy
arrays.  There is no CPU implementation behind it: without the built extension or without a CUDA
device every call raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("KSW2B_LIB") or os.path.join(PKG_DIR, "libksw2_b200.so")   # KSW2B_LIB: A/B builds (scripts/)
SOURCES = [os.path.join(CSRC, f) for f in ("ksw2_b200.cu", "ksw2_prim.cuh", "ksw2_tile.cuh", "ksw2_pair.cuh", "ksw2_params.h", "ksw2_scalar.cuh", "ksw2_rows.cuh", "ksw2_extf2.cuh", "ksw2_gg2.cuh")] + \
          [os.path.join(ROOT, "include", f) for f in ("ksw2.h", "ksw2_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550", "-diag-suppress", "128"]

EXTZ2, EXTD2, EXTS2 = 0, 1, 2
KIND = {"extz2": EXTZ2, "extd2": EXTD2, "exts2": EXTS2, "extz": 3, "extd": 4, "extf2": 5, "gg": 6, "gg2": 7, "gg2_sse": 8}


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a (cross-compiles without a GPU)."""
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in SOURCES):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "ksw2_b200.cu"), "-ldl"]
    subprocess.check_call(cmd)
    return LIB_PATH


CLI_SRC = os.path.join(ROOT, "tools", "ksw2b_test.cpp")
CLI_PATH = os.path.join(ROOT, "tools", "ksw2b-test")


def build_cli(force=False):
    """Compile tools/ksw2b-test (host code only: FASTA in, one batch through the C ABI, the reference CLI's output format)."""
    build()
    deps = [CLI_SRC, LIB_PATH, os.path.join(ROOT, "include", "ksw2_b200.h")]
    if not force and os.path.exists(CLI_PATH) and all(os.path.getmtime(CLI_PATH) >= os.path.getmtime(d) for d in deps):
        return CLI_PATH
    subprocess.check_call([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-o", CLI_PATH, CLI_SRC, "-L" + PKG_DIR, "-lksw2_b200", "-lz",
                           "-Wl,-rpath,$ORIGIN/../ksw2_b200"])
    return CLI_PATH


class Params(C.Structure):
    """mirror of ksw2b_params_t (include/ksw2_b200.h)"""
    _fields_ = [("kind", C.c_int), ("m", C.c_int), ("mat", C.POINTER(C.c_int8)),
                ("q", C.c_int), ("e", C.c_int), ("q2", C.c_int), ("e2", C.c_int),
                ("w", C.c_int), ("zdrop", C.c_int), ("end_bonus", C.c_int), ("flag", C.c_int),
                ("noncan", C.c_int), ("junc_bonus", C.c_int)]


RESULT_DTYPE = np.dtype([("max", "<i4"), ("zdropped", "<i4"), ("max_q", "<i4"), ("max_t", "<i4"), ("mqe", "<i4"), ("mqe_t", "<i4"),
                         ("mte", "<i4"), ("mte_q", "<i4"), ("score", "<i4"), ("reach_end", "<i4"), ("n_cigar", "<i4"),
                         ("tb_i", "<i4"), ("tb_j", "<i4"), ("n_diag", "<i4"), ("cigar_off", "<i8")])
assert RESULT_DTYPE.itemsize == 64


class ExtzT(C.Structure):
    """mirror of ksw_extz_t (include/ksw2.h; reference ksw2.h:33-42), sizeof == 56 on x86-64"""
    _fields_ = [("max_zd", C.c_uint32), ("max_q", C.c_int), ("max_t", C.c_int), ("mqe", C.c_int), ("mqe_t", C.c_int),
                ("mte", C.c_int), ("mte_q", C.c_int), ("score", C.c_int), ("m_cigar", C.c_int), ("n_cigar", C.c_int),
                ("reach_end", C.c_int), ("cigar", C.POINTER(C.c_uint32))]


_lib = None


def lib():
    """Load the extension; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run ksw2_b200.build() (nvcc) first; there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.ksw2b_create.restype = C.c_void_p; L.ksw2b_create.argtypes = [C.c_int]
        L.ksw2b_destroy.argtypes = [C.c_void_p]
        L.ksw2b_last_error.restype = C.c_char_p
        L.ksw2b_align.restype = C.c_int
        L.ksw2b_align.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.POINTER(C.POINTER(C.c_uint32))]
        L.ksw2b_plan_create.restype = C.c_void_p
        L.ksw2b_plan_create.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int64, C.c_void_p, C.c_void_p]
        if hasattr(L, "ksw2b_align_ex"):        # (A/B builds of older sources, scripts/ab.sh, may lack the newest entry points)
            L.ksw2b_align_ex.restype = C.c_int
            L.ksw2b_align_ex.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_uint32))]
            L.ksw2b_plan_create_ex.restype = C.c_void_p
            L.ksw2b_plan_create_ex.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
            L.ksw2b_set_timing.restype = None; L.ksw2b_set_timing.argtypes = [C.c_void_p, C.c_int]
            L.ksw2b_last_timing.restype = None
            L.ksw2b_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
            L.ksw2b_multi_create.restype = C.c_void_p; L.ksw2b_multi_create.argtypes = [C.c_void_p, C.c_int]
            L.ksw2b_multi_destroy.restype = None; L.ksw2b_multi_destroy.argtypes = [C.c_void_p]
            L.ksw2b_multi_devices.restype = C.c_int; L.ksw2b_multi_devices.argtypes = [C.c_void_p]
            L.ksw2b_multi_align.restype = C.c_int
            L.ksw2b_multi_align.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_uint32))]
            L.ksw2b_multi_last.restype = None; L.ksw2b_multi_last.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ksw2b_last_transfer_bytes.restype = None
        L.ksw2b_last_transfer_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        L.ksw2b_plan_run.restype = C.c_int
        L.ksw2b_plan_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ksw2b_plan_fetch.restype = C.c_int
        L.ksw2b_plan_fetch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)), C.c_void_p]
        L.ksw2b_plan_cells.restype = C.c_int64; L.ksw2b_plan_cells.argtypes = [C.c_void_p]
        L.ksw2b_plan_launches.restype = C.c_int; L.ksw2b_plan_launches.argtypes = [C.c_void_p]
        L.ksw2b_plan_device_results.restype = C.c_void_p; L.ksw2b_plan_device_results.argtypes = [C.c_void_p]
        L.ksw2b_plan_destroy.argtypes = [C.c_void_p]
        L.ksw2b_plan_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.ksw2b_plan_fill_ms.restype = C.c_double; L.ksw2b_plan_fill_ms.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ksw2b_host_alloc.restype = C.c_void_p; L.ksw2b_host_alloc.argtypes = [C.c_size_t]
        L.ksw2b_host_free.argtypes = [C.c_void_p]
        L.ksw2b_set_tuning.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ksw2b_set_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
        zargs = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int8, C.c_void_p]
        L.ksw_extz2_sse.restype = None
        L.ksw_extz2_sse.argtypes = zargs + [C.c_int8, C.c_int8, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ExtzT)]
        L.ksw_extd2_sse.restype = None
        L.ksw_extd2_sse.argtypes = zargs + [C.c_int8, C.c_int8, C.c_int8, C.c_int8, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ExtzT)]
        L.ksw_extz.restype = None
        L.ksw_extz.argtypes = zargs + [C.c_int8, C.c_int8, C.c_int, C.c_int, C.c_int, C.POINTER(ExtzT)]
        L.ksw_extd.restype = None
        L.ksw_extd.argtypes = zargs + [C.c_int8, C.c_int8, C.c_int8, C.c_int8, C.c_int, C.c_int, C.c_int, C.POINTER(ExtzT)]
        L.ksw_extf2_sse.restype = None
        L.ksw_extf2_sse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int8, C.c_int8, C.c_int8, C.c_int, C.c_int, C.POINTER(ExtzT)]
        for nm in ("ksw_gg", "ksw_gg2", "ksw_gg2_sse"):
            if hasattr(L, nm):              # (A/B builds of older sources, scripts/ab.sh, may lack the newest entry points)
                getattr(L, nm).restype = C.c_int
                getattr(L, nm).argtypes = zargs + [C.c_int8, C.c_int8, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint32))]
        L.ksw_exts2_sse.restype = None
        L.ksw_exts2_sse.argtypes = zargs + [C.c_int8, C.c_int8, C.c_int8, C.c_int8, C.c_int, C.c_int8, C.c_int, C.c_void_p, C.POINTER(ExtzT)]
        _lib = L
    return _lib


def make_params(kind, mat, m=5, q=4, e=2, q2=24, e2=1, w=-1, zdrop=-1, end_bonus=0, flag=0, noncan=0, junc_bonus=0):
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    P = Params(KIND[kind] if isinstance(kind, str) else kind, m, mat.ctypes.data_as(C.POINTER(C.c_int8)),
               q, e, q2, e2, w, zdrop, end_bonus, flag, noncan, junc_bonus)
    P._keep = mat
    return P


def pack(seqs):
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    cat = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
    if cat.size == 0:
        cat = np.zeros(1, np.uint8)
    return np.ascontiguousarray(cat), off


class Context:
    """ksw2b_ctx_t wrapper: one CUDA device, reusable buffers."""

    def __init__(self, device=-1):
        self.h = lib().ksw2b_create(device)
        if not self.h:
            raise RuntimeError("ksw2b_create failed: " + lib().ksw2b_last_error().decode())

    def close(self):
        if self.h:
            lib().ksw2b_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_timing(self, on=True):
        lib().ksw2b_set_timing(self.h, 1 if on else 0)

    def last_timing(self):
        """(fill_ms, fill_launches, span_ms, launches) of the last align call on this context (needs set_timing(True))"""
        a, b, c, d = C.c_double(0), C.c_int(0), C.c_double(0), C.c_int(0)
        lib().ksw2b_last_timing(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, c.value, d.value

    def last_transfer_bytes(self):
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        lib().ksw2b_last_transfer_bytes(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_tuning(self, panel=0, threads=0, ctas_per_sm=0):
        lib().ksw2b_set_tuning(self.h, panel, threads, ctas_per_sm)

    def set_mode(self, mode=0, warp_panel=0):
        """0 auto, 1 one thread per alignment, 2 one warp per alignment"""
        lib().ksw2b_set_mode(self.h, mode, warp_panel)

    def align_packed(self, P, qcat, qoff, tcat, toff, jcat=None, w=None, want_cigars=True):
        """host buffers in, (results[n] structured array, list of CIGAR arrays) out: the drop-in batch call.
        w: optional int32 band per pair (ksw2b_align_ex)"""
        n = len(qoff) - 1
        res = np.zeros(n, dtype=RESULT_DTYPE)
        cig = C.POINTER(C.c_uint32)()
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.int32)
            assert len(w) == n
            rc = lib().ksw2b_align_ex(self.h, C.byref(P), n, qcat.ctypes.data, qoff.ctypes.data, tcat.ctypes.data, toff.ctypes.data,
                                      jcat.ctypes.data if jcat is not None else None, w.ctypes.data, res.ctypes.data, C.byref(cig))
        else:
            rc = lib().ksw2b_align(self.h, C.byref(P), n, qcat.ctypes.data, qoff.ctypes.data, tcat.ctypes.data, toff.ctypes.data,
                                   jcat.ctypes.data if jcat is not None else None, res.ctypes.data, C.byref(cig))
        if rc != 0:
            raise RuntimeError(f"ksw2b_align rc={rc}: " + lib().ksw2b_last_error().decode())
        return res, (_collect(res, cig, P, n) if want_cigars else [])

    def align(self, P, queries, targets, juncs=None, w=None):
        qcat, qoff = pack(queries)
        tcat, toff = pack(targets)
        jcat = pack(juncs)[0] if juncs is not None else None
        return self.align_packed(P, qcat, qoff, tcat, toff, jcat, w)


def _collect(res, cig, P, n):
    cigs = []
    if not (P.flag & 1):
        tot = int((res["cigar_off"] + res["n_cigar"]).max()) if n else 0
        allc = np.ctypeslib.as_array(cig, shape=(tot,)).copy() if tot and cig else np.zeros(0, np.uint32)
        cigs = [allc[o:o + k] for o, k in zip(res["cigar_off"], res["n_cigar"])]
    return cigs


class MultiContext:
    """ksw2b_multi_t wrapper: several GPUs of one box driven from one caller; devices: list of ordinals or a count"""

    def __init__(self, devices):
        if isinstance(devices, int):
            self.n, arr = devices, None
        else:
            self.n = len(devices); arr = (C.c_int * self.n)(*devices)
        self.h = lib().ksw2b_multi_create(arr, self.n)
        if not self.h:
            raise RuntimeError("ksw2b_multi_create failed: " + lib().ksw2b_last_error().decode())

    def close(self):
        if self.h:
            lib().ksw2b_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def align_packed(self, P, qcat, qoff, tcat, toff, jcat=None, w=None, want_cigars=True):
        n = len(qoff) - 1
        res = np.zeros(n, dtype=RESULT_DTYPE)
        cig = C.POINTER(C.c_uint32)()
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.int32)
        rc = lib().ksw2b_multi_align(self.h, C.byref(P), n, qcat.ctypes.data, qoff.ctypes.data, tcat.ctypes.data, toff.ctypes.data,
                                     jcat.ctypes.data if jcat is not None else None, w.ctypes.data if w is not None else None,
                                     res.ctypes.data, C.byref(cig))
        if rc != 0:
            raise RuntimeError(f"ksw2b_multi_align rc={rc}: " + lib().ksw2b_last_error().decode())
        return res, (_collect(res, cig, P, n) if want_cigars else [])

    def last(self):
        pairs = np.zeros(self.n, np.int64); span = np.zeros(self.n, np.float64)
        lib().ksw2b_multi_last(self.h, pairs.ctypes.data, span.ctypes.data)
        return pairs, span
