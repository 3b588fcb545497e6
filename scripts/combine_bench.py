#!/usr/bin/env python
"""Throughput of the UNCHANGED one-pair-per-call API (ksw_extz2_sse) when T host threads call it concurrently:
the library combines the calls in flight into GPU batches.  Uses the multi-threaded C runner oracle/ref_driver.c (pthreads, one
call per pair, exactly like a minimap2-style caller) pointed at the PRODUCT library, and checks the results against the
reference build when it is there.  usage: combine_bench.py [pairs] [threads,threads,...]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
import ksw2_b200 as K
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
threads = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4, 16, 64, 256]
qcat, qoff, tcat, toff = bench.gen_c2(n, 150, 20260925)
mat = H.simple_mat(5, 2, 4)
P = H.make_params("extz2", mat, q=4, e=2, w=100, zdrop=100, end_bonus=0, flag=0x41)
K.lib()
cells = None
ref = None
if H.have_ref():
    cells_arr = np.zeros(n, np.int64)
    ref = H.run_cpu("ref", P, None, None, nthreads=os.cpu_count(), packed=(qcat, qoff, tcat, toff))[0]
K.lib().ksw2b_combine_stats.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
out = {}
for T in threads:
    c0 = C.c_ulonglong(0); b0 = C.c_ulonglong(0)
    K.lib().ksw2b_combine_stats(C.byref(c0), C.byref(b0))
    res = np.zeros((n, H.NF), np.int32); secs = C.c_double(0)
    rc = H.driver().ksd_run(K.LIB_PATH.encode(), b"ksw_extz2_sse", C.byref(P), n, qcat.ctypes.data, qoff.ctypes.data, tcat.ctypes.data, toff.ctypes.data,
                            None, T, 1, res.ctypes.data, None, None, 0, C.byref(secs), None)
    assert rc == 0
    ok = None if ref is None else bool(np.array_equal(ref[:, :9], res[:, :9]))
    calls = C.c_ulonglong(0); batches = C.c_ulonglong(0)
    K.lib().ksw2b_combine_stats(C.byref(calls), C.byref(batches))
    per = (calls.value - c0.value) / max(1, batches.value - b0.value)
    out[T] = {"calls_per_s": n / secs.value, "calls_per_batch": per, "parity_vs_reference": ok}
    print(f"threads={T:4d}: {n / secs.value:12.0f} calls/s  {n * 20050 / secs.value / 1e9:8.2f} GCUPS(full band)  "
          f"calls/batch in this run {per:8.1f}  parity_vs_reference={ok}", flush=True)
import json
print(json.dumps({"single_call": out, "lanes": os.environ.get("KSW2B_LANES", "4"), "linger_us": os.environ.get("KSW2B_LINGER_US", "0")}))
