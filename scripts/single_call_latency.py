import sys, os, ctypes as C, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, harness as H, ksw2_b200 as K
L = K.lib(); mat = H.simple_mat(5, 2, 4); rng = np.random.default_rng(1)
t = rng.integers(0, 4, 150).astype(np.uint8); q = t.copy(); ez = K.ExtzT()
for i in range(20):
    L.ksw_extz2_sse(None, 150, q.ctypes.data, 150, t.ctypes.data, 5, mat.ctypes.data, 4, 2, 100, 100, 0, 0x41, C.byref(ez))
os.environ["KSW2B_TIMING"] = "1"
t0 = time.perf_counter()
for i in range(5):
    L.ksw_extz2_sse(None, 150, q.ctypes.data, 150, t.ctypes.data, 5, mat.ctypes.data, 4, 2, 100, 100, 0, 0x41, C.byref(ez))
print("5 calls", (time.perf_counter() - t0) * 1e3, "ms")
os.environ.pop("KSW2B_TIMING")
t0 = time.perf_counter()
for i in range(2000):
    L.ksw_extz2_sse(None, 150, q.ctypes.data, 150, t.ctypes.data, 5, mat.ctypes.data, 4, 2, 100, 100, 0, 0x41, C.byref(ez))
print("per call us", (time.perf_counter() - t0) / 2000 * 1e6)
