"""Per-phase timeline of one scan launch (needs the -DMC_SCAN_TRACE build in tools/_trace)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["MESHCLUST_B200_LIB"] = os.path.join(ROOT, "tools", "_trace", "libmc_trace.so")
sys.path.insert(0, ROOT)
from meshclust_b200 import api

shape = sys.argv[1] if len(sys.argv) > 1 else "c2"
n, k, reps = {"c1": (10_000, 3, 4), "c2": (100_000, 4, 10), "c4": (1_000_000, 5, 1)}[shape]
nb = 4 ** k
rng = np.random.default_rng(1)
base = rng.integers(1, 8, (1000, nb), dtype=np.uint8)
hist = base[rng.integers(0, 1000, n * reps)]
ctx = api.Context(0)
ctx.load_histograms(hist, np.full(n * reps, 1500, np.uint64), k)
ctx.set_model(np.array([0, 0.5, 0, -1, 100.0]), np.array([100, 1, 4000, 1, 4000.0]), np.array([-1.0, 2, 1, 0.5, 0.5]), 4)
L = 6
cr = np.array([(i % reps) * n + (i * 7919) % n for i in range(L)], np.int64)
lo = np.array([(i % reps) * n for i in range(L)], np.int64)
ctx.scan_enqueue_many(cr, lo, lo + n - 1, False, 0)
ctx.sync()
buf = np.zeros(148 * 32 * 8, np.uint64)
lib = ctypes.CDLL(os.environ["MESHCLUST_B200_LIB"])
assert lib.mc_debug_scan_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
t = buf.reshape(148, 32, 8).astype(np.int64)
t0 = t[:, 0, 0][t[:, 0, 0] > 0].min()
names = ["cta start", "after barrier init+sync", "consumer before 1st full wait", "1st tile landed", "1st tile reduced", "consumer done (all tiles+epilogues)", "after final syncthreads", "last CTA wrote result"]
for s, nm in enumerate(names):
    v = t[:, :, s]
    v = v[v > 0] - t0
    if v.size:
        print(f"{nm:40s} min {v.min()/1e3:7.2f} us  median {np.median(v)/1e3:7.2f} us  max {v.max()/1e3:7.2f} us   (n={v.size})")
