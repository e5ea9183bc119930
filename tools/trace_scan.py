"""Per-phase timelines of 8 back-to-back scan launches (needs the -DMC_SCAN_TRACE build variant)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("MESHCLUST_B200_LIB", os.path.join(ROOT, "meshclust_b200", "libmeshclust_b200_trace.so"))   # MC_LIB_VARIANT=trace MC_EXTRA_NVCC_FLAGS=-DMC_SCAN_TRACE python -m meshclust_b200.build
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
L = 8
cr = np.array([(i % reps) * n + (i * 7919) % n for i in range(L)], np.int64)
lo = np.array([(i % reps) * n for i in range(L)], np.int64)
ctx.scan_enqueue_many(cr, lo, lo + n - 1, False, 0)   # warm-up
ctx.sync()
ctx.scan_enqueue_many(cr, lo, lo + n - 1, False, 0)   # slots 0..7 -> trace keys 0..7 (the slot stride is 160 records)
ctx.sync()
buf = np.zeros(8 * 148 * 32 * 8, np.uint64)
lib = ctypes.CDLL(os.environ["MESHCLUST_B200_LIB"])
assert lib.mc_debug_scan_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
t = buf.reshape(8, 148, 32, 8).astype(np.int64)
# order the keys by the time their launch started
starts = [t[key, :, 0, 0][t[key, :, 0, 0] > 0].min() if (t[key, :, 0, 0] > 0).any() else 0 for key in range(8)]
order = [k_ for k_ in np.argsort(starts) if starts[k_] > 0]
t0 = starts[order[0]]
names = ["cta start", "after barrier init+sync", "consumer before 1st full wait", "1st tile landed", "1st tile reduced", "consumer done", "after final syncthreads", "CTA partial written"]
print(f"== {shape}: timelines of {len(order)} back-to-back launches, us relative to the first CTA start of the first one (min / median / max over warps)")
for key in order:
    row = []
    for s_, nm in enumerate(names):
        v = t[key, :, :, s_]
        v = v[v > 0] - t0
        row.append(f"{nm}: {v.min()/1e3:6.2f}/{np.median(v)/1e3:6.2f}/{v.max()/1e3:6.2f}" if v.size else f"{nm}: -")
    print(" | ".join(row))
