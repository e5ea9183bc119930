"""Driver for timing / ncu captures of K2b (150 centers x all rows distance keys): python tools/prof_keys.py c4|c5|c2"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshclust_b200 import api  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "c4"
n, k = {"c2": (100_000, 4), "c4": (1_000_000, 5), "c5": (200_000, 6)}[shape]
nb = 4 ** k
rng = np.random.default_rng(1)
base = rng.integers(1, 8, (1000, nb), dtype=np.uint8)
hist = base[rng.integers(0, 1000, n)]
ctx = api.Context(0)
ctx.load_histograms(hist, np.full(n, 1000, np.uint64), k)
centers = rng.integers(0, n, 150).astype(np.int32)
for rep in range(3):
    t0 = time.perf_counter()
    keys = ctx.distance_keys(centers)
    print(f"{shape}: 150 x {n} keys in {1e3 * (time.perf_counter() - t0):.1f} ms (kernel + device->host copy of {keys.nbytes / 1e6:.0f} MB)", flush=True)
