"""Small driver for ncu captures of the scan kernel: python tools/prof_scan.py --shape c2|c4|c5 [--launches N]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshclust_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="c2")
ap.add_argument("--launches", type=int, default=6)
ap.add_argument("--chain", action="store_true", help="one launch per scan, each behind the one before it (MC_SCAN_CHAIN)")
a = ap.parse_args()
n, k, reps = {"c1": (10_000, 3, 64), "c2": (100_000, 4, 10), "c4": (1_000_000, 5, 1), "c5": (200_000, 6, 1)}[a.shape]
nb = 4 ** k
rng = np.random.default_rng(1)
base = rng.integers(1, 8, (1000, nb), dtype=np.uint8)
hist = base[rng.integers(0, 1000, n * reps)]
lens = np.full(n * reps, 1500, np.uint64)
ctx = api.Context(0)
ctx.load_histograms(hist, lens, k)
ctx.set_model(np.array([0, 0.5, 0, -1, 100.0]), np.array([100, 1, 4000, 1, 4000.0]), np.array([-1.0, 2, 1, 0.5, 0.5]), 4)
L = a.launches
cr = np.array([(i % reps) * n + (i * 7919) % n for i in range(L)], np.int64)
lo = np.array([(i % reps) * n for i in range(L)], np.int64)
hi = lo + n - 1
ctx.scan_enqueue_many(cr, lo, hi, api.MC_SCAN_CHAIN if a.chain else False, 0)
ctx.sync()
print(ctx.scan_collect(0, L)[:2])
