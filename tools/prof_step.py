"""Driver for timing / ncu captures of the fused Phase-A step (scan + accumulate tail):
python tools/prof_step.py --shape c2|c4 [--clusters N]
Synthetic clusters of n/1000 members each; runs the greedy loop for a few clusters and prints the
host-side time per step next to the device time of the step's two kernels (CUDA events)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshclust_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="c4")
ap.add_argument("--clusters", type=int, default=20)
a = ap.parse_args()
n, k = {"c2": (100_000, 4), "c4": (1_000_000, 5), "c5": (200_000, 6)}[a.shape]
nb = 4 ** k
rng = np.random.default_rng(1)
ntem = 1000
base = rng.integers(1, 12, (ntem, nb), dtype=np.uint8)
tmpl = rng.integers(0, ntem, n)
hist = base[tmpl]
noise = rng.integers(0, 50, (n, 1)) == 0
hist = np.where(noise & (rng.integers(0, 8, (n, nb)) == 0), hist + 1, hist).astype(np.uint8)
lens = np.full(n, 1000, np.uint64)
ctx = api.Context(0)
ctx.load_histograms(hist, lens, k)
# bounds / weights so that same-template rows are positives and nothing else is
raw, _ = ctx.pair_features(np.arange(2000, dtype=np.int32), np.arange(2000, dtype=np.int32)[::-1].copy())
mins, maxs = raw.min(0), np.maximum(raw.max(0), 1e-9)
mins[0], maxs[0] = 0, 10
ctx.set_model(mins, maxs, np.array([-0.9, 1.0, 0.0, 0.0, 0.0]), 3)
ctx.alive_reset()
seed, steps, t_host = 0, 0, 0.0
for c in range(a.clusters):
    ctx.alive_kill(np.array([seed]))
    last, restart = seed, True
    while True:
        t0 = time.perf_counter()
        r, rows = ctx.accumulate_step(last, 0, n - 1, restart)
        t_host += time.perf_counter() - t0
        steps += 1
        restart = False
        if r.scan.n_pos == 0:
            break
        last = r.nearest_row
    print(f"cluster {c}: members {r.n_members} next seed {r.scan.best_row}", flush=True)
    if r.scan.best_row < 0:
        break
    seed = int(r.scan.best_row)
print(f"{a.shape}: {steps} steps, {t_host / steps * 1e6:.1f} us per mc_accumulate_step (host wall, includes the synchronisation)")
