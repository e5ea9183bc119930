"""Time back-to-back scan launches with CUDA events (and host wall) for several shapes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meshclust_b200 import api

def run(shape, L=40):
    n, k, reps = {"c1": (10_000, 3, 64), "c2": (100_000, 4, 10), "c2x1": (100_000, 4, 1), "c4": (1_000_000, 5, 1), "c5": (200_000, 6, 1), "c2big": (400_000, 4, 4)}[shape]
    nb = 4 ** k
    rng = np.random.default_rng(1)
    base = rng.integers(1, 8, (1000, nb), dtype=np.uint8)
    hist = base[rng.integers(0, 1000, n * reps)]
    lens = np.full(n * reps, 1500, np.uint64)
    ctx = api.Context(0)
    ctx.load_histograms(hist, lens, k)
    ctx.set_model(np.array([0, 0.5, 0, -1, 100.0]), np.array([100, 1, 4000, 1, 4000.0]), np.array([-1.0, 2, 1, 0.5, 0.5]), 4)
    cr = np.array([(i % reps) * n + (i * 7919) % n for i in range(L)], np.int64)
    lo = np.array([(i % reps) * n for i in range(L)], np.int64)
    hi = lo + n - 1
    stream = torch.cuda.ExternalStream(ctx.stream)
    ctx.scan_enqueue_many(cr, lo, hi, False, 0); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    ctx.scan_enqueue_many(cr, lo, hi, False, 0)
    e1.record(stream)
    t_enq = time.perf_counter() - t0
    ctx.sync()
    t_all = time.perf_counter() - t0
    dev = e0.elapsed_time(e1) * 1e3 / L
    by = n * (nb + 33)
    print(f"{shape}: dev {dev:.2f} us/launch  host-enqueue {t_enq*1e6/L:.2f} us/launch  wall {t_all*1e6/L:.2f} us/launch  -> {by/dev/1e3:.0f} GB/s", flush=True)
    # one at a time with sync (latency of a single scan incl. result D2H)
    t0 = time.perf_counter()
    for i in range(20):
        ctx.scan(int(cr[i % L]), int(lo[i % L]), int(hi[i % L]), want_marks=False)
    print(f"   mc_scan sync round trip {1e6*(time.perf_counter()-t0)/20:.1f} us", flush=True)
    ctx.close()

for s in sys.argv[1:] or ["c1", "c2", "c2x1", "c4", "c5"]:
    run(s)
