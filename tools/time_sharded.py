"""Where does a sharded burst lose time?  One GPU, C2 shape, 10 scans per burst, CUDA events.
  A  plain mc_scan_enqueue_many (single-GPU path, one batch launch per burst)
  B  mc_scan_sharded_burst, world = 1 (self inbox), with fold / send / combine
  C  mc_scan_sharded_burst, world = 1, MC_BURST_NO_EXCHANGE (scan variant only)
  D  mc_scan_sharded_burst, world = 2, rank 0 only, MC_BURST_NO_EXCHANGE (half of the blocks of a 2n-row range)
python tools/time_sharded.py A|B|C|D"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1]
if mode in "CD":
    os.environ["MC_BURST_NO_EXCHANGE"] = "1"
import torch  # noqa: E402
from meshclust_b200 import api  # noqa: E402

n, k, S, R = 100_000, 4, 10, 10
world = 2 if mode == "D" else 1
N = n * world
nb = 4 ** k
rng = np.random.default_rng(1)
base = rng.integers(1, 8, (1000, nb), dtype=np.uint8)
hist = base[rng.integers(0, 1000, N * R)]
lens = np.full(N * R, 1500, np.uint64)
ctxs = [api.Context(0) for _ in range(world)]
for c in ctxs:
    c.load_histograms(hist, lens, k)
    c.set_model(np.array([0, 0.5, 0, -1, 100.0]), np.array([100, 1, 4000, 1, 4000.0]), np.array([-1.0, 2, 1, 0.5, 0.5]), 4)
if mode != "A":
    for r, c in enumerate(ctxs):
        c.comm_init(r, world)
    api.Context.comm_connect_local(ctxs)
ctx = ctxs[0]
stream = torch.cuda.ExternalStream(ctx.stream)
steps = 40
args = []
for st in range(steps + 3):
    reps = [(st * S + s) % R for s in range(S)]
    cr = np.array([r * N + (st * 7919 + s * 31) % N for s, r in enumerate(reps)], np.int64)
    lo = np.array([r * N for r in reps], np.int64)
    args.append((cr, lo, lo + N - 1))
e = np.zeros(0, np.int64)


def run(a, b):
    inflight = []
    for st in range(a, b):
        cr, lo, hi = args[st]
        if mode == "A":
            ctx.scan_enqueue_many(cr, lo, hi, False, 0)
            continue
        inflight.append(st)
        if len(inflight) > 2:
            old = inflight.pop(0)
            ctx.scan_sharded_burst(cr, lo, hi, False, (st % 3) * 16, (old % 3) * 16, S)
        else:
            ctx.scan_sharded_burst(cr, lo, hi, False, (st % 3) * 16, 0, 0)
    while inflight:
        old = inflight.pop(0)
        ctx.scan_sharded_burst(e, e, e, False, 0, (old % 3) * 16, S)


run(0, 3)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
run(3, steps + 3)
e1.record(stream)
ctx.sync()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / steps
print(f"{mode}: {us:.1f} us per burst of {S} scans ({us / S:.2f} us per scan), rows per scan on this rank {n}", flush=True)
