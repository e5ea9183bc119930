import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust_b200 import api, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
l, o, t = synth.generate_config(cfg, n)
segs = np.zeros(0, np.int32); so = np.zeros(n + 1, np.int64)
ctx = api.Context(0)
# upload as pre-encoded digits (no segments -> letters pass through upper-cased): fine for timing
ctx.load_sequences(l, o, segs, so)
rng = np.random.default_rng(0)
for m in (150, 3000, 3000, 20000):
    pa = rng.integers(0, n, m).astype(np.int32); pb = rng.integers(0, n, m).astype(np.int32)
    t0 = time.perf_counter(); ctx.align_pairs(pa, pb); dt = time.perf_counter() - t0
    cells = float((np.diff(o)[pa].astype(np.float64) * np.diff(o)[pb]).sum())
    print(f"{cfg}: {m} pairs  {dt*1e3:.1f} ms  {cells/dt/1e9:.1f} GCUPS", flush=True)
