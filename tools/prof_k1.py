"""Driver for timing / ncu captures of stage 0/1 (encode + k-mer histograms) on a BASELINE config:
python tools/prof_k1.py c2 [n]   -- prints per-kernel device times (CUDA events around each C-ABI call)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshclust_b200 import api, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
letters, offs, _ = synth.generate_config(cfg, n)
k = synth.CONFIGS[cfg].kmer
n = offs.size - 1
segs = np.stack([np.zeros(n, np.int32), (np.diff(offs) - 1).astype(np.int32)], 1).reshape(-1)   # no N in the synthetic configs
seg_off = np.arange(n + 1, dtype=np.int64)
ctx = api.Context(0)
for rep in range(3):
    t0 = time.perf_counter()
    ctx.load_sequences(letters, offs, segs, seg_off)
    t1 = time.perf_counter()
    used, mx = ctx.build_histograms(k, 0)
    t2 = time.perf_counter()
    print(f"{cfg}: n={n} bases={int(offs[-1])}  upload+encode {1e3 * (t1 - t0):.2f} ms  histograms {1e3 * (t2 - t1):.3f} ms "
          f"(k={k}, {used * 8}-bit, max count {mx})  -> {int(offs[-1]) / (t2 - t1) / 1e9:.1f} Gbases/s histogram stage incl. sync", flush=True)
