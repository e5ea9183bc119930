import os, sys, time
t0 = time.perf_counter()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust_b200 import api, synth
t1 = time.perf_counter()
ctx = api.Context(0)
t2 = time.perf_counter()
l, o, t = synth.generate(20000, 100, 1000, 0.03, 1)
t3 = time.perf_counter()
ctx.load_sequences(l, o)
t4 = time.perf_counter()
ctx.build_histograms(4, 0)
t5 = time.perf_counter()
ctx.build_histograms(4, 0)
t6 = time.perf_counter()
print(f"import {t1-t0:.2f}s  ctx_create {t2-t1:.2f}s  gen {t3-t2:.2f}s  load+encode {t4-t3:.3f}s  hist(first) {t5-t4:.3f}s  hist(again) {t6-t5:.4f}s")
