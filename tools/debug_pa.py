"""Debug aid: mc_accumulate_run vs the host-driven step loop on one of the parity-test scenarios; prints the first divergence."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pybvec
from meshclust_b200 import api
from test_gpu_parity import _rand_hists, _model

def scenario(dtype, hi, k, n, sim, bin_size, spread):
    rng = np.random.default_rng(900 + k + n)
    nb = 4 ** k
    H0 = _rand_hists(rng, n, nb, dtype, hi, clusters=max(3, n // 150))
    lens0 = (1000 + rng.integers(0, spread, n)).astype(np.uint64)
    lens0[rng.integers(0, n, n // 10)] = 1000 + spread // 2
    bounds, order, first = _pybvec.layout(lens0, bin_size)
    H, lens = np.ascontiguousarray(H0[order]), np.ascontiguousarray(lens0[order])
    mins, maxs, w = _model(3)
    maxs[0] = float(spread); maxs[2] = 4.0 * nb * (hi / 255.0)
    return H, lens, bounds, first, mins, maxs, w

def main():
    args = [(np.uint16, 3000, 4, 2500, 0.85, 70, 500), (np.uint8, 255, 2, 3000, 0.90, 200, 60), (np.uint8, 255, 3, 37, 0.90, 8, 30)]
    with api.Context(0) as ctx:
        for a in args:
            H, lens, bounds, first, mins, maxs, w = scenario(*a)
            n, k = H.shape[0], a[2]
            ctx.load_histograms(H, lens, k)
            ctx.set_model(mins, maxs, w, 3)
            sums = ctx.pair_classify(np.arange(n, dtype=np.int32), np.zeros(n, np.int32))[0]
            w = w.copy(); w[0] -= np.sort(sums)[int(n * (1 - 1.5 / max(3, n // 150)))]
            ctx.set_model(mins, maxs, w, 3)
            want = _pybvec.accumulate_by_steps(ctx, a[4], bounds, first, lens)
            runs = []
            for rep in range(3):
                try:
                    runs.append(ctx.accumulate_run(a[4], bounds, first))
                except Exception as e:
                    print(a[:5], "run", rep, "FAILED:", e); runs.append(None)
            for rep, r in enumerate(runs):
                if r is None: continue
                c, o, m, st = r
                same = np.array_equal(c, want[0]) and np.array_equal(o, want[1]) and np.array_equal(m, want[2])
                print(a[:5], "run", rep, "clusters", c.size, "want", want[0].size, "identical" if same else "DIFFERENT", (st.n_scans, st.n_evals, st.n_steps), want[3])
                if not same:
                    for ci in range(min(c.size, want[0].size)):
                        g = m[o[ci]:o[ci + 1]]; x = want[2][want[1][ci]:want[1][ci + 1]]
                        if c[ci] != want[0][ci] or not np.array_equal(g, x):
                            print("  first divergence at cluster", ci, "center", c[ci], "want", want[0][ci], "sizes", g.size, x.size)
                            print("   got ", g[:20], " want", x[:20])
                            sg, sx = set(g.tolist()), set(x.tolist())
                            print("   only got", sorted(sg - sx)[:10], "only want", sorted(sx - sg)[:10], "len(center)", lens[c[ci]], "bin of seed", np.searchsorted(first, x[0], side="right") - 1)
                            break
                    break
main()
