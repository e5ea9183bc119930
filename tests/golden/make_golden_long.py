"""Generate tests/golden/ref_align_long.npz with the COMPILED, UNMODIFIED reference
(oracle/_ref/libmcref.so): GlobAlignE triples (score, alignment length, matches) of long pairs --
BASELINE configs[4] lengths (9.5 - 10.5 kb, C5 generator: mutated templates), unequal lengths, an 'N'
byte, unrelated pairs, and one pair whose lengths add up to just under the 65 535 limit of the GPU
kernel's packed (length, matches) word.

Build container only:   python tests/golden/make_golden_long.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _oracle as O  # noqa: E402
from meshclust_b200 import synth  # noqa: E402


def main():
    r = O.ref()
    rng = np.random.default_rng(50505)
    seqs = []

    def mutated(t, mu):
        c, _ = synth._mutate(rng, t.copy(), np.array([0, t.size]), mu)
        return c

    # 20 pairs from C5-like templates (10 kb, 3 % mutation incl. indels), 4 of them at 10 % / 25 %
    for i in range(20):
        t = rng.integers(0, 4, int(rng.integers(9500, 10500)), dtype=np.uint8)
        mu = 0.03 if i < 16 else (0.10 if i < 18 else 0.25)
        a, b = mutated(t, mu), mutated(t, mu)
        if i % 5 == 0:   # unequal: one side loses its tail / gains a head
            b = b[: int(b.size * 0.93)]
        if i % 5 == 1:
            a = np.concatenate([rng.integers(0, 4, 700, dtype=np.uint8), a])[:10500]
        if i % 4 == 2:   # an 'N' byte (outside every segment the digits keep the letter, Chromosome.cpp:99-112)
            a = a.copy(); a[int(rng.integers(0, a.size))] = ord("N")
        if i % 7 == 3:
            b = b.copy(); b[int(rng.integers(0, b.size))] = ord("N")
        seqs += [a.tobytes(), b.tobytes()]
    # 3 unrelated pairs (identity ~ 0.5: the longest runs of gaps)
    for _ in range(3):
        seqs += [rng.integers(0, 4, int(rng.integers(9500, 10500)), dtype=np.uint8).tobytes(),
                 rng.integers(0, 4, int(rng.integers(9500, 10500)), dtype=np.uint8).tobytes()]
    # lengths just under the limit of the packed word: 32 700 + 32 800 = 65 500
    t = rng.integers(0, 4, 32750, dtype=np.uint8)
    a, b = mutated(t, 0.03), mutated(t, 0.03)
    seqs += [a[:32700].tobytes(), np.concatenate([b, rng.integers(0, 4, 200, dtype=np.uint8)])[:32800].tobytes()]
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    digits = np.frombuffer(b"".join(seqs), np.uint8).copy()
    pa = np.arange(0, len(seqs), 2, dtype=np.int32)
    pb = pa + 1
    # both argument orders (results depend on the order in corner cases)
    pa2, pb2 = np.concatenate([pa, pb[:6]]), np.concatenate([pb, pa[:6]])
    sc, ln, mt = r.globalign_batch(digits, offs, pa2, pb2)
    path = os.path.join(HERE, "ref_align_long.npz")
    np.savez_compressed(path, digits=digits, offs=offs, pa=pa2, pb=pb2, score=sc, alen=ln, matches=mt)
    print("wrote", path, os.path.getsize(path), "bytes;", pa2.size, "pairs; identities", np.round(mt / ln, 3))


if __name__ == "__main__":
    main()
