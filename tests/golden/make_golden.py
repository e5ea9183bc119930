"""Generate tests/golden/ref_vectors.npz from the COMPILED, UNMODIFIED reference
(oracle/_ref/libmcref.so, built by `make -C oracle ref` from /root/reference).

Run in the build container only (the GPU box has no /root/reference; it uses the committed file):

    python tests/golden/make_golden.py

Every array in the file is an output of reference code (GlobAlignE, ChromosomeOneDigit,
KmerHashTable/fill_table, DivergencePoint, Feature, Trainer::get_close arithmetic) on the seeded
inputs stored next to it.
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
    rng = np.random.default_rng(20261018)
    out = {}

    # ---- alignment: the Runner.cpp:304-313 pairs + corner cases + random pairs
    kat = [("GATCTCAG", "GACAG"), ("GACAG", "GATCAG"), ("GGAACCTT", "GGCCAATT"),
           ("GATCCATTACCG", "GATATTACCTT"),
           ("AGATGGTGCACGAACCGCGATTTGATGAATAACCTATTCGAACAGATTCCACCCCGTACTTAGATTCCACGGTAACAGTG",
            "AGATGGTgaCggacccaTTTaagAATtAACCTAcTCGacAGAtTCCAcCtCCGtctaGATTCCACGGTacAaagTGAAGG"),
           ("ACGT", ""), ("", "ACGT"), ("", ""), ("A", "A"), ("A", "C"), ("AAAA", "AAAAAAAAAAAA")]
    m = {"A": 0, "C": 1, "G": 2, "T": 3}
    seqs = []
    for a, b in kat:
        seqs.append(bytes(m[c] for c in a.upper()))
        seqs.append(bytes(m[c] for c in b.upper()))
    for _ in range(120):
        la = int(rng.integers(1, 200))
        a = rng.integers(0, 4, la, dtype=np.uint8)
        b = a.copy()
        mut = rng.random(la) < rng.choice([0.02, 0.1, 0.3])
        b[mut] = rng.integers(0, 4, int(mut.sum()), dtype=np.uint8)
        # indels
        keep = rng.random(la) > 0.03
        b = b[keep]
        if rng.random() < 0.3:
            b = np.concatenate([b, rng.integers(0, 4, int(rng.integers(1, 40)), dtype=np.uint8)])
        if rng.random() < 0.15 and la > 4:
            a[int(rng.integers(0, la))] = ord("N")
        if rng.random() < 0.15 and b.size > 4:
            b[int(rng.integers(0, b.size))] = ord("N")
        seqs.append(a.tobytes())
        seqs.append(b.tobytes())
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    digits = np.frombuffer(b"".join(seqs), np.uint8).copy()
    pa = np.arange(0, len(seqs), 2, dtype=np.int32)
    pb = pa + 1
    sc, ln, mt = r.globalign_batch(digits, offs, pa, pb)
    out.update(al_digits=digits, al_offs=offs, al_pa=pa, al_pb=pb, al_score=sc, al_len=ln, al_matches=mt)

    # ---- encode + histograms: edge-case records and a small synthetic batch
    recs = [b"ACGTACGTACGTACGTACGTACGTAC", b"ACGTACGTACGTTCGTACGAACGTACGG",
            b"acgtacgtacgtacgtacgtNNNNNacgtacgtacgtacgtacgtacgt", b"ACGTACGTAC",
            b"NNNNACGTACGTACGTACGTACGTACGTNNNNNNNNNNNNNNNNNNNNACGTRYMKSWHBVDXACGTACGTAAN",
            b"ACGTACGTACGTACGTACGTACGTACGTNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNA",
            b"ACGTNNACGTNNNNNNNNNNNNNNNNACGTACGT" + b"TTGACA" * 10,
            b"A" * 300 + b"C" * 30, b"N" + b"ACGGT" * 9]
    letters, loffs, _ = synth.generate(48, 6, 240, 0.05, 11)
    letters = letters.copy()
    letters[rng.integers(0, letters.size, 40)] = ord("N")
    letters[::5] |= 0x20
    allrec = recs + [letters[loffs[i]:loffs[i + 1]].tobytes() for i in range(48)]
    eoffs = np.zeros(len(allrec) + 1, np.int64)
    np.cumsum([len(s) for s in allrec], out=eoffs[1:])
    eletters = np.frombuffer(b"".join(allrec), np.uint8).copy()
    edigits = np.zeros_like(eletters)
    segs, seg_off = [], [0]
    for i, s in enumerate(allrec):
        d, sg = r.encode(s)
        assert d is not None
        edigits[eoffs[i]:eoffs[i + 1]] = d
        segs.append(sg.reshape(-1))
        seg_off.append(seg_off[-1] + len(sg))
    out.update(enc_letters=eletters, enc_offs=eoffs, enc_digits=edigits,
               enc_segs=np.concatenate(segs).astype(np.int32), enc_seg_off=np.array(seg_off, np.int64))
    for k in (1, 2, 3, 4, 5, 6):
        rc, h, _ = r.hist_batch(eletters, eoffs, k, 2)
        assert rc == 0
        out[f"hist_k{k}"] = h
    # records the reference rejects
    out["enc_bad"] = np.array([1 if r.encode(s)[0] is None else 0 for s in
                               (b"NNNNNN", b"", b"ACGTACGTACGTACGTACGTAC-GT", b"NNNNNA")], np.int32)

    # ---- pair arithmetic on seeded histograms (u8 and u16)
    for name, dt, hi in (("u8", np.uint8, 255), ("u16", np.uint16, 4000)):
        for nb in (16, 64, 256, 1024):
            n = 40
            base = np.minimum(rng.poisson(rng.choice([2, 6, 30]), nb) + 1, hi)
            H = np.minimum(np.maximum(base[None, :] + rng.integers(-3, 4, (n, nb)), 1), hi).astype(dt)
            H[n // 2:] = np.minimum(rng.poisson(5, (n - n // 2, nb)) + 1, hi).astype(dt)
            lens = (H.sum(1) + rng.integers(0, 30, n)).astype(np.uint64)
            raw = np.zeros((n, n, 5))
            dist = np.zeros((n, n), np.uint64)
            for i in range(n):
                for j in range(n):
                    raw[i, j], dist[i, j] = r.features(H[i], H[j], int(lens[i]), int(lens[j]))
            mean = r.mean(H[: n // 2])
            dd = np.array([r.distance_d(H[i], mean) for i in range(n)])
            out[f"pf_{name}_{nb}_H"] = H
            out[f"pf_{name}_{nb}_lens"] = lens
            out[f"pf_{name}_{nb}_raw"] = raw
            out[f"pf_{name}_{nb}_dist"] = dist
            out[f"pf_{name}_{nb}_mean"] = mean
            out[f"pf_{name}_{nb}_dd"] = dd

    # ---- get_close arithmetic (Trainer.cpp:81-106) on a u8 batch, 3 and 4 features
    nb, n = 256, 600
    base = np.minimum(rng.poisson(4, nb) + 1, 255)
    H = np.minimum(np.maximum(base[None, :] + rng.integers(-3, 4, (n, nb)), 1), 255).astype(np.uint8)
    H[n // 2:] = np.minimum(rng.poisson(4, (n - n // 2, nb)) + 1, 255).astype(np.uint8)
    lens = (H.sum(1) - nb + 3).astype(np.uint64)
    mins = np.array([0.0, 0.55, 12.0, -0.15, 180.0])
    maxs = np.array([55.0, 0.995, 800.0, 0.9, 420.0])
    out.update(sc_H=H, sc_lens=lens, sc_mins=mins, sc_maxs=maxs)
    for nfeat, w in ((3, [-2.2, 2.4, 1.3, 0.6]), (4, [-3.1, 2.4, 1.3, 0.6, 0.9])):
        s, f0, fl = r.scan(H, lens, H[7], int(lens[7]), mins, maxs, np.array(w), nfeat)
        out[f"sc_w{nfeat}"] = np.array(w)
        out[f"sc_sum{nfeat}"] = s
        out[f"sc_f0{nfeat}"] = f0
        out[f"sc_flag{nfeat}"] = fl

    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
