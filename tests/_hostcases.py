"""Small end-to-end cases for the host control flow (bin/meshclust vs the reference binary at
--threads 1).  The FASTA inputs are regenerated from seeds; the expected CLSTR files are committed
under tests/golden/clstr/ (made by tests/golden/make_host_golden.py with the compiled reference)."""
from __future__ import annotations

import gzip
import os

import numpy as np

from meshclust_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden", "clstr")


def var_len_fasta(path, n, ntemp, lmin, lmax, mu, seed, related=0.0, iupac=False, homopolymer=0, crlf=False):
    rng = np.random.default_rng(seed)
    temps = [rng.integers(0, 4, int(rng.integers(lmin, lmax)), dtype=np.uint8) for _ in range(ntemp)]
    if homopolymer:   # a long single-letter run in every template: some k-mer count exceeds 255 -> 16-bit histograms
        for t_i, t in enumerate(temps):
            s0 = int(rng.integers(0, max(1, t.size - homopolymer)))
            t[s0:s0 + homopolymer] = t_i % 4
    if related > 0:
        anc = temps[0]
        temps = [synth._mutate(rng, anc.copy(), np.array([0, anc.size]), related)[0] for _ in range(ntemp)]
    pieces, offs, tm = [], [0], []
    for i in range(n):
        t = i % ntemp
        c, _ = synth._mutate(rng, temps[t].copy(), np.array([0, temps[t].size]), mu)
        pieces.append(c)
        offs.append(offs[-1] + c.size)
        tm.append(t)
    letters = synth._ACGT[np.concatenate(pieces)].copy()
    if iupac:   # lower case, N runs (short and long) and IUPAC codes sprinkled in
        letters[::3] |= 0x20
        for s in rng.integers(0, letters.size - 40, n // 4):
            letters[s:s + int(rng.choice([1, 3, 12, 30]))] = ord("N")
        letters[rng.integers(0, letters.size, n)] = rng.choice(np.frombuffer(b"RYMKSWHBVD", np.uint8), n)
    synth.write_fasta(path, letters, np.array(offs, np.int64), [f">seq{i} template{tm[i]}" for i in range(n)])
    if crlf:   # DOS line ends, a blank line inside a record and no newline at the very end (safe_getline, ChromListMaker.cpp:23-47)
        data = open(path, "rb").read().replace(b"\n", b"\r\n")
        first = data.index(b"\r\n", data.index(b"\r\n") + 2)
        data = data[:first] + b"\r\n" + data[first:]
        open(path, "wb").write(data.rstrip(b"\r\n"))


# name -> (list of (file name, generator kwargs), CLI arguments)
CASES = {
    "A": ([("A.fa", dict(n=1500, ntemp=15, lmin=300, lmax=301, mu=0.03, seed=11))], ["--id", "0.90", "--kmer", "3"]),
    "B": ([("B.fa", dict(n=2500, ntemp=25, lmin=200, lmax=700, mu=0.03, seed=12))], ["--id", "0.90", "--kmer", "3"]),
    "D": ([("D2.fa", dict(n=800, ntemp=12, lmin=280, lmax=340, mu=0.04, seed=15)),
           ("D1.fa", dict(n=700, ntemp=10, lmin=300, lmax=330, mu=0.04, seed=14))],
          ["--id", "0.88", "--kmer", "3", "--delta", "2", "--iterations", "5"]),
    "E": ([("E.fa", dict(n=1200, ntemp=40, lmin=250, lmax=400, mu=0.02, seed=16, related=0.08))], ["--id", "0.93"]),
    # BASELINE.json configs[0] in full: 10k synthetic 1 kb sequences from 100 mutated templates
    "c1_full": ([("c1.fa", dict(config="c1"))], ["--id", "0.90", "--kmer", "3"]),
    # --align (forced) and automatic alignment below 60 % identity (Runner.cpp:32-34)
    "G": ([("G.fa", dict(n=500, ntemp=10, lmin=180, lmax=240, mu=0.08, seed=18))], ["--id", "0.75", "--align"]),
    "H": ([("H.fa", dict(n=400, ntemp=8, lmin=150, lmax=260, mu=0.15, seed=19))], ["--id", "0.55", "--delta", "3"]),
    # low-complexity runs: the largest k-mer count exceeds 255, so the run uses 16-bit histograms (Runner.cpp:75-89)
    "I": ([("I.fa", dict(n=900, ntemp=12, lmin=900, lmax=1100, mu=0.03, seed=21, homopolymer=330))], ["--id", "0.90", "--kmer", "3"]),
    # harder data (10-12 % mutation): the 3-feature model stays below 97.5 % accuracy, so the greedy selection
    # (Trainer.cpp:603-646) goes on to 4 features (KULCZYNSKI2 joins) and Phase B runs all its iterations
    "K": ([("K.fa", dict(n=2000, ntemp=40, lmin=300, lmax=500, mu=0.10, seed=31))], ["--id", "0.80", "--kmer", "3"]),
    "L": ([("L.fa", dict(n=2500, ntemp=60, lmin=200, lmax=900, mu=0.12, seed=33, related=0.2))], ["--id", "0.80", "--kmer", "4"]),
    # --delta 0 (every center only sees its own members, nothing merges), few iterations
    "M": ([("M.fa", dict(n=1500, ntemp=15, lmin=300, lmax=301, mu=0.03, seed=11))], ["--id", "0.90", "--kmer", "3", "--delta", "0", "--iterations", "3"]),
    # small inputs: fewer points than pivots x picks (duplicated pivots, one bvec bin); 90 sequences with automatic k
    # make the reference warn "Alignment may be too large for sampling" and end with a 4-feature model
    "N": ([("N.fa", dict(n=200, ntemp=5, lmin=250, lmax=320, mu=0.03, seed=41))], ["--id", "0.90", "--kmer", "3"]),
    "O": ([("O.fa", dict(n=90, ntemp=3, lmin=400, lmax=420, mu=0.02, seed=42))], ["--id", "0.95"]),
    # CRLF line ends, an empty line, no final newline: the serial FASTA parser (the parallel one hands such files over)
    "J": ([("J.fa", dict(n=900, ntemp=9, lmin=300, lmax=360, mu=0.03, seed=23, crlf=True))], ["--id", "0.90", "--kmer", "3"]),
    "F": ([("F.fa", dict(n=1300, ntemp=20, lmin=260, lmax=420, mu=0.03, seed=17, iupac=True))],
          ["--id", "0.90", "--kmer", "4", "--sample", "2000", "--pivot", "10"]),
    # 1024- and 4096-bin histograms (k = 5, k = 6: the shapes of BASELINE configs[3] and [4])
    "P": ([("P.fa", dict(n=3000, ntemp=30, lmin=900, lmax=1100, mu=0.03, seed=51))], ["--id", "0.90", "--kmer", "5"]),
    "Q": ([("Q.fa", dict(n=1500, ntemp=15, lmin=3000, lmax=3500, mu=0.03, seed=52))], ["--kmer", "6"]),
    # BASELINE.json configs[1] in full (100k x 1.5 kb, 1000 clusters; the reference needs ~10 min at --threads 1)
    "c2_full": ([("c2.fa", dict(config="c2"))], ["--id", "0.97", "--kmer", "4"]),
    # BASELINE.json configs[2] (forced alignment) on its first 2000 sequences (~30 min of reference time)
    "c3_2k": ([("c3.fa", dict(config="c3", n=2000))], ["--id", "0.70", "--align"]),
}

# cases whose CPU stand-in run (oracle arithmetic on a few host cores) would take many minutes: GPU tests only
GPU_ONLY = {"c2_full", "c3_2k"}
CPU_CASES = [c for c in CASES if c not in GPU_ONLY]


def make_inputs(name: str, workdir: str):
    files, args = CASES[name]
    paths = []
    for fname, kw in files:
        p = os.path.join(workdir, fname)
        if "config" in kw:
            letters, offs, tmpl = synth.generate_config(kw["config"], kw.get("n"))
            synth.write_fasta(p, letters, offs, synth.headers_for(offs.size - 1, tmpl))
        else:
            var_len_fasta(p, **kw)
        paths.append(p)
    return paths, list(args)


def golden_path(name: str) -> str:
    return os.path.join(GOLDEN_DIR, f"{name}.clstr.gz")


def read_golden(name: str) -> bytes:
    with gzip.open(golden_path(name), "rb") as f:
        return f.read()
