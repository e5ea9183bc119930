"""Seeded synthetic mutated-template DNA.

Templates are i.i.d. uniform ACGT of length L.  Sequence i is template ``i % Tn`` with a
per-base mutation rate ``mu`` split 80 % substitution / 10 % deletion / 10 % insertion.
Header is ``>seq{i} template{t}``; FASTA lines are 70 columns; upper case, no N.

The five BASELINE.json configs are in :data:`CONFIGS`.  Everything is vectorised numpy so the
1 M-sequence config is generated in seconds; the same seed always gives the same bytes.
"""
from __future__ import annotations

import dataclasses
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclasses.dataclass(frozen=True)
class Config:
    name: str
    n: int
    templates: int
    length: int
    mu: float
    seed: int
    identity: float
    kmer: int | None
    align: bool = False
    related: float = 0.0   # >0: templates are one ancestor mutated by this rate (16S-like)


CONFIGS = {
    "c1": Config("c1", 10_000, 100, 1000, 0.03, 1, 0.90, 3),
    "c2": Config("c2", 100_000, 1000, 1500, 0.01, 2, 0.97, 4, related=0.15),
    "c3": Config("c3", 20_000, 200, 500, 0.10, 3, 0.70, None, align=True),
    "c4": Config("c4", 1_000_000, 1000, 1000, 0.03, 4, 0.90, 5),
    "c5": Config("c5", 200_000, 500, 10_000, 0.03, 5, 0.90, 6),
}


def _mutate(rng: np.random.Generator, codes: np.ndarray, offs: np.ndarray, mu: float):
    """Mutate a batch of concatenated 0..3 code strings. Returns (codes, offsets)."""
    total = codes.size
    hit = rng.random(total) < mu
    kind = rng.random(total)            # <0.8 sub, <0.9 del, else ins
    sub = hit & (kind < 0.8)
    dele = hit & (kind >= 0.8) & (kind < 0.9)
    ins = hit & (kind >= 0.9)
    out = codes.copy()
    nsub = int(sub.sum())
    out[sub] = (out[sub] + rng.integers(1, 4, nsub, dtype=np.uint8)) & 3
    count = np.ones(total, dtype=np.int64)
    count[dele] = 0
    count[ins] = 2
    seq_id = np.repeat(np.arange(offs.size - 1), np.diff(offs))
    new_len = np.bincount(seq_id, weights=count, minlength=offs.size - 1).astype(np.int64)
    expanded = np.repeat(out, count)
    # second copy of an inserted base becomes a fresh random base
    pos = np.cumsum(count) - 1          # index of the LAST emitted copy of each source base
    ins_pos = pos[ins]
    expanded[ins_pos] = rng.integers(0, 4, ins_pos.size, dtype=np.uint8)
    new_offs = np.zeros(offs.size, dtype=np.int64)
    np.cumsum(new_len, out=new_offs[1:])
    return expanded, new_offs


def generate(n: int, templates: int, length: int, mu: float, seed: int, related: float = 0.0,
             batch: int = 1 << 16):
    """Returns (letters uint8[total], offsets int64[n+1], template_of int32[n])."""
    rng = np.random.default_rng(seed)
    if related > 0:
        anc = rng.integers(0, 4, length, dtype=np.uint8)
        tcodes = np.tile(anc, templates)
        toffs = np.arange(templates + 1, dtype=np.int64) * length
        tcodes, toffs = _mutate(rng, tcodes, toffs, related)
    else:
        tcodes = rng.integers(0, 4, templates * length, dtype=np.uint8)
        toffs = np.arange(templates + 1, dtype=np.int64) * length
    tlen = np.diff(toffs)
    pieces, lens = [], []
    template_of = (np.arange(n) % templates).astype(np.int32)
    for b0 in range(0, n, batch):
        tid = template_of[b0:b0 + batch]
        ln = tlen[tid]
        offs = np.zeros(tid.size + 1, dtype=np.int64)
        np.cumsum(ln, out=offs[1:])
        # gather template bases
        idx = np.repeat(toffs[tid] - offs[:-1], ln) + np.arange(offs[-1])
        codes = tcodes[idx]
        codes, offs = _mutate(rng, codes, offs, mu)
        pieces.append(codes)
        lens.append(np.diff(offs))
    codes = np.concatenate(pieces)
    lens = np.concatenate(lens)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    return _ACGT[codes], offsets, template_of


def generate_config(name: str, n: int | None = None):
    """Generate a BASELINE config (optionally truncated to the first n sequences' worth)."""
    c = CONFIGS[name]
    nn = c.n if n is None else n
    return generate(nn, min(c.templates, nn), c.length, c.mu, c.seed, c.related)


def headers_for(n: int, template_of: np.ndarray) -> list[str]:
    return [f">seq{i} template{int(template_of[i])}" for i in range(n)]


def write_fasta(path: str, letters: np.ndarray, offsets: np.ndarray, headers: list[str], width: int = 70):
    with open(path, "wb") as f:
        buf = []
        for i, h in enumerate(headers):
            s = letters[offsets[i]:offsets[i + 1]].tobytes()
            buf.append(h.encode() + b"\n")
            for j in range(0, len(s), width):
                buf.append(s[j:j + width] + b"\n")
            if len(buf) > 100000:
                f.write(b"".join(buf))
                buf = []
        f.write(b"".join(buf))
