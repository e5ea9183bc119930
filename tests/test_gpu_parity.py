"""GPU: every stage of the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs and against the committed golden vectors (outputs of the reference itself).

Bars: histograms, digit strings, distance keys, alignment (score, len, matches), scan counts and
argmax rows -- bit-exact.  Raw features -- bit-exact (they derive from exact integer reductions
through the same FP64 formulas).  GLM sum / f0 -- stated tolerance 1e-12 relative; decisions must
agree except for pairs with |sum| < 1e-9, which are counted.
"""
import numpy as np
import pytest

from meshclust_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-12
NEAR = 1e-9


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def _model(nfeat):
    mins = np.array([0.0, 0.55, 12.0, -0.15, 180.0])
    maxs = np.array([55.0, 0.995, 800.0, 0.9, 420.0])
    w = np.array([-2.2, 2.4, 1.3, 0.6] if nfeat == 3 else [-3.1, 2.4, 1.3, 0.6, 0.9])
    return mins, maxs, w


# ----------------------------------------------------------------------------------------------
# stage 0/1
# ----------------------------------------------------------------------------------------------
def test_encode_and_hist_golden(ctx, golden):
    letters, offs = golden["enc_letters"], golden["enc_offs"]
    ctx.load_sequences(letters, offs)
    assert np.array_equal(ctx.copy_digits(), golden["enc_digits"])
    for k in (1, 2, 3, 4, 5, 6):
        want = golden[f"hist_k{k}"]
        used, mx = ctx.build_histograms(k, 0)
        assert mx == int(want.max())
        assert used == (1 if mx <= 255 else 2)
        got = ctx.copy_histograms()
        assert np.array_equal(got.astype(np.uint16), want)
        ln, mg, sq = ctx.copy_point_stats()
        assert np.array_equal(ln, np.diff(offs).astype(np.uint64))
        assert np.array_equal(mg, want.astype(np.uint64).sum(1))
        assert np.array_equal(sq, (want.astype(np.uint64) ** 2).sum(1))
        # forced 16-bit bins give the same counts
        ctx.build_histograms(k, 2)
        assert np.array_equal(ctx.copy_histograms(), want)


def _fasta_text(rng, n, golden_like=True):
    """A FASTA file as bytes + what a parser makes of it: letters per record; ragged line widths, empty lines,
    N runs, IUPAC codes, lower case, records shorter than 20 letters, no newline at the very end."""
    recs, out = [], bytearray()
    spans = []
    for i in range(n):
        out += f">rec{i} some description {i * 7}\n".encode()
        kind = rng.integers(0, 6)
        L = int(rng.integers(2, 30)) if kind == 5 else int(rng.integers(20, 900))
        seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), L)
        if kind == 1:
            seq = seq | 0x20                                  # lower case
        if kind == 2 and L > 60:                               # N runs, short and long, also at the ends
            for _ in range(int(rng.integers(1, 4))):
                a = int(rng.integers(0, L - 1))
                seq[a:a + int(rng.integers(1, 40))] = ord("N") if rng.random() < 0.7 else ord("n")
        if kind == 3:                                          # IUPAC codes
            pos = rng.integers(0, L, 5)
            seq[pos] = rng.choice(np.frombuffer(b"RYMKSWHBVDX", np.uint8), 5)
        begin = len(out)
        width = int(rng.integers(1, 100))
        for a in range(0, L, width):
            out += seq[a:a + width].tobytes() + b"\n"
            if rng.random() < 0.05:
                out += b"\n"                                   # an empty line inside the record
        spans.append((begin, len(out)))
        recs.append(seq.copy())
    while out and out[-1:] == b"\n":
        out = out[:-1]
    spans[-1] = (spans[-1][0], len(out))
    return np.frombuffer(bytes(out), np.uint8), spans, recs


@pytest.mark.parametrize("n,k", [(700, 3), (64, 4), (1, 2)])
def test_ingest_fasta_vs_host_parser(ctx, oracle, n, k):
    """mc_ingest_fasta (raw file bytes + per-record spans -> letters, in a permuted row order, + N / non-ACGT flags
    on the device; SURVEY 8(f1), ChromListMaker.cpp:92-120, Chromosome.cpp:162-184) against the letters a parser
    extracts on the host, and the histograms built from them against the host-letter path and the oracle."""
    from meshclust_b200.api import McError, segments_for_batch
    rng = np.random.default_rng(4100 + n)
    raw, spans, recs = _fasta_text(rng, n)
    order = rng.permutation(n)                                 # row r holds record order[r]
    lens = np.array([recs[i].size for i in order], np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    sb = np.array([spans[i][0] for i in order], np.int64)
    se = np.array([spans[i][1] for i in order], np.int64)
    letters = np.concatenate([recs[i] for i in order]).astype(np.uint8)
    if n > 1:
        ctx.stage_fasta_bytes(raw, n)       # the bytes go ahead; the ingest below must find them (same buffer)
    flags = ctx.ingest_fasta(raw, sb, se, offs)
    assert np.array_equal(ctx.copy_letters(), letters)
    up = letters & 0xDF
    has_n = np.add.reduceat((up == ord("N")).astype(np.int64), offs[:-1]) > 0
    other = np.add.reduceat((~np.isin(up, np.frombuffer(b"ACGTN", np.uint8))).astype(np.int64), offs[:-1]) > 0
    assert np.array_equal((flags & 1) != 0, has_n) and np.array_equal((flags & 2) != 0, other)
    segs, seg_off = segments_for_batch(letters, offs)
    ctx.load_segments(segs, seg_off, validate=bool(other.any()))
    used, mx = ctx.build_histograms(k, 0)
    got = ctx.copy_histograms()
    rc, want, wmx = oracle.hist_batch(letters, offs, k, used)
    assert rc == 0 and mx == wmx and np.array_equal(got, want)
    ln = ctx.copy_point_stats()[0]
    assert np.array_equal(ln, lens.astype(np.uint64))
    # the digit strings the aligner reads come out of the same buffer
    ctx2_digits = ctx.copy_digits()
    ctx.load_sequences(letters, offs, segs, seg_off)
    assert np.array_equal(ctx.copy_digits(), ctx2_digits)
    # spans and letter counts that disagree are an input error, not a silent truncation
    long_rows = np.nonzero(lens >= 20)[0]
    if long_rows.size:
        r0 = int(long_rows[0])
        lens2 = lens.copy()
        lens2[r0] -= 1
        with pytest.raises(McError):
            ctx.ingest_fasta(raw, sb, se, np.concatenate([[0], np.cumsum(lens2)]).astype(np.int64))
        # an invalid letter is reported by mc_load_segments when validation is asked for (records with a segment only)
        raw2 = raw.copy()
        p0 = int(sb[r0])
        while raw2[p0] == 10:
            p0 += 1
        raw2[p0] = ord("!")
        fl2 = ctx.ingest_fasta(raw2, sb, se, offs)
        assert fl2[r0] & 2
        with pytest.raises(McError):
            ctx.load_segments(segs, seg_off, validate=True)


@pytest.mark.parametrize("cfg,n,k", [("c1", 3000, 3), ("c2", 2000, 4), ("c4", 2000, 5), ("c5", 300, 6), ("c3", 1500, 4)])
def test_hist_vs_oracle_configs(ctx, oracle, cfg, n, k):
    letters, offs, _ = synth.generate_config(cfg, n)
    rc, want, mx = oracle.hist_batch(letters, offs, k, 1)
    assert rc == 0 and mx <= 255
    got, gmx = ctx.kmer_histograms_host(letters, offs, k, 1)
    assert gmx == mx
    assert np.array_equal(got, want)


@pytest.mark.parametrize("k,width", [(7, 1), (7, 2), (1, 2), (2, 2)])
def test_hist_extreme_k(ctx, oracle, k, width):
    # the largest supported k (64 KB of counters per sequence: two warps per CTA) and the smallest, both bin widths;
    # ragged lengths incl. sequences shorter than a 16-letter chunk, lower case, IUPAC and N runs (segments)
    rng = np.random.default_rng(70 + k)
    seqs = []
    for i in range(60):
        L = int(rng.integers(21, 40)) if i % 7 == 0 else int(rng.integers(200, 3000))
        sq = rng.choice(np.frombuffer(b"ACGT", np.uint8), L)
        if i % 5 == 1:
            sq = sq | 0x20
        if i % 5 == 2 and L > 100:
            sq[30:55] = ord("N")
            sq[L // 2] = ord("R")
        seqs.append(sq.astype(np.uint8))
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([s.size for s in seqs], out=offs[1:])
    letters = np.concatenate(seqs)
    rc, want, mx = oracle.hist_batch(letters, offs, k, width)
    assert rc == 0
    got, gmx = ctx.kmer_histograms_host(letters, offs, k, width)
    assert gmx == mx and np.array_equal(got, want)


def test_hist_u16_homopolymer(ctx, oracle):
    # long homopolymer runs push one bin past 255 -> 16-bit histograms (Runner.cpp:75-89)
    rng = np.random.default_rng(0)
    seqs = [b"A" * 700 + bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 300)), b"ACGT" * 200 + b"T" * 400,
            bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 900))]
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    letters = np.frombuffer(b"".join(seqs), np.uint8)
    rc, want, mx = oracle.hist_batch(letters, offs, 3, 2)
    assert mx > 255
    ctx.load_sequences(letters, offs)
    used, gmx = ctx.build_histograms(3, 0)
    assert (used, gmx) == (2, mx)
    assert np.array_equal(ctx.copy_histograms(), want)
    with pytest.raises(Exception):
        ctx.build_histograms(3, 1)       # forcing 8-bit bins must fail loudly, not wrap


def test_invalid_input_is_rejected(ctx, built_lib):
    s = np.frombuffer(b"ACGTACGTACGTACGTACGTAC-GTACGT", np.uint8)
    with pytest.raises(built_lib.McError) as e:
        ctx.load_sequences(s, np.array([0, s.size], np.int64))
    assert e.value.code == built_lib.MC_ERR_INPUT
    for bad in (b"NNNNNNNN", b"A", b"NNNNNNNA"):
        with pytest.raises(built_lib.McError):
            ctx.load_sequences(np.frombuffer(bad, np.uint8), np.array([0, len(bad)], np.int64))


def test_ragged_and_tiny_sequences(ctx, oracle):
    # 2-base, < 20 bp, exactly 20 bp, unaligned neighbours, N runs at both ends
    # (a 1-base record is INVALID in the reference: its only run starts on the last character and
    # is never closed, Chromosome.cpp:162-184 -> segment->at(0) throws)
    seqs = [b"AC", b"ACGTACGTAC", b"ACGTACGTACGTACGTACGT", b"C" * 33, b"NNNACGTACGTACGTACGTACGTACGTNN", b"G" * 17,
            b"ACGTTGCAAC" * 13 + b"G", b"T" * 16, b"ACGGTCA" * 50]
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    letters = np.frombuffer(b"".join(seqs), np.uint8)
    for k in (1, 2, 4, 6):
        rc, want, _ = oracle.hist_batch(letters, offs, k, 1)
        assert rc == 0
        got, _ = ctx.kmer_histograms_host(letters, offs, k, 1)
        assert np.array_equal(got, want)
    import _oracle
    assert np.array_equal(ctx.copy_digits(), _oracle.encode_digits(letters, offs))


# ----------------------------------------------------------------------------------------------
# stage 2
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["u8", "u16"])
@pytest.mark.parametrize("nb", [16, 64, 256, 1024])
def test_pair_features_golden(ctx, golden, name, nb):
    H, lens = golden[f"pf_{name}_{nb}_H"], golden[f"pf_{name}_{nb}_lens"]
    n = H.shape[0]
    k = {16: 2, 64: 3, 256: 4, 1024: 5}[nb]
    ctx.load_histograms(H, lens, k)
    a, b = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    raw, dist = ctx.pair_features(a.reshape(-1), b.reshape(-1))
    assert np.array_equal(_bits(raw), _bits(golden[f"pf_{name}_{nb}_raw"].reshape(-1, 5)))
    assert np.array_equal(dist, golden[f"pf_{name}_{nb}_dist"].reshape(-1))
    keys = ctx.distance_keys(np.arange(n))
    assert np.array_equal(keys.astype(np.uint64), golden[f"pf_{name}_{nb}_dist"])
    # mean of the first half + nearest member (get_mean) against the golden distance_d values
    row, d = ctx.mean_nearest(np.arange(n // 2))
    dd = golden[f"pf_{name}_{nb}_dd"][: n // 2]
    assert row == int(np.argmin(dd)) and _bits(d) == _bits(dd.min())


@pytest.mark.parametrize("nfeat", [3, 4])
def test_scan_golden(ctx, golden, nfeat):
    H, lens = golden["sc_H"], golden["sc_lens"]
    n = H.shape[0]
    ctx.load_histograms(H, lens, 4)
    ctx.set_model(golden["sc_mins"], golden["sc_maxs"], golden[f"sc_w{nfeat}"], nfeat)
    want_sum, want_f0, want_flag = golden[f"sc_sum{nfeat}"], golden[f"sc_f0{nfeat}"], golden[f"sc_flag{nfeat}"]
    s, f0, fl, _ = ctx.pair_classify(np.arange(n), np.full(n, 7))
    assert np.allclose(s, want_sum, rtol=RTOL, atol=0)
    assert np.allclose(f0, want_f0, rtol=RTOL, atol=0)
    near = np.abs(want_sum) < NEAR
    assert np.array_equal(fl[~near], want_flag[~near])
    res, marks = ctx.scan(7, 0, n - 1)
    assert res.n_eval == n
    assert np.array_equal(marks[~near], want_flag[~near])
    assert res.n_pos == int(marks.sum())
    assert res.best_row == int(np.argmax(want_f0)) and res.best_f0 == want_f0.max()


def _rand_hists(rng, n, nb, dtype=np.uint8, hi=255, clusters=8):
    bases = np.minimum(rng.poisson(max(1, 1000 // nb) + 1, (clusters, nb)) + 1, hi)
    H = bases[rng.integers(0, clusters, n)] + rng.integers(-2, 3, (n, nb))
    return np.minimum(np.maximum(H, 1), hi).astype(dtype)


@pytest.mark.parametrize("k,n", [(1, 700), (2, 1000), (3, 5000), (4, 4000), (5, 3000), (6, 700), (7, 300)])
@pytest.mark.parametrize("nfeat", [3, 4])
def test_scan_vs_oracle(ctx, oracle, k, n, nfeat):
    rng = np.random.default_rng(100 + k)
    nb = 4 ** k
    H = _rand_hists(rng, n, nb)
    lens = (H.astype(np.int64).sum(1) - nb + k - 1 + rng.integers(0, 3, n)).astype(np.uint64)
    mins, maxs, w = _model(nfeat)
    maxs[2] = 4.0 * nb
    maxs[4] = 2.0 * nb
    mins[4] = 0.5 * nb
    ctx.load_histograms(H, lens, k)
    ctx.set_model(mins, maxs, w, nfeat)
    total_near = 0
    for center in (0, n // 3, n - 1):
        ctx.alive_reset()
        ws, wf0, wfl = oracle.scan(H, lens, H[center], int(lens[center]), mins, maxs, w, nfeat)
        near = np.abs(ws) < NEAR
        total_near += int(near.sum())
        # ragged sub-range + a few dead rows
        lo, hi = 5, n - 7
        dead = np.array([lo, lo + 17, hi, (lo + hi) // 2])
        ctx.alive_kill(dead)
        live = np.ones(n, bool)
        live[dead] = False
        live[:lo] = False
        live[hi + 1:] = False
        res, marks = ctx.scan(center, lo, hi)
        idx = np.nonzero(live)[0]
        assert res.n_eval == idx.size
        full = np.zeros(n, np.uint8)
        full[lo:hi + 1] = marks
        ok = live & ~near
        assert np.array_equal(full[ok], wfl[ok])
        assert not full[~live].any()
        cand = wf0[idx]
        best = idx[int(np.argmax(cand))] if cand.max() > -1 else -1
        assert res.best_row == best
        if best >= 0:
            assert np.isclose(res.best_f0, wf0[best], rtol=RTOL, atol=0)
        # marked rows left the alive set: a second scan sees only the survivors
        res2, marks2 = ctx.scan(center, lo, hi)
        assert res2.n_eval == idx.size - int(marks.sum())
        assert res2.n_pos == 0 or near.any()
    print(f"near-threshold pairs (|sum| < {NEAR}): {total_near}")


@pytest.mark.parametrize("nfeat", [3, 4])
@pytest.mark.parametrize("k", [3, 4, 5])
def test_scan_decisions_at_the_threshold(ctx, oracle, k, nfeat):
    """The scan decides most rows from an interval of the GLM sum (FP32 PEARSON / KULCZYNSKI2 with an error
    bound, mc_scan_decide) and runs the exact FP64 terms only when the interval touches the threshold.  Models
    whose bias puts the threshold ON a row's sum, and 1e-12 .. 1e-3 beside it, exercise both sides of that
    filter: every decision outside |sum| < 1e-9 must be the oracle's, and the near-threshold counter must
    count exactly the rows inside."""
    rng = np.random.default_rng(7000 + 10 * k + nfeat)
    nb, n = 4 ** k, 3000
    H = _rand_hists(rng, n, nb, clusters=5)
    lens = (H.astype(np.int64).sum(1) - nb + k - 1 + rng.integers(0, 3, n)).astype(np.uint64)
    mins, maxs, w = _model(nfeat)
    maxs[2] = 4.0 * nb
    maxs[4] = 2.0 * nb
    mins[4] = 0.5 * nb
    ctx.load_histograms(H, lens, k)
    center = 17
    base, _, _ = oracle.scan(H, lens, H[center], int(lens[center]), mins, maxs, w, nfeat)
    targets = rng.choice(n, 6, replace=False)
    flips = 0
    for t, eps in zip(targets, (0.0, 1e-12, -1e-10, 3e-8, -1e-6, 2e-4)):
        w2 = w.copy()
        w2[0] = w[0] - base[t] + eps      # row t now sums to ~eps; its same-cluster neighbours lie close by
        ws, wf0, wfl = oracle.scan(H, lens, H[center], int(lens[center]), mins, maxs, w2, nfeat)
        near = np.abs(ws) < NEAR
        ctx.set_model(mins, maxs, w2, nfeat)
        ctx.alive_reset()
        ctx.near_threshold_count(reset=True)
        res, marks = ctx.scan(center, 0, n - 1)
        got_near = ctx.near_threshold_count(reset=True)
        live = np.ones(n, bool)
        ok = live & ~near
        assert np.array_equal(marks[ok], wfl[ok]), f"decision differs from the oracle away from the threshold (eps {eps})"
        assert got_near == int((near & live).sum())
        assert res.n_pos == int(marks[live].sum())
        flips += int((np.abs(ws[live]) < 1e-3).sum())
    assert flips >= 6    # the models did put rows next to the threshold


@pytest.mark.parametrize("dtype,k,C", [(np.uint8, 4, 150), (np.uint8, 5, 37), (np.uint8, 6, 150), (np.uint16, 3, 33), (np.uint16, 4, 150), (np.uint16, 5, 5), (np.uint8, 3, 150)])
def test_distance_keys_tile_equals_pair_lists(ctx, dtype, k, C):
    """K2b as a tile (C centers staged in shared memory, every row read once; rows of 128 bytes and more) against
    the one-warp-per-pair kernel on ALL C x n pairs, for center counts that are not multiples of 32 and for more
    centers than shared memory holds at once (k = 6: 4 KB rows)."""
    rng = np.random.default_rng(8100 + k)
    nb, n = 4 ** k, 1111
    H = _rand_hists(rng, n, nb, dtype, 255 if dtype == np.uint8 else 2000, clusters=7)
    ctx.load_histograms(H, np.full(n, 900, np.uint64), k)
    centers = rng.integers(0, n, C).astype(np.int32)
    centers[0], centers[-1] = 0, n - 1
    keys = ctx.distance_keys(centers)
    a = np.repeat(centers, n).astype(np.int32)
    b = np.tile(np.arange(n, dtype=np.int32), C)
    _, dist = ctx.pair_features(b, a)
    assert np.array_equal(keys.astype(np.uint64).reshape(-1), dist)


def test_histograms_from_letters_and_from_digits(ctx, oracle):
    """K1 counts LETTERS in one pass; the digit strings the aligner reads are made lazily, in place.  A buffer
    that has been encoded already (mc_copy_digits forces it) must give the same histograms."""
    l, o, _ = synth.generate_config("c2", 300)
    l = l.copy()
    l[5:40] = ord("N")          # a run of N: two segments
    l[200] = ord("n")
    l[900:905] = np.frombuffer(b"RYKMS", np.uint8)
    l[1500:1600] |= 0x20        # lower case
    rc, want, _ = oracle.hist_batch(l, o, 4, 1)
    assert rc == 0
    ctx.load_sequences(l, o)
    ctx.build_histograms(4, 1)
    assert np.array_equal(ctx.copy_histograms(), want)
    ctx.copy_digits()           # in-place encode
    ctx.build_histograms(4, 1)
    assert np.array_equal(ctx.copy_histograms(), want)


@pytest.mark.parametrize("k,n", [(2, 3000), (4, 5000), (5, 2100), (6, 600)])
def test_scan_batch_launch_equals_single_scans(ctx, k, n, monkeypatch):
    """independent scans (nothing removed) share one launch (blockIdx.y = scan): same summaries as one
    launch per scan, for more scans than one launch carries, ragged ranges and an empty alive set"""
    rng = np.random.default_rng(1200 + k)
    nb = 4 ** k
    H = _rand_hists(rng, n, nb, clusters=6)
    lens = (900 + rng.integers(0, 60, n)).astype(np.uint64)
    mins, maxs, w = _model(4)
    maxs[2] = 4.0 * nb
    maxs[4] = 2.0 * nb
    mins[4] = 0.5 * nb
    ctx.load_histograms(H, lens, k)
    ctx.set_model(mins, maxs, w, 4)
    ctx.alive_kill(rng.integers(0, n, n // 5))
    m = 37
    cr = rng.integers(0, n, m)
    lo = rng.integers(0, n // 3, m)
    hi = n - 1 - rng.integers(0, n // 3, m)
    lo[3], hi[3] = 17, 17
    ctx.scan_enqueue_many(cr, lo, hi, False, 0)            # batched launches
    got = ctx.scan_collect(0, m)
    monkeypatch.setenv("MC_SCAN_NO_BATCH", "1")
    ctx.scan_enqueue_many(cr, lo, hi, False, 100)          # one launch per scan
    want = ctx.scan_collect(100, m)
    assert got == want
    one = [ctx.scan_enqueue(int(cr[i]), int(lo[i]), int(hi[i]), False, 300 + i) for i in range(5)]
    assert ctx.scan_collect(300, 5) == want[:5]
    monkeypatch.delenv("MC_SCAN_NO_BATCH")
    # the same scans as a dependent chain that removes nothing (MC_SCAN_CHAIN)
    from meshclust_b200 import api
    ctx.scan_enqueue_many(cr, lo, hi, api.MC_SCAN_CHAIN, 400)
    assert ctx.scan_collect(400, m) == want


@pytest.mark.parametrize("dtype,k,n,nc", [(np.uint8, 4, 20000, 10), (np.uint8, 5, 9000, 3), (np.uint16, 3, 6000, 16), (np.uint8, 4, 3000, 4), (np.uint8, 2, 5000, 20)])
def test_scan_host_pipeline_equals_resident_scans(ctx, dtype, k, n, nc):
    """mc_scan_host (chunked upload overlapped with the scans, every center with its own mark array)
    against mc_load_histograms + one mc_scan per center on the resident rows"""
    from meshclust_b200 import api
    rng = np.random.default_rng(1500 + k)
    nb = 4 ** k
    H = _rand_hists(rng, n, nb, dtype, 255 if dtype == np.uint8 else 1500, clusters=6)
    lens = (900 + rng.integers(0, 120, n)).astype(np.uint64)
    mins, maxs, w = _model(4)
    maxs[2] = 4.0 * nb
    maxs[4] = 2.0 * nb
    mins[4] = 0.5 * nb
    centers = rng.integers(0, n, nc)
    centers[0], centers[-1] = n - 1, 0
    ctx.set_model(mins, maxs, w, 4)
    marks = np.full((nc, n), 7, np.uint8)
    got = ctx.scan_host(H, lens, k, centers, marks)
    with api.Context(0) as c2:
        c2.load_histograms(H, lens, k)
        c2.set_model(mins, maxs, w, 4)
        for i, c in enumerate(centers):
            c2.alive_reset()
            want, wm = c2.scan(int(c), 0, n - 1)
            assert got[i] == want.as_tuple(), (i, got[i], want.as_tuple())
            assert np.array_equal(marks[i], wm)
    # the rows stay resident for later calls
    res, _ = ctx.scan(int(centers[1]), 0, n - 1)
    assert res.n_eval == n


def test_scan_first_max_wins_and_null(ctx, oracle):
    # identical rows tie on f0: the first one in row order must win; f0 <= -1 everywhere -> no seed
    rng = np.random.default_rng(9)
    nb, n = 256, 2000
    H = _rand_hists(rng, n, nb, clusters=2)
    H[100] = H[900]
    H[1500] = H[900]
    lens = np.full(n, 300, np.uint64)
    mins, maxs, w = _model(3)
    ctx.load_histograms(H, lens, 4)
    ctx.set_model(mins, maxs, w, 3)
    _, wf0, _ = oracle.scan(H, lens, H[900], 300, mins, maxs, w, 3)
    res, _ = ctx.scan(900, 0, n - 1)
    assert res.best_row == int(np.argmax(wf0)) == 100
    # length difference far beyond the training max drives LD' (and f0) below -1
    lens2 = np.full(n, 300, np.uint64)
    lens2[5] = 100000
    ctx.load_histograms(H, lens2, 4)
    ctx.set_model(mins, maxs, w, 3)
    ctx.alive_kill(np.array([5]))
    ws, wf0, wfl = oracle.scan(H, lens2, H[5], 100000, mins, maxs, w, 3)
    assert wf0[np.arange(n) != 5].max() <= -1
    res, _ = ctx.scan(5, 0, n - 1)
    assert res.best_row == -1 and res.best_f0 == -1.0
    assert res.n_pos == int(wfl.sum()) - int(wfl[5])
    # empty range
    res, _ = ctx.scan(5, 10, 9)
    assert res.as_tuple() == (0, 0, -1, -1.0)


@pytest.mark.parametrize("k", [3, 4, 5])
def test_distance_keys_vs_oracle(ctx, oracle, k):
    rng = np.random.default_rng(40 + k)
    nb, n = 4 ** k, 3000
    H = _rand_hists(rng, n, nb)
    lens = np.full(n, 1000, np.uint64)
    ctx.load_histograms(H, lens, k)
    centers = rng.integers(0, n, 9)
    keys = ctx.distance_keys(centers)
    for ci, c in enumerate(centers):
        rows = rng.integers(0, n, 200)
        want = [oracle.features(H[r], H[c], 1000, 1000)[1] for r in rows]
        assert keys[ci, rows].tolist() == want


# ----------------------------------------------------------------------------------------------
# stage 3
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,hi", [(np.uint8, 255), (np.uint16, 3000)])
@pytest.mark.parametrize("k", [2, 4, 6])
def test_mean_nearest_vs_oracle(ctx, oracle, dtype, hi, k):
    rng = np.random.default_rng(60 + k)
    nb, n = 4 ** k, 500
    H = _rand_hists(rng, n, nb, dtype, hi, clusters=3)
    ctx.load_histograms(H, np.full(n, 999, np.uint64), k)
    members = rng.permutation(n)[:120]
    members[7] = members[3]          # a duplicate row ties with itself: first position wins
    pieces = [members[:1], members[1:40], members[40:41], members[41:]]
    got = None
    for i, p in enumerate(pieces):   # accumulate() grows `current` and re-runs get_mean
        got = ctx.mean_nearest(p, append=i > 0)
        cur = members[: sum(len(q) for q in pieces[: i + 1])]
        mean = oracle.mean(H[cur])
        dd = np.array([oracle.distance_d(H[r], mean) for r in cur])
        assert got[0] == int(cur[int(np.argmin(dd))])
        assert _bits(got[1]) == _bits(dd.min())


@pytest.mark.parametrize("dtype,hi,k,n", [(np.uint8, 255, 3, 6000), (np.uint8, 255, 5, 3000), (np.uint16, 3000, 4, 2500), (np.uint8, 200, 6, 900)])
def test_accumulate_step_vs_unfused_and_oracle(ctx, oracle, dtype, hi, k, n):
    """mc_accumulate_step (scan + remove + get_mean in one submission) against the same greedy loop
    driven through mc_scan + mc_mean_nearest, and against the oracle's scan / mean / distance_d."""
    from meshclust_b200 import api
    rng = np.random.default_rng(300 + k)
    nb = 4 ** k
    H = _rand_hists(rng, n, nb, dtype, hi, clusters=7)
    lens = (1000 + rng.integers(0, 40, n)).astype(np.uint64)
    mins, maxs, w = _model(3)
    maxs[2] = 4.0 * nb * (hi / 255.0)
    # pick the bias so that a scan marks a sizeable but partial set of rows
    ctx.load_histograms(H, lens, k)
    ctx.set_model(mins, maxs, w, 3)
    s0, _, _ = oracle.scan(H, lens, H[0], int(lens[0]), mins, maxs, w, 3)
    w = w.copy()
    w[0] -= np.sort(s0)[int(0.9 * n)]
    ctx.set_model(mins, maxs, w, 3)

    def greedy(step_fn, reset_fn):
        reset_fn()
        alive = np.ones(n, bool)
        trace = []
        seed = 0
        for _ in range(6):
            alive[seed] = False
            last, restart, is_min = seed, True, False
            lo, hi_ = 3, n - 4
            while not is_min:
                scan, nearest, rows = step_fn(last, lo, hi_, restart)
                restart = False
                trace.append((last, scan[0], scan[1], scan[2], nearest, tuple(rows[:50]), len(rows)))
                alive[rows] = False
                is_min = scan[1] == 0
                if not is_min:
                    last = nearest
                if len(trace) > 60:
                    break
            nxt = scan[2]
            if nxt < 0:
                break
            seed = int(nxt)
        return trace

    # fused
    c1 = ctx

    def reset1():
        c1.alive_reset()

    def step1(last, lo, hi_, restart):
        if restart:
            c1.alive_kill(np.array([last]))
        r, rows = c1.accumulate_step(last, lo, hi_, restart)
        return r.scan.as_tuple(), r.nearest_row, rows

    t1 = greedy(step1, reset1)

    # unfused on a second context
    with api.Context(0) as c2:
        c2.load_histograms(H, lens, k)
        c2.set_model(mins, maxs, w, 3)
        state = {"cur": None}

        def reset2():
            c2.alive_reset()

        def step2(last, lo, hi_, restart):
            if restart:
                c2.alive_kill(np.array([last]))
                state["cur"] = [last]
            res, marks = c2.scan(last, lo, hi_)
            rows = np.nonzero(marks)[0] + lo
            nearest = -1
            if rows.size:
                state["cur"].extend(rows.tolist())
                nearest, _ = c2.mean_nearest(np.array(state["cur"], np.int64))
            return res.as_tuple(), nearest, rows

        t2 = greedy(step2, reset2)
    assert len(t1) == len(t2) and len(t1) >= 6
    for a, b in zip(t1, t2):
        assert a[:5] == b[:5] and a[6] == b[6] and a[5] == b[5], (a, b)
    assert sum(t[6] for t in t1) > 50, "the test model marks nothing: no mean was exercised"

    # oracle check of the first productive step: marks and the nearest member
    ctx.alive_reset()
    ctx.alive_kill(np.array([0]))
    r, rows = ctx.accumulate_step(0, 0, n - 1, True)
    ws, wf0, wfl = oracle.scan(H, lens, H[0], int(lens[0]), mins, maxs, w, 3)
    near = np.abs(ws) < NEAR
    want = np.nonzero(wfl.astype(bool) & (np.arange(n) != 0))[0]
    if not near.any():
        assert np.array_equal(rows, want)
    cur = np.concatenate([[0], rows])
    mean = oracle.mean(H[cur])
    dd = np.array([oracle.distance_d(H[x], mean) for x in cur])
    assert r.nearest_row == int(cur[int(np.argmin(dd))])
    assert r.n_members == cur.size
    # empty range: nothing evaluated, nothing marked, the cluster is just its seed
    r, rows = ctx.accumulate_step(5, 10, 9, True)
    assert r.scan.as_tuple() == (0, 0, -1, -1.0) and r.nearest_row == -1 and r.n_members == 1 and rows.size == 0


@pytest.mark.parametrize("dtype,hi,k,n,sim,bin_size,spread", [
    (np.uint8, 255, 4, 6000, 0.97, 1000, 40),     # every window covers all bins (the C2 situation)
    (np.uint8, 255, 3, 5000, 0.90, 100, 400),     # many bins, windows that end inside bins, bins that run empty
    (np.uint8, 255, 5, 3000, 0.80, 50, 900),      # very ragged lengths, 60 bins
    (np.uint8, 200, 6, 1200, 0.90, 64, 300),      # 8-row tiles (4 KB rows)
    (np.uint16, 3000, 4, 2500, 0.85, 70, 500),    # 16-bit bins
    (np.uint8, 255, 2, 3000, 0.90, 200, 60),      # 16-byte rows, many equal lengths
    (np.uint8, 255, 3, 37, 0.90, 8, 30),          # fewer rows than SMs
])
def test_accumulate_run_vs_step_loop(ctx, dtype, hi, k, n, sim, bin_size, spread):
    """mc_accumulate_run (the whole accumulate() loop of ClusterFactory.cpp:637-729 with the bvec on the
    device, one persistent kernel) against the same loop driven from the host: a literal restatement
    of bvec.cpp (tests/_pybvec.py) + one mc_accumulate_step per scan.  Clusters, centers, member order,
    scan and evaluation counts must be identical."""
    import _pybvec
    rng = np.random.default_rng(900 + k + n)
    nb = 4 ** k
    H0 = _rand_hists(rng, n, nb, dtype, hi, clusters=max(3, n // 150))
    lens0 = (1000 + rng.integers(0, spread, n)).astype(np.uint64)
    lens0[rng.integers(0, n, n // 10)] = 1000 + spread // 2      # runs of equal lengths
    bounds, order, first = _pybvec.layout(lens0, bin_size)
    H, lens = np.ascontiguousarray(H0[order]), np.ascontiguousarray(lens0[order])
    mins, maxs, w = _model(3)
    maxs[0] = float(spread)
    maxs[2] = 4.0 * nb * (hi / 255.0)
    ctx.load_histograms(H, lens, k)
    ctx.set_model(mins, maxs, w, 3)
    # bias: a scan marks roughly the rows of the center's own template
    sums = ctx.pair_classify(np.arange(n, dtype=np.int32), np.zeros(n, np.int32))[0]
    w = w.copy()
    w[0] -= np.sort(sums)[int(n * (1 - 1.5 / max(3, n // 150)))]
    ctx.set_model(mins, maxs, w, 3)
    ctx.near_threshold_count(reset=True)
    want = _pybvec.accumulate_by_steps(ctx, sim, bounds, first, lens)
    near_steps = ctx.near_threshold_count(reset=True)
    centers, offs, members, st = ctx.accumulate_run(sim, bounds, first)
    assert np.array_equal(np.sort(members), np.arange(n)), "every row belongs to exactly one cluster"
    assert np.array_equal(centers, want[0])
    assert np.array_equal(offs, want[1])
    assert np.array_equal(members, want[2])
    assert (st.n_scans, st.n_evals, st.n_steps) == want[3]
    assert st.n_near_threshold == near_steps == ctx.near_threshold_count()
    assert st.n_clusters == centers.size >= 2
    assert (np.diff(offs) > 1).sum() >= 2, "the test model never marks anything: no mean was exercised"
    # the same run with row compactions (the alive rows copied together whenever an eighth / half of them are gone;
    # inputs of this size do not compact by themselves): two row numberings inside the kernel, the same clusters outside
    import os
    for shift in ("3", "1"):
        os.environ["MC_PA_COMPACT_MIN"] = "64"
        os.environ["MC_PA_COMPACT_SHIFT"] = shift
        try:
            c2, o2, m2, st2 = ctx.accumulate_run(sim, bounds, first)
        finally:
            del os.environ["MC_PA_COMPACT_MIN"], os.environ["MC_PA_COMPACT_SHIFT"]
        assert np.array_equal(c2, centers) and np.array_equal(o2, offs) and np.array_equal(m2, members)
        assert (st2.n_scans, st2.n_evals, st2.n_steps, st2.n_near_threshold) == (st.n_scans, st.n_evals, st.n_steps, st.n_near_threshold)
        assert st2.n_compactions >= (1 if n >= 500 else 0) and st.n_compactions == 0


def test_accumulate_run_rejects_unsorted_bins(ctx):
    rng = np.random.default_rng(5)
    n, k = 500, 3
    H = _rand_hists(rng, n, 4 ** k)
    lens = np.sort((1000 + rng.integers(0, 50, n)).astype(np.uint64))
    lens[[10, 11]] = lens[[11, 10]] + np.array([5, 0], np.uint64)      # row 10 longer than row 11, same bin
    ctx.load_histograms(H, lens, k)
    ctx.set_model(*_model(3), 3)
    with pytest.raises(Exception) as e:
        ctx.accumulate_run(0.9, np.array([1000], np.uint64), np.array([0, n], np.int64))
    assert "non-decreasing" in str(e.value)


@pytest.mark.parametrize("world,k,n", [(2, 4, 5000), (3, 3, 4001), (4, 5, 3000)])
def test_sharded_scan_equals_single(ctx, world, k, n):
    """`world` ranks (here: contexts on one GPU, wired like ranks of one process) each scan
    their shard and exchange CTA partials through peer inboxes; every rank must end up with exactly
    the summary a single context computes over the whole range, scan after scan, with removal."""
    from meshclust_b200 import api, sharding
    rng = np.random.default_rng(500 + world)
    nb = 4 ** k
    H = _rand_hists(rng, n, nb, clusters=5)
    H[n // 2 + 3] = H[7]            # an exact tie on f0 across two shards: the smaller row must win
    lens = (1000 + rng.integers(0, 30, n)).astype(np.uint64)
    lens[n // 2 + 3] = lens[7]
    mins, maxs, w = _model(3)
    maxs[2] = 4.0 * nb
    ctx.load_histograms(H, lens, k)
    ctx.set_model(mins, maxs, w, 3)
    ranks = []
    try:
        for r in range(world):
            c = api.Context(0)
            c.load_histograms(H, lens, k)
            c.set_model(mins, maxs, w, 3)
            c.comm_init(r, world)
            ranks.append(c)
        api.Context.comm_connect_local(ranks)
        ctx.alive_reset()
        for c in ranks:
            c.alive_reset()
        centers = [7, n // 2 + 3, n - 1, 0, n // 3]
        for it, center in enumerate(centers):
            lo, hi = (0, n - 1) if it % 2 == 0 else (11, n - 13)
            want, _ = ctx.scan(center, lo, hi)
            for slot_base in (0,):
                for c in ranks:      # all ranks enqueue before anyone waits (one host thread drives them)
                    c.scan_sharded_enqueue(center, lo, hi, True, slot_base + it % 5)
                for c in ranks:
                    got = c.scan_sharded_collect(slot_base + it % 5, 1)[0]
                    assert got == want.as_tuple(), (it, got, want.as_tuple())
        # several scans in flight, no removal, collected together
        want = []
        for i, center in enumerate(centers):
            ctx.scan_enqueue(center, 0, n - 1, False, i)
        want = ctx.scan_collect(0, len(centers))
        for c in ranks:
            for i, center in enumerate(centers):
                c.scan_sharded_enqueue(center, 0, n - 1, False, 10 + i)
        for c in ranks:
            assert c.scan_sharded_collect(10, len(centers)) == want
        # the streaming form: bursts of scans on the scan stream, fold + send + combine on a second stream,
        # results one burst behind
        cr = np.array(centers, np.int64)
        lo_a = np.zeros(len(centers), np.int64)
        hi_a = np.full(len(centers), n - 1, np.int64)
        e = np.zeros(0, np.int64)
        for c in ranks:
            assert c.scan_sharded_burst(cr, lo_a, hi_a, False, 16, 0, 0) == []
        for c in ranks:
            assert c.scan_sharded_burst(cr[::-1].copy(), lo_a, hi_a, False, 32, 16, len(centers)) == want
        for c in ranks:
            assert c.scan_sharded_burst(e, e, e, False, 0, 32, len(centers)) == want[::-1]
        # a range that misses some shards entirely
        w1, _ = ctx.scan(3, 0, n // (2 * world))
        for c in ranks:
            c.scan_sharded_enqueue(3, 0, n // (2 * world), True, 63)
        for c in ranks:
            assert c.scan_sharded_collect(63, 1)[0] == w1.as_tuple()
        # protocol errors are reported, not hung on
        ranks[0].scan_sharded_enqueue(3, 0, 10, False, 5)
        with pytest.raises(api.McError):
            ranks[0].scan_sharded_enqueue(3, 0, 10, False, 5)
        for c in ranks[1:]:
            c.scan_sharded_enqueue(3, 0, 10, False, 5)
        for c in ranks:
            c.scan_sharded_collect(5, 1)
    finally:
        for c in ranks:
            c.close()


@pytest.mark.parametrize("world,k,n", [(2, 4, 4000), (3, 5, 2500)])
def test_accumulate_step_sharded_equals_single(ctx, world, k, n):
    """the sharded Phase-A step (ranks scan their shard, marks land in rank 0's array over peer
    memory, rank 0 folds the inboxes and runs the tail) against mc_accumulate_step on one context"""
    from meshclust_b200 import api
    rng = np.random.default_rng(700 + world)
    nb = 4 ** k
    H = _rand_hists(rng, n, nb, clusters=6)
    lens = (1000 + rng.integers(0, 40, n)).astype(np.uint64)
    mins, maxs, w = _model(3)
    maxs[2] = 4.0 * nb
    ctx.load_histograms(H, lens, k)
    ctx.set_model(mins, maxs, w, 3)
    s0 = ctx.pair_classify(np.arange(n, dtype=np.int32), np.zeros(n, np.int32))[0]
    w = w.copy()
    w[0] -= np.sort(s0)[int(0.9 * n)]
    ctx.set_model(mins, maxs, w, 3)
    ranks = []
    try:
        for r in range(world):
            c = api.Context(0)
            if r == 0:
                c.load_histograms(H, lens, k)
                c.set_model(mins, maxs, w, 3)
            else:
                c.clone_points_from(ranks[0])
            c.comm_init(r, world)
            ranks.append(c)
        api.Context.comm_connect_local(ranks)
        ctx.alive_reset()
        for c in ranks:
            c.alive_reset()
        seed, steps, marked_total = 1, 0, 0
        for cluster in range(5):
            kill = np.array([seed])
            ctx.alive_kill(kill)
            for c in ranks:
                c.alive_kill(kill)
            last, restart = seed, True
            while True:
                lo, hi = (0, n - 1) if steps % 2 == 0 else (9, n - 10)
                want, wrows = ctx.accumulate_step(last, lo, hi, restart)
                got, grows = api.Context.accumulate_step_sharded(ranks, last, lo, hi, restart)
                assert got.scan.as_tuple() == want.scan.as_tuple(), (cluster, steps)
                assert got.nearest_row == want.nearest_row and got.n_members == want.n_members
                assert np.array_equal(grows, wrows)
                restart = False
                steps += 1
                marked_total += wrows.size
                if want.scan.n_pos == 0 or steps > 40:
                    break
                last = want.nearest_row
            if want.scan.best_row < 0:
                break
            seed = int(want.scan.best_row)
        assert steps >= 6 and marked_total > 50
    finally:
        for c in ranks:
            c.close()


@pytest.mark.parametrize("dtype,k,n", [(np.uint8, 4, 3000), (np.uint16, 3, 1000), (np.uint8, 1, 500)])
def test_permute_rows(ctx, dtype, k, n):
    """mc_permute_rows re-numbers a prefix of the rows: histograms and constants follow, alive flags are
    set from n_alive, rows beyond the prefix stay, scans give the same answers in the new numbering"""
    from meshclust_b200 import api
    rng = np.random.default_rng(900 + k)
    nb = 4 ** k
    H = _rand_hists(rng, n, nb, dtype, 255 if dtype == np.uint8 else 2000, clusters=4)
    lens = (800 + rng.integers(0, 200, n)).astype(np.uint64)
    mins, maxs, w = _model(3)
    maxs[2] = 4.0 * nb
    ctx.load_histograms(H, lens, k)
    ctx.set_model(mins, maxs, w, 3)
    count, n_alive = n - 37, n // 2
    perm = rng.permutation(count)
    ctx.permute_rows(perm, n_alive)
    want = H.copy()
    want[:count] = H[perm]
    wl = lens.copy()
    wl[:count] = lens[perm]
    assert np.array_equal(ctx.copy_histograms(), want)
    ln, mg, sq = ctx.copy_point_stats()
    assert np.array_equal(ln, wl)
    assert np.array_equal(mg, want.astype(np.uint64).sum(1))
    assert np.array_equal(sq, (want.astype(np.uint64) ** 2).sum(1))
    # alive: rows < n_alive and the untouched tail (still alive from the load)
    res, marks = ctx.scan(0, 0, n - 1)
    assert res.n_eval == n_alive + (n - count)
    # same scan on a fresh context holding the permuted data, with the same rows alive
    with api.Context(0) as c2:
        c2.load_histograms(want, wl, k)
        c2.set_model(mins, maxs, w, 3)
        c2.alive_kill(np.arange(n_alive, count))
        res2, marks2 = c2.scan(0, 0, n - 1)
    assert res.as_tuple() == res2.as_tuple() and np.array_equal(marks, marks2)
    with pytest.raises(api.McError):
        ctx.permute_rows(np.zeros(5, np.int64), 2)      # not a permutation


@pytest.mark.parametrize("k", [3, 4, 5])
def test_update_centers_vs_oracle(ctx, oracle, k):
    rng = np.random.default_rng(80 + k)
    nb, n = 4 ** k, 1500
    H = _rand_hists(rng, n, nb, clusters=12)
    lens = (H.astype(np.int64).sum(1) - nb + k - 1).astype(np.uint64)
    mins, maxs, w = _model(4)
    maxs[2], maxs[4], mins[4] = 4.0 * nb, 2.0 * nb, 0.5 * nb
    ctx.load_histograms(H, lens, k)
    ctx.set_model(mins, maxs, w, 4)
    # 30 clusters of random members; center c sees clusters c-2..c+2 (delta = 2)
    perm = rng.permutation(n)
    cuts = np.sort(rng.choice(np.arange(1, n), 29, replace=False))
    bounds = np.concatenate([[0], cuts, [n]])
    centers = np.array([perm[bounds[c]] for c in range(30)])
    delta = 2
    cb = np.array([bounds[max(0, c - delta)] for c in range(30)])
    ce = np.array([bounds[min(29, c + delta) + 1] for c in range(30)])
    got = ctx.update_centers(centers, perm, cb, ce)
    nsurv = 0
    for c in range(30):
        cand = perm[cb[c]:ce[c]]
        s, f0, fl = oracle.scan(H[cand], lens[cand], H[centers[c]], int(lens[centers[c]]), mins, maxs, w, 4)
        assert not (np.abs(s) < NEAR).any()
        good = cand[fl == 1]
        nsurv += good.size
        if good.size == 0:
            assert got[c] == -1
            continue
        mean = oracle.mean(H[good])
        dd = np.array([oracle.distance_d(H[r], mean) for r in good])
        assert got[c] == int(good[int(np.argmin(dd))])
    assert nsurv > 0


# ----------------------------------------------------------------------------------------------
# stage 4
# ----------------------------------------------------------------------------------------------
def test_alignment_golden(ctx, golden):
    digits, offs = golden["al_digits"], golden["al_offs"]
    n = offs.size - 1
    # digits are already encoded: load as sequences without segments (left untouched, upper-cased
    # letters stay letters; bytes 0..3 and 'N' pass through)
    ctx.load_sequences(digits, offs, np.zeros(0, np.int32), np.zeros(n + 1, np.int64))
    assert np.array_equal(ctx.copy_digits(), digits)
    sc, ln, mt = ctx.align_pairs(golden["al_pa"], golden["al_pb"])
    assert np.array_equal(sc, golden["al_score"])
    assert np.array_equal(ln, golden["al_len"])
    assert np.array_equal(mt, golden["al_matches"])


def test_alignment_long_pairs_golden(ctx):
    """BASELINE configs[4] lengths: 30 pairs of 9.5 - 10.5 kb (79+ strips of 128 rows, scratch lines
    double-buffered across strips), unequal lengths, 'N' bytes, unrelated pairs, both argument orders,
    and one 32.7 kb x 32.8 kb pair just under the limit of the packed (length, matches) word -- against
    triples computed by the compiled reference (tests/golden/make_golden_long.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_align_long.npz"))
    digits, offs = g["digits"], g["offs"]
    n = offs.size - 1
    assert (np.diff(offs) >= 8900).all() and g["pa"].size >= 20
    ctx.load_sequences(digits, offs, np.zeros(0, np.int32), np.zeros(n + 1, np.int64))
    # a batch this small and long runs as teams of six warps per pair by default (strips of one pair pipelined over the
    # warps of a CTA, flow control through shared memory); teams of four and one warp per pair must give the same triples
    for teams in (None, "0", "4", "6"):
        if teams is None:
            os.environ.pop("MC_NW_TEAMS", None)
        else:
            os.environ["MC_NW_TEAMS"] = teams
        try:
            sc, ln, mt = ctx.align_pairs(g["pa"], g["pb"])
        finally:
            os.environ.pop("MC_NW_TEAMS", None)
        assert np.array_equal(sc, g["score"]), teams
        assert np.array_equal(ln, g["alen"]), teams
        assert np.array_equal(mt, g["matches"]), teams
    # the same pairs one at a time and in reverse batch order: batching must not matter
    order = np.arange(g["pa"].size)[::-1]
    sc2, ln2, mt2 = ctx.align_pairs(g["pa"][order], g["pb"][order])
    assert np.array_equal(sc2, g["score"][order]) and np.array_equal(ln2, g["alen"][order]) and np.array_equal(mt2, g["matches"][order])


@pytest.mark.parametrize("cfg,n,npairs", [("c3", 400, 3000), ("c1", 300, 1200), ("c2", 120, 300)])
def test_alignment_vs_oracle_configs(ctx, oracle, cfg, n, npairs):
    import _oracle
    letters, offs, tmpl = synth.generate_config(cfg, n)
    ctx.load_sequences(letters, offs)
    digits = _oracle.encode_digits(letters, offs)
    assert np.array_equal(ctx.copy_digits(), digits)
    rng = np.random.default_rng(12)
    pa = rng.integers(0, n, npairs).astype(np.int32)
    pb = rng.integers(0, n, npairs).astype(np.int32)
    same = rng.random(npairs) < 0.5        # half the pairs from the same template (high identity)
    pb[same] = (pa[same] + len(set(tmpl.tolist())) * rng.integers(0, 3, int(same.sum()))) % n
    want = oracle.globalign_batch(digits, offs, pa, pb)
    got = ctx.align_pairs(pa, pb)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_alignment_lengths_sweep(ctx, oracle):
    # lengths around the 32-row strip boundaries, empty strings, N bytes
    rng = np.random.default_rng(13)
    lens = [0, 1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 130, 257]
    seqs = []
    for L in lens:
        s = rng.integers(0, 4, L, dtype=np.uint8)
        if L > 40:
            s[L // 2] = ord("N")
        seqs.append(s)
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([s.size for s in seqs], out=offs[1:])
    digits = np.concatenate(seqs)
    n = len(seqs)
    ctx.load_sequences(digits, offs, np.zeros(0, np.int32), np.zeros(n + 1, np.int64))
    pa, pb = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    pa, pb = pa.reshape(-1).astype(np.int32), pb.reshape(-1).astype(np.int32)
    want = oracle.globalign_batch(digits, offs, pa, pb)
    got = ctx.align_pairs(pa, pb)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


# ----------------------------------------------------------------------------------------------
# size-independent properties at BASELINE sizes (the oracle would take too long there)
# ----------------------------------------------------------------------------------------------
def test_full_size_properties_c2(ctx, oracle):
    letters, offs, _ = synth.generate_config("c2")
    n = offs.size - 1
    hist, mx = ctx.kmer_histograms_host(letters, offs, 4, 1)
    assert mx <= 255
    ln, mg, sq = ctx.copy_point_stats()
    # every k-mer start is counted exactly once: mag = 4^k + (L - k + 1)
    assert np.array_equal(mg, (256 + np.diff(offs) - 3).astype(np.uint64))
    assert np.array_equal(hist.astype(np.uint64).sum(1), mg)
    # spot-check 300 rows bit-exactly against the oracle
    rows = np.random.default_rng(0).integers(0, n, 300)
    sub_offs = np.zeros(rows.size + 1, np.int64)
    np.cumsum(np.diff(offs)[rows], out=sub_offs[1:])
    sub = np.concatenate([letters[offs[r]:offs[r + 1]] for r in rows])
    rc, want, _ = oracle.hist_batch(sub, sub_offs, 4, 1)
    assert np.array_equal(hist[rows], want)
    # scan: a row is always similar to itself with f0 = its maximum possible value; symmetric keys
    mins = np.array([0.0, 0.5, 0.0, 0.0, 100.0])
    maxs = np.array([80.0, 1.0, 1200.0, 1.0, 600.0])
    ctx.set_model(mins, maxs, np.array([-1.0, 2.0, 1.0, 0.5]), 3)
    for c in (0, n // 2, n - 1):
        ctx.alive_reset()
        res, marks = ctx.scan(c, 0, n - 1)
        assert res.n_eval == n and marks[c] == 1 and res.n_pos == int(marks.sum())
        s, f0, fl = oracle.scan(hist[c:c + 1], ln[c:c + 1], hist[c], int(ln[c]), mins, maxs, np.array([-1.0, 2.0, 1.0, 0.5]), 3)
        assert res.best_f0 >= f0[0]
    keys = ctx.distance_keys(np.array([3, 77]))
    assert keys[0, 3] == 0 and keys[1, 77] == 0 and keys[0, 77] == keys[1, 3]


def test_alignment_teams_ragged(ctx, oracle):
    # teams of warps per pair (strips of 512 rows pipelined over the warps of a CTA) on lengths around the strip and
    # chunk boundaries: one strip (no pipeline), exactly full strips, fewer strips than warps, many strips, columns that
    # end inside / at a 32-column chunk, an empty partner; more pairs than resident teams (teams move on to the next
    # pair without a barrier)
    import os
    rng = np.random.default_rng(29)
    lens = [0, 1, 31, 33, 511, 512, 513, 1024, 1500, 2047, 2049, 3100, 6200]
    seqs = []
    for L in lens:
        s = rng.integers(0, 4, L, dtype=np.uint8)
        if L > 600:
            s[L // 3] = ord("N")
        seqs.append(s)
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([s.size for s in seqs], out=offs[1:])
    digits = np.concatenate(seqs)
    n = len(seqs)
    ctx.load_sequences(digits, offs, np.zeros(0, np.int32), np.zeros(n + 1, np.int64))
    pa, pb = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    pa, pb = np.tile(pa.reshape(-1), 4).astype(np.int32), np.tile(pb.reshape(-1), 4).astype(np.int32)   # 676 pairs > 444 teams
    want = oracle.globalign_batch(digits, offs, pa, pb)
    for teams in ("6", "4", "0"):
        os.environ["MC_NW_TEAMS"] = teams
        try:
            got = ctx.align_pairs(pa, pb)
        finally:
            os.environ.pop("MC_NW_TEAMS", None)
        for g, w in zip(got, want):
            assert np.array_equal(g, w), teams
