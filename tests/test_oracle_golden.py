"""CPU: the C oracle (oracle/mc_oracle.c) against the committed golden vectors, which were
produced by the compiled, unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest


def test_alignment_golden(oracle, golden):
    sc, ln, mt = oracle.globalign_batch(golden["al_digits"], golden["al_offs"], golden["al_pa"], golden["al_pb"])
    assert np.array_equal(sc, golden["al_score"])
    assert np.array_equal(ln, golden["al_len"])
    assert np.array_equal(mt, golden["al_matches"])


def test_alignment_known_answers(oracle):
    # SURVEY.md Appendix B (generated from GlobAlignE itself): score, length, identity
    m = {"A": 0, "C": 1, "G": 2, "T": 3}
    dig = lambda s: bytes(m[c] for c in s.upper())
    kat = [("GATCTCAG", "GACAG", 0, 8, 0.625), ("GACAG", "GATCAG", 2, 6, 0.8333333333333334),
           ("GGAACCTT", "GGCCAATT", 0, 8, 0.5), ("GATCCATTACCG", "GATATTACCTT", 1, 13, 0.6923076923076923),
           ("ACGT", "", -6, 4, 0.0), ("", "ACGT", -7, 4, 0.0), ("A", "A", 1, 1, 1.0), ("A", "C", -1, 1, 0.0),
           ("AAAA", "AAAAAAAAAAAA", -6, 12, 0.3333333333333333)]
    for a, b, score, length, ident in kat:
        sc, ln, mt = oracle.globalign(dig(a), dig(b))
        assert (sc, ln) == (score, length)
        assert mt / ln == ident
    assert oracle.globalign(b"", b"") == (0, 0, 0)   # identity 0/0 = NaN in the reference


def test_encode_golden(oracle, golden):
    offs = golden["enc_offs"]
    letters = golden["enc_letters"]
    seg_off = golden["enc_seg_off"]
    for i in range(offs.size - 1):
        d, segs = oracle.encode(letters[offs[i]:offs[i + 1]].tobytes())
        assert np.array_equal(d, golden["enc_digits"][offs[i]:offs[i + 1]])
        assert np.array_equal(segs.reshape(-1), golden["enc_segs"][2 * seg_off[i]:2 * seg_off[i + 1]])
    bad = [oracle.encode(s)[0] is None for s in (b"NNNNNN", b"", b"ACGTACGTACGTACGTACGTAC-GT", b"NNNNNA")]
    assert np.array_equal(np.array(bad, np.int32), golden["enc_bad"])


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_histograms_golden(oracle, golden, k):
    rc, h, mx = oracle.hist_batch(golden["enc_letters"], golden["enc_offs"], k, 2)
    assert rc == 0
    assert np.array_equal(h, golden[f"hist_k{k}"])
    assert mx == int(golden[f"hist_k{k}"].max())


def test_histogram_known_answers(oracle):
    # SURVEY.md Appendix B, k=2
    recs = {b"ACGTACGTACGTACGTACGTACGTAC": [1, 8, 1, 1, 1, 1, 7, 1, 1, 1, 1, 7, 7, 1, 1, 1],
            b"ACGTACGTACGTTCGTACGAACGTACGG": [2, 7, 1, 1, 1, 1, 8, 1, 2, 1, 2, 6, 5, 2, 1, 2],
            b"acgtacgtacgtacgtacgtNNNNNacgtacgtacgtacgtacgtacgt": [1, 12, 1, 1, 2, 5, 12, 1, 1, 1, 1, 12, 10, 2, 1, 1],
            b"ACGTACGTAC": [1] * 16}
    for s, want in recs.items():
        l = np.frombuffer(s, np.uint8)
        rc, h, _ = oracle.hist_batch(l, np.array([0, len(s)], np.int64), 2, 1)
        assert rc == 0 and h[0].tolist() == want


@pytest.mark.parametrize("name", ["u8", "u16"])
@pytest.mark.parametrize("nb", [16, 64, 256, 1024])
def test_pair_features_golden(oracle, golden, name, nb):
    H, lens = golden[f"pf_{name}_{nb}_H"], golden[f"pf_{name}_{nb}_lens"]
    raw, dist = golden[f"pf_{name}_{nb}_raw"], golden[f"pf_{name}_{nb}_dist"]
    n = H.shape[0]
    for i in range(n):
        for j in range(n):
            r, d = oracle.features(H[i], H[j], int(lens[i]), int(lens[j]))
            assert np.array_equal(r.view(np.uint64), raw[i, j].view(np.uint64)), (i, j)   # bit-exact
            assert d == dist[i, j]
    mean = oracle.mean(H[: n // 2])
    assert np.array_equal(mean.view(np.uint64), golden[f"pf_{name}_{nb}_mean"].view(np.uint64))
    dd = np.array([oracle.distance_d(H[i], mean) for i in range(n)])
    assert np.array_equal(dd.view(np.uint64), golden[f"pf_{name}_{nb}_dd"].view(np.uint64))


def test_pair_known_answers(oracle):
    # distance / intersection / manhattan / pearson / kulczynski2 / LD
    H = np.array([[1, 8, 1, 1, 1, 1, 7, 1, 1, 1, 1, 7, 7, 1, 1, 1], [2, 7, 1, 1, 1, 1, 8, 1, 2, 1, 2, 6, 5, 2, 1, 2],
                  [1, 12, 1, 1, 2, 5, 12, 1, 1, 1, 1, 12, 10, 2, 1, 1], [1] * 16], np.uint8)
    lens = [26, 28, 49, 10]
    want = {(0, 1): (2239, 0.8809523809523809, 10, 0.9551548037645808, 225.65173000567216, 2),
            (0, 2): (3901, 0.780952380952381, 23, 0.9584105906300393, 210.0, 23),
            (0, 3): (6848, 0.5614035087719298, 25, 0.0, 177.9512195121951, 16),
            (1, 2): (4685, 0.7289719626168224, 29, 0.925864717608711, 194.09302325581396, 21),
            (2, 3): (8400, 0.4, 48, 0.0, 160.0, 39)}
    for (i, j), (d, inter, man, pear, kul, ld) in want.items():
        r, dist = oracle.features(H[i], H[j], lens[i], lens[j])
        assert dist == d and r[1] == inter and r[2] == man and r[3] == pear and r[4] == kul and r[0] == ld
    mean = oracle.mean(H)
    dd = [oracle.distance_d(H[i], mean) for i in range(4)]
    assert dd[:3] == [1225.765101746515, 1653.7113244932184, 4448.2891195693965]
    assert abs(dd[3] - 6488.3401920438964) < 1e-9


@pytest.mark.parametrize("nfeat", [3, 4])
def test_scan_golden(oracle, golden, nfeat):
    H, lens = golden["sc_H"], golden["sc_lens"]
    s, f0, fl = oracle.scan(H, lens, H[7], int(lens[7]), golden["sc_mins"], golden["sc_maxs"], golden[f"sc_w{nfeat}"], nfeat)
    # stated tolerance for FP features / GLM sum: 1e-12 relative; on this path they are bit-equal
    assert np.allclose(s, golden[f"sc_sum{nfeat}"], rtol=1e-12, atol=0)
    assert np.allclose(f0, golden[f"sc_f0{nfeat}"], rtol=1e-12, atol=0)
    near = np.abs(golden[f"sc_sum{nfeat}"]) < 1e-9
    assert np.array_equal(fl[~near], golden[f"sc_flag{nfeat}"][~near])
    assert 0 < fl.sum() < fl.size


def test_alignment_long_pairs_golden(oracle):
    """The C restatement of GlobAlignE against reference triples at BASELINE configs[4] lengths
    (tests/golden/ref_align_long.npz; three 10 kb pairs here, the GPU test checks all of them)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_align_long.npz"))
    sel = np.array([0, 2, 20])      # unequal lengths, an 'N' byte, an unrelated pair
    sc, ln, mt = oracle.globalign_batch(g["digits"], g["offs"], g["pa"][sel], g["pb"][sel])
    assert np.array_equal(sc, g["score"][sel]) and np.array_equal(ln, g["alen"][sel]) and np.array_equal(mt, g["matches"][sel])
