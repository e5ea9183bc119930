"""CPU, build container only: the C oracle against the compiled unmodified reference
(oracle/_ref/libmcref.so) on fresh seeded inputs.  Skipped where /root/reference was never built."""
import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def test_alignment_random(oracle):
    r = O.ref()
    rng = np.random.default_rng(5)
    for _ in range(400):
        la, lb = int(rng.integers(0, 70)), int(rng.integers(0, 70))
        a = rng.integers(0, 4, la, dtype=np.uint8)
        b = rng.integers(0, 4, lb, dtype=np.uint8)
        if la and lb and rng.random() < 0.6:
            b[: min(la, lb)] = a[: min(la, lb)]
            mut = rng.random(lb) < 0.15
            b[mut] = rng.integers(0, 4, int(mut.sum()), dtype=np.uint8)
        assert oracle.globalign(a.tobytes(), b.tobytes()) == r.globalign(a.tobytes(), b.tobytes())


def test_encode_random(oracle):
    r = O.ref()
    rng = np.random.default_rng(6)
    alpha = np.frombuffer(b"ACGTacgtNnRYMKSWHBVDX-", dtype=np.uint8)
    for _ in range(500):
        L = int(rng.integers(1, 160))
        s = rng.choice(alpha[:8], L)
        for _ in range(int(rng.integers(0, 4))):
            st, ln = int(rng.integers(0, L)), int(rng.integers(1, 14))
            s[st:st + ln] = ord("N")
        if rng.random() < 0.3:
            s[rng.integers(0, L, 2)] = rng.choice(alpha[10:], 2)
        d1, s1 = oracle.encode(s.tobytes())
        d2, s2 = r.encode(s.tobytes())
        assert (d1 is None) == (d2 is None)
        if d1 is not None:
            assert np.array_equal(d1, d2) and np.array_equal(s1, s2)


@pytest.mark.parametrize("k", [2, 4, 6])
def test_hist_random(oracle, k):
    from meshclust_b200 import synth
    r = O.ref()
    letters, offs, _ = synth.generate(150, 9, 400, 0.05, 70 + k)
    rc1, h1, _ = oracle.hist_batch(letters, offs, k, 1)
    rc2, h2, _ = r.hist_batch(letters, offs, k, 1)
    assert rc1 == 0 and rc2 == 0 and np.array_equal(h1, h2)


def test_scan_random(oracle):
    r = O.ref()
    rng = np.random.default_rng(8)
    nb, n = 64, 500
    base = np.minimum(rng.poisson(10, nb) + 1, 255)
    H = np.minimum(np.maximum(base[None, :] + rng.integers(-4, 5, (n, nb)), 1), 255).astype(np.uint8)
    lens = (H.sum(1) - nb + 2).astype(np.uint64)
    mins = np.array([0.0, 0.6, 5.0, -0.3, 60.0])
    maxs = np.array([40.0, 0.99, 300.0, 0.95, 140.0])
    w = np.array([-2.5, 2.0, 1.0, 0.8, 0.5])
    s1, f1, g1 = oracle.scan(H, lens, H[3], int(lens[3]), mins, maxs, w, 4)
    s2, f2, g2 = r.scan(H, lens, H[3], int(lens[3]), mins, maxs, w, 4)
    assert np.allclose(s1, s2, rtol=1e-12, atol=0) and np.allclose(f1, f2, rtol=1e-12, atol=0)
    assert np.array_equal(g1, g2)
