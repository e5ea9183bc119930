"""ctypes bindings for the CPU oracle (oracle/libmcoracle.so) and, when it has been built, the
compiled unmodified reference (oracle/_ref/libmcref.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_ORACLE_SO = os.path.join(ORACLE_DIR, "libmcoracle.so")
_REF_SO = os.path.join(ORACLE_DIR, "_ref", "libmcref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "meshclust")

c_p = C.c_void_p
i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")


def build_oracle():
    src = os.path.join(ORACLE_DIR, "mc_oracle.c")
    if (not os.path.exists(_ORACLE_SO)) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(c_p)


class Oracle:
    """Wrapper with the same method names for both the C restatement (prefix mco_) and the
    reference harness (prefix ref_)."""

    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L = self.lib
        if prefix == "mco_":
            L.mco_encode.restype = C.c_int
            L.mco_hist_batch.restype = C.c_long
            L.mco_distance_d.restype = C.c_double
        else:
            L.ref_encode.restype = C.c_int
            L.ref_hist_batch.restype = C.c_int
            L.ref_distance_d.restype = C.c_double
            L.ref_max_threads.restype = C.c_int

    def f(self, name):
        return getattr(self.lib, self.prefix + name)

    # -- sequences ------------------------------------------------------------------------
    def encode(self, seq: bytes, max_segs: int = 4096):
        n = len(seq)
        digits = np.zeros(max(n, 1), dtype=np.uint8)
        segs = np.zeros(2 * max_segs, dtype=np.int32)
        ns = self.f("encode")(C.c_char_p(seq), C.c_long(n), _ptr(digits), _ptr(segs), C.c_int(max_segs))
        if ns < 0:
            return None, None
        return digits[:n].copy(), segs[: 2 * ns].reshape(-1, 2).copy()

    def hist_batch(self, letters: np.ndarray, offs: np.ndarray, k: int, tbytes: int = 1):
        n = offs.size - 1
        nb = 4 ** k
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[tbytes]
        out = np.zeros((n, nb), dtype=dt)
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        if self.prefix == "mco_":
            mx = C.c_uint64(0)
            rc = self.lib.mco_hist_batch(_ptr(letters), _ptr(offs), C.c_int(n), C.c_int(k),
                                         C.c_int(tbytes), _ptr(out), C.byref(mx))
            return rc, out, int(mx.value)
        rc = self.lib.ref_hist_batch(_ptr(letters), _ptr(offs), C.c_int(n), C.c_int(k), C.c_int(tbytes), _ptr(out))
        return rc, out, None

    # -- pair arithmetic ------------------------------------------------------------------
    def features(self, p: np.ndarray, q: np.ndarray, lp: int, lq: int):
        assert p.dtype == q.dtype and p.size == q.size
        out5 = np.zeros(5, dtype=np.float64)
        dist = C.c_uint64(0)
        self.f("features")(_ptr(p), _ptr(q), C.c_int(p.size), C.c_int(p.dtype.itemsize),
                           C.c_uint64(lp), C.c_uint64(lq), _ptr(out5), C.byref(dist))
        return out5, int(dist.value)

    def distance_d(self, p: np.ndarray, mean: np.ndarray) -> float:
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        return float(self.f("distance_d")(_ptr(p), C.c_int(p.size), C.c_int(p.dtype.itemsize), _ptr(mean)))

    def mean(self, hists: np.ndarray) -> np.ndarray:
        m, nb = hists.shape
        out = np.zeros(nb, dtype=np.float64)
        self.f("mean")(_ptr(np.ascontiguousarray(hists)), C.c_int(nb), C.c_int(hists.dtype.itemsize), C.c_int(m), _ptr(out))
        return out

    def scan(self, hists, lens, center, center_len, mins, maxs, weights, nfeat):
        n, nb = hists.shape
        hists = np.ascontiguousarray(hists)
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        center = np.ascontiguousarray(center)
        mins = np.ascontiguousarray(mins, dtype=np.float64)
        maxs = np.ascontiguousarray(maxs, dtype=np.float64)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        s = np.zeros(n, np.float64)
        f0 = np.zeros(n, np.float64)
        fl = np.zeros(n, np.uint8)
        if self.prefix == "mco_":
            self.lib.mco_scan(_ptr(hists), _ptr(lens), C.c_int(n), C.c_int(nb), C.c_int(hists.dtype.itemsize),
                              _ptr(center), C.c_uint64(center_len), _ptr(mins), _ptr(maxs), _ptr(weights),
                              C.c_int(nfeat), _ptr(s), _ptr(f0), _ptr(fl))
        else:
            assert hists.dtype == np.uint8
            self.lib.ref_scan_u8(_ptr(hists), _ptr(lens), C.c_int(n), C.c_int(nb), _ptr(center),
                                 C.c_uint64(center_len), _ptr(mins), _ptr(maxs), _ptr(weights),
                                 C.c_int(nfeat), _ptr(s), _ptr(f0), _ptr(fl))
        return s, f0, fl

    # -- reference-only: points held as objects, the way the reference holds them ---------
    def pointset(self, hists: np.ndarray, lens: np.ndarray):
        assert self.prefix == "ref_" and hists.dtype == np.uint8
        hists = np.ascontiguousarray(hists)
        lens = np.ascontiguousarray(lens, np.uint64)
        self.lib.ref_pointset_create.restype = c_p
        h = self.lib.ref_pointset_create(_ptr(hists), _ptr(lens), C.c_int(hists.shape[0]), C.c_int(hists.shape[1]))
        return c_p(h)

    def pointset_set_model(self, ps, mins, maxs, nfeat):
        mins = np.ascontiguousarray(mins, np.float64)
        maxs = np.ascontiguousarray(maxs, np.float64)
        self.lib.ref_pointset_set_model(ps, _ptr(mins), _ptr(maxs), C.c_int(nfeat))

    def pointset_scan(self, ps, center: int, weights, flags: np.ndarray | None = None):
        weights = np.ascontiguousarray(weights, np.float64)
        bi, bf = C.c_long(-1), C.c_double(-1)
        self.lib.ref_pointset_scan.restype = C.c_long
        npos = self.lib.ref_pointset_scan(ps, C.c_int(center), _ptr(weights), None if flags is None else _ptr(flags), C.byref(bi), C.byref(bf))
        return int(npos), int(bi.value), float(bf.value)

    def pointset_destroy(self, ps):
        self.lib.ref_pointset_destroy(ps)

    # -- alignment ------------------------------------------------------------------------
    def globalign(self, s1: bytes, s2: bytes, params=(1, -1, 2, 1)):
        sc, ln, mt = C.c_int(0), C.c_int(0), C.c_int(0)
        if self.prefix == "mco_":
            self.lib.mco_globalign(C.c_char_p(s1), C.c_int(len(s1)), C.c_char_p(s2), C.c_int(len(s2)),
                                   *[C.c_int(x) for x in params], C.byref(sc), C.byref(ln), C.byref(mt))
        else:
            idn = C.c_double(0)
            self.lib.ref_globalign(C.c_char_p(s1), C.c_int(len(s1)), C.c_char_p(s2), C.c_int(len(s2)),
                                   *[C.c_int(x) for x in params], C.byref(sc), C.byref(ln), C.byref(mt), C.byref(idn))
        return sc.value, ln.value, mt.value

    def globalign_batch(self, digits: np.ndarray, offs: np.ndarray, pa: np.ndarray, pb: np.ndarray):
        digits = np.ascontiguousarray(digits, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        pa = np.ascontiguousarray(pa, dtype=np.int32)
        pb = np.ascontiguousarray(pb, dtype=np.int32)
        m = pa.size
        sc = np.zeros(m, np.int32)
        ln = np.zeros(m, np.int32)
        mt = np.zeros(m, np.int32)
        self.f("globalign_batch")(_ptr(digits), _ptr(offs), _ptr(pa), _ptr(pb), C.c_int(m), _ptr(sc), _ptr(ln), _ptr(mt))
        return sc, ln, mt


_oracle = None
_ref = None


def oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        build_oracle()
        _oracle = Oracle(_ORACLE_SO, "mco_")
    return _oracle


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref() -> Oracle:
    global _ref
    if _ref is None:
        _ref = Oracle(_REF_SO, "ref_")
    return _ref


def encode_digits(letters: np.ndarray, offs: np.ndarray):
    """Digit strings of a batch via the oracle (segment-less sequences stay upper-case letters)."""
    o = oracle()
    out = np.zeros_like(letters)
    for i in range(offs.size - 1):
        d, _ = o.encode(letters[offs[i]:offs[i + 1]].tobytes())
        out[offs[i]:offs[i + 1]] = d
    return out
