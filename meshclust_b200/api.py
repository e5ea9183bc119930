"""ctypes mirror of include/meshclust_b200.h (same names, same argument meaning, same errors).

Nothing here computes: every method forwards host numpy buffers to the C-ABI.  The library must
have been built in-tree (``python -m meshclust_b200.build``); a missing library is an ImportError,
never a silent fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# MESHCLUST_B200_LIB selects another BUILD of the same CUDA library (e.g. an instrumented one)
LIB_PATH = os.environ.get("MESHCLUST_B200_LIB") or os.path.join(_PKG, "libmeshclust_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m meshclust_b200.build` "
        "(meshclust_b200 has no CPU fallback)")

_lib = C.CDLL(LIB_PATH)

MC_OK = 0
MC_ERR_CUDA, MC_ERR_ARG, MC_ERR_STATE, MC_ERR_INPUT, MC_ERR_UNSUPPORTED = -1, -2, -3, -4, -5

# every symbol include/meshclust_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "mc_version", "mc_last_error", "mc_device_count", "mc_ctx_create", "mc_ctx_destroy", "mc_stream",
    "mc_sync", "mc_launch_count", "mc_host_segments", "mc_load_sequences", "mc_ingest_fasta", "mc_stage_fasta_bytes", "mc_reserve_scratch", "mc_load_segments", "mc_copy_letters", "mc_copy_digits",
    "mc_build_histograms", "mc_load_histograms", "mc_copy_histograms", "mc_copy_point_stats",
    "mc_set_model", "mc_distance_keys", "mc_pair_features", "mc_pair_classify", "mc_alive_reset",
    "mc_alive_kill", "mc_scan", "mc_scan_enqueue", "mc_scan_collect", "mc_scan_enqueue_many", "mc_scan_fold_dev", "mc_set_stream", "mc_mean_nearest", "mc_accumulate_step", "mc_accumulate_run", "mc_near_threshold_count", "mc_permute_rows", "mc_reserve_permute", "mc_comm_init", "mc_comm_connect", "mc_comm_connect_local",
    "mc_scan_sharded_enqueue", "mc_scan_sharded_enqueue_many", "mc_scan_sharded_collect", "mc_scan_sharded_combine", "mc_scan_sharded_wait", "mc_scan_sharded_burst", "mc_clone_points", "mc_clone_sequences", "mc_accumulate_step_sharded", "mc_update_centers", "mc_align_pairs",
    "mc_kmer_histograms_host", "mc_scan_host",
]


MC_SCAN_KEEP, MC_SCAN_REMOVE, MC_SCAN_CHAIN = 0, 1, 2


class ScanResult(C.Structure):
    _fields_ = [("n_eval", C.c_int64), ("n_pos", C.c_int64), ("best_row", C.c_int64), ("best_f0", C.c_double)]

    def as_tuple(self):
        return (self.n_eval, self.n_pos, self.best_row, self.best_f0)


class StepResult(C.Structure):
    _fields_ = [("scan", ScanResult), ("nearest_row", C.c_int64), ("n_members", C.c_int64)]


class RunStats(C.Structure):
    _fields_ = [("n_clusters", C.c_int64), ("n_scans", C.c_int64), ("n_evals", C.c_int64), ("n_near_threshold", C.c_int64),
                ("n_steps", C.c_int64), ("device_seconds", C.c_double), ("n_compactions", C.c_int64)]


_lib.mc_version.restype = C.c_char_p
_lib.mc_last_error.restype = C.c_char_p
_lib.mc_stream.restype = C.c_void_p
_lib.mc_stream.argtypes = [C.c_void_p]
_lib.mc_launch_count.restype = C.c_int64
_lib.mc_launch_count.argtypes = [C.c_void_p]
_lib.mc_ctx_destroy.restype = None
_lib.mc_ctx_destroy.argtypes = [C.c_void_p]


class McError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"meshclust_b200 error {code}: {msg}")
        self.code = code


def _check(rc: int):
    if rc != MC_OK:
        raise McError(rc, _lib.mc_last_error().decode(errors="replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def version() -> str:
    return _lib.mc_version().decode()


def device_count() -> int:
    return int(_lib.mc_device_count())


def host_segments(letters: np.ndarray | bytes, max_segs: int = 1 << 16):
    """Chromosome::removeN/mergeSegments/makeSegmentList on the host; None when the reference throws."""
    a = np.frombuffer(letters, dtype=np.uint8) if isinstance(letters, (bytes, bytearray)) else np.ascontiguousarray(letters, np.uint8)
    segs = np.zeros(2 * max_segs, np.int32)
    ns = _lib.mc_host_segments(_p(a), C.c_int64(a.size), _p(segs), C.c_int(max_segs))
    if ns < 0:
        return None
    return segs[: 2 * ns].reshape(-1, 2).copy()


def segments_for_batch(letters: np.ndarray, offsets: np.ndarray):
    segs, seg_off = [], np.zeros(offsets.size, np.int64)
    for i in range(offsets.size - 1):
        s = host_segments(letters[offsets[i]:offsets[i + 1]])
        if s is None:
            raise McError(MC_ERR_INPUT, f"sequence {i} has no non-N run")
        segs.append(s)
        seg_off[i + 1] = seg_off[i] + len(s)
    segs = np.concatenate(segs).astype(np.int32) if segs else np.zeros((0, 2), np.int32)
    return np.ascontiguousarray(segs.reshape(-1)), seg_off


class Context:
    """One GPU.  Mirrors ``mc_ctx`` (replaces the reference's OpenMP runtime set-up, Runner.cpp:201-214)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(_lib.mc_ctx_create(C.byref(self._h), C.c_int(device)))
        self.n = 0
        self.k = 0
        self.tbytes = 0
        self.total_bases = 0

    def close(self):
        if self._h:
            _lib.mc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing -------------------------------------------------------------------------
    @property
    def stream(self) -> int:
        return int(_lib.mc_stream(self._h) or 0)

    def sync(self):
        _check(_lib.mc_sync(self._h))

    @property
    def launches(self) -> int:
        return int(_lib.mc_launch_count(self._h))

    # -- stage 0/1 ------------------------------------------------------------------------
    def load_sequences(self, letters: np.ndarray, offsets: np.ndarray, segs=None, seg_offsets=None):
        letters = np.ascontiguousarray(letters, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        if segs is None:
            segs, seg_offsets = segments_for_batch(letters, offsets)
        segs = np.ascontiguousarray(segs, np.int32)
        seg_offsets = np.ascontiguousarray(seg_offsets, np.int64)
        self.n = offsets.size - 1
        self.total_bases = int(offsets[-1])
        _check(_lib.mc_load_sequences(self._h, _p(letters), _p(offsets), C.c_int64(self.n), _p(segs), _p(seg_offsets)))

    def stage_fasta_bytes(self, raw: np.ndarray, n_records: int):
        assert raw.dtype == np.uint8 and raw.flags.c_contiguous
        _check(_lib.mc_stage_fasta_bytes(self._h, _p(raw), C.c_int64(raw.size), C.c_int64(n_records)))

    def ingest_fasta(self, raw: np.ndarray, span_begin: np.ndarray, span_end: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        """mc_ingest_fasta: raw file bytes + the byte span of every record's sequence lines -> letters on the device;
        returns the per-record flags (bit 0: has N, bit 1: has a letter other than ACGTN)."""
        raw = np.ascontiguousarray(raw, np.uint8)
        span_begin = np.ascontiguousarray(span_begin, np.int64)
        span_end = np.ascontiguousarray(span_end, np.int64)
        offsets = np.ascontiguousarray(offsets, np.int64)
        self.n = offsets.size - 1
        self.total_bases = int(offsets[-1])
        flags = np.zeros(self.n, np.uint8)
        _check(_lib.mc_ingest_fasta(self._h, _p(raw), C.c_int64(raw.size), _p(span_begin), _p(span_end), _p(offsets), C.c_int64(self.n), _p(flags)))
        return flags

    def load_segments(self, segs: np.ndarray, seg_offsets: np.ndarray, validate: bool = True):
        segs = np.ascontiguousarray(segs, np.int32)
        seg_offsets = np.ascontiguousarray(seg_offsets, np.int64)
        _check(_lib.mc_load_segments(self._h, _p(segs), _p(seg_offsets), C.c_int(1 if validate else 0)))

    def copy_letters(self) -> np.ndarray:
        out = np.zeros(self.total_bases, np.uint8)
        _check(_lib.mc_copy_letters(self._h, _p(out)))
        return out

    def copy_digits(self) -> np.ndarray:
        out = np.zeros(self.total_bases, np.uint8)
        _check(_lib.mc_copy_digits(self._h, _p(out)))
        return out

    def build_histograms(self, k: int, tbytes: int = 0):
        used, mx = C.c_int(0), C.c_uint64(0)
        _check(_lib.mc_build_histograms(self._h, C.c_int(k), C.c_int(tbytes), C.byref(used), C.byref(mx)))
        self.k, self.tbytes = k, used.value
        return used.value, int(mx.value)

    def load_histograms(self, hists: np.ndarray, lens: np.ndarray, k: int):
        hists = np.ascontiguousarray(hists)
        assert hists.dtype in (np.uint8, np.uint16) and hists.shape[1] == 4 ** k
        lens = np.ascontiguousarray(lens, np.uint64)
        self.n, self.k, self.tbytes = hists.shape[0], k, hists.dtype.itemsize
        _check(_lib.mc_load_histograms(self._h, _p(hists), C.c_int(self.tbytes), C.c_int(k), _p(lens), C.c_int64(self.n)))

    def copy_histograms(self) -> np.ndarray:
        out = np.zeros((self.n, 4 ** self.k), np.uint8 if self.tbytes == 1 else np.uint16)
        _check(_lib.mc_copy_histograms(self._h, _p(out)))
        return out

    def copy_point_stats(self):
        ln, mg, sq = (np.zeros(self.n, np.uint64) for _ in range(3))
        _check(_lib.mc_copy_point_stats(self._h, _p(ln), _p(mg), _p(sq)))
        return ln, mg, sq

    # -- stage 2 --------------------------------------------------------------------------
    def set_model(self, mins, maxs, weights, nfeat: int):
        mins = np.ascontiguousarray(mins, np.float64)
        maxs = np.ascontiguousarray(maxs, np.float64)
        weights = np.ascontiguousarray(weights, np.float64)
        _check(_lib.mc_set_model(self._h, _p(mins), _p(maxs), _p(weights), C.c_int(nfeat)))

    def distance_keys(self, center_rows) -> np.ndarray:
        cr = np.ascontiguousarray(center_rows, np.int32)
        out = np.zeros((cr.size, self.n), np.uint16)
        _check(_lib.mc_distance_keys(self._h, _p(cr), C.c_int(cr.size), _p(out)))
        return out

    def pair_features(self, a, b):
        a = np.ascontiguousarray(a, np.int32)
        b = np.ascontiguousarray(b, np.int32)
        raw = np.zeros((a.size, 5), np.float64)
        dist = np.zeros(a.size, np.uint64)
        _check(_lib.mc_pair_features(self._h, _p(a), _p(b), C.c_int64(a.size), _p(raw), _p(dist)))
        return raw, dist

    def pair_classify(self, a, b):
        a = np.ascontiguousarray(a, np.int32)
        b = np.ascontiguousarray(b, np.int32)
        s = np.zeros(a.size, np.float64)
        f0 = np.zeros(a.size, np.float64)
        fl = np.zeros(a.size, np.uint8)
        feats = np.zeros((a.size, 4), np.float64)
        _check(_lib.mc_pair_classify(self._h, _p(a), _p(b), C.c_int64(a.size), _p(s), _p(f0), _p(fl), _p(feats)))
        return s, f0, fl, feats

    def alive_reset(self):
        _check(_lib.mc_alive_reset(self._h))

    def alive_kill(self, rows):
        r = np.ascontiguousarray(rows, np.int64)
        _check(_lib.mc_alive_kill(self._h, _p(r), C.c_int64(r.size)))

    def scan(self, center_row: int, lo: int, hi: int, want_marks: bool = True):
        res = ScanResult()
        marks = np.zeros(max(hi - lo + 1, 0), np.uint8) if want_marks else None
        _check(_lib.mc_scan(self._h, C.c_int64(center_row), C.c_int64(lo), C.c_int64(hi), C.byref(res), _p(marks)))
        return res, marks

    def scan_enqueue(self, center_row: int, lo: int, hi: int, remove_marked: bool, slot: int):
        _check(_lib.mc_scan_enqueue(self._h, C.c_int64(center_row), C.c_int64(lo), C.c_int64(hi),
                                    C.c_int(1 if remove_marked else 0), C.c_int(slot)))

    def scan_enqueue_many(self, center_rows, lo, hi, remove_marked: bool, slot0: int = 0):
        cr = np.ascontiguousarray(center_rows, np.int64)
        lo = np.ascontiguousarray(lo, np.int64)
        hi = np.ascontiguousarray(hi, np.int64)
        # remove_marked: False / True, or MC_SCAN_CHAIN (2): a dependent chain that removes nothing
        _check(_lib.mc_scan_enqueue_many(self._h, _p(cr), _p(lo), _p(hi), C.c_int(cr.size),
                                         C.c_int(int(remove_marked)), C.c_int(slot0)))

    def scan_fold_dev(self, slot0: int, nslots: int, out_dev_ptr: int):
        """fold nslots scans into mc_scan_result records at a DEVICE address (e.g. tensor.data_ptr())"""
        _check(_lib.mc_scan_fold_dev(self._h, C.c_int(slot0), C.c_int(nslots), C.c_void_p(out_dev_ptr)))

    def set_stream(self, stream_ptr: int):
        _check(_lib.mc_set_stream(self._h, C.c_void_p(stream_ptr)))

    def scan_collect(self, slot0: int, nslots: int):
        res = (ScanResult * nslots)()
        _check(_lib.mc_scan_collect(self._h, C.c_int(slot0), C.c_int(nslots), C.byref(res)))
        return [r.as_tuple() for r in res]

    # -- multi-GPU: sharded scans ---------------------------------------------------------
    def comm_init(self, rank: int, world: int) -> bytes:
        """allocate this rank's inbox; returns its CUDA IPC handle (64 bytes) for the other ranks"""
        h = (C.c_uint8 * 64)()
        _check(_lib.mc_comm_init(self._h, C.c_int(rank), C.c_int(world), h))
        return bytes(h)

    def comm_connect(self, handles):
        """handles: the 64-byte IPC handles of all ranks in rank order (other processes)"""
        buf = b"".join(handles)
        arr = (C.c_uint8 * len(buf)).from_buffer_copy(buf)
        _check(_lib.mc_comm_connect(self._h, arr))

    @staticmethod
    def comm_connect_local(contexts):
        arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
        _check(_lib.mc_comm_connect_local(arr, C.c_int(len(contexts))))

    def scan_sharded_enqueue(self, center_row: int, lo: int, hi: int, remove_marked: bool, slot: int):
        _check(_lib.mc_scan_sharded_enqueue(self._h, C.c_int64(center_row), C.c_int64(lo), C.c_int64(hi),
                                            C.c_int(1 if remove_marked else 0), C.c_int(slot)))

    def scan_sharded_enqueue_many(self, center_rows, lo, hi, remove_marked: bool, slot0: int = 0):
        cr = np.ascontiguousarray(center_rows, np.int64)
        lo = np.ascontiguousarray(lo, np.int64)
        hi = np.ascontiguousarray(hi, np.int64)
        _check(_lib.mc_scan_sharded_enqueue_many(self._h, _p(cr), _p(lo), _p(hi), C.c_int(cr.size),
                                                 C.c_int(1 if remove_marked else 0), C.c_int(slot0)))

    def scan_sharded_collect(self, slot0: int, nslots: int):
        res = (ScanResult * nslots)()
        _check(_lib.mc_scan_sharded_collect(self._h, C.c_int(slot0), C.c_int(nslots), C.byref(res)))
        return [r.as_tuple() for r in res]

    def scan_sharded_burst(self, cr, lo, hi, remove_marked: bool, slot0: int, prev_slot0: int, prev_count: int):
        """arrays must already be contiguous int64 (this is the per-step call of a streaming caller)"""
        res = (ScanResult * max(prev_count, 1))()
        _check(_lib.mc_scan_sharded_burst(self._h, _p(cr), _p(lo), _p(hi), C.c_int(cr.size),
                                          C.c_int(int(remove_marked)), C.c_int(slot0), C.c_int(prev_slot0),
                                          C.c_int(prev_count), C.byref(res)))
        return [res[i].as_tuple() for i in range(prev_count)]

    def scan_sharded_combine(self, slot0: int, nslots: int):
        _check(_lib.mc_scan_sharded_combine(self._h, C.c_int(slot0), C.c_int(nslots)))

    def scan_sharded_wait(self, slot0: int, nslots: int):
        res = (ScanResult * nslots)()
        _check(_lib.mc_scan_sharded_wait(self._h, C.c_int(slot0), C.c_int(nslots), C.byref(res)))
        return [r.as_tuple() for r in res]

    def clone_points_from(self, src: "Context"):
        _check(_lib.mc_clone_points(self._h, src._h))
        self.n, self.k, self.tbytes = src.n, src.k, src.tbytes

    @staticmethod
    def accumulate_step_sharded(contexts, center_row: int, lo: int, hi: int, restart: bool):
        arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
        res = StepResult()
        rows = np.zeros(max(contexts[0].n, 1), np.int64)
        _check(_lib.mc_accumulate_step_sharded(arr, C.c_int(len(contexts)), C.c_int64(center_row), C.c_int64(lo), C.c_int64(hi),
                                               C.c_int(1 if restart else 0), C.byref(res), _p(rows), C.c_int64(rows.size)))
        return res, rows[: res.scan.n_pos].copy()

    # -- stage 3 --------------------------------------------------------------------------
    def mean_nearest(self, rows, append: bool = False):
        r = np.ascontiguousarray(rows, np.int64)
        row, dist = C.c_int64(-1), C.c_double(0)
        _check(_lib.mc_mean_nearest(self._h, _p(r), C.c_int64(r.size), C.c_int(1 if append else 0), C.byref(row), C.byref(dist)))
        return int(row.value), float(dist.value)

    def accumulate_step(self, center_row: int, lo: int, hi: int, restart: bool):
        """one accumulate() iteration: scan + remove + get_mean; returns (StepResult, marked rows ascending)"""
        res = StepResult()
        rows = getattr(self, "_rows_buf", None)
        if rows is None or rows.size < max(self.n, 1):
            rows = self._rows_buf = np.empty(max(self.n, 1), np.int64)
        _check(_lib.mc_accumulate_step(self._h, C.c_int64(center_row), C.c_int64(lo), C.c_int64(hi), C.c_int(1 if restart else 0),
                                       C.byref(res), _p(rows), C.c_int64(rows.size)))
        return res, rows[: res.scan.n_pos].copy()

    def accumulate_run(self, similarity: float, bin_bounds, bin_first_row):
        """ClusterFactory::MS's first phase (all accumulate() calls) in one persistent kernel.
        Returns (center_rows[nc], cluster_offsets[nc + 1], member_rows[n], RunStats)."""
        bounds = np.ascontiguousarray(bin_bounds, np.uint64)
        first = np.ascontiguousarray(bin_first_row, np.int64)
        n = int(self.n)
        centers = np.zeros(max(n, 1), np.int64)
        offs = np.zeros(n + 1, np.int64)
        members = np.zeros(max(n, 1), np.int64)
        st = RunStats()
        _check(_lib.mc_accumulate_run(self._h, C.c_double(similarity), _p(bounds), _p(first), C.c_int64(bounds.size),
                                      _p(centers), _p(offs), _p(members), C.byref(st)))
        nc = int(st.n_clusters)
        return centers[:nc].copy(), offs[:nc + 1].copy(), members[:int(offs[nc])].copy(), st

    def near_threshold_count(self, reset: bool = False) -> int:
        out = C.c_int64(0)
        _check(_lib.mc_near_threshold_count(self._h, C.byref(out), C.c_int(1 if reset else 0)))
        return int(out.value)

    def permute_rows(self, old_of_new, n_alive: int):
        o = np.ascontiguousarray(old_of_new, np.int64)
        _check(_lib.mc_permute_rows(self._h, _p(o), C.c_int64(o.size), C.c_int64(n_alive)))

    def update_centers(self, center_rows, cand_rows, cand_begin, cand_end) -> np.ndarray:
        cr = np.ascontiguousarray(center_rows, np.int64)
        cand = np.ascontiguousarray(cand_rows, np.int64)
        cb = np.ascontiguousarray(cand_begin, np.int64)
        ce = np.ascontiguousarray(cand_end, np.int64)
        out = np.zeros(cr.size, np.int64)
        _check(_lib.mc_update_centers(self._h, _p(cr), C.c_int64(cr.size), _p(cand), C.c_int64(cand.size), _p(cb), _p(ce), _p(out)))
        return out

    # -- stage 4 --------------------------------------------------------------------------
    def align_pairs(self, a, b):
        a = np.ascontiguousarray(a, np.int32)
        b = np.ascontiguousarray(b, np.int32)
        sc = np.zeros(a.size, np.int32)
        ln = np.zeros(a.size, np.int32)
        mt = np.zeros(a.size, np.int32)
        _check(_lib.mc_align_pairs(self._h, _p(a), _p(b), C.c_int64(a.size), _p(sc), _p(ln), _p(mt)))
        return sc, ln, mt

    # -- one-shot host-buffer calls -------------------------------------------------------
    def kmer_histograms_host(self, letters: np.ndarray, offsets: np.ndarray, k: int, tbytes: int = 1, out: np.ndarray | None = None):
        letters = np.ascontiguousarray(letters, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = offsets.size - 1
        if out is None:
            out = np.zeros((n, 4 ** k), np.uint8 if tbytes == 1 else np.uint16)
        mx = C.c_uint64(0)
        _check(_lib.mc_kmer_histograms_host(self._h, _p(letters), _p(offsets), C.c_int64(n), C.c_int(k), C.c_int(tbytes), _p(out), C.byref(mx)))
        self.n, self.k, self.tbytes, self.total_bases = n, k, tbytes, int(offsets[-1])
        return out, int(mx.value)

    def scan_host(self, hists: np.ndarray, lens: np.ndarray, k: int, center_rows, marks_out: np.ndarray | None = None):
        hists = np.ascontiguousarray(hists)
        lens = np.ascontiguousarray(lens, np.uint64)
        cr = np.ascontiguousarray(center_rows, np.int64)
        res = (ScanResult * cr.size)()
        self.n, self.k, self.tbytes = hists.shape[0], k, hists.dtype.itemsize
        _check(_lib.mc_scan_host(self._h, _p(hists), C.c_int(self.tbytes), C.c_int(k), _p(lens), C.c_int64(self.n),
                                 _p(cr), C.c_int(cr.size), C.byref(res), _p(marks_out)))
        return [r.as_tuple() for r in res]
