"""TEST INFRASTRUCTURE: a literal Python restatement of the reference's length-binned container
(src/cluster/src/bvec.cpp, bvec_iterator.h) on (row, length) entries, used to drive
mc_accumulate_step from the host and compare the outcome with mc_accumulate_run."""
from __future__ import annotations

import numpy as np

U64 = (1 << 64) - 1


def layout(lens: np.ndarray, bin_size: int = 1000):
    """bvec::bvec + insert + insert_finalize (bvec.cpp:10-24,152-177,209-218): returns
    (bounds, order, first_rows): order[r] = index into `lens` of the point that becomes row r."""
    s = np.sort(lens)
    bounds = [int(s[i]) for i in range(0, len(s), bin_size)]
    bins = [[] for _ in bounds]
    for i, ln in enumerate(lens):
        f, b = index_of(bounds, int(ln))
        sizes = [len(bins[j]) for j in range(f, b + 1)]
        m = min(sizes)
        mins = [j for j in range(f, b + 1) if len(bins[j]) == m]
        bins[mins[len(mins) // 2]].append((i, int(ln)))
    order, first = [], [0]
    for bn in bins:
        bn.sort(key=lambda e: e[1])   # any order among equal lengths is a valid bvec
        order += [e[0] for e in bn]
        first.append(len(order))
    return np.array(bounds, np.uint64), np.array(order, np.int64), np.array(first, np.int64)


def index_of(bounds, point):
    """bvec::index_of (bvec.cpp:123-149), the literal linear walk."""
    low, high = len(bounds) - 1, 0
    for i in range(len(bounds)):
        prev = bounds[i - 1] if i > 0 else 0
        pi = i - 1 if i > 0 else 0
        if prev <= point <= bounds[i]:
            low, high = min(low, pi), max(high, pi)
    if point >= bounds[-1]:
        high = max(high, len(bounds) - 1)
    return low, high


class LitBvec:
    def __init__(self, bounds, first_rows, lens_by_row):
        self.bounds = [int(b) for b in bounds]
        self.data = [[(r, int(lens_by_row[r])) for r in range(int(first_rows[b]), int(first_rows[b + 1]))]
                     for b in range(len(self.bounds))]

    def pop(self):
        for bn in self.data:
            if bn:
                return bn.pop(0)[0]
        return -1

    def erase_row(self, row):
        for bn in self.data:
            for i, e in enumerate(bn):
                if e[0] == row:
                    del bn[i]
                    return

    def inner(self, length, idx, want_front):
        """bvec::inner_index_of (bvec.cpp:52-120); returns (bin, pos or None when nothing was set)."""
        if not self.data[idx]:
            rng = range(len(self.data)) if want_front else range(len(self.data) - 1, -1, -1)
            for i in rng:
                if self.data[i]:
                    return i, 0
            return idx, None
        bn = self.data[idx]
        front = back = 0
        low, high = 0, len(bn) - 1
        while low <= high:
            mid = (low + high) // 2
            d = bn[mid][1]
            if d == length:
                front = back = mid
                break
            elif length < d:
                high = mid
            else:
                low = mid + 1
            if low == high:
                front, back = low, high
                break
        if want_front:
            i = front
            while i >= 0 and bn[i][1] == length:
                front = i
                i -= 1
            return idx, front
        i = back
        while i < len(bn) and bn[i][1] == length:
            back = i
            i += 1
        return idx, back

    def diff(self, a, r):
        """bvec_iterator::operator- (bvec_iterator.h:61-76) on (bin, pos) with size_t positions."""
        if a[0] < r[0] or (a[0] == r[0] and a[1] < r[1]):
            return -self.diff(r, a)

        def s64(v):
            v &= U64
            return v - (1 << 64) if v >> 63 else v
        if a[0] == r[0]:
            return s64(a[1] - r[1])
        t = s64(a[1]) + s64(len(self.data[r[0]]) - r[1])
        for i in range(r[0] + 1, a[0]):
            t += len(self.data[i])
        return t

    def get_range(self, begin_len, end_len):
        """bvec::get_range (bvec.cpp:247-278) -> (lo_row, hi_row, front_bin, back_bin) or None when the
        OpenMP trip count of `for (it = front; it <= back; ++it)` is <= 0."""
        fbin, _ = index_of(self.bounds, begin_len)
        _, bbin = index_of(self.bounds, end_len)
        fpos, bpos = 0, (len(self.data[-1]) - 1) & U64
        fbin, p = self.inner(begin_len, fbin, True)
        if p is not None:
            fpos = p
        bbin, p = self.inner(end_len, bbin, False)
        if p is not None:
            bpos = p
        if self.diff((bbin, bpos), (fbin, fpos)) + 1 <= 0:
            return None
        return self.data[fbin][fpos][0], self.data[bbin][bpos][0], fbin, bbin

    def remove_rows(self, rows):
        rs = set(int(r) for r in rows)
        for b in range(len(self.data)):
            self.data[b] = [e for e in self.data[b] if e[0] not in rs]


def accumulate_by_steps(ctx, sim, bounds, first_rows, lens_by_row):
    """ClusterFactory.cpp:722-729 + accumulate (:637-714), one mc_accumulate_step per scan."""
    bv = LitBvec(bounds, first_rows, lens_by_row)
    ctx.alive_reset()
    centers, offs, members = [], [0], []
    scans = evals = steps = 0
    last = bv.pop()
    if last >= 0:
        ctx.alive_kill([last])
    while last >= 0:
        current = [last]
        restart, is_min, seed = True, False, -1
        while not is_min:
            ln = int(lens_by_row[last])
            rng = bv.get_range(int(np.float64(ln) * np.float64(sim)), int(np.float64(ln) / np.float64(sim)))
            lo, hi = (rng[0], rng[1]) if rng else (0, -1)
            res, rows = ctx.accumulate_step(last, lo, hi, restart)
            restart = False
            steps += 1
            if hi >= lo:
                scans += 1
                evals += res.scan.n_eval
            is_min = res.scan.n_pos == 0
            if is_min:
                if res.scan.best_row < 0:
                    seed = bv.pop()
                else:
                    seed = res.scan.best_row
                    bv.erase_row(seed)
                if seed >= 0:
                    ctx.alive_kill([seed])
            else:
                bv.remove_rows(rows)
                current += [int(r) for r in rows]
                last = int(res.nearest_row)
        centers.append(last)
        members += current
        offs.append(len(members))
        last = seed
    return np.array(centers, np.int64), np.array(offs, np.int64), np.array(members, np.int64), (scans, evals, steps)
