// Phase A of the clustering -- ClusterFactory::accumulate (ClusterFactory.cpp:637-714) with its bvec
// bookkeeping (bvec.cpp:27-38 pop, :52-120 inner_index_of, :123-149 index_of, :247-278 get_range,
// :281-285 erase, :290-317 remove_available), Trainer::get_close (Trainer.cpp:34-114) and get_mean
// (ClusterFactory.cpp:382-425) -- as ONE persistent cooperative kernel: the greedy loop
//     range of the center's length window -> scan -> (marks: extend `current`, mean, nearest -> new
//     center | no marks: close the cluster, arg-max becomes the next seed)
// never returns to the host.  The host launches once and reads the clusters back.
//
// One CTA per SM; every CTA runs the same control flow on the same values (read from global memory
// behind grid-wide barriers), so there is no controller CTA and no broadcast:
//   control   two warps: the length window of the center as an inclusive row range.  The bvec lives on
//             the device as (a) immutable per-row search records computed once (which bin the window
//             starts / ends in and how many of the bin's entries are shorter / not longer than the
//             bound), (b) an alive bitmap in global memory, (c) per-bin alive counts in shared memory,
//             kept identical in every CTA.  The reference's binary search over the i-th ALIVE entry of
//             a bin is replayed on ranks (popcounts of the bitmap), without touching the lengths.
//   scan      the TMA-staged streaming scan of scan_tma.cu (one producer warp feeding per-consumer
//             rings with cp.async.bulk, VABSDIFF4 / DP4A reductions, FP64 feature + GLM epilogue on all
//             lanes); a CTA owns a contiguous run of tiles, so the rows it marks are ordered.  Marked rows
//             are added to per-CTA bin sums in shared memory; the bitmap is READ-ONLY during control and
//             scan (a fast CTA may be scanning while a slow one still derives the range of the same step).
//   barrier 1 (after the CTA partials and the bin sums have been flushed)
//   fold      every CTA folds all partials.  No positives: the cluster is closed, the arg-max (or the
//             first alive row) becomes the next seed -- no second barrier.
//   tail      positives: every CTA clears the bits of its marked rows, writes them into the cluster's member
//             list at the offset its predecessors' counts give, derives the truncated mean from the global bin sums and
//             evaluates distance_d for its own new members and its share of the older ones
//   barrier 2, fold of the nearest-member partials -> the new center.
// HBM traffic per step = the scan's algorithmic bytes; everything else is a few KB out of L2.
#include "pair_core.cuh"
#include "tma_utils.cuh"

constexpr int PA_MAX_CONSUMERS = 15;   // 16 warps = 512 threads: 128 registers per thread
constexpr int PA_WARPS = 1 + PA_MAX_CONSUMERS;
constexpr int PA_THREADS = 32 * PA_WARPS;
constexpr int PA_MAX_STAGES = 32;
constexpr unsigned long long PA_TIMEOUT_NS = 20000000000ull;   // a barrier that does not complete is an error, not a hang

struct PaPartial {   // one per CTA and step
	mc_scan_result s;
	long long n_near;   // evaluated pairs with |GLM sum| < 1e-9 (north_star: "reported by count")
	long long pad;
};

struct PaNear {
	long long pos;   // position in `current` (first minimum wins, ClusterFactory.cpp:412-418)
	double dist;
	long long row;
	long long pad;
};

__device__ __forceinline__ void pa_near_merge(PaNear &a, const PaNear &b) {
	if (b.pos >= 0 && (a.pos < 0 || b.dist < a.dist || (b.dist == a.dist && b.pos < a.pos))) a = b;
}

// Immutable search record of a row used as a center: bvec::get_range for its length window
// [len * id, len / id] needs, per end, the bin index_of() picks and -- for the in-bin binary search --
// how many entries of that bin (alive or not) are shorter than / not longer than the bound.
struct __align__(32) PaRange {
	int fb, bb;        // bins of the lower / upper bound (bvec::index_of)
	int f_lt, f_le;    // entries of bin fb with length < / <= the lower bound
	int b_lt, b_le;    // entries of bin bb with length < / <= the upper bound
	int pad0, pad1;
};

struct PaArgs {
	const uint8_t *hist;
	const McRowAux *aux;
	long long n;
	const unsigned long long *bounds;   // nb bin bounds (bvec.cpp:10-24)
	const int *row0;                    // nb + 1: first row of every bin
	int nb;
	int qmax;                           // capacity of the per-CTA tile tables
	const PaRange *range_tab;           // n
	uint32_t *alive_bits;               // bit r = row r is still in the bvec
	unsigned long long *g_sum;          // 3 x pa_gsum_words(bins): running bin sums of `current`, rotating per cluster
	unsigned long long *recs;           // 2 x grid x PA_REC_WORDS: scan summaries of the CTAs as tagged words
	unsigned long long *near_recs;      // 2 x grid x PA_REC_WORDS: nearest-member candidates of the CTAs
	unsigned long long *bar;            // [1] abort code
	int *members;                       // n: the clusters' member rows (ORIGINAL row numbers), cluster after cluster
	int *mcur;                          // n: the same members as rows of the current numbering (see compaction)
	// row compaction: staging copies of the alive rows, ping-pong (capacities see mc_accumulate_run); null = off
	uint8_t *st_hist[2];
	McRowAux *st_aux[2];
	int *st_orig[2];
	PaRange *st_range[2];
	uint32_t *st_bits[2];
	long long st_cap[2];
	long long compact_min;              // no compaction below this many current rows
	int compact_shift;                  // compaction once at most n_cur - (n_cur >> shift) of the current rows are alive
	double sim;
	int *cl_center;                     // n
	int *cl_off;                        // n + 1
	long long *stats;                   // [0] clusters [1] scans [2] evals [3] near-threshold pairs [4] steps [5] ns
	unsigned long long *trace;          // optional: PA_TRACE_SLOTS words per step of CTA 0
	int trace_steps;
	int ns, ncw, d;                     // ring geometry: stages, consumer warps, stages per consumer
	McModel model;
};

// ---- bvec::index_of (bvec.cpp:123-149) on the sorted bounds ------------------------------------
__device__ __forceinline__ void pa_index_of(const unsigned long long *b, int nb, unsigned long long point, int *pfront, int *pback) {
	int low = nb - 1, high = 0;
	int lo = 0, hi = nb;
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (b[mid] < point) lo = mid + 1; else hi = mid; }
	const int i1 = lo;   // first bound >= point
	if (i1 < nb) {
		hi = nb;
		while (lo < hi) { const int mid = (lo + hi) >> 1; if (b[mid] <= point) lo = mid + 1; else hi = mid; }
		const int imax = lo < nb - 1 ? lo : nb - 1;   // number of bounds <= point, capped
		const int l = i1 > 0 ? i1 - 1 : 0, h = imax > 0 ? imax - 1 : 0;
		low = low < l ? low : l;
		high = high > h ? high : h;
	}
	if (point >= b[nb - 1]) high = high > nb - 1 ? high : nb - 1;
	*pfront = low;
	*pback = high;
}

// The search record of row r (see PaRange) for bins that start at row0[] (the rows of a bin in non-decreasing
// length order, bvec::insert_finalize, bvec.cpp:209-218)
template <class Row0>
__device__ __forceinline__ PaRange pa_make_range(const McRowAux *__restrict__ aux, long long r, const unsigned long long *bounds,
                                                 const Row0 *row0, int nb, double sim) {
	const unsigned long long len = __ldcg(&aux[r].len);
	// ClusterFactory.cpp:651-652: get_range(len * id, len / id), both converted to uint64
	const unsigned long long begin_len = (unsigned long long)((double)len * sim);
	const unsigned long long end_len = (unsigned long long)((double)len / sim);
	PaRange g;
	int dummy;
	pa_index_of(bounds, nb, begin_len, &g.fb, &dummy);
	pa_index_of(bounds, nb, end_len, &dummy, &g.bb);
	auto count_below = [&](int bin, unsigned long long key, bool inclusive) {
		int lo = (int)row0[bin], hi = (int)row0[bin + 1];
		const int base = lo;
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			const unsigned long long v = __ldcg(&aux[mid].len);
			if (inclusive ? v <= key : v < key) lo = mid + 1; else hi = mid;
		}
		return lo - base;
	};
	g.f_lt = count_below(g.fb, begin_len, false);
	g.f_le = count_below(g.fb, begin_len, true);
	g.b_lt = count_below(g.bb, end_len, false);
	g.b_le = count_below(g.bb, end_len, true);
	g.pad0 = 0; g.pad1 = 0;
	return g;
}

// One thread per row: the search records.  Also checks what the kernel relies on: rows of a bin are in
// non-decreasing length order.
__global__ void pa_prepare_kernel(const McRowAux *__restrict__ aux, long long n, const unsigned long long *__restrict__ bounds,
                                  const int *__restrict__ row0, int nb, double sim, PaRange *__restrict__ tab,
                                  unsigned int *__restrict__ err) {
	for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
		tab[r] = pa_make_range(aux, r, bounds, row0, nb, sim);
		// sortedness inside the bin of r
		int lo = 0, hi = nb;
		while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long)row0[mid + 1] <= r) lo = mid + 1; else hi = mid; }
		if (r > row0[lo] && aux[r - 1].len > aux[r].len) atomicExch(err, 1u);
	}
}

// ---- alive bitmap helpers (warp-cooperative; every lane gets the result) ------------------------
// alive rows among the first pa / pb rows of [r0, r1) and in all of it
// `skip`: a row that has left the bvec although its bit may still be set (the center of the running step)
__device__ __forceinline__ void pa_rank3(const uint32_t *bits, long long r0, long long r1, long long pa, long long pb, long long skip,
                                         int lane, unsigned &ra, unsigned &rb, unsigned &tot) {
	unsigned a = 0, b = 0, t = 0;
	if (r1 > r0) {
		const long long w0 = r0 >> 5, w1 = (r1 - 1) >> 5;
		for (long long w = w0 + lane; w <= w1; w += 32) {
			uint32_t v = __ldcg(bits + w);
			if (w == w0) v &= 0xffffffffu << (r0 & 31);
			if (w == w1 && (r1 & 31)) v &= 0xffffffffu >> (32 - (int)(r1 & 31));
			if (w == (skip >> 5)) v &= ~(1u << (int)(skip & 31));
			auto below = [&](long long lim) {
				const long long dlt = lim - (w << 5);
				return dlt <= 0 ? 0u : (dlt >= 32 ? 0xffffffffu : ((1u << (int)dlt) - 1u));
			};
			a += __popc(v & below(r0 + pa));
			b += __popc(v & below(r0 + pb));
			t += __popc(v);
		}
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		a += __shfl_xor_sync(MC_FULL_MASK, a, o);
		b += __shfl_xor_sync(MC_FULL_MASK, b, o);
		t += __shfl_xor_sync(MC_FULL_MASK, t, o);
	}
	ra = a; rb = b; tot = t;
}

// row of the pos-th (0-based) alive row of [r0, r1), or -1
__device__ __forceinline__ long long pa_select(const uint32_t *bits, long long r0, long long r1, unsigned long long pos, long long skip, int lane) {
	if (r1 <= r0) return -1;
	const long long w0 = r0 >> 5, w1 = (r1 - 1) >> 5;
	unsigned long long seen = 0;
	for (long long wb = w0; wb <= w1; wb += 32) {
		const long long w = wb + lane;
		uint32_t v = 0;
		if (w <= w1) {
			v = __ldcg(bits + w);
			if (w == w0) v &= 0xffffffffu << (r0 & 31);
			if (w == w1 && (r1 & 31)) v &= 0xffffffffu >> (32 - (int)(r1 & 31));
			if (w == (skip >> 5)) v &= ~(1u << (int)(skip & 31));
		}
		const unsigned pc = __popc(v);
		unsigned incl = pc;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned u = __shfl_up_sync(MC_FULL_MASK, incl, o);
			if (lane >= o) incl += u;
		}
		const unsigned tot = __shfl_sync(MC_FULL_MASK, incl, 31);
		if (pos < seen + tot) {
			const unsigned long long before = seen + incl - pc;
			const bool mine = before <= pos && pos < seen + incl;
			const unsigned ball = __ballot_sync(MC_FULL_MASK, mine);
			const int src = __ffs(ball) - 1;
			long long row = -1;
			if (mine) row = (w << 5) + __fns(v, 0, (int)(pos - before) + 1);
			return __shfl_sync(MC_FULL_MASK, row, src);
		}
		seen += tot;
	}
	return -1;
}

// bvec::inner_index_of's binary search (bvec.cpp:52-120) over the ALIVE entries of a bin, replayed on
// ranks: alive entries [0, a_lt) are shorter than the bound, [a_lt, a_le) equal, [a_le, A) longer.
__device__ __forceinline__ void pa_inner_search(unsigned long long A, unsigned long long a_lt, unsigned long long a_le,
                                                unsigned long long &front, unsigned long long &back) {
	front = 0; back = 0;
	unsigned long long low = 0, high = A - 1;
	while (low <= high) {
		const unsigned long long mid = (low + high) / 2;
		if (mid >= a_lt && mid < a_le) { front = back = mid; break; }   // d == length
		else if (mid >= a_le) high = mid;                               // length < d
		else low = mid + 1;
		if (low == high) { front = low; back = high; break; }
	}
	// the walks over equal lengths (bvec.cpp:100-118)
	if (front >= a_lt && front < a_le) front = a_lt;
	if (back >= a_lt && back < a_le) back = a_le - 1;
}

// One end of bvec::get_range inside the bin that owns rows [r0, r1): of the bin's entries (alive or not)
// the first p_lt are shorter than the bound and the first p_le not longer; the in-bin search of the
// reference runs over the ALIVE entries and ends at a position (front or back flavour) whose row comes
// back.  Bins of up to 32 * PA_KW words are searched out of registers: one load of the bitmap words,
// then ranks, search and select without touching memory again.
constexpr int PA_KW = 4;

// 32-bit flavour of pa_inner_search (bins hold fewer than 2^31 entries)
__device__ __forceinline__ void pa_inner_search32(unsigned A, unsigned a_lt, unsigned a_le, unsigned &front, unsigned &back) {
	front = 0; back = 0;
	unsigned low = 0, high = A - 1;
	while (low <= high) {
		const unsigned mid = (low + high) >> 1;   // A < 2^31: no overflow
		if (mid >= a_lt && mid < a_le) { front = back = mid; break; }
		else if (mid >= a_le) high = mid;
		else low = mid + 1;
		if (low == high) { front = low; back = high; break; }
	}
	if (front >= a_lt && front < a_le) front = a_lt;
	if (back >= a_lt && back < a_le) back = a_le - 1;
}

__device__ __forceinline__ void pa_locate(const uint32_t *bits, long long r0l, long long r1l, int p_lt, int p_le, bool want_back,
                                          long long skipl, int lane, unsigned long long &pos_out, long long &row_out) {
	const int r0 = (int)r0l, r1 = (int)r1l, skip = (int)skipl;
	const int w0 = r0 >> 5, w1 = (r1 - 1) >> 5;
	const int nw = w1 - w0 + 1;
	if (nw > 32 * PA_KW) {
		unsigned a_lt, a_le, tot;
		pa_rank3(bits, r0l, r1l, p_lt, p_le, skipl, lane, a_lt, a_le, tot);
		unsigned long long f, b;
		pa_inner_search(tot, a_lt, a_le, f, b);
		pos_out = want_back ? b : f;
		row_out = pos_out < tot ? pa_select(bits, r0l, r1l, pos_out, skipl, lane) : -1;
		return;
	}
	uint32_t v[PA_KW];
	unsigned a = 0, b = 0, t = 0;
	const int lim_lt = r0 + p_lt, lim_le = r0 + p_le;
#pragma unroll
	for (int j = 0; j < PA_KW; j++) {
		const int w = w0 + j * 32 + lane;
		uint32_t x = 0;
		if (w <= w1) {
			x = __ldcg(bits + w);
			if (w == w0) x &= 0xffffffffu << (r0 & 31);
			if (w == w1 && (r1 & 31)) x &= 0xffffffffu >> (32 - (r1 & 31));
			if (w == (skip >> 5)) x &= ~(1u << (skip & 31));
		}
		v[j] = x;
		auto below = [&](int lim) {   // bits of word w that belong to rows < lim
			const int dlt = lim - (w << 5);
			return dlt <= 0 ? 0u : (dlt >= 32 ? 0xffffffffu : ((1u << dlt) - 1u));
		};
		a += __popc(x & below(lim_lt));
		b += __popc(x & below(lim_le));
		t += __popc(x);
	}
	a = __reduce_add_sync(MC_FULL_MASK, a);
	b = __reduce_add_sync(MC_FULL_MASK, b);
	t = __reduce_add_sync(MC_FULL_MASK, t);
	if (t == 0) { pos_out = 0; row_out = -1; return; }
	unsigned f, bk;
	pa_inner_search32(t, a, b, f, bk);
	const unsigned pos = want_back ? bk : f;
	pos_out = pos;
	int row = -1;
	unsigned seen = 0;
#pragma unroll
	for (int j = 0; j < PA_KW; j++) {
		if (j * 32 < nw) {
			const unsigned pc = __popc(v[j]);
			unsigned incl = pc;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const unsigned u = __shfl_up_sync(MC_FULL_MASK, incl, o);
				if (lane >= o) incl += u;
			}
			const unsigned tot = __shfl_sync(MC_FULL_MASK, incl, 31);
			if (row < 0 && pos >= seen && pos < seen + tot) {
				const unsigned before = seen + incl - pc;
				const bool mine = before <= pos && pos < before + pc;
				const int src = __ffs(__ballot_sync(MC_FULL_MASK, mine)) - 1;
				int r = -1;
				if (mine) {
					// the (pos - before)-th set bit of v[j], by halving
					uint32_t x = v[j];
					unsigned k = pos - before;
					int bit = 0;
					unsigned c = __popc(x & 0xffffu); if (k >= c) { k -= c; x >>= 16; bit += 16; }
					c = __popc(x & 0xffu); if (k >= c) { k -= c; x >>= 8; bit += 8; }
					c = __popc(x & 0xfu); if (k >= c) { k -= c; x >>= 4; bit += 4; }
					c = __popc(x & 0x3u); if (k >= c) { k -= c; x >>= 2; bit += 2; }
					if (k >= (x & 1u)) bit += 1;
					r = ((w0 + j * 32 + lane) << 5) + bit;
				}
				row = __shfl_sync(MC_FULL_MASK, r, src);
			}
			seen += tot;
		}
	}
	row_out = row;
}

// ---- grid-wide exchange: tagged records instead of a barrier -------------------------------------
// What the CTAs owe each other at the two synchronisation points of a step is one small record each
// (the scan summary of the CTA; its nearest-member candidate).  A record is PA_REC_WORDS 8-byte words
// {u32 data, u32 tag}, tag = step + 1: an 8-byte store is single-copy atomic, so a reader that sees
// the tag of this step in a word also has its data -- no counter, no atomic, one L2 round trip.  Every
// CTA polls the records of ALL CTAs, so passing the poll is also the barrier: nobody is past it before
// everybody has arrived.  Writers fence before they store (their bin-sum atomics, member lists and
// bitmap updates are visible to whoever sees the record), readers fence after the poll.  Records are
// double-buffered by step parity (a CTA is at most one exchange ahead of the slowest one).
constexpr int PA_REC_WORDS = 8;   // 64 bytes
constexpr int PA_RPL = 5;         // records a lane of the reading warp takes: grids of up to 160 CTAs

__device__ __forceinline__ void pa_ll_store(unsigned long long *p, uint32_t data, uint32_t tag) {
	const unsigned long long v = ((unsigned long long)tag << 32) | (unsigned long long)data;
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// one shot at the NW words of a record (pairs of words per load); true when every word carries `tag`
template <int NW>
__device__ __forceinline__ bool pa_rec_try(const unsigned long long *rec, uint32_t tag, uint32_t (&data)[NW]) {
	static_assert(NW % 2 == 0, "records are read as pairs of words");
	bool ok = true;
#pragma unroll
	for (int i = 0; i < NW; i += 2) {
		unsigned long long a, b;
		asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(rec + i) : "memory");
		ok = ok && (uint32_t)(a >> 32) == tag && (uint32_t)(b >> 32) == tag;
		data[i] = (uint32_t)a;
		data[i + 1] = (uint32_t)b;
	}
	return ok;
}

// the record of a CTA that is late: spin on ONE word (every other CTA is spinning on the same line), then
// read all of it; false = the run was aborted (a peer never wrote: time-out, not a hang)
template <int NW>
__device__ __forceinline__ bool pa_rec_wait(const unsigned long long *rec, uint32_t tag, uint32_t (&data)[NW], unsigned long long *abort_word) {
	unsigned long long t0 = 0;
	for (unsigned spins = 0;; spins++) {
		unsigned long long a;
		asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(rec) : "memory");
		if ((uint32_t)(a >> 32) == tag && pa_rec_try<NW>(rec, tag, data)) return true;
		if ((spins & 0xff) == 0xff) {
			unsigned long long t1, ab;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
			asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(ab) : "l"(abort_word) : "memory");
			if (ab) return false;
			if (t0 == 0) t0 = t1;
			else if (t1 - t0 > PA_TIMEOUT_NS) { atomicExch(abort_word, 1ull); return false; }
		}
	}
}

// Release without a fence.  What a CTA publishes before its record are read-modify-write operations on
// global memory (bin-sum additions, bitmap bits, member rows as exchanges).  An atomic whose RESULT has come
// back has been performed in L2, the point of coherence of the GPU; the record is stored after every such
// result has been consumed (pa_retire: a real instruction that needs the value, then a block barrier), so a
// CTA that sees the record reads the updated words (ld.cg / ld.relaxed.gpu go to L2).  A gpu-scope fence in
// the same place costs 0.7 us per exchange on B200 (-DPA_RELEASE_BY_FENCE=1 builds that variant).
#ifndef PA_RELEASE_BY_FENCE
#define PA_RELEASE_BY_FENCE 0
#endif
__device__ __forceinline__ void pa_retire(unsigned long long v, int *sink) {
#if !PA_RELEASE_BY_FENCE
	unsigned z;
	asm volatile("{ .reg .b64 t; and.b64 t, %1, 0; cvt.u32.u64 %0, t; }" : "=r"(z) : "l"(v));
	if (z) *sink = 1;   // never true; the branch needs the result of the atomic
#endif
}
__device__ __forceinline__ void pa_release_fence() {
#if PA_RELEASE_BY_FENCE
	__threadfence();
#endif
}

// grid-wide barrier on a monotonic counter (compactions only: a few per run): barrier e is passed when the
// counter reaches e * gridDim.x.  false = the run was aborted (time-out).
__device__ __forceinline__ bool pa_grid_barrier(unsigned long long *bar, unsigned long long target, int *s_flag) {
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ull) : "memory");
		unsigned long long t0 = 0;
		int ok = 1;
		for (unsigned spins = 0;; spins++) {
			unsigned long long v;
			asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
			if (v >= target) break;
			if ((spins & 0x3ff) == 0x3ff) {
				unsigned long long t1, ab;
				asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
				asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(ab) : "l"(bar + 1) : "memory");
				if (ab) { ok = 0; break; }
				if (t0 == 0) t0 = t1;
				else if (t1 - t0 > PA_TIMEOUT_NS) { atomicExch(bar + 1, 1ull); ok = 0; break; }
			}
		}
		*s_flag = ok;
	}
	__syncthreads();
	return *s_flag != 0;
}


// The running bin sums of `current` are the one thing every CTA adds to at the same time.  2 KB of
// consecutive 64-bit counters live in FOUR L2 slices (the slice hash ignores most of the low address bits),
// and ~100 CTAs x 4^k reductions queue up there for microseconds; one 32-byte sector (4 bins) per KB puts
// every sector into a slice of its own.
constexpr int PA_GSUM_STRIDE = 128;   // 64-bit words between the sectors of 4 bins
__host__ __device__ __forceinline__ size_t pa_gsum_idx(int b) { return (size_t)(b >> 2) * PA_GSUM_STRIDE + (size_t)(b & 3); }
__host__ __device__ __forceinline__ size_t pa_gsum_words(int nbins) { return (size_t)((nbins + 3) / 4) * PA_GSUM_STRIDE; }

// the warp's scan summary with warp-wide reductions: counts are sums; the arg-max is the largest f0 and,
// among equal f0, the smallest row (mc_scan_merge's rule): f0 as an order-preserving 64-bit key, high word
// first.  Every lane returns the warp's summary.
__device__ __forceinline__ void pa_scan_fold_redux(mc_scan_result &v) {
	v.n_eval = __reduce_add_sync(MC_FULL_MASK, (unsigned)v.n_eval);
	v.n_pos = __reduce_add_sync(MC_FULL_MASK, (unsigned)v.n_pos);
	const bool has = v.best_row >= 0;
	unsigned long long k = (unsigned long long)__double_as_longlong(v.best_f0);
	if (k == 0x8000000000000000ull) k = 0;              // -0.0 == +0.0 for the comparison the reference makes
	k = (k >> 63) ? ~k : (k | 0x8000000000000000ull);   // total order of the doubles (no NaN: f0 > best never holds for one)
	const unsigned hi = has ? (unsigned)(k >> 32) : 0u, lo = (unsigned)k;
	const unsigned mhi = __reduce_max_sync(MC_FULL_MASK, hi);
	const bool c1 = has && hi == mhi;
	const unsigned mlo = __reduce_max_sync(MC_FULL_MASK, c1 ? lo : 0u);
	const bool c2 = c1 && lo == mlo;
	const unsigned mrow = __reduce_min_sync(MC_FULL_MASK, c2 ? (unsigned)v.best_row : 0xffffffffu);
	if (__any_sync(MC_FULL_MASK, has)) {
		unsigned long long kk = ((unsigned long long)mhi << 32) | mlo;
		kk = (kk >> 63) ? (kk & 0x7fffffffffffffffull) : ~kk;
		v.best_f0 = __longlong_as_double((long long)kk);
		v.best_row = (long long)mrow;
	} else {
		v.best_f0 = -1.0;
		v.best_row = -1;
	}
}

// rows per tile: 32 (one row per lane in the epilogue) while the tile fits 32 KB
template <int RB>
struct PaTile {
	static constexpr int RT = (RB * 32 <= 32 * 1024) ? 32 : (32 * 1024) / RB;
	static constexpr int ROW_BYTES = RT * RB;
	static constexpr int AUX_BYTES = RT * 32;
	static constexpr int STAGE_BYTES = ((ROW_BYTES + AUX_BYTES + 127) / 128) * 128;
};

// trace slots of a step (CTA 0): 0 start, 1 range known, 2 scan done, 3 summaries of all CTAs in, 5 tail done,
// 6 candidates of all CTAs in, 7 end; 8 / 9: how long after CTA 0 the LAST CTA started / finished its scan (ns)
constexpr int PA_TRACE_SLOTS = 32;
#define PA_TRACE(slot)                                                                      \
	do {                                                                                    \
		if (A.trace && blockIdx.x == 0 && threadIdx.x == 0 && tstep < A.trace_steps) {      \
			unsigned long long _t;                                                          \
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                          \
			A.trace[(size_t)tstep * PA_TRACE_SLOTS + (slot)] = _t;                          \
		}                                                                                   \
	} while (0)

// the same, but not before `dep` has been produced (a load result, say)
// consumer warp 0 of CTA 0
#define PA_TRACE_C(slot, dep)                                                               \
	do {                                                                                    \
		if (A.trace && blockIdx.x == 0 && threadIdx.x == 32 && tstep < A.trace_steps) {     \
			unsigned long long _t;                                                          \
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t) : "r"((int)(dep)));        \
			A.trace[(size_t)tstep * PA_TRACE_SLOTS + (slot)] = _t;                          \
		}                                                                                   \
	} while (0)
#define PA_TRACE_DEP(slot, dep)                                                             \
	do {                                                                                    \
		if (A.trace && blockIdx.x == 0 && threadIdx.x == 0 && tstep < A.trace_steps) {      \
			unsigned long long _t;                                                          \
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t) : "r"((int)(dep)));        \
			A.trace[(size_t)tstep * PA_TRACE_SLOTS + (slot)] = _t;                          \
		}                                                                                   \
	} while (0)

// an index that would leave its array ends the run with a code (the host reports it) instead of a fault
#define PA_CHECK(cond, code)                                                     \
	do {                                                                         \
		if (!(cond)) {                                                           \
			atomicExch(A.bar + 1, (unsigned long long)(code));                   \
			return;                                                              \
		}                                                                        \
	} while (0)

template <int TB, int RB>
__global__ void __launch_bounds__(PA_THREADS, 1) phase_a_kernel(const __grid_constant__ PaArgs A) {
	using C = RowCfg<RB>;
	using T = PaTile<RB>;
	constexpr int NB = RB / TB;
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) uint64_t full_bar[PA_MAX_STAGES];
	__shared__ __align__(8) uint64_t empty_bar[PA_MAX_STAGES];
	__shared__ PaPartial warp_part[PA_MAX_CONSUMERS];
	__shared__ PaNear warp_near[PA_WARPS];
	__shared__ PaPartial s_tot;
	__shared__ long long s_base, s_lo, s_hi, s_front_row, s_back_row, s_new_center, s_seed;
	__shared__ unsigned long long s_front_pos, s_back_pos;
	__shared__ unsigned long long s_red[PA_WARPS];
	__shared__ int s_front_bin, s_back_bin, s_first_live, s_last_live, s_any_marks, s_ok, s_cta_marks, s_sink;
	__shared__ unsigned s_t_scan0;
	__shared__ unsigned long long s_wend[PA_WARPS];

	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int G = gridDim.x, cta = blockIdx.x;
	const int nb = A.nb;
	const int NCW = A.ncw, D = A.d, NS = A.ns;

	// ---- shared memory carve
	uint8_t *sp = smem;
	unsigned long long *s_bounds = reinterpret_cast<unsigned long long *>(sp); sp += (size_t)nb * 8;
	uint32_t *s_row0 = reinterpret_cast<uint32_t *>(sp); sp += (size_t)(nb + 1) * 4;
	uint32_t *s_alive = reinterpret_cast<uint32_t *>(sp); sp += (size_t)nb * 4;
	uint32_t *s_row0n = reinterpret_cast<uint32_t *>(sp); sp += (size_t)(nb + 1) * 4;   // bin starts after a compaction
	uint32_t *s_sum = reinterpret_cast<uint32_t *>(sp); sp += (size_t)NB * 4;
	uint32_t *s_marks = reinterpret_cast<uint32_t *>(sp); sp += (size_t)A.qmax * 4;
	uint32_t *s_mpref = reinterpret_cast<uint32_t *>(sp); sp += (size_t)A.qmax * 4;
	sp = reinterpret_cast<uint8_t *>(((uintptr_t)sp + 15) & ~(uintptr_t)15);
	uint8_t *s_tq = sp; sp += RB;
	sp = reinterpret_cast<uint8_t *>(((uintptr_t)sp + 127) & ~(uintptr_t)127);
	uint8_t *ring = sp;

	for (int i = threadIdx.x; i < nb; i += PA_THREADS) {
		s_bounds[i] = A.bounds[i];
		s_alive[i] = (uint32_t)(A.row0[i + 1] - A.row0[i]);
	}
	for (int i = threadIdx.x; i <= nb; i += PA_THREADS) s_row0[i] = (uint32_t)A.row0[i];
	for (int i = threadIdx.x; i < NB; i += PA_THREADS) s_sum[i] = 0;
	if (threadIdx.x == 0) {
		for (int s = 0; s < NS; s++) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		s_first_live = 0;
		s_last_live = nb - 1;
		s_any_marks = 0;
		s_ok = 1;
		s_sink = 0;
	}
	__syncthreads();

	// ---- the rows the scans stream.  They start as the caller's arrays; once at most half of them are alive
	// the alive ones are copied, in order, into a staging buffer (compact(), at a cluster boundary) and the scans
	// read that: two numberings from then on -- ORIGINAL rows (members, centers' histograms, the output) and
	// CURRENT rows (tiles, bitmap, bins, search records); cur_orig maps the second to the first.
	const uint8_t *cur_hist = A.hist;
	const McRowAux *cur_aux = A.aux;
	uint32_t *cur_bits = A.alive_bits;
	const PaRange *cur_range = A.range_tab;
	const int *cur_orig = nullptr;     // null: the numberings coincide
	long long n_cur = A.n, alive_cnt = A.n;
	int n_compact = 0;
	unsigned long long cbar = 0;       // grid barriers passed (compactions only)
	auto to_orig = [&](long long r) -> long long { return cur_orig ? (long long)__ldcg(cur_orig + r) : r; };

	auto bin_of = [&](long long row) {   // bin that holds `row`
		int lo = 0, hi = nb;
		while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long)s_row0[mid + 1] <= row) lo = mid + 1; else hi = mid; }
		return lo;
	};
	// first alive row in iteration order (bvec::pop, bvec.cpp:27-38), -1 when the bvec is empty; warp 0 only.
	// `skip`: the center of the running step (its bit may still be set, see below)
	auto pop_row = [&](long long skip) -> long long {
		int fl = s_first_live;
		while (fl < nb && s_alive[fl] == 0) fl++;
		__syncwarp();
		if (lane == 0) s_first_live = fl;
		if (fl >= nb) return -1;
		return pa_select(cur_bits, s_row0[fl], s_row0[fl + 1], 0, skip, lane);
	};

	// ---- loop state, identical in every thread of every CTA
	long long step = 0;               // scans issued (parity and tag of the records)
	long long cluster = 0, cl_begin = 0, m0 = 1;
	long long center, center_cur;     // the center as an original row (histogram, members) and as a current row (bitmap, search record)
	bool first_step = true;
	long long cl_seed = -1;   // `current` starts as {seed} (ClusterFactory.cpp:641): its histogram is added when the mean is taken
	long long st_scans = 0, st_evals = 0, st_near = 0;
	unsigned long long t_start = 0;
	if (cta == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

	// running count of tiles consumed so far by (this consumer warp | the consumer this producer lane feeds)
	// (kept modulo 2 * D: slot = cnt % D, mbarrier parity = (cnt / D) & 1)
	unsigned ring_cnt = 0;
	unsigned slot_uses = 0;   // producer lane: copies issued into its slot so far (only "none yet" and the parity matter)
	bool slot_used = false;

	// A seed leaves the bvec when it is chosen (bvec::pop / erase, bvec.cpp:27-38,281-285).  Its bit in the
	// global bitmap is cleared LATER -- behind the exchange of the cluster's first scan, when no CTA can still
	// be looking for the same seed -- and until then every reader of the bitmap masks the center's bit itself.
	// The per-bin counts (private to the CTA) are updated at once.
	if (wib == 0) {
		const long long r = pop_row(-1);   // first seed: bvec::pop (ClusterFactory.cpp:722)
		if (lane == 0) {
			s_seed = r;
			if (r >= 0) s_alive[bin_of(r)]--;
		}
	}
	__syncthreads();
	center = center_cur = s_seed;
	if (center >= 0) alive_cnt--;
	if (center >= 0 && cta == 0 && threadIdx.x == 0) { A.members[0] = (int)center; A.mcur[0] = (int)center; A.cl_off[0] = 0; }

	while (center >= 0) {
		const int par = (int)(step & 1);
		const uint32_t tag = (uint32_t)(step + 1);
		if (first_step) cl_seed = center;
		const long long tstep = step;
		unsigned long long *my_rec = A.recs + ((size_t)par * G + cta) * PA_REC_WORDS;
		unsigned long long *my_near = A.near_recs + ((size_t)par * G + cta) * PA_REC_WORDS;
		PA_TRACE(0);
		// ================= control: bvec::get_range of the center's length window =================
		PA_CHECK(center_cur >= 0 && center_cur < n_cur && center < A.n, 101);   // (every thread: the whole CTA leaves)
		if (wib < 2) {
			const PaRange rec = cur_range[center_cur];
			const bool back = wib == 1;
			int bin = back ? rec.bb : rec.fb;
			PA_TRACE_DEP(10, bin);
			unsigned long long pos = 0;
			long long row = -1;
			if (s_alive[bin] == 0) {
				if (!back) {
					int fl = s_first_live;
					while (fl < nb && s_alive[fl] == 0) fl++;
					if (fl < nb) bin = fl;   // position 0 of the first non-empty bin
				} else {
					pos = (unsigned long long)s_alive[nb - 1] - 1ull;   // wraps on an empty last bin, as in the reference
					int ll = s_last_live;
					while (ll >= 0 && s_alive[ll] == 0) ll--;
					if (ll >= 0) { bin = ll; pos = 0; }   // position 0 of the LAST non-empty bin (bvec.cpp:68-77)
				}
				if (pos < s_alive[bin]) row = pa_select(cur_bits, s_row0[bin], s_row0[bin + 1], pos, center_cur, lane);
			} else {
				pa_locate(cur_bits, s_row0[bin], s_row0[bin + 1], back ? rec.b_lt : rec.f_lt, back ? rec.b_le : rec.f_le, back,
				          center_cur, lane, pos, row);
			}
			PA_TRACE_DEP(11, row);
			if (lane == 0) {
				if (back) { s_back_bin = bin; s_back_pos = pos; s_back_row = row; }
				else { s_front_bin = bin; s_front_pos = pos; s_front_row = row; }
			}
		}
		__syncthreads();
		if (wib == 0) {
			// trip count of `for (it = front; it <= back; ++it)` = (back - front) + 1 with
			// bvec_iterator::operator- (bvec_iterator.h:61-76); <= 0 means no iteration
			int abin = s_back_bin, rbin = s_front_bin;
			unsigned long long apos = s_back_pos, rpos = s_front_pos;
			bool neg = false;
			if (abin < rbin || (abin == rbin && apos < rpos)) {
				neg = true;
				const int tb = abin; abin = rbin; rbin = tb;
				const unsigned long long tp = apos; apos = rpos; rpos = tp;
			}
			long long dlt;
			if (abin == rbin) dlt = (long long)(apos - rpos);
			else {
				long long mid = 0;
				for (int i = rbin + 1 + lane; i < abin; i += 32) mid += s_alive[i];
#pragma unroll
				for (int o = 16; o; o >>= 1) mid += __shfl_xor_sync(MC_FULL_MASK, mid, o);
				dlt = (long long)apos + (long long)((unsigned long long)s_alive[rbin] - rpos) + mid;
			}
			if (neg) dlt = -dlt;
			if (lane == 0) {
				if (dlt + 1 > 0 && s_front_row >= 0 && s_back_row >= s_front_row) { s_lo = s_front_row; s_hi = s_back_row; }
				else { s_lo = 0; s_hi = -1; }
				unsigned long long t = 0;
				if (A.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
				s_t_scan0 = (unsigned)t;
			}
		}
		__syncthreads();
		const long long lo = s_lo, hi = s_hi;
		PA_TRACE(1);

		// ================= scan: Trainer::get_close over the alive rows of [lo, hi] =================
		// tiles on an absolute grid (tile t = rows [t*RT, (t+1)*RT)); this CTA owns a contiguous run
		const long long t0 = lo / T::RT, ntiles = hi >= lo ? hi / T::RT - t0 + 1 : 0;
		// (32-bit divisions while the products fit: the 64-bit one is a long routine on the critical path)
		const bool small = ntiles < (1ll << 24);
		const long long cbeg = t0 + (small ? (long long)((unsigned)ntiles * (unsigned)cta / (unsigned)G) : ntiles * cta / G);
		const long long cend = t0 + (small ? (long long)((unsigned)ntiles * (unsigned)(cta + 1) / (unsigned)G) : ntiles * (cta + 1) / G);
		const int qc = (int)(cend - cbeg);   // <= qmax
		PaPartial mine;
		mc_scan_init(mine.s);
		mine.n_near = 0; mine.pad = 0;
		if (wib == 0) {
			// ----- producer: lane l owns ring slot l = (consumer l / D, ring position l % D)
			if (lane < NS) {
				const int w = lane / D, dpos = lane % D;
				const unsigned nw = qc > w ? (unsigned)(qc - w + NCW - 1) / (unsigned)NCW : 0u;   // tiles of consumer w in this step
				// the first tile of this step that lands in this lane's slot, then every D-th
				unsigned u = ((unsigned)dpos + (unsigned)D - ring_cnt % (unsigned)D) % (unsigned)D;
				for (; u < nw; u += (unsigned)D) {
					if (slot_used) mbar_wait_or_trap(&empty_bar[lane], (slot_uses - 1u) & 1u);
					slot_used = true;
					slot_uses++;
					const long long r0 = (cbeg + w + (long long)u * NCW) * T::RT;
					long long nr = n_cur - r0;
					if (nr > T::RT) nr = T::RT;
					uint8_t *dst = ring + (size_t)lane * T::STAGE_BYTES;
					mbar_expect_tx(&full_bar[lane], (uint32_t)(nr * RB + nr * 32));
					tma_bulk_g2s(dst, cur_hist + (size_t)r0 * RB, (uint32_t)(nr * RB), &full_bar[lane]);
					tma_bulk_g2s(dst + T::ROW_BYTES, cur_aux + r0, (uint32_t)(nr * 32), &full_bar[lane]);
				}
				ring_cnt = (ring_cnt + nw) % (2u * (unsigned)D);
			}
		} else if (wib - 1 < NCW) {
			// ----- consumers
			const int cw = wib - 1;
			const int g = lane / C::LPP, r = lane % C::LPP;
			CenterRegs<RB> cen;
			cen.load(A.hist + (size_t)center * RB, r);
			const McRowAux caux = A.aux[center];
			const uint64_t lq = caux.len, mq = caux.mag, sq = caux.sq;
			PA_TRACE_C(22, cen.w[0][0] + (unsigned)lq);
			unsigned u = 0;
			unsigned rslot = ring_cnt % (unsigned)D, rpar = (ring_cnt / (unsigned)D) & 1u;
			for (int jj = cw; jj < qc; jj += NCW, u++) {
				const long long tile = cbeg + jj;
				const long long row_mine = tile * T::RT + lane;
				const long long word = (tile * T::RT) >> 5;
				const int shift = (int)((tile * T::RT) & 31);
				const uint32_t aw = __ldcg(cur_bits + word);
				const bool have_row = lane < T::RT && row_mine >= lo && row_mine <= hi && ((aw >> (shift + lane)) & 1u) && row_mine != center_cur;
				const int slot = cw * D + (int)rslot;
				mbar_wait_or_trap(&full_bar[slot], rpar);
				if (++rslot == (unsigned)D) { rslot = 0; rpar ^= 1u; }
				if (u == 0) PA_TRACE_C(23, 0);
				const uint8_t *st = ring + (size_t)slot * T::STAGE_BYTES;
				McRowAux my_aux;
				my_aux.len = 0; my_aux.mag = 0; my_aux.sq = 0; my_aux.alive = 0; my_aux.pad = 0;
				if (have_row) my_aux = *reinterpret_cast<const McRowAux *>(st + T::ROW_BYTES + (size_t)lane * 32);
				PairAcc<TB> part[C::LPP];
#pragma unroll
				for (int it = 0; it < C::LPP; it++) {
					const int p = g * C::LPP + it;
					PairAcc<TB> acc;
					if (T::RT == 32 || p < T::RT) {
						const uint8_t *row = st + (size_t)p * RB;
#pragma unroll
						for (int c = 0; c < C::CH; c++) {
							const uint4 v = *reinterpret_cast<const uint4 *>(row + (size_t)(c * C::LPP + r) * 16);
							acc.add(v.x, cen.w[c][0]); acc.add(v.y, cen.w[c][1]);
							acc.add(v.z, cen.w[c][2]); acc.add(v.w, cen.w[c][3]);
						}
					}
					part[it] = acc;
				}
				__syncwarp();
				if (lane == 0) mbar_arrive(&empty_bar[slot]);   // the stage can be refilled during the epilogue
				const PairAcc<TB> tot = mc_transpose_reduce<C::LPP>(part, r);
				unsigned flag = 0;
				if (have_row) {
					double f0;
					const unsigned dec = mc_scan_decide<TB>(A.model, tot.summin(my_aux.mag, mq), tot.dot(), my_aux.len, my_aux.mag, my_aux.sq, lq, mq, sq, NB, 1.0 / NB, f0);
					flag = dec & 1u;
					mine.s.n_eval++;
					mine.s.n_pos += flag;
					mine.n_near += dec >> 1;
					if (f0 > mine.s.best_f0) { mine.s.best_f0 = f0; mine.s.best_row = row_mine; }
				}
				const unsigned mask = __ballot_sync(MC_FULL_MASK, flag != 0);
				if (u == 0) PA_TRACE_C(24, mask);
				if (lane == 0) s_marks[jj] = mask;
				if (mask) {
					// bvec::remove_available (bvec.cpp:290-317): the rows leave the bvec -- their bits are
					// cleared in the tail, behind the exchange: a slower CTA may still be reading the bitmap for
					// the range of THIS step -- and join `current`: their histograms go into this CTA's bin sums
					if (lane == 0) s_any_marks = 1;
					for (unsigned mm = mask; mm; mm &= mm - 1) {
						const long long mrow = tile * T::RT + (__ffs(mm) - 1);
						const uint32_t *src = reinterpret_cast<const uint32_t *>(cur_hist + (size_t)mrow * RB);
						for (int wd = lane; wd < RB / 4; wd += 32) {
							const uint32_t v = __ldg(src + wd);
							if (TB == 1) {
								atomicAdd(&s_sum[wd * 4 + 0], v & 0xffu); atomicAdd(&s_sum[wd * 4 + 1], (v >> 8) & 0xffu);
								atomicAdd(&s_sum[wd * 4 + 2], (v >> 16) & 0xffu); atomicAdd(&s_sum[wd * 4 + 3], v >> 24);
							} else {
								atomicAdd(&s_sum[wd * 2 + 0], v & 0xffffu); atomicAdd(&s_sum[wd * 2 + 1], v >> 16);
							}
						}
					}
				}
			}
			ring_cnt = (ring_cnt + u) % (2u * (unsigned)D);
			PA_TRACE_C(25, u);
			if (A.trace && cta == 0 && lane == 0) {
				unsigned long long t;
				asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
				s_wend[wib] = t;
			}
			pa_scan_fold_redux(mine.s);
			mine.n_near = __reduce_add_sync(MC_FULL_MASK, (unsigned)mine.n_near);
			if (lane == 0) warp_part[cw] = mine;
		}
		__syncthreads();
		PA_TRACE(2);
		if (A.trace && cta == 0 && threadIdx.x == 0 && tstep < A.trace_steps) {
			unsigned long long m = 0;
			for (int w = 1; w <= NCW; w++) m = s_wend[w] > m ? s_wend[w] : m;
			A.trace[(size_t)tstep * PA_TRACE_SLOTS + 21] = m;
		}
		// ----- the bin sums of this CTA's marked rows join the cluster's global sums
		const int gbuf = (int)(cluster % 3);
		{
			if (s_any_marks) {   // uniform in the CTA
				unsigned long long *gs = A.g_sum + (size_t)gbuf * pa_gsum_words(NB);
				unsigned long long ret = 0;
				for (int b = threadIdx.x; b < NB; b += PA_THREADS) {
					const unsigned long long v = s_sum[b];
					if (v) ret |= atomicAdd(gs + pa_gsum_idx(b), v);
					s_sum[b] = 0;
				}
				pa_retire(ret, &s_sink);
				__syncthreads();
			}
		}
		PA_TRACE(14);
		// ================= exchange 1: this CTA's summary out, everybody's summaries in =================
		if (wib == 0) {
			PaPartial b;
			mc_scan_init(b.s);
			b.n_near = 0; b.pad = 0;
			if (lane < NCW) b = warp_part[lane];
			pa_scan_fold_redux(b.s);
			b.n_near = __reduce_add_sync(MC_FULL_MASK, (unsigned)b.n_near);
			unsigned long long tnow = 0;
			if (A.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
			pa_release_fence();
			PA_TRACE(4);
			{
				const unsigned long long f0b = (unsigned long long)__double_as_longlong(b.s.best_f0);
				uint32_t wv = 0;
				switch (lane) {
				case 0: wv = (uint32_t)b.s.n_pos; break;
				case 1: wv = (uint32_t)b.s.n_eval; break;
				case 2: wv = (uint32_t)(int)b.s.best_row; break;
				case 3: wv = (uint32_t)f0b; break;
				case 4: wv = (uint32_t)(f0b >> 32); break;
				case 5: wv = (uint32_t)b.n_near; break;
				case 6: wv = s_t_scan0; break;
				case 7: wv = (uint32_t)tnow; break;
				default: break;
				}
				if (lane < PA_REC_WORDS) pa_ll_store(my_rec + lane, wv, tag);
			}
			PA_TRACE(12);
			// every CTA reads the records of all CTAs (lane l: CTAs l, l + 32, ...): first one try at each of them
			// with all loads in flight, then the late ones one by one
			uint32_t d[PA_RPL][PA_REC_WORDS];
			unsigned have = 0;
#pragma unroll
			for (int j = 0; j < PA_RPL; j++) {
				const int i = lane + 32 * j;
				if (i >= G) {
					have |= 1u << j;
					d[j][0] = 0; d[j][1] = 0; d[j][2] = 0xffffffffu; d[j][3] = 0; d[j][4] = 0xbff00000u; d[j][5] = 0;   // nothing, arg-max (-1, none)
					d[j][6] = s_t_scan0; d[j][7] = (uint32_t)tnow;
				} else if (pa_rec_try<PA_REC_WORDS>(A.recs + ((size_t)par * G + i) * PA_REC_WORDS, tag, d[j])) have |= 1u << j;
			}
			bool ok = true;
#pragma unroll
			for (int j = 0; j < PA_RPL; j++)
				if (ok && !((have >> j) & 1u)) ok = pa_rec_wait<PA_REC_WORDS>(A.recs + ((size_t)par * G + lane + 32 * j) * PA_REC_WORDS, tag, d[j], A.bar + 1);
			PaPartial t;
			mc_scan_init(t.s);
			t.n_near = 0; t.pad = 0;
			unsigned base = 0, late0 = 0, late1 = 0;
#pragma unroll
			for (int j = 0; j < PA_RPL; j++) {
				mc_scan_result q;
				q.n_pos = d[j][0]; q.n_eval = d[j][1]; q.best_row = (long long)(int)d[j][2];
				q.best_f0 = __longlong_as_double((long long)(((unsigned long long)d[j][4] << 32) | d[j][3]));
				mc_scan_merge(t.s, q);
				t.n_near += d[j][5];
				if (lane + 32 * j < cta) base += d[j][0];
				const unsigned l0 = d[j][6] - s_t_scan0, l1 = d[j][7] - (unsigned)tnow;
				if ((int)l0 > (int)late0) late0 = l0;
				if ((int)l1 > (int)late1) late1 = l1;
			}
			ok = __all_sync(MC_FULL_MASK, ok);
			pa_scan_fold_redux(t.s);
			t.n_near = __reduce_add_sync(MC_FULL_MASK, (unsigned)t.n_near);
			base = __reduce_add_sync(MC_FULL_MASK, base);
			if (A.trace && cta == 0 && tstep < A.trace_steps) {
				late0 = __reduce_max_sync(MC_FULL_MASK, late0);
				late1 = __reduce_max_sync(MC_FULL_MASK, late1);
				if (lane == 0) {
					A.trace[(size_t)tstep * PA_TRACE_SLOTS + 8] = late0;
					A.trace[(size_t)tstep * PA_TRACE_SLOTS + 9] = late1;
				}
			}
			if (lane == 0) { s_tot = t; s_base = base; s_ok = ok ? 1 : 0; s_any_marks = 0; }
		}
		__syncthreads();
		if (!s_ok) return;
		PA_TRACE(3);
		const mc_scan_result tot = s_tot.s;
		if (hi >= lo) { st_scans++; st_evals += tot.n_eval; st_near += s_tot.n_near; }
		// no CTA is looking for this cluster's seed any more: its bit goes (returning form: performed before
		// the next block-wide barrier lets any thread of this CTA read the word again without the mask)
		if (first_step && wib == PA_WARPS - 1 && lane == 0) {
			const unsigned old = atomicAnd(cur_bits + (center_cur >> 5), ~(1u << (center_cur & 31)));
			unsigned z;
			asm volatile("and.b32 %0, %1, 0;" : "=r"(z) : "r"(old));
			if (z) s_sink = 1;   // never true: waits for the atomic
		}
		const bool was_first = first_step;
		step++;
		first_step = false;

		if (tot.n_pos == 0) {
			// ----- is_min: no close point left (ClusterFactory.cpp:693-711): the cluster is closed, the
			// arg-max of f0 becomes the next seed, or the first point of the bvec when there is none
			if (wib == 0) {
				long long r = tot.best_row;
				if (r < 0) r = pop_row(was_first ? center_cur : -1);
				if (lane == 0) {
					s_seed = r;
					if (r >= 0) s_alive[bin_of(r)]--;
				}
			}
			if (cta == 0) {
				if (threadIdx.x == 32) {
					A.cl_center[cluster] = (int)center;
					A.cl_off[cluster + 1] = (int)(cl_begin + m0);
				}
				// the sums buffer of the cluster after the next one (last read two clusters ago)
				unsigned long long *gz = A.g_sum + (size_t)((cluster + 2) % 3) * pa_gsum_words(NB);
				unsigned long long ret = 0;
				for (int b = threadIdx.x; b < NB; b += PA_THREADS) {
#if PA_RELEASE_BY_FENCE
					gz[pa_gsum_idx(b)] = 0;
#else
					ret |= atomicExch(gz + pa_gsum_idx(b), 0ull);
#endif
				}
				pa_retire(ret, &s_sink);   // performed before this CTA's next record (block barrier below)
			}
			__syncthreads();
			const long long seed = s_seed;   // a current row
			cluster++;
			cl_begin += m0;
			m0 = 1;
			first_step = true;
			center_cur = seed;
			PA_CHECK(seed < n_cur, 107);
			center = seed >= 0 ? to_orig(seed) : -1;
			PA_CHECK(center < A.n, 108);
			if (seed >= 0) alive_cnt--;
			// ---------- row compaction (ClusterFactory.cpp has nothing like it: the bvec forgets removed points, the row
			// arrays here do not).  At a cluster boundary, once at most 1 - 2^-shift of the current rows are still set in
			// the bitmap, the set ones are copied in order into the next staging buffer and renumbered.  The seed's bit is
			// still set (its clearing is pending), so it travels like any other row: nothing of the protocol changes.
			if (seed >= 0 && A.st_hist[0] && n_cur >= A.compact_min && alive_cnt + 1 <= n_cur - (n_cur >> A.compact_shift) &&
			    alive_cnt + 1 <= A.st_cap[n_compact & 1]) {
				const int sb = n_compact & 1;
				uint8_t *nh = A.st_hist[sb];
				McRowAux *na = A.st_aux[sb];
				int *no = A.st_orig[sb];
				PaRange *nr = A.st_range[sb];
				uint32_t *nbits = A.st_bits[sb];
				const int seed_bin = bin_of(seed);
				// new bin starts: the alive rows of every bin, plus the seed in its own
				if (wib == 0) {
					unsigned carry = 0;
					for (int i0 = 0; i0 <= nb; i0 += 32) {
						const int i = i0 + lane;
						const unsigned v = i < nb ? s_alive[i] + (i == seed_bin ? 1u : 0u) : 0u;
						unsigned incl = v;
#pragma unroll
						for (int o = 1; o < 32; o <<= 1) {
							const unsigned u = __shfl_up_sync(MC_FULL_MASK, incl, o);
							if (lane >= o) incl += u;
						}
						if (i <= nb) s_row0n[i] = carry + incl - v;
						carry += __shfl_sync(MC_FULL_MASK, incl, 31);
					}
				}
				__syncthreads();
				const long long n_new_rows = s_row0n[nb];
				// this CTA's share of the OLD rows: whole bitmap words
				const long long W = (n_cur + 31) >> 5;
				const long long W0 = W * cta / G, W1 = W * (cta + 1) / G;
				const int nwords = (int)(W1 - W0);   // <= qmax (a tile is at most 32 rows)
				if (wib == 0) {
					// rows set before this share: whole bins from the new starts, the rest of the first bin by popcount
					unsigned long long rank0 = 0;
					if (nwords > 0) {
						const long long r0 = W0 << 5;
						const int b0 = bin_of(r0 < n_cur ? r0 : n_cur - 1);
						unsigned ra, rb, tot = 0;
						if (r0 > (long long)s_row0[b0]) pa_rank3(cur_bits, s_row0[b0], r0, 0, 0, -1, lane, ra, rb, tot);
						rank0 = (unsigned long long)s_row0n[b0] + tot;
					}
					unsigned carry = 0;
					for (int i0 = 0; i0 < nwords; i0 += 32) {
						const int i = i0 + lane;
						uint32_t v = 0;
						if (i < nwords) {
							v = __ldcg(cur_bits + W0 + i);
							const long long rbase = (W0 + i) << 5;
							if (rbase + 32 > n_cur) v &= (n_cur > rbase) ? (0xffffffffu >> (32 - (int)(n_cur - rbase))) : 0u;
							s_marks[i] = v;
						}
						const unsigned pc = __popc(v);
						unsigned incl = pc;
#pragma unroll
						for (int o = 1; o < 32; o <<= 1) {
							const unsigned u = __shfl_up_sync(MC_FULL_MASK, incl, o);
							if (lane >= o) incl += u;
						}
						if (i < nwords) s_mpref[i] = carry + incl - pc;
						carry += __shfl_sync(MC_FULL_MASK, incl, 31);
					}
					if (lane == 0) s_base = (long long)rank0;
				}
				__syncthreads();
				const long long rank0 = s_base;
				// copy: one warp per bitmap word; the 16-byte units of the word's set rows as one flat index space, eight
				// loads per lane in flight (lane r knows where the r-th set row is)
				for (int i = wib; i < nwords; i += PA_WARPS) {
					const uint32_t mm = s_marks[i];
					const int cnt = __popc(mm);
					if (cnt == 0) continue;
					const long long dst0 = rank0 + s_mpref[i], src0 = (W0 + i) << 5;
					PA_CHECK(src0 + (31 - __clz(mm)) < n_cur && dst0 + cnt <= A.st_cap[sb] && dst0 + cnt <= n_new_rows, 102);
					const int pos = lane < cnt ? (int)__fns(mm, 0, lane + 1) : 0;
					constexpr int U = RB / 16;   // a power of two
					const int total = cnt * U;
					for (int base = 0; base < total; base += 32 * 8) {
						uint4 v[8];
#pragma unroll
						for (int j = 0; j < 8; j++) {
							const int idx = base + j * 32 + lane;
							const int r = idx / U < cnt ? idx / U : cnt - 1;
							const int srow = __shfl_sync(MC_FULL_MASK, pos, r);
							if (idx < total) v[j] = __ldcg(reinterpret_cast<const uint4 *>(cur_hist + (size_t)(src0 + srow) * RB) + (idx % U));
						}
#pragma unroll
						for (int j = 0; j < 8; j++) {
							const int idx = base + j * 32 + lane;
							if (idx < total) reinterpret_cast<uint4 *>(nh + (size_t)(dst0 + idx / U) * RB)[idx % U] = v[j];
						}
					}
					if (lane < cnt) {
						const uint4 *pa = reinterpret_cast<const uint4 *>(cur_aux + src0 + pos);
						const uint4 a0 = __ldcg(pa), a1 = __ldcg(pa + 1);
						uint4 *qa = reinterpret_cast<uint4 *>(na + dst0 + lane);
						qa[0] = a0; qa[1] = a1;
						no[dst0 + lane] = (int)to_orig(src0 + pos);
					}
				}
				// the new bitmap: every row set
				{
					const long long Wn = (n_new_rows + 31) >> 5;
					for (long long w = Wn * cta / G + threadIdx.x; w < Wn * (cta + 1) / G; w += PA_THREADS) {
						const long long left = n_new_rows - (w << 5);
						nbits[w] = left >= 32 ? 0xffffffffu : (0xffffffffu >> (32 - (int)left));
					}
				}
				// where the seed went: the rows set before it in its bin
				if (wib == 1) {
					unsigned ra, rb, tot = 0;
					if (seed > (long long)s_row0[seed_bin]) pa_rank3(cur_bits, s_row0[seed_bin], seed, 0, 0, -1, lane, ra, rb, tot);
					if (lane == 0) s_new_center = (long long)s_row0n[seed_bin] + tot;
				}
				asm volatile("fence.proxy.async;" ::: "memory");   // (writer side: generic stores -> bulk copies of other SMs)
				__threadfence();
				if (!pa_grid_barrier(A.bar, ++cbar * (unsigned long long)G, &s_ok)) return;
				// the search records of the new rows, over all threads of the grid
				for (long long j = (long long)cta * PA_THREADS + threadIdx.x; j < n_new_rows; j += (long long)G * PA_THREADS)
					nr[j] = pa_make_range(na, j, s_bounds, s_row0n, nb, A.sim);
				__threadfence();
				if (!pa_grid_barrier(A.bar, ++cbar * (unsigned long long)G, &s_ok)) return;
				asm volatile("fence.proxy.async;" ::: "memory");   // the bulk copies of the next scan read what generic stores wrote
				for (int i = threadIdx.x; i <= nb; i += PA_THREADS) s_row0[i] = s_row0n[i];
				cur_hist = nh; cur_aux = na; cur_orig = no; cur_range = nr; cur_bits = nbits;
				n_cur = n_new_rows;
				center_cur = s_new_center;
				PA_CHECK(center_cur >= 0 && center_cur < n_cur, 103);
				n_compact++;
				__syncthreads();
			}
			if (seed >= 0 && cta == 0 && threadIdx.x == 0) { A.members[cl_begin] = (int)center; A.mcur[cl_begin] = (int)center_cur; }
			PA_TRACE(7);
			continue;
		}

		// ================= tail: `current` += marked rows, get_mean, nearest member =================
		const long long n_new = tot.n_pos, m_all = m0 + n_new;
		if (wib == 0) {
			// exclusive prefix of the marks per tile of this CTA's run
			unsigned carry = 0;
			for (int i0 = 0; i0 < qc; i0 += 32) {
				const int i = i0 + lane;
				const unsigned pc = i < qc ? __popc(s_marks[i]) : 0u;
				unsigned incl = pc;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const unsigned v = __shfl_up_sync(MC_FULL_MASK, incl, o);
					if (lane >= o) incl += v;
				}
				if (i < qc) s_mpref[i] = carry + incl - pc;
				carry += __shfl_sync(MC_FULL_MASK, incl, 31);
			}
			if (lane == 0) s_cta_marks = (int)carry;
		}
		{
			// truncated mean: floor(sum / |current|) per bin and its magnitude (DivergencePoint.cpp:53-65,155-173)
			const unsigned long long *gs = A.g_sum + (size_t)gbuf * pa_gsum_words(NB);
			unsigned long long local = 0;
			const uint8_t *seed_hist = A.hist + (size_t)cl_seed * RB;
			for (int b = threadIdx.x; b < NB; b += PA_THREADS) {
				const unsigned long long sb = TB == 1 ? (unsigned long long)seed_hist[b] : (unsigned long long)reinterpret_cast<const uint16_t *>(seed_hist)[b];
				const unsigned long long tot_b = __ldcg(gs + pa_gsum_idx(b)) + sb;
				// 32-bit division whenever the operands allow (the 64-bit one is a ~100-instruction routine)
				const unsigned long long v = ((tot_b | (unsigned long long)m_all) >> 32) ? tot_b / (unsigned long long)m_all
				                                                                       : (unsigned long long)((uint32_t)tot_b / (uint32_t)m_all);
				if (TB == 1) s_tq[b] = (uint8_t)v; else reinterpret_cast<uint16_t *>(s_tq)[b] = (uint16_t)v;
				local += v;
			}
#pragma unroll
			for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(MC_FULL_MASK, local, o);
			if (lane == 0) s_red[wib] = local;
		}
		__syncthreads();
		PA_TRACE(16);
		unsigned long long magc = 0;
#pragma unroll
		for (int w = 0; w < PA_WARPS; w++) magc += s_red[w];
		PaNear best;
		best.pos = -1; best.dist = 0; best.row = -1; best.pad = 0;
		// a candidate: its histogram row (either numbering's array holds the same bytes), its position in `current`,
		// and its row in both numberings (the winner becomes the center)
		auto consider = [&](const uint8_t *hrow, uint64_t mp, long long row_orig, long long row_cur, long long pos) {
			const PairAcc<TB> pa = mc_warp_pair_reduce<TB>(hrow, s_tq, RB, lane);
			PaNear cnd;
			cnd.pos = pos; cnd.row = row_orig; cnd.pad = row_cur;
			cnd.dist = mc_distance_d(pa.summin(mp, magc), mp, magc);
			pa_near_merge(best, cnd);   // NaN never replaces, like the reference's `<`
		};
		// this CTA's new members, in row order behind the ones of the CTAs before it; the k-th one goes to warp
		// k mod PA_WARPS (rows of one length are neighbours: a tile can hold many of them)
		unsigned long long ret = 0;
		for (int i = threadIdx.x; i < qc; i += PA_THREADS) {
			const unsigned mask = s_marks[i];
			if (mask) {   // the rows leave the bvec (visible to every CTA behind exchange 2)
				const long long trow = (cbeg + i) * T::RT;
				ret |= atomicAnd(cur_bits + (trow >> 5), ~(mask << (int)(trow & 31)));
			}
		}
		for (int k = wib; k < s_cta_marks; k += PA_WARPS) {
			int tl = 0, th = qc - 1;   // last tile whose prefix is <= k
			while (tl < th) { const int mid = (tl + th + 1) >> 1; if ((int)s_mpref[mid] <= k) tl = mid; else th = mid - 1; }
			const long long row = (cbeg + tl) * T::RT + __fns(s_marks[tl], 0, k - (int)s_mpref[tl] + 1);   // a current row
			const long long pos = m0 + s_base + k;
			PA_CHECK(row >= 0 && row < n_cur, 104);
			const long long row_orig = to_orig(row);
			PA_CHECK(row_orig >= 0 && row_orig < A.n, 105);
			if (lane == 0) {
#if PA_RELEASE_BY_FENCE
				A.members[cl_begin + pos] = (int)row_orig;
				A.mcur[cl_begin + pos] = (int)row;
#else
				ret |= (unsigned)atomicExch(A.members + cl_begin + pos, (int)row_orig);
				ret |= (unsigned)atomicExch(A.mcur + cl_begin + pos, (int)row);
#endif
			}
			consider(cur_hist + (size_t)row * RB, cur_aux[row].mag, row_orig, row, pos);
		}
		// its share of the members `current` already had
		for (long long idx = (long long)cta * PA_WARPS + wib; idx < m0; idx += (long long)G * PA_WARPS) {
			const long long ro = __ldcg(A.members + cl_begin + idx), rc = __ldcg(A.mcur + cl_begin + idx);
			PA_CHECK(ro >= 0 && ro < A.n && rc >= 0 && rc < n_cur, 106);
			consider(A.hist + (size_t)ro * RB, A.aux[ro].mag, ro, rc, idx);
		}
		if (lane == 0) warp_near[wib] = best;
		pa_retire(ret, &s_sink);
		PA_TRACE_DEP(17, best.row);
		__syncthreads();
		PA_TRACE(18);
		// ================= exchange 2: nearest-member candidates =================
		if (wib == 0) {
			PaNear b;
			b.pos = -1; b.dist = 0; b.row = -1; b.pad = 0;
			if (lane < PA_WARPS) b = warp_near[lane];
			auto fold_near = [&](PaNear &x) {
#pragma unroll
				for (int o = 16; o; o >>= 1) {
					PaNear q;
					q.pos = __shfl_xor_sync(MC_FULL_MASK, x.pos, o);
					q.dist = __shfl_xor_sync(MC_FULL_MASK, x.dist, o);
					q.row = __shfl_xor_sync(MC_FULL_MASK, x.row, o);
					q.pad = __shfl_xor_sync(MC_FULL_MASK, x.pad, o);
					pa_near_merge(x, q);
				}
			};
			fold_near(b);
			pa_release_fence();   // member rows and cleared bits of the whole CTA precede the record
			PA_TRACE(19);
			{
				const unsigned long long db = (unsigned long long)__double_as_longlong(b.dist);
				uint32_t wv = 0;
				switch (lane) {
				case 0: wv = (uint32_t)(int)b.pos; break;
				case 1: wv = (uint32_t)db; break;
				case 2: wv = (uint32_t)(db >> 32); break;
				case 3: wv = (uint32_t)(int)b.row; break;
				case 4: wv = (uint32_t)(int)b.pad; break;   // the same row in the current numbering
				default: break;
				}
				if (lane < 6) pa_ll_store(my_near + lane, wv, tag);
			}
			PA_TRACE(5);
			uint32_t d[PA_RPL][6];
			unsigned have = 0;
#pragma unroll
			for (int j = 0; j < PA_RPL; j++) {
				const int i = lane + 32 * j;
				if (i >= G) { have |= 1u << j; d[j][0] = 0xffffffffu; d[j][1] = 0; d[j][2] = 0; d[j][3] = 0xffffffffu; d[j][4] = 0xffffffffu; d[j][5] = 0; }
				else if (pa_rec_try<6>(A.near_recs + ((size_t)par * G + i) * PA_REC_WORDS, tag, d[j])) have |= 1u << j;
			}
			bool ok = true;
#pragma unroll
			for (int j = 0; j < PA_RPL; j++)
				if (ok && !((have >> j) & 1u)) ok = pa_rec_wait<6>(A.near_recs + ((size_t)par * G + lane + 32 * j) * PA_REC_WORDS, tag, d[j], A.bar + 1);
			PaNear t;
			t.pos = -1; t.dist = 0; t.row = -1; t.pad = 0;
#pragma unroll
			for (int j = 0; j < PA_RPL; j++) {
				PaNear q;
				q.pos = (long long)(int)d[j][0];
				q.dist = __longlong_as_double((long long)(((unsigned long long)d[j][2] << 32) | d[j][1]));
				q.row = (long long)(int)d[j][3];
				q.pad = (long long)(int)d[j][4];
				pa_near_merge(t, q);
			}
			ok = __all_sync(MC_FULL_MASK, ok);
			fold_near(t);
			if (lane == 0) { s_new_center = t.row; s_seed = t.pad; s_ok = ok ? 1 : 0; }
		}
		__syncthreads();
		if (!s_ok) return;
		PA_TRACE(6);
		// every CTA keeps its own bin counts: the new members leave their bins
		for (long long i = threadIdx.x; i < n_new; i += PA_THREADS)
			atomicSub(&s_alive[bin_of(__ldcg(A.mcur + cl_begin + m0 + i))], 1u);
		center = s_new_center;   // get_mean always finds a member: `current` is never empty
		center_cur = s_seed;
		alive_cnt -= n_new;
		m0 = m_all;
		__syncthreads();
		PA_TRACE(7);
	}

	if (cta == 0 && threadIdx.x == 0) {
		unsigned long long t_end;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
		A.stats[0] = cluster;
		A.stats[1] = st_scans;
		A.stats[2] = st_evals;
		A.stats[3] = st_near;
		A.stats[4] = step;
		A.stats[5] = (long long)(t_end - t_start);
		A.stats[6] = n_compact;
	}
}

// ---- launcher ----------------------------------------------------------------------------------
struct PaLaunch {
	PaArgs args;
	int grid;
	size_t smem;
};

template <int TB, int RB>
static int pa_launch_t(mc_ctx *ctx, PaArgs &a, int grid) {
	using T = PaTile<RB>;
	constexpr int NB = RB / TB;
	int dev_smem = 0;
	MC_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
	const size_t fixed = (size_t)a.nb * 8 + 2 * (size_t)(a.nb + 1) * 4 + (size_t)a.nb * 4 + (size_t)NB * 4 + (size_t)a.qmax * 8 + 16 + RB + 128 + 128;
	const size_t static_smem = 4096;   // barriers, warp partials, control words (upper bound)
	MC_REQUIRE(fixed + static_smem + 2 * (size_t)T::STAGE_BYTES <= (size_t)dev_smem, MC_ERR_UNSUPPORTED,
	           "mc_accumulate_run: %d bvec bins / %lld rows need more shared memory than one CTA has", a.nb, (long long)a.n);
	int ns = (int)(((size_t)dev_smem - static_smem - fixed) / T::STAGE_BYTES);
	if (ns > PA_MAX_STAGES) ns = PA_MAX_STAGES;
	a.ncw = ns < PA_MAX_CONSUMERS ? ns : PA_MAX_CONSUMERS;
	a.d = ns / a.ncw;
	a.ns = a.ncw * a.d;
	const size_t smem = fixed + (size_t)a.ns * T::STAGE_BYTES;
	const void *fn = (const void *)phase_a_kernel<TB, RB>;
	MC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	MC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PA_THREADS, smem));
	MC_REQUIRE(per_sm >= 1, MC_ERR_UNSUPPORTED, "mc_accumulate_run: the persistent kernel does not fit an SM (%zu bytes of shared memory)", smem);
	void *params[] = {&a};
	MC_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(PA_THREADS), params, smem, ctx->stream));
	ctx->launches++;
	return MC_OK;
}

int mc_pa_rows_per_tile(int tbytes, int nbins) {
	const int rb = tbytes * nbins;
	return rb * 32 <= 32 * 1024 ? 32 : (32 * 1024) / rb;
}

size_t mc_pa_exchange_bytes(int grid) { return (size_t)4 * grid * PA_REC_WORDS * 8; }
int mc_pa_max_grid() { return 32 * PA_RPL; }
int mc_pa_trace_slots() { return PA_TRACE_SLOTS; }
size_t mc_pa_gsum_bytes(int nbins) { return 3 * pa_gsum_words(nbins) * 8; }
size_t mc_pa_range_bytes() { return sizeof(PaRange); }

int mc_launch_pa_prepare(mc_ctx *ctx, const unsigned long long *bounds_dev, const int *row0_dev, int nb, double sim,
                         void *range_tab_dev, unsigned int *err_dev) {
	int64_t blocks = (ctx->n + 255) / 256;
	if (blocks > (int64_t)ctx->num_sms * 8) blocks = (int64_t)ctx->num_sms * 8;
	pa_prepare_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(ctx->d_aux, ctx->n, bounds_dev, row0_dev, nb, sim, (PaRange *)range_tab_dev, err_dev);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

// shapes of the staged scan: rows of 16 bytes .. 4 KB
bool mc_pa_shape_supported(int tbytes, int nbins) {
	const int rb = tbytes * nbins;
	if (tbytes == 1) return rb == 16 || rb == 64 || rb == 256 || rb == 1024 || rb == 4096;
	return rb == 32 || rb == 128 || rb == 512 || rb == 2048;
}

int mc_launch_phase_a(mc_ctx *ctx, const unsigned long long *bounds_dev, const int *row0_dev, int nb, const void *range_tab_dev,
                      uint32_t *alive_bits_dev, unsigned long long *g_sum_dev, void *exch_dev,
                      unsigned long long *bar_dev, int *members_dev, int *cl_center_dev, int *cl_off_dev, long long *stats_dev,
                      unsigned long long *trace_dev, int trace_steps, int grid, int qmax, int *mcur_dev, double sim,
                      void *const *staging /* [2][5]: hist, aux, orig, range, bits; null = no compaction */, const long long *staging_cap,
                      long long compact_min, int compact_shift) {
	PaArgs a{};
	a.hist = (const uint8_t *)ctx->d_hist;
	a.aux = ctx->d_aux;
	a.n = ctx->n;
	a.bounds = bounds_dev;
	a.row0 = row0_dev;
	a.nb = nb;
	a.qmax = qmax;
	a.range_tab = (const PaRange *)range_tab_dev;
	a.alive_bits = alive_bits_dev;
	a.g_sum = g_sum_dev;
	{
		unsigned long long *x = (unsigned long long *)exch_dev;
		a.recs = x; x += 2 * (size_t)grid * PA_REC_WORDS;
		a.near_recs = x;
	}
	a.bar = bar_dev;
	a.members = members_dev;
	a.mcur = mcur_dev;
	a.sim = sim;
	a.compact_min = compact_min;
	a.compact_shift = compact_shift;
	for (int b = 0; b < 2; b++) {
		a.st_hist[b] = staging ? (uint8_t *)staging[b * 5 + 0] : nullptr;
		a.st_aux[b] = staging ? (McRowAux *)staging[b * 5 + 1] : nullptr;
		a.st_orig[b] = staging ? (int *)staging[b * 5 + 2] : nullptr;
		a.st_range[b] = staging ? (PaRange *)staging[b * 5 + 3] : nullptr;
		a.st_bits[b] = staging ? (uint32_t *)staging[b * 5 + 4] : nullptr;
		a.st_cap[b] = staging ? staging_cap[b] : 0;
	}
	a.cl_center = cl_center_dev;
	a.cl_off = cl_off_dev;
	a.stats = stats_dev;
	a.trace = trace_dev;
	a.trace_steps = trace_steps;
	a.model = ctx->model;
	const int rb = ctx->tbytes * ctx->nbins;
	if (ctx->tbytes == 1) {
		switch (rb) {
		case 16: return pa_launch_t<1, 16>(ctx, a, grid);
		case 64: return pa_launch_t<1, 64>(ctx, a, grid);
		case 256: return pa_launch_t<1, 256>(ctx, a, grid);
		case 1024: return pa_launch_t<1, 1024>(ctx, a, grid);
		case 4096: return pa_launch_t<1, 4096>(ctx, a, grid);
		default: break;
		}
	} else {
		switch (rb) {
		case 32: return pa_launch_t<2, 32>(ctx, a, grid);
		case 128: return pa_launch_t<2, 128>(ctx, a, grid);
		case 512: return pa_launch_t<2, 512>(ctx, a, grid);
		case 2048: return pa_launch_t<2, 2048>(ctx, a, grid);
		default: break;
		}
	}
	mc_set_error("mc_accumulate_run: histogram rows of %d bytes are not supported by the persistent kernel", rb);
	return MC_ERR_UNSUPPORTED;
}
