// Host control flow of bin/meshclust.  Mirrors, step by step, what the reference does between its
// hot loops (citations inline); every hot loop is a call into the C-ABI of include/meshclust_b200.h.
// Compiled with -ffp-contract=off; see matrix.hpp for the two places the reference binary fuses.
#include "pipeline.hpp"

#include <libgen.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <future>
#include <thread>
#include <unordered_map>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "bvec.hpp"
#include "fasta.hpp"
#include "lazy_sort.hpp"
#include "matrix.hpp"
#include "meshclust_b200.h"

namespace mch {

namespace {

struct Timer {
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	double lap() {
		auto t1 = std::chrono::steady_clock::now();
		double s = std::chrono::duration<double>(t1 - t0).count();
		t0 = t1;
		return s;
	}
};

[[noreturn]] void die_gpu(const char *what) {
	fprintf(stderr, "meshclust: %s failed: %s\n", what, mc_last_error());
	exit(2);
}
#define GPU(call)                          \
	do {                                   \
		if ((call) != MC_OK) die_gpu(#call); \
	} while (0)

// ----------------------------------------------------------------------------------------------
struct Dataset {
	FastaBatch fa;                    // file order = point id order (Runner.cpp:345-349)
	int64_t n = 0;
	std::vector<uint64_t> len;        // by id
	std::vector<int64_t> row_of_id;   // GPU row (= initial bvec iteration order) of each id
	std::vector<int64_t> id_of_row;
};

struct Model {
	int nfeat = 0;                    // 3 or 4 combos
	double mins[5] = {0, 0, 0, 0, 0}, maxs[5] = {1, 1, 1, 1, 1};
	double w[5] = {0, 0, 0, 0, 0};
	bool align = false;               // --align: single FEAT_ALIGN feature, weights [-id, 1]
	double cutoff = 0;
};

struct Ctx {
	mc_ctx *gpu = nullptr;            // rank 0: holds the sequences, runs training, Phase B and the Phase-A tail
	std::vector<mc_ctx *> ranks;      // all GPUs that share the Phase-A scans (ranks[0] == gpu)
	std::future<std::vector<int>> by_length;   // Trainer::split's first sort, started as soon as the lengths are known
	std::thread ranks_thread;         // creates the contexts of ranks 1.. in the background
	int ranks_rc = MC_OK;
	std::string ranks_err;
	bool shard_phase_a = false;       // the Phase-A scans are shared by all ranks (inputs of gigabytes)
	bool seq_cloned = false;          // ranks 1.. hold copies of the sequences (alignments are split by pairs)
	long align_split_calls = 0;       // alignment batches that were split over the GPUs
	Options opt;
	Dataset ds;
	Model model;
	int k = 0;
	int tbytes = 1;
};

// ----------------------------------------------------------------------------------------------
// Trainer (Trainer.cpp)
// ----------------------------------------------------------------------------------------------
using Pair = std::pair<int, int>;   // point ids (first has the smaller header, Trainer.cpp:746,755)

struct HeaderPairLess {
	const std::vector<std::string> *h;
	bool operator()(const Pair &a, const Pair &b) const {
		const int c = (*h)[a.first].compare((*h)[b.first]);
		if (c < 0) return true;
		return (*h)[a.first] == (*h)[b.first] && (*h)[a.second].compare((*h)[b.second]) < 0;
	}
};
struct ScoredLess {
	HeaderPairLess base;
	bool operator()(const std::pair<Pair, double> &a, const std::pair<Pair, double> &b) const { return base(a.first, b.first); }
};

// the contexts of ranks 1.. (created in the background since start-up) and their copies of the sequences
void join_ranks(Ctx &c) {
	if (c.ranks_thread.joinable()) c.ranks_thread.join();
	if (c.ranks_rc != MC_OK) {
		fprintf(stderr, "meshclust: mc_ctx_create failed on an additional GPU: %s\n", c.ranks_err.c_str());
		exit(2);
	}
}

void ensure_rank_sequences(Ctx &c) {
	if (c.seq_cloned) return;
	Timer t;
	join_ranks(c);
	std::vector<std::thread> th;
	std::vector<int> rcs(c.ranks.size(), MC_OK);
	// (with the sequences, a scratch buffer that holds the boundary lines of a full batch of alignments: growing it in
	// the middle of the first large batch is a cudaFree + cudaMalloc of hundreds of MB on every GPU at once -- 0.2 s)
	uint64_t lmax = 0;
	for (uint64_t l : c.ds.len) lmax = std::max(lmax, l);
	const int64_t nw_scratch = (int64_t)std::min(4.3e9, 4736.0 * 2.0 * ((double)lmax + 34.0) * 24.0 + 64e6);
	for (size_t r = 1; r < c.ranks.size(); r++)
		th.emplace_back([&, r]() {
			rcs[r] = mc_clone_sequences(c.ranks[r], c.gpu);
			if (rcs[r] == MC_OK) mc_reserve_scratch(c.ranks[r], nw_scratch);
		});
	for (auto &x : th) x.join();
	for (size_t r = 1; r < c.ranks.size(); r++)
		if (rcs[r] != MC_OK) die_gpu("mc_clone_sequences");
	c.seq_cloned = true;
	printf("  [sequences copied to %zu more GPUs %.2fs]\n", c.ranks.size() - 1, t.lap());
}

// mc_align_pairs over row pairs, on all GPUs of the run when the batch is worth it.
// K4 is embarrassingly parallel over pairs.  Cells, not pairs, are the unit of work: the batch is cut
// into contiguous runs of equal cell counts, one per GPU, each driven by its own host thread; the other GPUs get
// their copy of the sequences the first time this happens.
void align_rows(Ctx &c, const std::vector<int32_t> &a, const std::vector<int32_t> &b, std::vector<int32_t> &sc, std::vector<int32_t> &ln,
                std::vector<int32_t> &mt) {
	const size_t m = a.size();
	sc.resize(m); ln.resize(m); mt.resize(m);
	if (m == 0) return;
	const int world = (int)c.ranks.size();
	double cells = 0;
	std::vector<double> cell_prefix;
	if (world > 1 && m >= (size_t)(4 * world)) {
		cell_prefix.resize(m + 1, 0.0);
		for (size_t i = 0; i < m; i++)
			cell_prefix[i + 1] = cell_prefix[i] + (double)c.ds.len[(size_t)c.ds.id_of_row[(size_t)a[i]]] * (double)c.ds.len[(size_t)c.ds.id_of_row[(size_t)b[i]]];
		cells = cell_prefix[m];
	}
	static const double split_min_cells = getenv("MC_ALIGN_SPLIT_MIN_CELLS") ? atof(getenv("MC_ALIGN_SPLIT_MIN_CELLS")) : 4e9;
	if (world > 1 && cells >= split_min_cells) {
		ensure_rank_sequences(c);
		std::vector<size_t> cut((size_t)world + 1, m);
		cut[0] = 0;
		for (int r = 1; r < world; r++)
			cut[(size_t)r] = (size_t)(std::lower_bound(cell_prefix.begin(), cell_prefix.end(), cells * r / world) - cell_prefix.begin());
		for (int r = 1; r <= world; r++) cut[(size_t)r] = std::min(m, std::max(cut[(size_t)r], cut[(size_t)r - 1]));
		std::vector<int> rcs((size_t)world, MC_OK);
		std::vector<std::string> errs((size_t)world);
		std::vector<std::thread> th;
		for (int r = 0; r < world; r++) {
			const size_t lo = cut[(size_t)r], hi = cut[(size_t)r + 1];
			if (hi <= lo) continue;
			th.emplace_back([&, r, lo, hi]() {
				rcs[(size_t)r] = mc_align_pairs(c.ranks[(size_t)r], a.data() + lo, b.data() + lo, (int64_t)(hi - lo), sc.data() + lo, ln.data() + lo, mt.data() + lo);
				if (rcs[(size_t)r] != MC_OK) errs[(size_t)r] = mc_last_error();
			});
		}
		for (auto &t : th) t.join();
		for (int r = 0; r < world; r++)
			if (rcs[(size_t)r] != MC_OK) {
				fprintf(stderr, "meshclust: mc_align_pairs failed on GPU %d: %s\n", r, errs[(size_t)r].c_str());
				exit(2);
			}
		c.align_split_calls++;
	} else GPU(mc_align_pairs(c.gpu, a.data(), b.data(), (int64_t)m, sc.data(), ln.data(), mt.data()));
}

// GlobAlignE identity of (a, b) id pairs: matches / length as double (GlobAlignE.cpp:301-305)
std::vector<double> align_ids(Ctx &c, const std::vector<Pair> &pairs) {
	const size_t m = pairs.size();
	std::vector<int32_t> a(m), b(m), sc, ln, mt;
	for (size_t i = 0; i < m; i++) {
		a[i] = (int32_t)c.ds.row_of_id[pairs[i].first];
		b[i] = (int32_t)c.ds.row_of_id[pairs[i].second];
	}
	align_rows(c, a, b, sc, ln, mt);
	std::vector<double> id(m);
	for (size_t i = 0; i < m; i++) id[i] = (double)mt[i] / ln[i];
	return id;
}

// point ids in the order Trainer::split's first std::sort leaves them (Trainer.cpp:672-675): unstable
// sort of (length, id) records by length, ids initially ascending
std::vector<int> sort_ids_by_length(const std::vector<uint64_t> &len) {
	struct LenId { uint64_t key; int id; };
	std::vector<LenId> rec(len.size());
	for (size_t i = 0; i < len.size(); i++) rec[i] = {len[i], (int)i};
	// (std::sort's permutation, with the host threads: host/lazy_sort.hpp)
	parallel_std_sort(rec, [](const LenId &a, const LenId &b) { return a.key < b.key; });
	std::vector<int> ids(len.size());
	for (size_t i = 0; i < len.size(); i++) ids[i] = rec[i].id;
	return ids;
}

// Trainer::split (Trainer.cpp:653-783)
std::vector<Pair> trainer_split(Ctx &c) {
	const Dataset &ds = c.ds;
	const int64_t n = ds.n;
	const double cutoff = c.opt.similarity;
	const size_t n_points = (size_t)c.opt.sample_size, max_pts_from_one = (size_t)c.opt.pivots;
	std::vector<int> points((size_t)n);
	for (int64_t i = 0; i < n; i++) points[i] = (int)i;
	Timer st;
	// The reference sorts Point* arrays with comparators that look the key up through the pointer.
	// std::sort's permutation depends only on the comparison outcomes, so sorting (key, id) records
	// with the same comparator on the key yields the identical order without the random accesses.
	struct KeyId { uint16_t key; int id; };
	// :672-675 unstable sort by length, then the median-length point (sorted on a helper thread while
	// the sequences were uploaded and counted: it needs nothing but the lengths)
	const bool dbg_t = getenv("MC_DEBUG_TIMING") != nullptr;
	Timer tsub;
	auto sub = [&](const char *what) { if (dbg_t) fprintf(stderr, "  [split: %-32s %.3f s]\n", what, tsub.lap()); };
	if (c.by_length.valid()) points = c.by_length.get();
	else points = sort_ids_by_length(ds.len);
	sub("ids by length (wait)");
	const int begin_pt = points[points.size() / 2];
	// :681-684 sort by distance to it
	{
		std::vector<uint16_t> key1((size_t)n);
		int32_t r = (int32_t)ds.row_of_id[begin_pt];
		GPU(mc_distance_keys(c.gpu, &r, 1, key1.data()));
		sub("keys of the median point");
		std::vector<KeyId> rec((size_t)n);
#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < n; i++) rec[i] = {key1[ds.row_of_id[points[i]]], points[i]};
		sub("records");
		parallel_std_sort(rec, [](const KeyId &a, const KeyId &b) { return a.key < b.key; });
		sub("sort by distance");
#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < n; i++) points[i] = rec[i].id;
	}
	// :685-690 pivots at even ranks
	const int num_iterations = (int)std::ceil(((double)n_points) / max_pts_from_one) - 1;
	std::vector<int> pivots;
	for (int i = 0; i <= num_iterations; i++) {
		const int idx = (int)((size_t)i * (points.size() - 1) / (size_t)num_iterations);
		pivots.push_back(points[idx]);
	}
	printf("Point pairs: %zu\n", pivots.size());
	const size_t np = pivots.size();
	const size_t to_add_each = max_pts_from_one / 2;

	// distance keys of every point against every pivot (the reference evaluates them inside the
	// sort comparators, 2 per comparison)
	std::vector<int32_t> prow(np);
	for (size_t i = 0; i < np; i++) prow[i] = (int32_t)ds.row_of_id[pivots[i]];
	// (not a std::vector: value-initialising 300 MB of keys that the copy overwrites was 0.1 s of "first sorts" at C4)
	RawBytes keys_raw;
	keys_raw.resize(np * (size_t)n * sizeof(uint16_t));
	uint16_t *keys = reinterpret_cast<uint16_t *>(keys_raw.data());
	{
		// first touch by all threads: the device-to-host copy into never-touched pages faults them in one by one
		// (0.13 s for the 300 MB of C4), and so did the value-initialisation
		uint8_t *kb = keys_raw.data();
		const size_t bytes = keys_raw.size();
#pragma omp parallel for schedule(static)
		for (long long off = 0; off < (long long)bytes; off += 4096) kb[off] = 0;
	}
	const double t_first = st.lap();
	GPU(mc_distance_keys(c.gpu, prow.data(), (int)np, keys));
	const double t_keys = st.lap();

	// :694-701 per pivot: copy + unstable sort by distance to the pivot.  Independent per pivot, so
	// the host threads can share them without changing any permutation.
	// Only ~40 positions of every sorted array are ever looked at (the search below, twenty picks at the end): the
	// sorts run lazily (host/lazy_sort.hpp: libstdc++'s introsort restricted to the ranges that hold a position asked
	// for -- the same permutation among equal keys, a fraction of the work).  MC_SPLIT_FULL_SORT=1: plain std::sort.
	struct ByKey { bool operator()(const KeyId &a, const KeyId &b) const { return a.key < b.key; } };
	std::vector<LazySort<KeyId, ByKey>> sorted(np);
	std::vector<int32_t> row_of_point((size_t)n);
	for (int64_t i = 0; i < n; i++) row_of_point[i] = (int32_t)ds.row_of_id[points[i]];
	const bool full_sort = getenv("MC_SPLIT_FULL_SORT") != nullptr;
#pragma omp parallel for schedule(dynamic)
	for (long i = 0; i < (long)np; i++) {
		const uint16_t *kk = keys + (size_t)i * n;
		std::vector<KeyId> rec((size_t)n);
		for (int64_t j = 0; j < n; j++) rec[j] = {kk[row_of_point[j]], points[j]};
		if (full_sort) std::sort(rec.begin(), rec.end(), ByKey());
		sorted[(size_t)i] = LazySort<KeyId, ByKey>(std::move(rec), ByKey(), full_sort ? 0 : -1);
		if (full_sort) sorted[(size_t)i].mark_sorted();
	}
	keys_raw.release();
	double t_sorts = st.lap();
	int rounds = 0;
	// positions of sorted[i] resolved by the host threads, pivot by pivot, before the serial code reads them
	double t_resolve = 0;
	auto resolve = [&](const std::vector<std::vector<size_t>> &ask) {
		Timer tr;
#pragma omp parallel for schedule(dynamic)
		for (long i = 0; i < (long)np; i++)
			for (size_t p : ask[(size_t)i]) sorted[(size_t)i].at(p);
		t_resolve += tr.lap();
	};

	// :703-721 binary search with alignment, all pivots in lock step (each search is independent).
	// One round of the reference is one alignment per pivot: 150 pairs cannot fill a GPU, and a
	// search is a chain of ~log2(n/4) dependent rounds.  So every batch aligns, per pivot, all the
	// positions the next SPEC rounds can reach (2^SPEC - 1 of them), and the rounds are then replayed
	// from the answers exactly as the reference takes them; the unused answers are discarded.
	// How far to look ahead: a round of short pairs is launch latency, so 4 levels (15 positions per pivot);
	// long pairs fill the GPU with far fewer of them (one warp per pair, ~1800 resident warps), and every
	// position beyond that is a second wave of work most of which is thrown away: 3 levels then (7 per pivot).
	int SPEC = 4;
	{
		double mean_len = 0;
		for (size_t i = 0; i < np; i++) mean_len += (double)(ds.fa.offsets[(size_t)pivots[i] + 1] - ds.fa.offsets[(size_t)pivots[i]]);
		mean_len /= (double)std::max<size_t>(np, 1);
		if (mean_len * mean_len > 1e7) {
			// long pairs: a batch that leaves every GPU at most two teams per SM (296 pairs on a B200) is aligned by
			// teams of six warps per pair, and a round then lasts a sixth of one pair's single-warp latency -- fewer,
			// speculative positions per round buy nothing beyond that.  One GPU: no look-ahead (150 pairs per round);
			// every GPU of the run takes its share of a round, so with more of them look further ahead (fewer rounds).
			SPEC = 6;
			while (SPEC > 1 && np * ((size_t)(1 << SPEC) - 1) > 296 * c.ranks.size()) SPEC--;
		}
		if (getenv("MC_SPLIT_SPEC")) SPEC = std::max(1, std::min(6, atoi(getenv("MC_SPLIT_SPEC"))));
	}
	std::vector<size_t> offset(np, (size_t)n / 4), pos(np, 2 * ((size_t)n / 4));
	std::vector<char> active(np, 1);
	for (size_t i = 0; i < np; i++) active[i] = offset[i] > 0;
	for (;;) {
		std::vector<Pair> q;
		std::vector<char> reach;             // per node: will the search ever align it?
		std::vector<size_t> first(np + 1, 0);
		// breadth-first over the decision tree of every active pivot: node 0 = current position; children of
		// node t are 2t+1 (answer below the cutoff: pos - off) and 2t+2 (above: pos + off), off halving per level
		const size_t tree = (size_t)(1 << SPEC) - 1;
		std::vector<size_t> npos(np * tree, 0), noff(np * tree, 0);
		std::vector<std::vector<size_t>> ask(np);
		for (size_t i = 0; i < np; i++) {
			if (!active[i]) continue;
			size_t *ps = npos.data() + i * tree, *of = noff.data() + i * tree;
			ps[0] = pos[i]; of[0] = offset[i];
			for (size_t t = 0; t < tree; t++) {
				const bool reachable = of[t] > 0;   // offset 0 ends the search before this alignment
				if (reachable) ask[i].push_back(ps[t]);
				if (2 * t + 2 < tree) {
					ps[2 * t + 1] = reachable ? ps[t] - of[t] : 0; of[2 * t + 1] = reachable ? of[t] / 2 : 0;
					ps[2 * t + 2] = reachable ? ps[t] + of[t] : 0; of[2 * t + 2] = reachable ? of[t] / 2 : 0;
				}
			}
		}
		resolve(ask);
		for (size_t i = 0; i < np; i++) {
			first[i] = q.size();
			if (!active[i]) continue;
			const size_t *ps = npos.data() + i * tree, *of = noff.data() + i * tree;
			for (size_t t = 0; t < tree; t++) {
				const bool reachable = of[t] > 0;
				q.push_back({pivots[i], reachable ? sorted[i].at(ps[t]).id : pivots[i]});
				reach.push_back(reachable ? 1 : 0);
			}
		}
		first[np] = q.size();
		if (q.empty()) break;
		// unreachable placeholders are not aligned
		std::vector<Pair> real;
		std::vector<size_t> where(q.size(), (size_t)-1);
		for (size_t t = 0; t < q.size(); t++)
			if (reach[t]) { where[t] = real.size(); real.push_back(q[t]); }
		const std::vector<double> algn = align_ids(c, real);
		rounds++;
		for (size_t i = 0; i < np; i++) {
			if (!active[i]) continue;
			size_t t = 0;
			for (int level = 0; level < SPEC && active[i]; level++) {
				const size_t slot = first[i] + t;
				// (an unreachable node is never visited: offset 0 deactivates the search one level above it)
				const double a = algn[where[slot]];
				if (a < cutoff) { pos[i] -= offset[i]; t = 2 * t + 1; }
				else if (a > cutoff) { pos[i] += offset[i]; t = 2 * t + 2; }
				else { active[i] = 0; break; }   // break: offset is not halved, pos stays
				offset[i] /= 2;
				if (offset[i] == 0) active[i] = 0;
			}
		}
	}

	{
		const double t_rounds = st.lap() - t_resolve;   // (the lazy sorts' share of the rounds is booked with the sorts)
		t_sorts += t_resolve;
		t_resolve = 0;
		printf("  [split: first sorts %.3fs, %zu x n distance keys %.3fs, pivot sorts %.3fs, %d alignment rounds %.3fs]\n", t_first, np, t_keys, t_sorts, rounds, t_rounds);
	}
	// :723-765 ten picks below and ten above the boundary at evenly strided ranks
	int aerr = 0;
	HeaderPairLess less{&ds.fa.headers};
	std::set<Pair, HeaderPairLess> pairs(less);
	{
		// the picks' positions, resolved by all threads first (the same arithmetic as the loop below)
		std::vector<std::vector<size_t>> ask(np);
		for (size_t i = 0; i < np; i++) {
			const size_t pivot = pos[i], sz = sorted[i].size();
			const double before_inc = (double)pivot / to_add_each, after_inc = ((double)(sz - pivot)) / to_add_each;
			double before_start = 0, after_start = (double)pivot;
			for (size_t t = 0; t < to_add_each; t++) { ask[i].push_back((size_t)(int)std::round(before_start)); before_start += before_inc; }
			for (size_t t = 0; t < to_add_each && std::round(after_start) < sz; t++) { ask[i].push_back((size_t)(int)std::round(after_start)); after_start += after_inc; }
		}
		resolve(ask);
	}
	for (size_t i = 0; i < np; i++) {
		struct Pts {
			LazySort<KeyId, ByKey> &s;
			size_t size() const { return s.size(); }
			int operator[](size_t k) const { return s.at(k).id; }
		} pts{sorted[i]};
		const int p = pivots[i];
		const size_t pivot = pos[i];
		const double before_inc = (double)pivot / to_add_each;
		const double after_inc = ((double)(pts.size() - pivot)) / to_add_each;
		if (before_inc < 1) aerr = 1;
		else if (after_inc < 1) aerr = -1;
		double before_start = 0, after_start = (double)pivot;
		std::vector<Pair> buf;
		auto ordered = [&](int other) {
			return ds.fa.headers[p].compare(ds.fa.headers[other]) < 0 ? Pair{p, other} : Pair{other, p};
		};
		for (size_t t = 0; t < to_add_each; t++) {
			const int idx = (int)std::round(before_start);
			buf.push_back(ordered(pts[idx]));
			before_start += before_inc;
		}
		for (size_t t = 0; t < to_add_each && std::round(after_start) < pts.size(); t++) {
			const int idx = (int)std::round(after_start);
			buf.push_back(ordered(pts[idx]));
			after_start += after_inc;
		}
		pairs.insert(buf.begin(), buf.end());
	}
	if (aerr < 0) fprintf(stderr, "Warning: Alignment may be too small for sampling\n");
	else if (aerr > 0) fprintf(stderr, "Warning: Alignment may be too large for sampling\n");
	return std::vector<Pair>(pairs.begin(), pairs.end());
}

using Scored = std::pair<Pair, double>;

// resize_vec (Trainer.cpp:201-243): note that it can return MORE than new_size items and repeats
// items when a bin runs dry -- kept as is
std::vector<Scored> resize_vec(const std::vector<Scored> &vec, size_t new_size, double min_align, double max_align, int num_bins) {
	if (new_size == vec.size()) return vec;
	std::vector<std::vector<Scored>> bins((size_t)num_bins);
	auto get_bin = [&](double x) {
		if (x >= max_align) return num_bins - 1;
		if (x <= min_align) return 0;
		return (int)(num_bins * (x - min_align) / (max_align - min_align));
	};
	for (const auto &p : vec) bins.at((size_t)get_bin(p.second)).push_back(p);
	std::vector<Scored> data;
	while (data.size() < new_size) {
		const int items_left = (int)(new_size - data.size());
		const int take = (int)std::ceil((double)items_left / num_bins);
		for (int i = (int)bins.size() - 1; i >= 0; i--)
			for (int j = 0; j < (int)std::min((size_t)take, bins[i].size()); j++) data.push_back(bins[i][j]);
	}
	return data;
}

// bin_data (Trainer.cpp:490-526): ten identity bins, alternate train/test, parity flips per bin
void bin_data(const std::vector<Scored> &vec, double min_align, double max_align, std::vector<Pair> &train, std::vector<Pair> &test) {
	const int n_bins = 10;
	auto get_bin = [&](double x) {
		if (x >= max_align) return n_bins - 1;
		if (x <= min_align) return 0;
		return (int)(n_bins * (x - min_align) / (max_align - min_align));
	};
	std::vector<std::vector<Scored>> bins((size_t)n_bins);
	for (const auto &d : vec) bins.at((size_t)get_bin(d.second)).push_back(d);
	int last = 0;
	for (const auto &bin : bins) {
		for (int i = 0; i < (int)bin.size(); i++) {
			if (i % 2 == last) train.push_back(bin[i].first);
			else test.push_back(bin[i].first);
		}
		last = !last;
	}
}

// Trainer::get_labels (Trainer.cpp:253-333)
void trainer_get_labels(Ctx &c, std::vector<Pair> vec, std::vector<Scored> &pos_out, std::vector<Scored> &neg_out) {
	const double cutoff = c.opt.similarity;
	// struct rng: srand(0), rand() % n; libstdc++ random_shuffle(first, last, rng)
	srand(0);
	for (size_t i = 1; i < vec.size(); i++) {
		const size_t j = (size_t)(rand() % (int)(i + 1));
		if (i != j) std::swap(vec[i], vec[j]);
	}
	const std::vector<double> algn = align_ids(c, vec);
	ScoredLess sless{HeaderPairLess{&c.ds.fa.headers}};
	std::set<Scored, ScoredLess> buf_pos(sless), buf_neg(sless);
	for (size_t i = 0; i < vec.size(); i++) {
		if (algn[i] >= cutoff) buf_pos.insert({vec[i], algn[i]});
		else buf_neg.insert({vec[i], algn[i]});
	}
	printf("positive=%zu negative=%zu\n", buf_pos.size(), buf_neg.size());
	if (buf_pos.empty() || buf_neg.empty()) {
		printf("Identity value does not match sampled data: %s\n",
		       buf_pos.empty() ? "Too many sequences below identity" : "Too many sequences above identity");
		exit(0);   // Trainer.cpp:306-315
	}
	const size_t m_size = std::min(buf_pos.size(), buf_neg.size());
	const std::vector<Scored> vpos(buf_pos.begin(), buf_pos.end()), vneg(buf_neg.begin(), buf_neg.end());
	pos_out = resize_vec(vpos, m_size, cutoff, 1, 5);
	neg_out = resize_vec(vneg, m_size, 0.4, cutoff, 5);
	printf("positive=%zu negative=%zu\n", pos_out.size(), neg_out.size());
}

// raw features of id pairs in lookup order [LD, INTERSECTION, MANHATTAN, PEARSON, KULCZYNSKI2]
std::vector<double> raw_features(Ctx &c, const std::vector<Pair> &pairs) {
	const size_t m = pairs.size();
	std::vector<int32_t> a(m), b(m);
	for (size_t i = 0; i < m; i++) {
		a[i] = (int32_t)c.ds.row_of_id[pairs[i].first];
		b[i] = (int32_t)c.ds.row_of_id[pairs[i].second];
	}
	std::vector<double> raw(m * 5);
	if (m) GPU(mc_pair_features(c.gpu, a.data(), b.data(), (int64_t)m, raw.data(), nullptr));
	return raw;
}

// generate_feat_mat (Trainer.cpp:367-414): rows = positives then negatives, leading 1 column
void feature_matrix(Ctx &c, const std::vector<Pair> &pos, const std::vector<Pair> &neg, int ncols, Mat &X, Mat &y) {
	std::vector<Pair> all(pos);
	all.insert(all.end(), neg.begin(), neg.end());
	const size_t m = all.size();
	std::vector<int32_t> a(m), b(m);
	for (size_t i = 0; i < m; i++) {
		a[i] = (int32_t)c.ds.row_of_id[all[i].first];
		b[i] = (int32_t)c.ds.row_of_id[all[i].second];
	}
	std::vector<double> feats(m * 4);
	GPU(mc_pair_classify(c.gpu, a.data(), b.data(), (int64_t)m, nullptr, nullptr, nullptr, feats.data()));
	X = Mat((int)m, ncols);
	y = Mat((int)m, 1);
	for (size_t i = 0; i < m; i++) {
		X.at((int)i, 0) = 1;
		for (int col = 1; col < ncols; col++) X.at((int)i, col) = feats[i * 4 + (col - 1)];
		y.at((int)i, 0) = i < pos.size() ? 1 : -1;
	}
}

// Trainer::train (Trainer.cpp:527-651)
void trainer_train(Ctx &c) {
	Model &M = c.model;
	M.cutoff = c.opt.similarity;
	if (c.opt.align) {
		// :570-577 single FEAT_ALIGN feature with bounds [0,1]; res == 1  <=>  identity >= id
		M.align = true;
		M.nfeat = 1;
		M.w[0] = -1 * M.cutoff;
		M.w[1] = 1;
		return;
	}
	printf("Splitting data\n");
	Timer tm;
	std::vector<Pair> sample = trainer_split(c);
	printf("  [split %.2fs, %zu pairs]\n", tm.lap(), sample.size());
	std::vector<Scored> pos, neg;
	trainer_get_labels(c, sample, pos, neg);
	printf("  [labels %.2fs]\n", tm.lap());
	std::vector<Pair> train_pos, test_pos, train_neg, test_neg;
	bin_data(pos, M.cutoff, 1, train_pos, test_pos);
	bin_data(neg, 0, M.cutoff, train_neg, test_neg);
	printf("training positive: %zu\ntraining negative: %zu\ntesting positive: %zu\ntesting negative: %zu\n",
	       train_pos.size(), train_neg.size(), test_pos.size(), test_neg.size());
	if (test_pos.empty() || test_neg.empty()) {
		fprintf(stderr, "terminate called after throwing an instance of 'char const*' (not enough points to sample)\n");
		abort();
	}

	// :584-587 feature combos in order; lookups appear in the order add_feature() meets them:
	// LD, INTERSECTION | MANHATTAN | PEARSON | KULCZYNSKI2  ->  bounds for lookups 0..3 are fixed by
	// the 3-feature round, lookup 4 by the 4-feature round (Feature::normalize skips finalized ones)
	const double raw_init_min = DBL_MAX, raw_init_max = DBL_MIN;   // Feature.cpp:21-22 (DBL_MIN > 0!)
	double mins[5], maxs[5];
	for (int i = 0; i < 5; i++) { mins[i] = raw_init_min; maxs[i] = raw_init_max; }
	const std::vector<double> raw_pos = raw_features(c, train_pos), raw_neg = raw_features(c, train_neg);
	auto fold = [&](int lo, int hi) {
		for (int i = lo; i < hi; i++) {
			double small = mins[i], big = maxs[i];
			for (const std::vector<double> *raw : {&raw_pos, &raw_neg})
				for (size_t j = 0; j < raw->size() / 5; j++) {
					const double v = (*raw)[j * 5 + i];
					if (v < small) small = v;
					if (v > big) big = v;
				}
			mins[i] = small;
			maxs[i] = big;
		}
	};
	double prev_acc = -10000;
	struct Snap { int nfeat; double mins[5], maxs[5], w[5]; };
	std::vector<Snap> snaps;
	for (int nf = 3; nf <= 4; nf++) {
		if (nf == 3) fold(0, 4); else fold(4, 5);
		for (int i = 0; i < (nf == 3 ? 4 : 5); i++) printf("bounds[%d]: %g to %g\n", i, mins[i], maxs[i]);
		const double w0[5] = {0, 0, 0, 0, 0};
		GPU(mc_set_model(c.gpu, mins, maxs, w0, nf));
		Mat Xtr, ytr, Xte, yte;
		feature_matrix(c, train_pos, train_neg, nf + 1, Xtr, ytr);
		feature_matrix(c, test_pos, test_neg, nf + 1, Xte, yte);
		const Mat w = glm_train(Xtr, ytr);
		const double acc = glm_accuracy(Xte, w, yte);
		glm_accuracy(Xtr, w, ytr);
		if (acc - prev_acc <= 1 && acc >= 90.0) {   // :633-638 keep the previous, smaller model
			printf("feat size is %d\n", snaps.back().nfeat);
			break;
		}
		Snap s;
		s.nfeat = nf;
		for (int i = 0; i < 5; i++) { s.mins[i] = mins[i]; s.maxs[i] = maxs[i]; s.w[i] = i <= nf ? w.at(i, 0) : 0.0; }
		snaps.push_back(s);
		prev_acc = acc;
		if (acc >= 97.5) { printf("breaking from acc cutoff\n"); break; }
	}
	const Snap &s = snaps.back();
	M.nfeat = s.nfeat;
	for (int i = 0; i < 5; i++) { M.mins[i] = s.mins[i]; M.maxs[i] = s.maxs[i]; M.w[i] = s.w[i]; }
	printf("Using %d features\n", M.nfeat);
	GPU(mc_set_model(c.gpu, M.mins, M.maxs, M.w, M.nfeat));
	printf("  [glm %.2fs]\n", tm.lap());
}

// ----------------------------------------------------------------------------------------------
// ClusterFactory::MS (ClusterFactory.cpp:717-761)
// ----------------------------------------------------------------------------------------------
struct Cluster {
	int64_t center_row;              // the row the center was cloned from (Center.h:14)
	std::vector<int64_t> rows;       // members in the reference's order
	bool removed = false;
};

// --align: identity cache of Feature::align (Feature.cpp:222-243), keyed by the id pair
// Open-addressing table keyed by the unordered id pair: only find / insert are ever used (never
// iteration), so it stands in for the reference's std::map<pair<id,id>,double> at a fraction of
// the cost (a --align run makes millions of look-ups and inserts).
class AlignCache {
public:
	AlignCache() { keys_.assign(1 << 16, EMPTY); vals_.resize(1 << 16); }
	static uint64_t key_of(int64_t a, int64_t b) { return a < b ? ((uint64_t)a << 32) | (uint64_t)b : ((uint64_t)b << 32) | (uint64_t)a; }
	const double *find(uint64_t key) const {
		for (size_t i = slot(key);; i = (i + 1) & (keys_.size() - 1)) {
			if (keys_[i] == key) return &vals_[i];
			if (keys_[i] == EMPTY) return nullptr;
		}
	}
	void set(uint64_t key, double v) {
		if ((used_ + 1) * 5 > keys_.size() * 3) grow();
		for (size_t i = slot(key);; i = (i + 1) & (keys_.size() - 1)) {
			if (keys_[i] == key) { vals_[i] = v; return; }
			if (keys_[i] == EMPTY) { keys_[i] = key; vals_[i] = v; used_++; return; }
		}
	}
	// ids that have been the center of a scan: a pair (point, center) can only be in the table when the
	// center has been one before -- a point that is still alive has never been a center itself
	std::vector<char> was_center;

private:
	static constexpr uint64_t EMPTY = ~0ull;
	size_t slot(uint64_t k) const {
		k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
		return (size_t)k & (keys_.size() - 1);
	}
	void grow() {
		std::vector<uint64_t> ok;
		std::vector<double> ov;
		ok.swap(keys_); ov.swap(vals_);
		keys_.assign(ok.size() * 2, EMPTY);
		vals_.resize(ok.size() * 2);
		used_ = 0;
		for (size_t i = 0; i < ok.size(); i++) if (ok[i] != EMPTY) set(ok[i], ov[i]);
	}
	std::vector<uint64_t> keys_;
	std::vector<double> vals_;
	size_t used_ = 0;
};

// one get_close over [lo,hi] in --align mode: every alive row of the range is aligned against the
// center (point = seq1, center = seq2, Feature.cpp:231-235), res == 1 <=> identity >= id
void align_scan(Ctx &c, BVec &bv, AlignCache &cache, int64_t center_row, int64_t lo, int64_t hi,
                std::vector<uint8_t> &alive, mc_scan_result &res, std::vector<uint8_t> &marks) {
	const Dataset &ds = c.ds;
	std::vector<int64_t> rows;
	for (int64_t r = lo; r <= hi; r++) if (alive[r]) rows.push_back(r);
	std::vector<int32_t> a, b;
	std::vector<size_t> need;
	std::vector<double> ident(rows.size());
	const int64_t cid = ds.id_of_row[center_row];
	if (cache.was_center.empty()) cache.was_center.assign((size_t)ds.n, 0);
	const bool seen_before = cache.was_center[(size_t)cid] != 0;
	cache.was_center[(size_t)cid] = 1;
	for (size_t i = 0; i < rows.size(); i++) {
		const int64_t pid = ds.id_of_row[rows[i]];
		const double *hit = seen_before ? cache.find(AlignCache::key_of(pid, cid)) : nullptr;
		if (hit) ident[i] = *hit;
		else { need.push_back(i); a.push_back((int32_t)rows[i]); b.push_back((int32_t)center_row); }
	}
	if (!need.empty()) {
		std::vector<int32_t> sc, ln, mt;
		align_rows(c, a, b, sc, ln, mt);
		for (size_t t = 0; t < need.size(); t++) {
			const double v = (double)mt[t] / ln[t];
			ident[need[t]] = v;
			cache.set(AlignCache::key_of(ds.id_of_row[rows[need[t]]], cid), v);
		}
	}
	res.n_eval = (int64_t)rows.size();
	res.n_pos = 0;
	res.best_row = -1;
	res.best_f0 = -1;
	marks.assign((size_t)(hi - lo + 1), 0);
	for (size_t i = 0; i < rows.size(); i++) {
		// normalised FEAT_ALIGN = (v - 0) / (1 - 0); sum = -id + 1 * v
		const double v = (ident[i] - 0.0) / (1.0 - 0.0);
		const double sum = std::fma(c.model.w[1], v, c.model.w[0]);
		const double r = std::round(1.0 / (1 + std::exp(-sum)));
		if (v > res.best_f0) { res.best_f0 = v; res.best_row = rows[i]; }
		if (r == 1.0) { marks[(size_t)(rows[i] - lo)] = 1; res.n_pos++; alive[rows[i]] = 0; }
	}
	(void)bv;
}

void write_clstr(const Ctx &c, const std::vector<Cluster> &part) {
	// print_output (ClusterFactory.cpp:495-520)
	printf("Printing output\n");
	FILE *f = fopen(c.opt.output.c_str(), "w");
	if (!f) { fprintf(stderr, "cannot open %s\n", c.opt.output.c_str()); exit(1); }
	// decimal conversion by hand into big buffers (a million fprintf calls are a visible slice of a run), formatted by
	// all host threads: the non-empty clusters are numbered first, then cut into runs of about equal member counts,
	// one buffer per run, written in order.  Rounds of at most ~4 M members bound the memory.
	std::vector<size_t> live;
	for (size_t i = 0; i < part.size(); i++) if (!part[i].rows.empty()) live.push_back(i);
	int threads = 1;
#ifdef _OPENMP
	threads = omp_get_max_threads();
#endif
	auto format = [&](size_t first, size_t last, std::string &buf) {   // clusters live[first, last)
		char num[32];
		auto put_u = [&](unsigned long long v) {
			int k = 0;
			do { num[k++] = (char)('0' + v % 10); v /= 10; } while (v);
			while (k) buf.push_back(num[--k]);
		};
		for (size_t ci = first; ci < last; ci++) {
			const Cluster &cl = part[live[ci]];
			buf += ">Cluster ";
			put_u((unsigned long long)ci);
			buf.push_back('\n');
			unsigned long long pt = 0;
			for (int64_t r : cl.rows) {
				const int64_t id = c.ds.id_of_row[r];
				put_u(pt);
				buf.push_back('\t');
				put_u((unsigned long long)c.ds.len[id]);
				buf += "nt, ";
				buf += c.ds.fa.headers[id];
				buf += "... ";
				if (r == cl.center_row) buf.push_back('*');
				buf.push_back('\n');
				pt++;
			}
		}
	};
	size_t done = 0;
	while (done < live.size()) {
		// this round: clusters [done, end) with at most ~4 M members in total
		size_t end = done, members = 0;
		while (end < live.size() && (end == done || members + part[live[end]].rows.size() <= ((size_t)4 << 20))) members += part[live[end++]].rows.size();
		std::vector<size_t> cut((size_t)threads + 1, end);
		cut[0] = done;
		{
			size_t acc = 0, t = 1;
			for (size_t ci = done; ci < end && t < (size_t)threads; ci++) {
				acc += part[live[ci]].rows.size();
				while (t < (size_t)threads && acc >= members * t / (size_t)threads) cut[t++] = ci + 1;
			}
		}
		std::vector<std::string> bufs((size_t)threads);
#pragma omp parallel for schedule(static, 1)
		for (int t = 0; t < threads; t++) {
			if (cut[(size_t)t + 1] <= cut[(size_t)t]) continue;
			bufs[(size_t)t].reserve((members / (size_t)threads + 1024) * 48);
			format(cut[(size_t)t], cut[(size_t)t + 1], bufs[(size_t)t]);
		}
		for (const std::string &bf : bufs) if (!bf.empty()) fwrite(bf.data(), 1, bf.size(), f);
		done = end;
	}
	fclose(f);
}

void mean_shift(Ctx &c, BVec &bv) {
	const Dataset &ds = c.ds;
	const double sim = c.opt.similarity;
	const int delta = c.opt.delta;
	std::vector<Cluster> part;
	Timer tm;
	AlignCache cache;
	std::vector<uint8_t> alive;   // host mirror, only needed by the --align scans
	if (c.model.align) alive.assign((size_t)ds.n, 1);
	// several GPUs share the Phase-A scans only when one scan is long (see run_pipeline); otherwise the additional
	// GPUs have served the alignments of the training stage and rank 0 runs the persistent kernel alone
	const int world = c.shard_phase_a ? (int)c.ranks.size() : 1;
	if (world > 1) join_ranks(c);
	// ---------------- Phase A on the device (mc_accumulate_run) ---------------------------------
	// The whole `while (last) accumulate(...)` loop (ClusterFactory.cpp:722-729, :637-714) with its bvec
	// bookkeeping runs as one persistent kernel; the host only reads the clusters back.  --align (the
	// decision is an alignment, not a scan), histogram shapes the staged scan kernel does not take and
	// MC_PHASE_A_STEPS=1 (tests) go through the step-by-step loop below, and so do runs whose scans are sharded
	// over several GPUs (inputs of gigabytes, see run_pipeline).
	bool phase_a_done = false;
	if (!c.model.align && !getenv("MC_PHASE_A_STEPS") && world == 1) {
		const std::vector<int64_t> first = bv.first_rows();
		std::vector<int64_t> centers((size_t)ds.n), offs((size_t)ds.n + 1), members((size_t)ds.n);
		mc_run_stats st;
		const int rc = mc_accumulate_run(c.gpu, sim, bv.bounds().data(), first.data(), (int64_t)bv.nbins(), centers.data(), offs.data(), members.data(), &st);
		if (rc == MC_OK) {
			part.resize((size_t)st.n_clusters);
#pragma omp parallel for schedule(dynamic, 64)
			for (long ci = 0; ci < (long)st.n_clusters; ci++) {
				part[(size_t)ci].center_row = centers[(size_t)ci];
				part[(size_t)ci].rows.assign(members.begin() + offs[(size_t)ci], members.begin() + offs[(size_t)ci + 1]);
			}
			printf("Accumulation: %zu clusters, %lld scans, %lld evals, on the device in %.4fs (%.2f us per step), %lld row compactions  [%.2fs]\n", part.size(),
			       (long long)st.n_scans, (long long)st.n_evals, st.device_seconds, st.n_steps ? st.device_seconds * 1e6 / (double)st.n_steps : 0.0,
			       (long long)st.n_compactions, tm.lap());
			phase_a_done = true;
		} else if (rc != MC_ERR_UNSUPPORTED) {
			die_gpu("mc_accumulate_run");
		}
	}
	if (!phase_a_done) {
	if (world > 1 && !c.model.align) {
		// rows are replicated once (device-to-device), scan work and alive flags are sharded
		// block-interleaved (~256 KB of consecutive rows per block, blocks round-robin over the GPUs);
		// summaries and marks cross GPUs inside the scan kernel
		Timer ts;
		for (int r = 1; r < world; r++) GPU(mc_clone_points(c.ranks[r], c.gpu));
		for (int r = 0; r < world; r++) GPU(mc_comm_init(c.ranks[r], r, world, nullptr));
		GPU(mc_comm_connect_local(c.ranks.data(), world));
		printf("  [points replicated to %d GPUs, peer inboxes connected %.2fs]\n", world, ts.lap());
	}
	const bool sharded = world > 1 && !c.model.align;
	for (int r = 0; r < (sharded ? world : 1); r++) GPU(mc_alive_reset(c.ranks[r]));
	const int64_t compact_min = getenv("MC_COMPACT_MIN_ROWS") ? atoll(getenv("MC_COMPACT_MIN_ROWS")) : 32768;
	if (!c.model.align && ds.n >= compact_min)   // staging for the row compactions: allocated before the clock of Phase A starts
		for (int r = 0; r < (sharded ? world : 1); r++) GPU(mc_reserve_permute(c.ranks[r]));
	tm.lap();

	auto kill_row = [&](int64_t row) {
		for (int r = 0; r < (sharded ? world : 1); r++) GPU(mc_alive_kill(c.ranks[r], &row, 1));
		if (c.model.align) alive[row] = 0;
	};

	// ---------------- Phase A: accumulate (ClusterFactory.cpp:637-714, :722-729) -----------------
	int64_t last = bv.pop();
	if (last >= 0) kill_row(last);
	std::vector<uint8_t> marks;
	std::vector<int64_t> marked_rows((size_t)ds.n);
	int64_t scans = 0, evals = 0;
	// Row compaction: once at most 60 % of the rows the scans still stream are alive, the alive rows
	// are moved to the front (same order) on the GPU(s) and every row number the host holds is
	// translated.  Scans then read alive rows only; nothing observable depends on the numbering.
	int64_t prefix = ds.n;                      // rows [0, prefix) may still be alive
	int64_t n_alive = (int64_t)bv.size();
	int compactions = 0;
	double compact_host_s = 0, compact_gpu_s = 0;
	auto compact_rows = [&](int64_t &seed) {
		Timer tc;
		std::vector<int64_t> old_of_new;
		old_of_new.reserve((size_t)prefix);
		bv.compact(old_of_new);                 // survivors, in iteration order
		const int64_t alive_now = (int64_t)old_of_new.size();
		std::vector<int64_t> new_of_old((size_t)prefix, -1);
		for (int64_t i = 0; i < alive_now; i++) new_of_old[(size_t)old_of_new[(size_t)i]] = i;
		for (int64_t o = 0; o < prefix; o++)    // then the rows that have left, in their old order
			if (new_of_old[(size_t)o] < 0) { new_of_old[(size_t)o] = (int64_t)old_of_new.size(); old_of_new.push_back(o); }
		compact_host_s += tc.lap();
		for (int r = 0; r < (sharded ? world : 1); r++) GPU(mc_permute_rows(c.ranks[r], old_of_new.data(), prefix, alive_now));
		compact_gpu_s += tc.lap();
		auto tr = [&](int64_t &row) { if (row >= 0 && row < prefix) row = new_of_old[(size_t)row]; };
#pragma omp parallel for schedule(dynamic, 16)
		for (long ci = 0; ci < (long)part.size(); ci++) { Cluster &cl = part[(size_t)ci]; tr(cl.center_row); for (int64_t &r : cl.rows) tr(r); }
		tr(seed);
		std::vector<int64_t> ids((size_t)prefix);
#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < prefix; i++) ids[(size_t)i] = c.ds.id_of_row[(size_t)old_of_new[(size_t)i]];
#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < prefix; i++) { c.ds.id_of_row[(size_t)i] = ids[(size_t)i]; c.ds.row_of_id[(size_t)ids[(size_t)i]] = i; }
		prefix = alive_now;
		compactions++;
		compact_host_s += tc.lap();
	};
	while (last >= 0) {
		if (!c.model.align && prefix >= compact_min && n_alive * 5 <= prefix * 3) compact_rows(last);
		std::vector<int64_t> current{last};
		bool is_min = false, first_mean = true;
		int64_t next_seed = -1;
		while (!is_min) {
			const uint64_t len = ds.len[ds.id_of_row[last]];
			const auto bounds = bv.get_range((uint64_t)(len * sim), (uint64_t)(len / sim));
			mc_scan_result res;
			res.n_eval = 0; res.n_pos = 0; res.best_row = -1; res.best_f0 = -1;
			int64_t lo = 0, hi = -1, nearest = -1;
			if (bv.trip_count(bounds.first, bounds.second) > 0) {
				lo = bv.row_at(bounds.first);
				hi = bv.row_at(bounds.second);
			}
			if (c.model.align) {
				if (hi >= lo) {
					marks.resize((size_t)(hi - lo + 1));
					align_scan(c, bv, cache, last, lo, hi, alive, res, marks);
				}
			} else {
				// get_close + remove_available + get_mean in one submission: the marked rows come back as
				// a list, `current` and its running bin sums stay in HBM
				mc_step_result sr;
				if (sharded) GPU(mc_accumulate_step_sharded(c.ranks.data(), world, last, lo, hi, first_mean ? 1 : 0, &sr, marked_rows.data(), (int64_t)marked_rows.size()));
				else GPU(mc_accumulate_step(c.gpu, last, lo, hi, first_mean ? 1 : 0, &sr, marked_rows.data(), (int64_t)marked_rows.size()));
				res = sr.scan;
				nearest = sr.nearest_row;
			}
			if (hi >= lo) { scans++; evals += res.n_eval; }
			is_min = res.n_pos == 0;
			if (is_min) {
				// no close point left: the arg-max of f0 becomes the next seed (or the first point)
				if (res.best_row < 0) next_seed = bv.pop();
				else { next_seed = res.best_row; bv.erase_row(res.best_row); }
				if (next_seed >= 0) { kill_row(next_seed); n_alive--; }
			} else if (c.model.align) {
				const size_t prev = current.size();
				bv.remove_marked(bounds.first.bin, bounds.second.bin,
				                 [&](int64_t r) { return r >= lo && r <= hi && marks[(size_t)(r - lo)] != 0; }, current);
				// get_mean over all of `current` (ClusterFactory.cpp:382-425)
				if (first_mean) GPU(mc_mean_nearest(c.gpu, current.data(), (int64_t)current.size(), 0, &nearest, nullptr));
				else GPU(mc_mean_nearest(c.gpu, current.data() + prev, (int64_t)(current.size() - prev), 1, &nearest, nullptr));
				first_mean = false;
				last = nearest;
			} else {
				bv.remove_rows(marked_rows.data(), (size_t)res.n_pos);
				n_alive -= res.n_pos;
				current.insert(current.end(), marked_rows.begin(), marked_rows.begin() + res.n_pos);
				first_mean = false;
				last = nearest;
			}
		}
		Cluster cl;
		cl.center_row = last;
		cl.rows.swap(current);
		part.push_back(std::move(cl));
		last = next_seed;
	}
	printf("Accumulation: %zu clusters, %lld scans, %lld evals, %d row compactions (host %.3fs, gpu calls %.3fs)  [%.2fs]\n", part.size(), (long long)scans, (long long)evals, compactions, compact_host_s, compact_gpu_s, tm.lap());
	}

	// ---------------- Phase B: update + merge (ClusterFactory.cpp:733-753) -----------------------
	int iters_run = 0;
	for (int iter = 0; iter < c.opt.iterations; iter++) {
		const int64_t nc = (int64_t)part.size();
		if (nc == 0) break;
		// mean_shift_update for every center (Jacobi sweep: each j only rewrites its own center)
		std::vector<int64_t> cand, off((size_t)nc + 1, 0), centers((size_t)nc), cb((size_t)nc), ce((size_t)nc), next((size_t)nc, -1);
		for (int64_t j = 0; j < nc; j++) {
			off[j + 1] = off[j] + (int64_t)part[j].rows.size();
			cand.insert(cand.end(), part[j].rows.begin(), part[j].rows.end());
			centers[j] = part[j].center_row;
		}
		for (int64_t j = 0; j < nc; j++) {
			cb[j] = off[std::max<int64_t>(0, j - delta)];
			ce[j] = off[std::min<int64_t>(j + delta, nc - 1) + 1];
		}
		bool changed = false;
		if (!c.model.align) {
			GPU(mc_update_centers(c.gpu, centers.data(), nc, cand.data(), (int64_t)cand.size(), cb.data(), ce.data(), next.data()));
			for (int64_t j = 0; j < nc; j++)
				if (next[j] >= 0 && next[j] != part[j].center_row) { part[j].center_row = next[j]; changed = true; }
		} else {
			// --align: a Center holds a CLONE, and DivergencePoint::clone() does not
			// copy the sequence (DivergencePoint.h:37-43, Center.h:14).  Feature::align therefore aligns
			// a member against an EMPTY string (identity 0/len = 0) unless that id pair was cached in
			// Phase A, and caches whatever it computed.  No alignment is needed here, only the cache.
			for (int64_t j = 0; j < nc; j++) {
				const int64_t cid = ds.id_of_row[part[j].center_row];
				std::vector<int64_t> good;
				for (int64_t q = cb[j]; q < ce[j]; q++) {
					const int64_t pid = ds.id_of_row[cand[q]];
					const uint64_t key = AlignCache::key_of(pid, cid);
					const double *hit = cache.find(key);
					double v;
					if (hit) v = *hit;
					else { v = ds.len[pid] > 0 ? 0.0 : std::nan(""); cache.set(key, v); }
					const double sum = std::fma(c.model.w[1], (v - 0.0) / (1.0 - 0.0), c.model.w[0]);
					if (std::round(1.0 / (1 + std::exp(-sum))) == 1.0) good.push_back(cand[q]);
				}
				if (good.empty()) continue;
				int64_t nearest = -1;
				GPU(mc_mean_nearest(c.gpu, good.data(), (int64_t)good.size(), 0, &nearest, nullptr));
				if (nearest >= 0 && nearest != part[j].center_row) part[j].center_row = nearest;
			}
		}
		// merge (ClusterFactory.cpp:427-493 with Trainer::merge, Trainer.cpp:129-157): the pair
		// evaluations of a pass do not depend on the merges of that pass, so they go in one batch
		std::vector<int32_t> pa, pb;
		std::vector<int64_t> first_pair((size_t)nc + 1, 0);
		for (int64_t i = 0; i < nc; i++) {
			const int64_t lastj = std::min<int64_t>(nc - 1, i + delta);
			for (int64_t t = i + 1; t <= lastj; t++) { pa.push_back((int32_t)part[t].center_row); pb.push_back((int32_t)part[i].center_row); }
			first_pair[i + 1] = (int64_t)pa.size();
		}
		std::vector<double> f0(pa.size());
		std::vector<uint8_t> fl(pa.size());
		if (!pa.empty() && !c.model.align) GPU(mc_pair_classify(c.gpu, pa.data(), pb.data(), (int64_t)pa.size(), nullptr, f0.data(), fl.data(), nullptr));
		if (c.model.align) {
			// Trainer::merge on two clones: both strings are empty -> 0/0 = NaN unless the pair is cached
			for (size_t q = 0; q < pa.size(); q++) {
				const int64_t ia = ds.id_of_row[pa[q]], ib = ds.id_of_row[pb[q]];
				const uint64_t key = AlignCache::key_of(ia, ib);
				const double *hit = cache.find(key);
				double v;
				if (hit) v = *hit;
				else { v = std::nan(""); cache.set(key, v); }
				const double sum = std::fma(c.model.w[1], (v - 0.0) / (1.0 - 0.0), c.model.w[0]);
				f0[q] = v;
				fl[q] = std::round(1.0 / (1 + std::exp(-sum))) == 1.0;
			}
		}
		for (int64_t i = 0; i < nc; i++) {
			std::pair<long, double> best(0, DBL_MIN);   // Trainer.cpp:135 initial value: DBL_MIN is positive
			for (int64_t q = first_pair[i]; q < first_pair[i + 1]; q++) {
				const long t = (long)(i + 1 + (q - first_pair[i]));
				if (fl[q]) best = best.second > f0[q] ? best : std::make_pair(t, f0[q]);
			}
			if (best.first > i) {
				auto &to_add = part[best.first].rows;
				to_add.insert(to_add.end(), part[i].rows.begin(), part[i].rows.end());
				part[i].removed = true;
				changed = true;
			}
		}
		part.erase(std::remove_if(part.begin(), part.end(), [](const Cluster &p) { return p.removed; }), part.end());
		// An iteration is a pure function of (centers, member lists): once one changes nothing, the
		// remaining ones cannot either, and the reference's fixed count of iterations (:733) is spent on
		// identical no-ops.  (--align keeps going: its id-pair cache changes between iterations.)
		if (!changed && !c.model.align) { iters_run = iter + 1; break; }
		iters_run = iter + 1;
	}
	printf("Update: %zu clusters after %d of %d iterations (fixed point)  [%.2fs]\n", part.size(), iters_run, c.opt.iterations, tm.lap());
	if (!c.model.align) {
		// the only decisions a differently rounded exp() could turn: GLM sums within 1e-9 of the threshold
		int64_t near = 0;
		GPU(mc_near_threshold_count(c.gpu, &near, 0));
		printf("Pairs within 1e-9 of the decision threshold: %lld\n", (long long)near);
	}
	write_clstr(c, part);
}

}  // namespace

// ----------------------------------------------------------------------------------------------
int run_pipeline(Options opt) {
	Ctx c;
	c.opt = opt;
	Timer total, tm;
#ifdef _OPENMP
	if (opt.threads > 0) omp_set_num_threads(opt.threads);
#endif
	// the CUDA context (driver start-up, ~1 s) is created on a helper thread while the files are read
	int ctx_rc = MC_OK;
	std::string ctx_err;
	double ctx_s = 0;
	// --gpus N shards the Phase-A scans (host-driven steps, summaries over NVLink): that pays when one scan is
	// long -- hundreds of microseconds, i.e. gigabytes of histograms.  Below that the single persistent kernel on
	// one GPU is faster than any exchange per step, and every additional CUDA context costs start-up time: small
	// inputs stay on one GPU.  The input size is known before the first CUDA call (file sizes), the histogram
	// size is not: the rule is on the bytes of FASTA (MC_SHARD_MIN_BYTES, default 3 GB; tests force sharding
	// with MC_PHASE_A_STEPS).
	if (opt.gpus > 1) {
		unsigned long long total_bytes = 0;
		for (const std::string &f : opt.files) {
			struct stat stt;
			if (stat(f.c_str(), &stt) == 0) total_bytes += (unsigned long long)stt.st_size;
		}
		const unsigned long long min_bytes = getenv("MC_SHARD_MIN_BYTES") ? strtoull(getenv("MC_SHARD_MIN_BYTES"), nullptr, 10) : 3000000000ull;
		c.shard_phase_a = getenv("MC_PHASE_A_STEPS") != nullptr || total_bytes >= min_bytes;
		// The alignments of the training stage (K4) are split by pairs over the GPUs; that pays when a pair is
		// millions of cells, i.e. for sequences of several thousand letters.  The record length is estimated from
		// the head of the first file (bytes per '>' line) -- all that is known before the first CUDA call.
		double est_len = 0;
		if (!opt.files.empty()) {
			FILE *f = fopen(opt.files[0].c_str(), "rb");
			if (f) {
				std::vector<char> head((size_t)1 << 20);
				const size_t got = fread(head.data(), 1, head.size(), f);
				fclose(f);
				size_t recs = 0;
				for (size_t i = 0; i < got; i++) recs += head[i] == '>' && (i == 0 || head[i - 1] == '\n');
				est_len = recs ? (double)got / (double)recs : (double)got;
			}
		}
		// (the training stage aligns ~6000 pairs whatever n is: 0.9 s on one B200 at 10 kb per record, growing with the
		// square of the length.  A second CUDA context costs 0.3 - 3 s to create and slows the first GPU's allocations
		// while it comes up: measured on C5 (10 kb), two GPUs saved 0.36 s of alignments and lost 0.5 s elsewhere.)
		const double min_len = getenv("MC_ALIGN_SHARD_MIN_LEN") ? atof(getenv("MC_ALIGN_SHARD_MIN_LEN")) : 20000.0;
		const bool share_alignments = est_len >= min_len || opt.align;
		if (!c.shard_phase_a && !share_alignments) {
			printf("  [--gpus %d: %.2f GB of input, records of ~%.0f letters: one GPU (persistent Phase-A kernel, alignment batches too small to split)]\n",
			       opt.gpus, (double)total_bytes * 1e-9, est_len);
			opt.gpus = 1;
			c.opt.gpus = 1;
		} else if (!c.shard_phase_a) {
			printf("  [--gpus %d: alignments are split over the GPUs, Phase A stays on one (%.2f GB of input)]\n", opt.gpus, (double)total_bytes * 1e-9);
		}
	}
	// CUDA start-up time grows with the number of GPUs the driver has to initialise: expose only the
	// ones this run uses (unless the user has chosen a set already)
	if (!getenv("CUDA_VISIBLE_DEVICES")) {
		std::string vis;
		for (int r = 0; r < std::max(1, opt.gpus); r++) vis += (r ? "," : "") + std::to_string(opt.device + r);
		setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
		opt.device = 0;
	}
	c.ranks.assign((size_t)std::max(1, opt.gpus), nullptr);
	std::promise<void> ctx_up;
	std::shared_future<void> ctx_ready = ctx_up.get_future().share();
	std::thread ctx_thread([&]() {
		Timer t;
		const int ndev = std::max(1, mc_device_count());
		// rank r sits on the r-th GPU after --device; with fewer GPUs than ranks they share devices
		ctx_rc = mc_ctx_create(&c.ranks[0], opt.device % ndev);
		if (ctx_rc != MC_OK) ctx_err = mc_last_error();
		c.gpu = c.ranks[0];
		ctx_s = t.lap();
		ctx_up.set_value();
	});
	// the other ranks are needed only when the first large alignment batch or a sharded Phase A starts: their
	// contexts come up in the background and are joined there -- AFTER rank 0's, which everything waits for
	// (contexts created at the same time share the driver's start-up lock: rank 0 took 5.8 s instead of 1.6 s)
	if (c.ranks.size() > 1)
		c.ranks_thread = std::thread([&c, opt, ctx_ready]() {
			ctx_ready.wait();
			const int ndev = std::max(1, mc_device_count());
			std::vector<std::thread> more;
			std::vector<int> rcs(c.ranks.size(), MC_OK);
			std::vector<std::string> errs(c.ranks.size());
			for (size_t r = 1; r < c.ranks.size(); r++)
				more.emplace_back([&, r]() {
					rcs[r] = mc_ctx_create(&c.ranks[r], (opt.device + (int)r) % ndev);
					if (rcs[r] != MC_OK) errs[r] = mc_last_error();
				});
			for (auto &th : more) th.join();
			for (size_t r = 1; r < c.ranks.size(); r++)
				if (rcs[r] != MC_OK && c.ranks_rc == MC_OK) { c.ranks_rc = rcs[r]; c.ranks_err = errs[r]; }
		});
	struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{ctx_thread};
	// ---- read (Runner.cpp:43-52, ChromListMaker) ------------------------------------------------
	std::vector<size_t> file_first;
	for (const std::string &f : opt.files) {
		if (access(f.c_str(), F_OK) == -1) {
			fprintf(stderr, "File \"%s\" does not exist\n", f.c_str());
			ctx_thread.join();
			exit(1);
		}
	}
	// Well-formed files with LF line ends are only INDEXED here (headers, where every record's sequence lines are,
	// how many letters they hold): the bytes go to the GPU as they are and mc_ingest_fasta squeezes the line feeds
	// out there.  Anything else (CR line ends, malformed records) takes the host parser, which owns the error
	// semantics.  MC_HOST_PARSE=1 forces the host parser.
	FastaIndex fidx;
	const bool indexed = !getenv("MC_HOST_PARSE") && index_fasta_files(opt.files, fidx);
	// the raw bytes go up on a helper thread as soon as the context exists, while this thread derives the row order
	std::thread stage_thread;
	int stage_rc = MC_OK;
	double stage_s = 0;
	if (indexed) {
		const int64_t nrec = (int64_t)fidx.size();   // (the headers move out of the index below)
		stage_thread = std::thread([&, nrec]() {
			ctx_ready.wait();
			if (ctx_rc != MC_OK) return;
			Timer t;
			// one scratch allocation for the whole run where its size can be foreseen: Phase A's working set with the
			// staging copies of the row compactions is ~1.7 x the histograms (8-bit bins assumed), the ingest needs the
			// file bytes
			if (opt.k >= 1 && opt.k <= 6 && !opt.align) {
				const double rows_bytes = (double)nrec * ((double)(1ll << (2 * opt.k)) + 128.0);
				mc_reserve_scratch(c.gpu, (int64_t)std::max(1.75 * rows_bytes + 64e6, (double)fidx.raw.size() + 32.0 * (double)nrec + 1e6));
			}
			stage_rc = mc_stage_fasta_bytes(c.gpu, fidx.raw.data(), (int64_t)fidx.raw.size(), nrec);
			stage_s = t.lap();
		});
	}
	if (indexed) {
		file_first = fidx.file_first;
		c.ds.fa.headers.swap(fidx.headers);
		c.ds.fa.offsets.assign(fidx.letters.size() + 1, 0);
		for (size_t i = 0; i < fidx.letters.size(); i++) c.ds.fa.offsets[i + 1] = c.ds.fa.offsets[i] + fidx.letters[i];
	} else {
		fidx.clear();
		for (const std::string &f : opt.files) {
			std::string msg;
			file_first.push_back(c.ds.fa.size());
			if (!read_fasta(f, c.ds.fa, msg)) {
				fprintf(stderr, "meshclust: %s\n", msg.c_str());
				ctx_thread.join();
				abort();   // the reference dies with an uncaught exception on malformed input
			}
		}
		file_first.push_back(c.ds.fa.size());
	}
	Dataset &ds = c.ds;
	ds.n = (int64_t)ds.fa.size();
	ds.len.resize((size_t)ds.n);
	for (int64_t i = 0; i < ds.n; i++) ds.len[i] = (uint64_t)(ds.fa.offsets[i + 1] - ds.fa.offsets[i]);
	printf("Read %lld sequences  [%.2fs]\n", (long long)ds.n, tm.lap());
	if (!opt.align && opt.similarity >= 0.6) c.by_length = std::async(std::launch::async, [&ds]() { return sort_ids_by_length(ds.len); });

	// ---- k (Runner.cpp:265-292 find_k) ----------------------------------------------------------
	c.k = opt.k;
	if (c.k == -1) {
		unsigned long long length = 0;
		for (size_t fi = 0; fi + 1 < file_first.size(); fi++) {
			unsigned long long l = 0;
			for (size_t i = file_first[fi]; i < file_first[fi + 1]; i++) l += ds.len[i];
			l /= (file_first[fi + 1] - file_first[fi]);
			length += l;
		}
		length /= opt.files.size();
		c.k = (int)std::ceil(std::log((double)length) / std::log(4)) - 1;
		printf("avg length: %llu\nRecommended K: %d\n", length, c.k);
	}
	if (opt.similarity < 0.6) c.opt.align = true;   // Runner.cpp:32-34
	if (c.opt.sample_size == 0) c.opt.sample_size = 3000;
	srand(10);

	// ---- bvec layout from the lengths alone (Runner.cpp:342-350; bvec.cpp) ----------------------
	const bool dbg_t = getenv("MC_DEBUG_TIMING") != nullptr;
	Timer tsub;
	auto sub = [&](const char *what) { if (dbg_t) fprintf(stderr, "  [rows: %-28s %.3f s]\n", what, tsub.lap()); };
	BVec bv(ds.len, 1000);
	sub("bvec bounds (sort lengths)");
	for (int64_t i = 0; i < ds.n; i++) bv.insert(i, ds.len[i]);
	sub("bvec insert");
	bv.finalize();
	sub("bvec per-bin sorts");
	ds.id_of_row = bv.assign_rows();
	ds.row_of_id.assign((size_t)ds.n, -1);
	for (int64_t r = 0; r < ds.n; r++) ds.row_of_id[ds.id_of_row[r]] = r;
	sub("row numbering");

	// ---- upload in row order, encode, histograms (K1) -------------------------------------------
	{
		std::vector<int64_t> offs((size_t)ds.n + 1, 0), seg_off((size_t)ds.n + 1, 0);
		for (int64_t r = 0; r < ds.n; r++) offs[r + 1] = offs[r] + (int64_t)ds.len[ds.id_of_row[r]];
		auto join_ctx = [&]() {
			ctx_thread.join();
			if (ctx_rc != MC_OK) {
				fprintf(stderr, "meshclust: mc_ctx_create failed: %s\n", ctx_err.c_str());
				exit(2);
			}
			printf("  [gpu context %.2fs on a helper thread, waited %.2fs]\n", ctx_s, tm.lap());
		};
		auto no_sequence = [&](int64_t bad_row) {
			fprintf(stderr, "meshclust: record \"%s\" has no usable sequence (the reference throws std::out_of_range)\n", ds.fa.headers[ds.id_of_row[bad_row]].c_str());
			if (ctx_thread.joinable()) ctx_thread.join();
			abort();
		};
		if (indexed) {
			// ---- device-side ingest: spans in row order, letters squeezed and permuted on the GPU
			std::vector<int64_t> sb((size_t)ds.n), se((size_t)ds.n);
			for (int64_t r = 0; r < ds.n; r++) { sb[r] = fidx.span_begin[ds.id_of_row[r]]; se[r] = fidx.span_end[ds.id_of_row[r]]; }
			printf("  [row order %.2fs]\n", tm.lap());
			join_ctx();
			stage_thread.join();
			if (stage_rc == MC_OK) printf("  [file bytes sent ahead in %.2fs on a helper thread, waited %.2fs]\n", stage_s, tm.lap());
			else fprintf(stderr, "meshclust: sending the file bytes ahead failed (%s); they go up with the ingest\n", mc_last_error());
			std::vector<uint8_t> rflags((size_t)ds.n);
			if (mc_ingest_fasta(c.gpu, fidx.raw.data(), (int64_t)fidx.raw.size(), sb.data(), se.data(), offs.data(), ds.n, rflags.data()) != MC_OK) die_gpu("mc_ingest_fasta");
			// segments (Chromosome.cpp:162-258): a record without N is one run -- kept from 20 letters on, cut at 1 Mbp --
			// and needs no look at its letters; the others are rebuilt from their lines and go through mc_host_segments
			// (two passes over the rows, shared by the host threads: count, prefix sum, fill)
			bool validate = false;
			int64_t bad_row = -1;
			std::vector<int32_t> nseg((size_t)ds.n, 0);
			auto whole_run = [](int64_t len, int32_t *out) -> int {   // segments of a record without N; out may be null (count only)
				if (len < 20) return 0;
				if (len <= 1000000) {
					if (out) { out[0] = 0; out[1] = (int32_t)(len - 1); }
					return 1;
				}
				const int64_t frag = len / 1000000;
				for (int64_t h = 0; h < frag; h++) {
					const int64_t fs = h * 1000000, fe = (h == frag - 1) ? len - 1 : fs + 1000000 - 1;
					if (out) { out[2 * h] = (int32_t)fs; out[2 * h + 1] = (int32_t)fe; }
				}
				return (int)frag;
			};
			auto letters_of = [&](int64_t r, std::vector<uint8_t> &tmp) {
				const int64_t len = offs[r + 1] - offs[r];
				tmp.resize((size_t)len);
				size_t w = 0;
				const uint8_t *rawp = fidx.raw.data();
				for (int64_t p = sb[r]; p < se[r]; p++)
					if (rawp[p] != '\n') tmp[w++] = rawp[p];
			};
#pragma omp parallel
			{
				std::vector<uint8_t> tmp;
				bool val_local = false;
				int64_t bad_local = -1;
#pragma omp for schedule(dynamic, 4096)
				for (int64_t r = 0; r < ds.n; r++) {
					const int64_t len = offs[r + 1] - offs[r];
					if (rflags[r] & 2) val_local = true;
					int ns;
					if (!(rflags[r] & 1)) ns = len <= 1 ? -1 : whole_run(len, nullptr);   // (mc_host_segments: an empty or one-letter record never closes a run)
					else {
						letters_of(r, tmp);
						ns = mc_host_segments(tmp.data(), len, nullptr, 0);
					}
					if (ns < 0) { if (bad_local < 0 || r < bad_local) bad_local = r; ns = 0; }
					nseg[(size_t)r] = ns;
				}
#pragma omp critical
				{
					if (val_local) validate = true;
					if (bad_local >= 0 && (bad_row < 0 || bad_local < bad_row)) bad_row = bad_local;
				}
			}
			if (bad_row >= 0) no_sequence(bad_row);
			for (int64_t r = 0; r < ds.n; r++) seg_off[r + 1] = seg_off[r] + nseg[(size_t)r];
			std::vector<int32_t> segs((size_t)seg_off[ds.n] * 2);
#pragma omp parallel
			{
				std::vector<uint8_t> tmp;
#pragma omp for schedule(dynamic, 4096)
				for (int64_t r = 0; r < ds.n; r++) {
					if (nseg[(size_t)r] == 0) continue;
					int32_t *dst = segs.data() + 2 * seg_off[r];
					const int64_t len = offs[r + 1] - offs[r];
					if (!(rflags[r] & 1)) whole_run(len, dst);
					else {
						letters_of(r, tmp);
						mc_host_segments(tmp.data(), len, dst, nseg[(size_t)r]);
					}
				}
			}
			if (mc_load_segments(c.gpu, segs.data(), seg_off.data(), validate ? 1 : 0) != MC_OK) {
				fprintf(stderr, "meshclust: %s\n", mc_last_error());
				abort();   // InvalidInputException in the reference
			}
			fidx.clear();
		} else {
		// letters in row order + the non-N segment list of every row; rows are independent, so the
		// host threads share them (segments: first pass counts, second pass fills)
		RawBytes letters;
		letters.resize(ds.fa.letters.size());
		constexpr int INLINE_SEGS = 4;
		std::vector<int32_t> seg_inline((size_t)ds.n * 2 * INLINE_SEGS);
		std::vector<int32_t> nseg((size_t)ds.n);
		int64_t bad_row = -1;
#pragma omp parallel for schedule(static)
		for (int64_t r = 0; r < ds.n; r++) {
			const int64_t id = ds.id_of_row[r];
			const uint8_t *src = ds.fa.letters.data() + ds.fa.offsets[id];
			const int64_t len = (int64_t)ds.len[id];
			memcpy(letters.data() + offs[r], src, (size_t)len);
			nseg[r] = mc_host_segments(src, len, seg_inline.data() + (size_t)r * 2 * INLINE_SEGS, INLINE_SEGS);
			if (nseg[r] < 0) {
#pragma omp critical
				if (bad_row < 0 || r < bad_row) bad_row = r;
			}
		}
		if (bad_row >= 0) no_sequence(bad_row);
		sub("letters to row order + segments");
		for (int64_t r = 0; r < ds.n; r++) seg_off[r + 1] = seg_off[r] + nseg[r];
		std::vector<int32_t> segs((size_t)seg_off[ds.n] * 2);
#pragma omp parallel for schedule(static)
		for (int64_t r = 0; r < ds.n; r++) {
			int32_t *dst = segs.data() + 2 * seg_off[r];
			if (nseg[r] <= INLINE_SEGS) memcpy(dst, seg_inline.data() + (size_t)r * 2 * INLINE_SEGS, (size_t)nseg[r] * 2 * sizeof(int32_t));
			else {
				const int64_t id = ds.id_of_row[r];
				mc_host_segments(ds.fa.letters.data() + ds.fa.offsets[id], (int64_t)ds.len[id], dst, nseg[r]);
			}
		}
		printf("  [row order + segments %.2fs]\n", tm.lap());
		join_ctx();
		if (mc_load_sequences(c.gpu, letters.data(), offs.data(), ds.n, segs.data(), seg_off.data()) != MC_OK) {
			fprintf(stderr, "meshclust: %s\n", mc_last_error());
			abort();   // InvalidInputException in the reference
		}
		}
	}
	printf("  [upload + encode %.2fs]\n", tm.lap());
	uint64_t largest = 0;
	GPU(mc_build_histograms(c.gpu, c.k, 0, &c.tbytes, &largest));
	printf("Using %d bit histograms\n", c.tbytes * 8);   // Runner.cpp:75-89
	printf("Counted %d-mers, largest count %llu  [%.2fs]\n", c.k, (unsigned long long)largest, tm.lap());

	// ---- train, cluster, write ------------------------------------------------------------------
	trainer_train(c);
	if (!opt.dump_model.empty()) {
		FILE *f = fopen(opt.dump_model.c_str(), "w");
		if (f) {
			fprintf(f, "nfeat %d\n", c.model.nfeat);
			for (int i = 0; i < 5; i++) fprintf(f, "bounds %d %.17g %.17g\n", i, c.model.mins[i], c.model.maxs[i]);
			for (int i = 0; i < 5; i++) fprintf(f, "weight %d %.17g\n", i, c.model.w[i]);
			fclose(f);
		}
	}
	mean_shift(c, bv);
	if (c.ranks.size() > 1) printf("  [alignment batches split over the %zu GPUs: %ld]\n", c.ranks.size(), c.align_split_calls);
	printf("Total %.2fs\n", total.lap());
	if (getenv("MC_CLEAN_EXIT")) { join_ranks(c); for (mc_ctx *g : c.ranks) mc_ctx_destroy(g); return 0; }
	// the output file is closed: skip the teardown of the CUDA context and of GBs of host vectors
	fflush(stdout);
	fflush(stderr);
	_exit(0);
}

// ----------------------------------------------------------------------------------------------
void usage(const std::string &prog) {
	printf("Usage: %s *.fasta [--id 0.90] [--kmer 3] [--delta 5] [--output output.clstr] [--iterations 20] [--align] [--sample 3000] [--pivot 40] [--threads TMAX]\n\n", prog.c_str());
	printf("meshclust_b200: B200-native hot path behind the MeShClust 1.2.0 command line.\n\n"
	       "--id          identity threshold (below 0.6 alignment is used automatically)\n"
	       "--kmer        k-mer size (default: from the average sequence length)\n"
	       "--delta       clusters looked at on each side in the final stage (default 5)\n"
	       "--output      output file, CD-HIT CLSTR format (default output.clstr)\n"
	       "--iterations  update/merge iterations (default 15)\n"
	       "--align       force alignment instead of k-mer features\n"
	       "--sample      number of sampled training pairs (default 3000)\n"
	       "--pivot       pairs per pivot sequence (default 20)\n"
	       "--threads     host threads\n"
	       "Any other argument is an input file.\n\n");
}

Options parse_options(int argc, char **argv) {
	Options o;
	auto need_long = [&](int &i, const char *what, bool allow_zero) -> long {
		errno = 0;
		const long v = strtol(argv[i + 1], NULL, 10);
		if (errno) { perror(argv[i + 1]); exit(EXIT_FAILURE); }
		if (allow_zero ? v < 0 : v <= 0) { fprintf(stderr, "%s must be greater than 0.\n", what); exit(EXIT_FAILURE); }
		i++;
		return v;
	};
	for (int i = 1; i < argc; i++) {
		const std::string arg = argv[i];
		const bool more = i + 1 < argc;
		if (arg == "--id" && more) {
			try {
				o.similarity = std::stod(argv[i + 1]);
				if (o.similarity <= 0 || o.similarity >= 1) throw std::invalid_argument("");
			} catch (const std::exception &) {
				fprintf(stderr, "Similarity must be between 0 and 1\n");
				exit(EXIT_FAILURE);
			}
			i++;
		} else if ((arg == "-k" || arg == "--kmer") && more) o.k = (int)need_long(i, "K", false);
		else if ((arg == "-o" || arg == "--output") && more) o.output = argv[++i];
		else if (arg == "-a" || arg == "--align") o.align = true;
		else if ((arg == "-s" || arg == "--sample") && more) o.sample_size = (int)need_long(i, "Sample size", false);
		else if ((arg == "-p" || arg == "--pivot") && more) o.pivots = (int)need_long(i, "Points per pivot", false);
		else if ((arg == "-t" || arg == "--threads") && more) {
			try {
				o.threads = std::stoi(argv[i + 1]);
				if (o.threads <= 0) throw std::invalid_argument("");
			} catch (const std::exception &) {
				fprintf(stderr, "Number of threads must be greater than 0.\n");
				exit(1);
			}
			i++;
		} else if ((arg == "-d" || arg == "--delta") && more) o.delta = (int)need_long(i, "Delta", true);
		else if ((arg == "-i" || arg == "--iter" || arg == "--iterations") && more) o.iterations = (int)need_long(i, "Iterations", false);
		else if (arg == "--device" && more) o.device = atoi(argv[++i]);          // extension: GPU ordinal
		else if (arg == "--gpus" && more) {                                      // extension: GPUs sharing the scans
			o.gpus = atoi(argv[++i]);
			if (o.gpus < 1 || o.gpus > 8) { fprintf(stderr, "--gpus must be between 1 and 8\n"); exit(EXIT_FAILURE); }
		}
		else if (arg == "--dump-model" && more) o.dump_model = argv[++i];        // extension: test hook
		else {
			struct stat st;
			if (stat(argv[i], &st) == 0 && S_ISREG(st.st_mode)) o.files.push_back(argv[i]);
			else { usage(argv[0]); exit(EXIT_FAILURE); }
		}
	}
	if (o.files.empty()) { usage(argv[0]); exit(EXIT_FAILURE); }
	// Runner.cpp:253-262: files ordered by basename
	std::sort(o.files.begin(), o.files.end(), [](const std::string &a, const std::string &b) {
		std::string as = a, bs = b;
		const std::string ab = basename(&as[0]), bb = basename(&bs[0]);
		return ab < bb;
	});
	return o;
}

}  // namespace mch
