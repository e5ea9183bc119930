// Elements of the array std::sort would leave, without sorting all of it.
//
// Trainer::split (Trainer.cpp:694-721) sorts all n points by their distance to each of ~150 pivots with an
// UNSTABLE std::sort and then looks at ~40 positions of every sorted array (a binary search driven by alignments,
// twenty strided picks).  Which of several points at the same distance sits at a position is decided by the
// permutation libstdc++'s introsort applies, so the positions cannot come from a selection algorithm of our own --
// but they can come from introsort itself, run lazily: std::sort is
//     __introsort_loop: while (size > 16) { depth 0 -> heapsort the range; else median-of-3 to the front,
//                       Hoare partition around it, recurse into the right part, continue with the left }
//     __final_insertion_sort: every element moves left past strictly larger ones
// and the partition of a range depends on nothing but the range's own content.  A position is therefore resolved
// by partitioning only the ranges that contain it, down to a leaf of at most 16 elements; the final insertion sort
// never moves an element out of its leaf (left part <= pivot <= right part, and it stops at equal elements), so a
// stable insertion sort of the leaf is all that is left of it.  Ranges partitioned once stay partitioned: the
// tree of cuts is kept, later positions reuse it.  tests/units/lazy_sort_selfcheck.cpp compares every position
// with std::sort on the same input.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <utility>
#include <vector>

namespace mch {

namespace detail {
// std::__unguarded_partition(first + 1, last, pivot = *first) without its data-dependent branches.
// The library's loop -- scan up to an element that is not less than the pivot, scan down to one the pivot is not less
// than, swap, repeat until the scans meet -- mispredicts on every other element of random keys (~6 ns per element).
// Its k-th swap always exchanges the k-th element from the left that is not less than the pivot with the k-th from the
// right that the pivot is not less than (elements between the two scans are still where they were), so both position
// lists can be written down first, in two passes without a branch on the data, and the swaps replayed from them:
//   up[k]   ascending positions i of [first + 1, last) with !(v[i] < pivot)
//   down[k] descending positions j of the same range with !(pivot < v[j])
// scan k stops at lo = min(up[k], down[k-1]) (the element the previous swap put at down[k-1] is a sentinel for it) and
// hi = max(down[k], up[k-1]) (likewise; `first`, the pivot itself, before any swap); !(lo < hi) ends the loop and lo
// is the cut.  tests/units/lazy_sort_selfcheck.cpp compares whole sorts built on this with std::sort.
template <class Rec, class Less>
size_t partition_by_lists(std::vector<Rec> &v, size_t first, size_t last, const Less &less, std::vector<uint32_t> &up, std::vector<uint32_t> &down) {
	const size_t n = last - (first + 1);
	if (up.size() < n + 1) up.resize(n + 1);
	if (down.size() < n + 1) down.resize(n + 1);
	const Rec piv = v[first];
	size_t nu = 0, nd = 0;
	uint32_t *pu = up.data(), *pd = down.data();
	for (size_t i = first + 1; i < last; i++) { pu[nu] = (uint32_t)i; nu += less(v[i], piv) ? 0 : 1; }
	for (size_t j = last; j-- > first + 1;) { pd[nd] = (uint32_t)j; nd += less(piv, v[j]) ? 0 : 1; }
	size_t prev_up = first;            // the pivot's own position stops the downward scan when nothing else does
	size_t prev_down = (size_t)-1;     // nothing stops the upward scan before the first swap but up[0] (median of three: it exists)
	for (size_t k = 0;; k++) {
		const size_t u = k < nu ? pu[k] : (size_t)-1, d = k < nd ? pd[k] : first;
		const size_t lo = u < prev_down ? u : prev_down;
		const size_t hi = d > prev_up ? d : prev_up;
		if (!(lo < hi)) return lo;
		std::swap(v[lo], v[hi]);
		prev_up = lo;
		prev_down = hi;
	}
}
// scratch lists of the calling thread (a partition is always run by one thread from start to end)
inline std::vector<uint32_t> &tls_up() { static thread_local std::vector<uint32_t> b; return b; }
inline std::vector<uint32_t> &tls_down() { static thread_local std::vector<uint32_t> b; return b; }
constexpr size_t LIST_PARTITION_MIN = 4096;   // below this the library's loop is as fast (the lists do not pay for themselves)
inline bool lists_enabled() { static const bool on = !(getenv("MC_LAZY_LISTS") && getenv("MC_LAZY_LISTS")[0] == '0'); return on; }
}  // namespace detail

template <class Rec, class Less>
class LazySort {
public:
	LazySort() = default;
	// depth_limit < 0: std::sort's own (2 * floor(log2 n)); tests pass small values to reach the heapsort branch
	LazySort(std::vector<Rec> &&v, Less less, int depth_limit = -1) : v_(std::move(v)), less_(less) {
		if (!v_.empty()) {
			int lg = 0;
			for (size_t s = v_.size(); s > 1; s >>= 1) lg++;   // std::__lg
			nodes_.push_back({0, v_.size(), depth_limit < 0 ? 2 * lg : depth_limit, UNTOUCHED, 0, -1, -1});
		}
	}
	size_t size() const { return v_.size(); }
	// the caller has sorted the array itself (with std::sort)
	void mark_sorted() { if (!nodes_.empty()) nodes_[0].state = SORTED; }

	// the element std::sort(begin, end, less) would put at `pos`
	const Rec &at(size_t pos) {
		int node = 0;
		for (;;) {
			// (indices, not references: the vector of nodes grows)
			const size_t first = nodes_[(size_t)node].first, last = nodes_[(size_t)node].last;
			if (nodes_[(size_t)node].state == SORTED) return v_[pos];
			if (last - first <= 16) {
				insertion_sort(first, last);
				nodes_[(size_t)node].state = SORTED;
				return v_[pos];
			}
			if (nodes_[(size_t)node].state == UNTOUCHED) {
				if (nodes_[(size_t)node].depth == 0) {
					std::partial_sort(v_.begin() + (ptrdiff_t)first, v_.begin() + (ptrdiff_t)last, v_.begin() + (ptrdiff_t)last, less_);   // __partial_sort: heap select + sort heap
					nodes_[(size_t)node].state = SORTED;
					return v_[pos];
				}
				const size_t cut = partition_pivot(first, last);
				const int d = nodes_[(size_t)node].depth - 1;
				const int l = (int)nodes_.size();
				nodes_.push_back({first, cut, d, UNTOUCHED, 0, -1, -1});
				nodes_.push_back({cut, last, d, UNTOUCHED, 0, -1, -1});
				Node &nd = nodes_[(size_t)node];
				nd.state = SPLIT; nd.cut = cut; nd.left = l; nd.right = l + 1;
			}
			const Node &nd = nodes_[(size_t)node];
			node = pos < nd.cut ? nd.left : nd.right;
		}
	}

	// everything: the rest of the sort (for callers that want the whole array after all)
	const std::vector<Rec> &all() {
		for (size_t i = 0; i < nodes_.size(); i++) {   // (the vector grows while leaves are resolved)
			const size_t first = nodes_[i].first, last = nodes_[i].last;
			if (nodes_[i].state == UNTOUCHED && last > first) at(first);
		}
		// at(first) resolves one path only; sweep until nothing is left untouched
		bool again = true;
		while (again) {
			again = false;
			for (size_t i = 0; i < nodes_.size(); i++)
				if (nodes_[i].state == UNTOUCHED && nodes_[i].last > nodes_[i].first) { at(nodes_[i].first); again = true; }
		}
		return v_;
	}

private:
	enum State { UNTOUCHED = 0, SPLIT = 1, SORTED = 2 };
	struct Node {
		size_t first, last;
		int depth;       // introsort's depth_limit on entry to this range
		int state;
		size_t cut;
		int left, right;
	};

	// std::__unguarded_partition_pivot
	size_t partition_pivot(size_t first, size_t last) {
		const size_t mid = first + (last - first) / 2;
		move_median_to_first(first, first + 1, mid, last - 1);
		if (last - first >= detail::LIST_PARTITION_MIN && last < ((size_t)1 << 32) && detail::lists_enabled())
			return detail::partition_by_lists(v_, first, last, less_, detail::tls_up(), detail::tls_down());
		// std::__unguarded_partition(first + 1, last, pivot = first)
		size_t lo = first + 1, hi = last;
		for (;;) {
			while (less_(v_[lo], v_[first])) ++lo;
			--hi;
			while (less_(v_[first], v_[hi])) --hi;
			if (!(lo < hi)) return lo;
			std::swap(v_[lo], v_[hi]);
			++lo;
		}
	}
	// std::__move_median_to_first
	void move_median_to_first(size_t result, size_t a, size_t b, size_t c) {
		if (less_(v_[a], v_[b])) {
			if (less_(v_[b], v_[c])) std::swap(v_[result], v_[b]);
			else if (less_(v_[a], v_[c])) std::swap(v_[result], v_[c]);
			else std::swap(v_[result], v_[a]);
		} else if (less_(v_[a], v_[c])) std::swap(v_[result], v_[a]);
		else if (less_(v_[b], v_[c])) std::swap(v_[result], v_[c]);
		else std::swap(v_[result], v_[b]);
	}
	// what __final_insertion_sort does to a leaf: stable, elements move left past strictly larger ones
	void insertion_sort(size_t first, size_t last) {
		for (size_t i = first + 1; i < last; i++) {
			Rec val = v_[i];
			size_t j = i;
			while (j > first && less_(val, v_[j - 1])) { v_[j] = v_[j - 1]; j--; }
			v_[j] = val;
		}
	}

	std::vector<Rec> v_;
	Less less_;
	std::vector<Node> nodes_;
};

// ---- the whole array, with the host threads: std::sort's result (ties included), not just a sorted array --------
// The same decomposition: after a range has been partitioned its two parts are independent, so they go to different
// threads (OpenMP tasks above a size threshold); leaves get their insertion sort where they fall.  Used for the two
// full sorts of Trainer::split, whose whole permutation is the input order of what follows.
namespace detail {
template <class Rec, class Less>
void introsort_tasks(std::vector<Rec> &v, size_t first, size_t last, int depth, const Less &less) {
	while (last - first > 16) {
		if (depth == 0) {
			std::partial_sort(v.begin() + (ptrdiff_t)first, v.begin() + (ptrdiff_t)last, v.begin() + (ptrdiff_t)last, less);
			return;
		}
		--depth;
		// std::__unguarded_partition_pivot
		const size_t mid = first + (last - first) / 2, a = first + 1, c = last - 1;
		if (less(v[a], v[mid])) {
			if (less(v[mid], v[c])) std::swap(v[first], v[mid]);
			else if (less(v[a], v[c])) std::swap(v[first], v[c]);
			else std::swap(v[first], v[a]);
		} else if (less(v[a], v[c])) std::swap(v[first], v[a]);
		else if (less(v[mid], v[c])) std::swap(v[first], v[c]);
		else std::swap(v[first], v[mid]);
		// (the library's loop here: with all threads partitioning at once the position lists cost more memory traffic
		// than the mispredictions they save -- 0.014 s per million records against 0.025 s)
		size_t lo = first + 1, hi = last;
		for (;;) {
			while (less(v[lo], v[first])) ++lo;
			--hi;
			while (less(v[first], v[hi])) --hi;
			if (!(lo < hi)) break;
			std::swap(v[lo], v[hi]);
			++lo;
		}
		const size_t cut = lo;
		if (last - cut > 32768 && cut - first > 32768) {
#pragma omp task default(none) shared(v, less) firstprivate(cut, last, depth)
			introsort_tasks(v, cut, last, depth, less);
		} else introsort_tasks(v, cut, last, depth, less);
		last = cut;
	}
	for (size_t i = first + 1; i < last; i++) {   // the leaf's share of __final_insertion_sort
		Rec val = v[i];
		size_t j = i;
		while (j > first && less(val, v[j - 1])) { v[j] = v[j - 1]; j--; }
		v[j] = val;
	}
}
}  // namespace detail

template <class Rec, class Less>
void parallel_std_sort(std::vector<Rec> &v, Less less) {
	if (v.size() < 2) return;
	int lg = 0;
	for (size_t s = v.size(); s > 1; s >>= 1) lg++;
	if (v.size() < 100000) { detail::introsort_tasks(v, 0, v.size(), 2 * lg, less); return; }
#pragma omp parallel
#pragma omp single nowait
	detail::introsort_tasks(v, 0, v.size(), 2 * lg, less);
}

}  // namespace mch
