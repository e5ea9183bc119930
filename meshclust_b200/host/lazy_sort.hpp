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
#include <utility>
#include <vector>

namespace mch {

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
