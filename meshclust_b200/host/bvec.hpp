// Host-side index bookkeeping of the reference's length-binned container (bvec.cpp, bvec_iterator.h).
// It holds ROW numbers only; the histograms live in HBM in exactly this container's initial
// iteration order (bin by bin, position by position), so a bvec range is a contiguous row range on
// the GPU and "first in iteration order" is "smallest row".  Every quirk that decides which rows a
// scan sees is kept: least-filled candidate bin on insert with the middle one
// on ties, an in-bin binary search that converges to *an* element, empty-bin fall-backs that land
// on position 0 of the first / LAST non-empty bin, inclusive ranges.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <functional>
#include <vector>

#include "lazy_sort.hpp"

namespace mch {

struct BvIdx {
	size_t bin = 0, pos = 0;
};

// Layout: a bin keeps its entries where finalize() put them for good; removal clears a bit in the
// bin's alive bitmap, and "element i of the bin" (what the reference indexes after its erases) is
// the i-th set bit.  Removing the rows a scan marked therefore costs O(marked), not O(range), and
// the GPU row of an entry (= its rank in the initial iteration order) never changes.
class BVec {
public:
	// entries are point ids while the container is being filled, rows afterwards
	struct Entry {
		int64_t v;
		uint64_t len;
	};

	// bvec::bvec (bvec.cpp:10-24): bounds are every bin_size-th sorted length
	BVec(std::vector<uint64_t> lengths, uint64_t bin_size = 1000) {
		parallel_std_sort(lengths, std::less<uint64_t>());   // (plain values: any correct sort gives this array)
		for (uint64_t i = 0; i < lengths.size(); i += bin_size) bounds_.push_back(lengths[i]);
		data_.resize(bounds_.size());
	}

	// bvec::index_of (bvec.cpp:123-149).  The reference walks every bound and keeps the smallest and
	// largest i-1 (0 for i = 0) over all i with bounds[i-1] <= point <= bounds[i] (bounds[-1] = 0);
	// the bounds are sorted, so those i form the interval [first bound >= point, #bounds <= point]
	// and two binary searches give the same answer (index_of_linear is the literal loop, kept for
	// the self-check in tests/test_host_units.py).
	void index_of(uint64_t point, size_t *pfront, size_t *pback) const {
		const size_t nb = bounds_.size();
		size_t low = nb - 1, high = 0;
		const size_t i1 = (size_t)(std::lower_bound(bounds_.begin(), bounds_.end(), point) - bounds_.begin());
		if (i1 < nb) {
			const size_t ub = (size_t)(std::upper_bound(bounds_.begin(), bounds_.end(), point) - bounds_.begin());
			const size_t imax = std::min(ub, nb - 1);
			low = std::min(low, i1 > 0 ? i1 - 1 : 0);
			high = std::max(high, imax > 0 ? imax - 1 : 0);
		}
		if (point >= bounds_.back()) high = std::max(high, nb - 1);
		if (pfront) *pfront = low;
		if (pback) *pback = high;
	}
	void index_of_linear(uint64_t point, size_t *pfront, size_t *pback) const {
		size_t low = bounds_.size() - 1, high = 0;
		for (size_t i = 0; i < bounds_.size(); i++) {
			const size_t prev = i > 0 ? bounds_[i - 1] : 0;
			const size_t prev_index = i > 0 ? i - 1 : 0;
			if (point >= prev && point <= bounds_[i]) {
				low = std::min(low, prev_index);
				high = std::max(high, prev_index);
			}
		}
		if (point >= bounds_.back()) high = std::max(high, bounds_.size() - 1);
		if (pfront) *pfront = low;
		if (pback) *pback = high;
	}

	// bvec::insert (bvec.cpp:152-177)
	void insert(int64_t id, uint64_t len) {
		size_t front = 0, back = 0;
		index_of(len, &front, &back);
		// least-filled candidate bin, the middle one among equals (the reference collects the minima in
		// a vector and takes mins[mins.size() / 2]); two passes instead of a vector per insert
		// (the fill counts in one flat array: equal lengths make the candidate range tens of bins wide)
		if (fill_.size() != data_.size()) { fill_.resize(data_.size()); for (size_t i = 0; i < data_.size(); i++) fill_[i] = (uint32_t)data_[i].items.size(); }
		// three passes without a branch on the data (minimum, how many, where the middle one is): the single pass with
		// its `if (smaller) ... else if (equal)` mispredicts on every few bins
		const uint32_t *fl = fill_.data();
		uint32_t minimum = std::numeric_limits<uint32_t>::max();
#pragma omp simd reduction(min : minimum)
		for (size_t i = front; i <= back; i++) minimum = fl[i] < minimum ? fl[i] : minimum;
		uint32_t nmin = 0;
#pragma omp simd reduction(+ : nmin)
		for (size_t i = front; i <= back; i++) nmin += fl[i] == minimum ? 1u : 0u;
		// front > back leaves no candidate: the reference prints an error and then indexes an empty
		// vector (undefined); it cannot happen for bounds taken from the same lengths
		size_t pick = data_.size(), seen = 0;
		const size_t want = (size_t)nmin / 2;
		size_t i = front;
		for (; i + 16 <= back + 1; i += 16) {   // whole blocks of 16 bins that end before the wanted one
			uint32_t c = 0;
#pragma omp simd reduction(+ : c)
			for (size_t j = 0; j < 16; j++) c += fl[i + j] == minimum ? 1u : 0u;
			if (seen + c > want) break;
			seen += c;
		}
		for (; i <= back; i++) {
			const bool is_min = fl[i] == minimum;
			if (is_min && seen == want) { pick = i; break; }
			seen += is_min;
		}
		data_.at(pick).items.push_back({id, len});
		fill_[pick]++;
	}

	// bvec::insert_finalize (bvec.cpp:209-218): per-bin std::sort by length (unstable: the same
	// libstdc++ introsort on the same sequence and comparator gives the same permutation)
	void finalize() {
		// (bins are independent: the host threads share them)
#pragma omp parallel for schedule(dynamic, 8)
		for (long b = 0; b < (long)data_.size(); b++) {
			auto &bin = data_[(size_t)b];
			std::sort(bin.items.begin(), bin.items.end(), [](const Entry &a, const Entry &b) { return a.len < b.len; });
		}
	}

	// after finalize(): rename the entries to their position in iteration order; returns id per row
	std::vector<int64_t> assign_rows() {
		std::vector<int64_t> id_of_row;
		row0_.clear();
		for (auto &bin : data_) {
			row0_.push_back((int64_t)id_of_row.size());
			for (auto &e : bin.items) {
				id_of_row.push_back(e.v);
				e.v = (int64_t)id_of_row.size() - 1;
			}
			bin.alive = bin.items.size();
			bin.bits.assign((bin.items.size() + 63) / 64, ~0ull);
			if (bin.items.size() % 64) bin.bits.back() = (1ull << (bin.items.size() % 64)) - 1;
		}
		return id_of_row;
	}

	size_t size() const {
		size_t t = 0;
		for (auto &b : data_) t += b.alive;
		return t;
	}

	// bvec::pop (bvec.cpp:27-38): first element of the first non-empty bin, or -1
	int64_t pop() {
		for (; first_live_ < data_.size(); first_live_++) {
			Bin &bin = data_[first_live_];
			if (bin.alive) {
				const size_t at = bin.select(0);
				bin.clear(at);
				return bin.items[at].v;
			}
		}
		return -1;
	}

	// bvec::inner_index_of (bvec.cpp:52-120)
	void inner_index_of(uint64_t length, size_t &idx, size_t *pfront, size_t *pback) const {
		if (data_.at(idx).alive == 0) {
			if (pfront)
				for (size_t i = 0; i < data_.size(); i++)
					if (data_[i].alive) { idx = i; *pfront = 0; break; }
			if (pback)
				for (long i = (long)data_.size() - 1; i >= 0; i--)
					if (data_[i].alive) { idx = (size_t)i; *pback = 0; break; }
			return;
		}
		const Bin &bin = data_[idx];
		size_t front = 0, back = 0, low = 0, high = bin.alive - 1;
		while (low <= high) {
			const size_t mid = (low + high) / 2;
			const uint64_t d = bin.at(mid).len;
			if (d == length) { front = back = mid; break; }
			else if (length < d) high = mid;
			else low = mid + 1;
			if (low == high) { front = low; back = high; break; }
		}
		if (pfront) {
			for (long i = (long)front; i >= 0 && bin.at((size_t)i).len == length; i--) front = (size_t)i;
			*pfront = front;
		}
		if (pback) {
			for (size_t i = back; i < bin.alive && bin.at(i).len == length; i++) back = i;
			*pback = back;
		}
	}

	// bvec::get_range (bvec.cpp:247-278); both ends inclusive
	std::pair<BvIdx, BvIdx> get_range(uint64_t begin_len, uint64_t end_len) const {
		BvIdx front, back;
		back.bin = data_.size() - 1;
		back.pos = data_[back.bin].alive - 1;   // wraps to SIZE_MAX on an empty last bin, as in the reference
		index_of(begin_len, &front.bin, nullptr);
		index_of(end_len, nullptr, &back.bin);
		inner_index_of(begin_len, front.bin, &front.pos, nullptr);
		inner_index_of(end_len, back.bin, nullptr, &back.pos);
		return {front, back};
	}

	// trip count of `for (it = front; it <= back; ++it)` as OpenMP computes it from
	// bvec_iterator::operator- (bvec_iterator.h:61-76): (back - front) + 1, <= 0 means no iteration
	int64_t trip_count(const BvIdx &f, const BvIdx &b) const {
		return diff(b, f) + 1;
	}

	// row of an element; the caller guarantees it exists
	int64_t row_at(const BvIdx &i) const { return data_[i.bin].at(i.pos).v; }

	// bvec::erase (bvec.cpp:281-285) by row: the argmax element of the last scan
	void erase_row(int64_t row) {
		const size_t b = bin_of_row(row);
		data_[b].clear((size_t)(row - row0_[b]));
	}

	// bvec::remove_available (bvec.cpp:290-317) when the marked rows are already known as a list
	// (ascending rows = bin order, position order: the serial order of the reference)
	void remove_rows(const int64_t *rows, size_t m) {
		size_t b = 0;
		for (size_t i = 0; i < m; i++) {
			if (!(b < data_.size() && rows[i] >= row0_[b] && rows[i] < row0_[b] + (int64_t)data_[b].items.size())) b = bin_of_row(rows[i]);
			data_[b].clear((size_t)(rows[i] - row0_[b]));
		}
	}

	// bvec::remove_available (bvec.cpp:290-317) for bins [a,b]: drop every marked row; appends them
	// to `out` in bin order, position order (the serial order of the reference)
	template <class IsMarked>
	void remove_marked(size_t a, size_t b, IsMarked marked, std::vector<int64_t> &out) {
		for (size_t i = a; i <= b && i < data_.size(); i++) {
			Bin &bin = data_[i];
			for (size_t w = 0; w < bin.bits.size(); w++) {
				uint64_t word = bin.bits[w];
				while (word) {
					const size_t at = w * 64 + (size_t)__builtin_ctzll(word);
					word &= word - 1;
					if (marked(bin.items[at].v)) { out.push_back(bin.items[at].v); bin.clear(at); }
				}
			}
		}
	}

	// Drop the removed entries for good and number the survivors 0, 1, 2, ... in iteration order
	// (what assign_rows did at the start).  Appends, for every survivor in order, its previous row to
	// `old_rows`: old_rows[new row] = old row.  Nothing observable changes: element i of a bin is the
	// same point as before, only its row number differs.
	void compact(std::vector<int64_t> &old_rows) {
		int64_t next = (int64_t)old_rows.size();
		for (size_t b = 0; b < data_.size(); b++) {
			Bin &bin = data_[b];
			row0_[b] = next;
			size_t w = 0;
			for (size_t j = 0; j < bin.items.size(); j++) {
				if (!((bin.bits[j / 64] >> (j % 64)) & 1)) continue;
				old_rows.push_back(bin.items[j].v);
				bin.items[w] = bin.items[j];
				bin.items[w].v = next++;
				w++;
			}
			bin.items.resize(w);
			bin.alive = w;
			bin.bits.assign((w + 63) / 64, ~0ull);
			if (w % 64) bin.bits.back() = (1ull << (w % 64)) - 1;
		}
		first_live_ = 0;
	}

	size_t nbins() const { return data_.size(); }
	// the layout a device-resident copy needs (mc_accumulate_run): bin bounds and, after assign_rows(),
	// the first row of every bin followed by the number of rows
	const std::vector<uint64_t> &bounds() const { return bounds_; }
	std::vector<int64_t> first_rows() const {
		std::vector<int64_t> r(row0_);
		r.push_back(row0_.empty() ? 0 : row0_.back() + (int64_t)data_.back().items.size());
		return r;
	}

private:
	std::vector<uint32_t> fill_;   // items per bin while the container is being filled (insert)
	struct Bin {
		std::vector<Entry> items;     // fixed after finalize(); items[j].v = row0 + j
		std::vector<uint64_t> bits;   // alive bitmap over items
		size_t alive = 0;
		// index in items of the i-th alive entry
		size_t select(size_t i) const {
			for (size_t w = 0; w < bits.size(); w++) {
				const size_t c = (size_t)__builtin_popcountll(bits[w]);
				if (i < c) {
					uint64_t word = bits[w];
					for (; i; i--) word &= word - 1;
					return w * 64 + (size_t)__builtin_ctzll(word);
				}
				i -= c;
			}
			return items.size();   // out of range: the caller asked for an element that is not there
		}
		const Entry &at(size_t i) const { return items.at(select(i)); }
		void clear(size_t at) {
			uint64_t &w = bits[at / 64];
			const uint64_t m = 1ull << (at % 64);
			if (w & m) { w &= ~m; alive--; }
		}
	};

	size_t bin_of_row(int64_t row) const {
		return (size_t)(std::upper_bound(row0_.begin(), row0_.end(), row) - row0_.begin()) - 1;
	}

	int64_t diff(const BvIdx &a, const BvIdx &rhs) const {   // a - rhs
		if (a.bin < rhs.bin || (a.bin == rhs.bin && a.pos < rhs.pos)) return -diff(rhs, a);
		if (a.bin == rhs.bin) return (int64_t)(a.pos - rhs.pos);
		int64_t sum = (int64_t)a.pos;
		sum += (int64_t)(data_.at(rhs.bin).alive - rhs.pos);
		for (size_t i = rhs.bin + 1; i < a.bin; i++) sum += (int64_t)data_[i].alive;
		return sum;
	}

	std::vector<Bin> data_;
	std::vector<uint64_t> bounds_;
	std::vector<int64_t> row0_;   // first row of each bin
	size_t first_live_ = 0;       // bins before this one are empty (pop() only ever moves forward)
};

}  // namespace mch
