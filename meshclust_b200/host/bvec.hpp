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
#include <vector>

namespace mch {

struct BvIdx {
	size_t bin = 0, pos = 0;
};

class BVec {
public:
	// entries are point ids while the container is being filled, rows afterwards
	struct Entry {
		int64_t v;
		uint64_t len;
	};

	// bvec::bvec (bvec.cpp:10-24): bounds are every bin_size-th sorted length
	BVec(std::vector<uint64_t> lengths, uint64_t bin_size = 1000) {
		std::sort(lengths.begin(), lengths.end());
		for (uint64_t i = 0; i < lengths.size(); i += bin_size) bounds_.push_back(lengths[i]);
		data_.resize(bounds_.size());
	}

	// bvec::index_of (bvec.cpp:123-149)
	void index_of(uint64_t point, size_t *pfront, size_t *pback) const {
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
		std::vector<size_t> mins;
		size_t minimum = std::numeric_limits<size_t>::max();
		for (size_t i = front; i <= back; i++) {
			const size_t sz = data_[i].size();
			if (sz < minimum) { minimum = sz; mins.clear(); mins.push_back(i); }
			else if (sz == minimum) mins.push_back(i);
		}
		// front > back leaves no candidate: the reference prints an error and then indexes an empty
		// vector (undefined); it cannot happen for bounds taken from the same lengths
		data_.at(mins.at(mins.size() / 2)).push_back({id, len});
	}

	// bvec::insert_finalize (bvec.cpp:209-218): per-bin std::sort by length (unstable: the same
	// libstdc++ introsort on the same sequence and comparator gives the same permutation)
	void finalize() {
		for (auto &bin : data_)
			std::sort(bin.begin(), bin.end(), [](const Entry &a, const Entry &b) { return a.len < b.len; });
	}

	// after finalize(): rename the entries to their position in iteration order; returns id per row
	std::vector<int64_t> assign_rows() {
		std::vector<int64_t> id_of_row;
		for (auto &bin : data_)
			for (auto &e : bin) {
				id_of_row.push_back(e.v);
				e.v = (int64_t)id_of_row.size() - 1;
			}
		return id_of_row;
	}

	size_t size() const {
		size_t t = 0;
		for (auto &b : data_) t += b.size();
		return t;
	}

	// bvec::pop (bvec.cpp:27-38): first element of the first non-empty bin, or -1
	int64_t pop() {
		for (auto &bin : data_)
			if (!bin.empty()) {
				const int64_t r = bin.front().v;
				bin.erase(bin.begin());
				return r;
			}
		return -1;
	}

	// bvec::inner_index_of (bvec.cpp:52-120)
	void inner_index_of(uint64_t length, size_t &idx, size_t *pfront, size_t *pback) const {
		if (data_.at(idx).empty()) {
			if (pfront)
				for (size_t i = 0; i < data_.size(); i++)
					if (!data_[i].empty()) { idx = i; *pfront = 0; break; }
			if (pback)
				for (long i = (long)data_.size() - 1; i >= 0; i--)
					if (!data_[i].empty()) { idx = (size_t)i; *pback = 0; break; }
			return;
		}
		const auto &bin = data_[idx];
		size_t front = 0, back = 0, low = 0, high = bin.size() - 1;
		while (low <= high) {
			const size_t mid = (low + high) / 2;
			const uint64_t d = bin[mid].len;
			if (d == length) { front = back = mid; break; }
			else if (length < d) high = mid;
			else low = mid + 1;
			if (low == high) { front = low; back = high; break; }
		}
		if (pfront) {
			for (long i = (long)front; i >= 0 && bin[i].len == length; i--) front = (size_t)i;
			*pfront = front;
		}
		if (pback) {
			for (size_t i = back; i < bin.size() && bin[i].len == length; i++) back = i;
			*pback = back;
		}
	}

	// bvec::get_range (bvec.cpp:247-278); both ends inclusive
	std::pair<BvIdx, BvIdx> get_range(uint64_t begin_len, uint64_t end_len) const {
		BvIdx front, back;
		back.bin = data_.size() - 1;
		back.pos = data_[back.bin].size() - 1;   // wraps to SIZE_MAX on an empty last bin, as in the reference
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
	int64_t row_at(const BvIdx &i) const { return data_[i.bin][i.pos].v; }

	// the reference iterates `trip_count` steps from `front` with operator++ (which skips empty
	// bins); this returns the row reached by the last step, i.e. the inclusive upper row bound
	int64_t last_row_of_walk(BvIdx f, int64_t steps) const {
		size_t r = f.bin, c = f.pos;
		for (int64_t i = 1; i < steps; i++) {
			if (c + 1 < data_[r].size()) c++;
			else {
				r++; c = 0;
				while (r < data_.size() && data_[r].empty()) r++;
			}
		}
		return data_[r][c].v;
	}

	// bvec::erase (bvec.cpp:281-285) by row: the argmax element of the last scan
	void erase_row(int64_t row) {
		for (auto &bin : data_) {
			if (bin.empty() || bin.front().v > row || bin.back().v < row) continue;
			auto it = std::lower_bound(bin.begin(), bin.end(), row, [](const Entry &e, int64_t r) { return e.v < r; });
			if (it != bin.end() && it->v == row) { bin.erase(it); return; }
		}
	}

	// bvec::remove_available (bvec.cpp:290-317) for bins [a,b]: drop every marked row; appends them
	// to `out` in bin order, position order (the serial order of the reference)
	template <class IsMarked>
	void remove_marked(size_t a, size_t b, IsMarked marked, std::vector<int64_t> &out) {
		for (size_t i = a; i <= b && i < data_.size(); i++) {
			auto &bin = data_[i];
			size_t w = 0;
			for (size_t j = 0; j < bin.size(); j++) {
				if (marked(bin[j].v)) out.push_back(bin[j].v);
				else bin[w++] = bin[j];
			}
			bin.resize(w);
		}
	}

	size_t nbins() const { return data_.size(); }

private:
	int64_t diff(const BvIdx &a, const BvIdx &rhs) const {   // a - rhs
		if (a.bin < rhs.bin || (a.bin == rhs.bin && a.pos < rhs.pos)) return -diff(rhs, a);
		if (a.bin == rhs.bin) return (int64_t)(a.pos - rhs.pos);
		int64_t sum = (int64_t)a.pos;
		sum += (int64_t)(data_.at(rhs.bin).size() - rhs.pos);
		for (size_t i = rhs.bin + 1; i < a.bin; i++) sum += (int64_t)data_[i].size();
		return sum;
	}

	std::vector<std::vector<Entry>> data_;
	std::vector<uint64_t> bounds_;
};

}  // namespace mch
