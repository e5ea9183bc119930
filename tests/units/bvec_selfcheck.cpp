// Self-check of host/bvec.hpp: the binary-search index_of against the literal loop of the
// reference (bvec.cpp:123-149), and the bitmap-backed bins against a plain erase-based model.
#include <cstdio>
#include <random>

#include "../../meshclust_b200/host/bvec.hpp"

int main() {
	std::mt19937_64 rng(12345);
	long checks = 0;
	for (int round = 0; round < 300; round++) {
		const int n = 1 + (int)(rng() % 6000);
		const uint64_t span = 1 + rng() % (round % 3 == 0 ? 5 : 3000);
		std::vector<uint64_t> len(n);
		for (auto &l : len) l = 100 + rng() % span;
		const uint64_t bin = 1 + rng() % 400;
		mch::BVec bv(len, bin);
		for (uint64_t p = 0; p < 100 + span + 50; p++) {
			size_t f0, b0, f1, b1;
			bv.index_of(p, &f0, &b0);
			bv.index_of_linear(p, &f1, &b1);
			if (f0 != f1 || b0 != b1) { printf("index_of mismatch point=%llu: (%zu,%zu) vs (%zu,%zu)\n", (unsigned long long)p, f0, b0, f1, b1); return 1; }
			checks++;
		}
		for (int i = 0; i < n; i++) bv.insert(i, len[i]);
		bv.finalize();
		std::vector<int64_t> id_of_row = bv.assign_rows();
		{
			// bvec::insert (bvec.cpp:152-177) literally: the minima of the candidate bins collected in a vector, the
			// middle one taken; then insert_finalize's per-bin std::sort -- the container must hold the same rows
			std::vector<std::vector<mch::BVec::Entry>> bins(bv.nbins());
			for (int i = 0; i < n; i++) {
				size_t f, b;
				bv.index_of_linear(len[i], &f, &b);
				size_t minimum = (size_t)-1;
				for (size_t j = f; j <= b; j++) minimum = std::min(minimum, bins[j].size());
				std::vector<size_t> mins;
				for (size_t j = f; j <= b; j++) if (bins[j].size() == minimum) mins.push_back(j);
				bins[mins[mins.size() / 2]].push_back({(int64_t)i, len[i]});
			}
			std::vector<int64_t> want;
			for (auto &bin : bins) {
				std::sort(bin.begin(), bin.end(), [](const mch::BVec::Entry &a, const mch::BVec::Entry &b) { return a.len < b.len; });
				for (auto &e : bin) want.push_back(e.v);
			}
			if (want != id_of_row) { printf("insert / finalize differ from the literal rule in round %d\n", round); return 1; }
			checks++;
		}
		// model: rows alive in order, with their lengths
		std::vector<char> alive(n, 1);
		size_t live = n;
		for (int step = 0; step < 40 && live > 0; step++) {
			// remove a random batch of rows, a pop and an erase, then compare sizes and a range walk
			std::vector<int64_t> rows;
			for (int r = 0; r < n; r++) if (alive[r] && rng() % 7 == 0) rows.push_back(r);
			bv.remove_rows(rows.data(), rows.size());
			for (auto r : rows) { alive[r] = 0; live--; }
			if (live == 0) break;
			const int64_t p = bv.pop();
			int64_t want = -1;
			for (int r = 0; r < n; r++) if (alive[r]) { want = r; break; }
			if (p != want) { printf("pop mismatch %lld vs %lld\n", (long long)p, (long long)want); return 1; }
			alive[p] = 0; live--;
			if (bv.size() != live) { printf("size mismatch\n"); return 1; }
			if (live == 0) break;
			const uint64_t a = 100 + rng() % span, b = a + rng() % span;
			const auto rg = bv.get_range(a, b);
			const int64_t trip = bv.trip_count(rg.first, rg.second);
			if (trip > 0) {
				const int64_t lo = bv.row_at(rg.first), hi = bv.row_at(rg.second);
				int64_t cnt = 0;
				for (int64_t r = lo; r <= hi; r++) cnt += alive[r];
				if (cnt != trip) { printf("trip count %lld but %lld alive rows in [%lld,%lld]\n", (long long)trip, (long long)cnt, (long long)lo, (long long)hi); return 1; }
				checks++;
			}
		}
	}
	printf("ok %ld checks\n", checks);
	return 0;
}
