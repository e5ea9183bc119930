// Self-check of host/lazy_sort.hpp: every position LazySort resolves must hold the element std::sort puts there
// (same libstdc++, same comparator, same input order), for inputs with many equal keys, sorted / reversed / constant
// inputs, and with small depth limits that force introsort's heapsort branch (compared with the library's own
// __introsort_loop + __final_insertion_sort run at that depth).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <random>
#include <vector>

#include "../../meshclust_b200/host/lazy_sort.hpp"

struct KeyId { uint16_t key; int id; };
struct ByKey { bool operator()(const KeyId &a, const KeyId &b) const { return a.key < b.key; } };

int main() {
	std::mt19937_64 rng(99);
	long checks = 0;
	for (int round = 0; round < 400; round++) {
		const size_t n = round < 40 ? (size_t)round : (size_t)(rng() % (round % 10 == 0 ? 300000 : 5000));
		const int kind = round % 7;
		const unsigned span = kind == 0 ? 1 : (kind == 1 ? 3 : (kind == 2 ? 50 : 10001));
		std::vector<KeyId> v(n);
		for (size_t i = 0; i < n; i++) v[i] = {(uint16_t)(rng() % span), (int)i};
		if (kind == 4) std::sort(v.begin(), v.end(), ByKey());
		if (kind == 5) { std::sort(v.begin(), v.end(), ByKey()); std::reverse(v.begin(), v.end()); }
		if (kind == 6) for (size_t i = 0; i < n; i++) v[i].key = (uint16_t)(i < n / 2 ? i % 1000 : (n - i) % 1000);   // organ pipes
		const int depth = (round % 5 == 3) ? (int)(rng() % 6) : -1;
		std::vector<KeyId> want = v;
		if (depth < 0) std::sort(want.begin(), want.end(), ByKey());
		else if (!want.empty()) {
			std::__introsort_loop(want.begin(), want.end(), (long)depth, __gnu_cxx::__ops::__iter_comp_iter(ByKey()));
			std::__final_insertion_sort(want.begin(), want.end(), __gnu_cxx::__ops::__iter_comp_iter(ByKey()));
		}
		if (depth < 0) {
			// the task-parallel full sort: the same array as std::sort, element by element
			std::vector<KeyId> par = v;
			mch::parallel_std_sort(par, ByKey());
			for (size_t p = 0; p < n; p++)
				if (par[p].key != want[p].key || par[p].id != want[p].id) { printf("round %d (n=%zu kind=%d): parallel_std_sort differs at %zu\n", round, n, kind, p); return 1; }
			checks += (long)n;
		}
		mch::LazySort<KeyId, ByKey> lazy(std::vector<KeyId>(v), ByKey(), depth);
		// a few positions in the order a binary search would ask for them, then strided ones, then all
		std::vector<size_t> ask;
		if (n) {
			size_t pos = 2 * (n / 4), off = n / 4;
			while (off) { ask.push_back(pos); if (rng() & 1) pos -= off; else pos += off; off /= 2; if (pos >= n) pos = n - 1; }
			for (int t = 0; t < 20; t++) ask.push_back((size_t)(rng() % n));
		}
		for (size_t p : ask) {
			const KeyId &g = lazy.at(p);
			if (g.key != want[p].key || g.id != want[p].id) { printf("round %d (n=%zu kind=%d depth=%d): position %zu holds (%u,%d), std::sort has (%u,%d)\n", round, n, kind, depth, p, g.key, g.id, want[p].key, want[p].id); return 1; }
			checks++;
		}
		if (round % 3 == 0) {
			for (size_t p = 0; p < n; p++) {
				const KeyId &g = lazy.at(p);
				if (g.key != want[p].key || g.id != want[p].id) { printf("round %d full sweep: position %zu differs\n", round, p); return 1; }
			}
			checks += (long)n;
		}
	}
	// what it buys: 40 positions of 1 M records against the full sort
	{
		const size_t n = 1000000;
		std::vector<KeyId> v(n);
		for (size_t i = 0; i < n; i++) v[i] = {(uint16_t)(rng() % 10001), (int)i};
		auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
		std::vector<KeyId> w = v;
		const double t0 = now();
		std::sort(w.begin(), w.end(), ByKey());
		const double t1 = now();
		mch::LazySort<KeyId, ByKey> lazy(std::move(v), ByKey());
		size_t pos = 2 * (n / 4), off = n / 4;
		long bad = 0;
		while (off) { bad += lazy.at(pos).id != w[pos].id; if (rng() & 1) pos -= off; else pos += off; off /= 2; }
		for (int t = 0; t < 10; t++) { const size_t p = (size_t)((double)pos * t / 10); bad += lazy.at(p).id != w[p].id; }
		for (int t = 0; t < 10; t++) { const size_t p = pos + (size_t)((double)(n - pos) * t / 10); bad += lazy.at(p).id != w[p].id; }
		const double t2 = now();
		if (bad) { printf("1 M records: %ld positions differ\n", bad); return 1; }
		std::vector<KeyId> par(n);
		for (size_t i = 0; i < n; i++) par[i] = {(uint16_t)((i * 2654435761u) % 10001), (int)i};
		std::vector<KeyId> ref = par;
		std::sort(ref.begin(), ref.end(), ByKey());
		const double t3 = now();
		mch::parallel_std_sort(par, ByKey());
		const double t4 = now();
		for (size_t i = 0; i < n; i++) if (par[i].id != ref[i].id) { printf("1 M records: parallel_std_sort differs at %zu\n", i); return 1; }
		printf("ok %ld checks; 1 M records: std::sort %.3f s, 40 positions lazily %.3f s, parallel full sort %.3f s\n", checks, t1 - t0, t2 - t1, t4 - t3);
	}
	return 0;
}
