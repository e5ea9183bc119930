// TEST INFRASTRUCTURE ONLY -- a CPU stand-in for libmeshclust_b200.so that implements the same
// C-ABI (include/meshclust_b200.h) with the oracle (oracle/mc_oracle.c).  It exists so that the
// HOST control flow of bin/meshclust (sampling, GLM, bvec, accumulate/update/merge bookkeeping) can
// be exercised against the reference binary in a container without a GPU.  It is linked only into
// tests/_build/meshclust_hostlogic by tests/test_host_logic.py; the product binary never sees it.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/meshclust_b200.h"
#include "../../oracle/mc_oracle.h"

struct mc_ctx {
	int64_t n = 0;
	std::vector<uint8_t> digits;
	std::vector<uint8_t> letters;   // mc_ingest_fasta -> mc_load_segments
	std::vector<int64_t> offs, seg_off;
	std::vector<int32_t> segs;
	int k = 0, nbins = 0, tbytes = 1;
	std::vector<uint8_t> hist;   // n * nbins * tbytes
	std::vector<uint64_t> len;
	std::vector<uint8_t> alive;
	double mins[5], maxs[5], w[5];
	int nfeat = 0;
	std::vector<int64_t> members;
	int64_t launches = 0;
	int64_t near = 0;   // decisions with |sum| < 1e-9
};

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

extern "C" {
const char *mc_version(void) { return "meshclust_b200 MOCK (CPU oracle; tests only)"; }
const char *mc_last_error(void) { return g_err; }
int mc_device_count(void) { return 0; }
int mc_ctx_create(mc_ctx **out, int) { *out = new mc_ctx(); return MC_OK; }
void mc_ctx_destroy(mc_ctx *c) { delete c; }
void *mc_stream(mc_ctx *) { return nullptr; }
int mc_sync(mc_ctx *) { return MC_OK; }
int64_t mc_launch_count(mc_ctx *c) { return c->launches; }

int mc_host_segments(const uint8_t *letters, int64_t len, int32_t *segs, int max_segs) {
	std::vector<char> d((size_t)len + 1);
	return mco_encode((const char *)letters, (long)len, d.data(), segs, max_segs) < 0 && 0 ? -1 : [&]() {
		// mco_encode also validates letters; segments alone must not fail on them
		int ns = mco_encode((const char *)letters, (long)len, d.data(), segs, max_segs);
		if (ns >= 0) return ns;
		std::vector<uint8_t> clean(letters, letters + len);
		for (auto &ch : clean) if ((ch | 0x20) != 'n') ch = 'A';
		return mco_encode((const char *)clean.data(), (long)len, d.data(), segs, max_segs);
	}();
}

int mc_load_sequences(mc_ctx *c, const uint8_t *letters, const int64_t *offsets, int64_t n, const int32_t *segs, const int64_t *seg_offsets) {
	c->n = n;
	c->offs.assign(offsets, offsets + n + 1);
	c->seg_off.assign(seg_offsets, seg_offsets + n + 1);
	c->segs.assign(segs, segs + 2 * seg_offsets[n]);
	c->digits.resize((size_t)offsets[n]);
	for (int64_t i = 0; i < n; i++) {
		std::vector<int> sg(2 * 4096);
		int ns = mco_encode((const char *)letters + offsets[i], (long)(offsets[i + 1] - offsets[i]), (char *)c->digits.data() + offsets[i], sg.data(), 4096);
		if (ns < 0) return fail(MC_ERR_INPUT, "Invalid nucleotide in input");
	}
	return MC_OK;
}
int mc_stage_fasta_bytes(mc_ctx *, const uint8_t *, int64_t, int64_t) { return MC_OK; }
int mc_reserve_scratch(mc_ctx *, int64_t) { return MC_OK; }
int mc_ingest_fasta(mc_ctx *c, const uint8_t *raw, int64_t raw_bytes, const int64_t *span_begin, const int64_t *span_end, const int64_t *offsets,
                    int64_t n, uint8_t *rec_flags_out) {
	(void)raw_bytes;
	c->n = n;
	c->offs.assign(offsets, offsets + n + 1);
	c->letters.assign((size_t)offsets[n], 0);
	for (int64_t i = 0; i < n; i++) {
		int64_t w = offsets[i];
		uint8_t fl = 0;
		for (int64_t p = span_begin[i]; p < span_end[i]; p++) {
			const uint8_t ch = raw[p];
			if (ch == '\n') continue;
			if (w >= offsets[i + 1]) return fail(MC_ERR_INPUT, "mc_ingest_fasta: a record's span does not hold the number of letters its offsets announce");
			c->letters[(size_t)w++] = ch;
			const uint8_t u = ch & 0xdf;
			if (u == 'N') fl |= 1;
			else if (u != 'A' && u != 'C' && u != 'G' && u != 'T') fl |= 2;
		}
		if (w != offsets[i + 1]) return fail(MC_ERR_INPUT, "mc_ingest_fasta: a record's span does not hold the number of letters its offsets announce");
		rec_flags_out[i] = fl;
	}
	return MC_OK;
}
int mc_load_segments(mc_ctx *c, const int32_t *segs, const int64_t *seg_offsets, int validate) {
	(void)validate;   // the CPU encode below always validates: a stricter check than asked for can only fail where the real one would be asked
	std::vector<uint8_t> letters;
	letters.swap(c->letters);
	const std::vector<int64_t> offs = c->offs;
	return mc_load_sequences(c, letters.data(), offs.data(), c->n, segs, seg_offsets);
}
int mc_copy_letters(mc_ctx *c, uint8_t *out) { memcpy(out, c->letters.data(), c->letters.size()); return MC_OK; }
int mc_copy_digits(mc_ctx *c, uint8_t *out) { memcpy(out, c->digits.data(), c->digits.size()); return MC_OK; }

int mc_build_histograms(mc_ctx *c, int k, int tbytes, int *tbytes_out, uint64_t *max_count_out) {
	c->k = k; c->nbins = 1 << (2 * k);
	std::vector<uint64_t> h((size_t)c->nbins);
	std::vector<uint64_t> all((size_t)c->n * c->nbins);
	uint64_t mx = 0;
	for (int64_t i = 0; i < c->n; i++) {
		mco_hist_digits((const char *)c->digits.data() + c->offs[i], c->segs.data() + 2 * c->seg_off[i], (int)(c->seg_off[i + 1] - c->seg_off[i]), k, h.data());
		for (int b = 0; b < c->nbins; b++) { all[(size_t)i * c->nbins + b] = h[b]; if (h[b] > mx) mx = h[b]; }
	}
	const int need = mx <= 255 ? 1 : 2;
	c->tbytes = tbytes ? tbytes : need;
	if (mx > 65535 || c->tbytes < need) return fail(MC_ERR_UNSUPPORTED, "count too large");
	c->hist.resize((size_t)c->n * c->nbins * c->tbytes);
	for (size_t i = 0; i < all.size(); i++) {
		if (c->tbytes == 1) c->hist[i] = (uint8_t)all[i]; else ((uint16_t *)c->hist.data())[i] = (uint16_t)all[i];
	}
	c->len.resize((size_t)c->n);
	for (int64_t i = 0; i < c->n; i++) c->len[i] = (uint64_t)(c->offs[i + 1] - c->offs[i]);
	c->alive.assign((size_t)c->n, 1);
	if (tbytes_out) *tbytes_out = c->tbytes;
	if (max_count_out) *max_count_out = mx;
	return MC_OK;
}
int mc_load_histograms(mc_ctx *c, const void *hists, int tbytes, int k, const uint64_t *lens, int64_t n) {
	c->n = n; c->k = k; c->nbins = 1 << (2 * k); c->tbytes = tbytes;
	c->hist.assign((const uint8_t *)hists, (const uint8_t *)hists + (size_t)n * c->nbins * tbytes);
	c->len.assign(lens, lens + n);
	c->alive.assign((size_t)n, 1);
	return MC_OK;
}
int mc_copy_histograms(mc_ctx *c, void *out) { memcpy(out, c->hist.data(), c->hist.size()); return MC_OK; }
int mc_copy_point_stats(mc_ctx *c, uint64_t *len, uint64_t *mag, uint64_t *sq) {
	for (int64_t i = 0; i < c->n; i++) {
		uint64_t m, s;
		mco_point_stats(c->hist.data() + (size_t)i * c->nbins * c->tbytes, c->nbins, c->tbytes, &m, &s);
		if (len) len[i] = c->len[i];
		if (mag) mag[i] = m;
		if (sq) sq[i] = s;
	}
	return MC_OK;
}
int mc_set_model(mc_ctx *c, const double *mins, const double *maxs, const double *w, int nfeat) {
	const int nl = nfeat >= 4 ? 5 : 4;
	for (int i = 0; i < 5; i++) { c->mins[i] = i < nl ? mins[i] : 0; c->maxs[i] = i < nl ? maxs[i] : 1; c->w[i] = i <= nfeat ? w[i] : 0; }
	c->nfeat = nfeat;
	return MC_OK;
}
static const uint8_t *row(mc_ctx *c, int64_t r) { return c->hist.data() + (size_t)r * c->nbins * c->tbytes; }

int mc_distance_keys(mc_ctx *c, const int32_t *centers, int C, uint16_t *keys) {
#pragma omp parallel for schedule(static)
	for (int64_t t = 0; t < (int64_t)C * c->n; t++) {
		const int ci = (int)(t / c->n);
		const int64_t r = t % c->n;
		double raw[5]; uint64_t d;
		mco_features(row(c, r), row(c, centers[ci]), c->nbins, c->tbytes, c->len[r], c->len[centers[ci]], raw, &d);
		keys[t] = (uint16_t)d;
	}
	return MC_OK;
}
int mc_pair_features(mc_ctx *c, const int32_t *a, const int32_t *b, int64_t m, double *out5, uint64_t *dist) {
	for (int64_t i = 0; i < m; i++) {
		double raw[5]; uint64_t d;
		mco_features(row(c, a[i]), row(c, b[i]), c->nbins, c->tbytes, c->len[a[i]], c->len[b[i]], raw, &d);
		if (out5) memcpy(out5 + i * 5, raw, sizeof(raw));
		if (dist) dist[i] = d;
	}
	return MC_OK;
}
static void eval_one(mc_ctx *c, int64_t p, int64_t q, double *sum, double *f0, uint8_t *flag, double *feats) {
	double raw[5]; uint64_t d;
	mco_features(row(c, p), row(c, q), c->nbins, c->tbytes, c->len[p], c->len[q], raw, &d);
	static const int is_sim[5] = {0, 1, 0, 0, 1};
	double cc[5], f[4];
	for (int j = 0; j < 5; j++) { double v = (raw[j] - c->mins[j]) / (c->maxs[j] - c->mins[j]); cc[j] = is_sim[j] ? v : 1 - v; }
	f[0] = (1.0 * cc[0]) * cc[1];
	f[1] = (1.0 * (cc[0] * cc[0])) * (cc[2] * cc[2]);
	f[2] = 1.0 * cc[3];
	f[3] = (1.0 * (cc[0] * cc[0])) * (cc[4] * cc[4]);
	double s = c->w[0];
	for (int j = 0; j < c->nfeat; j++) s = fma(c->w[j + 1], f[j], s);
	if (sum) *sum = s;
	if (f0) *f0 = f[0];
	if (flag) {
		*flag = round(1.0 / (1 + exp(-s))) == 1.0;
		if (fabs(s) < 1e-9) {
#pragma omp atomic
			c->near++;
		}
	}
	if (feats) memcpy(feats, f, sizeof(f));
}
int mc_pair_classify(mc_ctx *c, const int32_t *a, const int32_t *b, int64_t m, double *sum, double *f0, uint8_t *flag, double *feats) {
	for (int64_t i = 0; i < m; i++) eval_one(c, a[i], b[i], sum ? sum + i : nullptr, f0 ? f0 + i : nullptr, flag ? flag + i : nullptr, feats ? feats + i * 4 : nullptr);
	return MC_OK;
}
int mc_alive_reset(mc_ctx *c) { c->alive.assign((size_t)c->n, 1); return MC_OK; }
int mc_alive_kill(mc_ctx *c, const int64_t *rows, int64_t m) { for (int64_t i = 0; i < m; i++) c->alive[rows[i]] = 0; return MC_OK; }
int mc_scan(mc_ctx *c, int64_t center, int64_t lo, int64_t hi, mc_scan_result *res, uint8_t *marks) {
	res->n_eval = 0; res->n_pos = 0; res->best_row = -1; res->best_f0 = -1;
	if (hi < lo) return MC_OK;
	const int64_t m = hi - lo + 1;
	std::vector<double> f0v((size_t)m);
	std::vector<uint8_t> fl((size_t)m, 0);
#pragma omp parallel for schedule(static)
	for (int64_t r = lo; r <= hi; r++) {
		if (!c->alive[r]) continue;
		eval_one(c, r, center, nullptr, &f0v[r - lo], &fl[r - lo], nullptr);
	}
	for (int64_t r = lo; r <= hi; r++) {
		if (marks) marks[r - lo] = 0;
		if (!c->alive[r]) continue;
		res->n_eval++;
		if (f0v[r - lo] > res->best_f0) { res->best_f0 = f0v[r - lo]; res->best_row = r; }
		if (fl[r - lo]) { res->n_pos++; c->alive[r] = 0; if (marks) marks[r - lo] = 1; }
	}
	c->launches++;
	return MC_OK;
}
int mc_scan_enqueue(mc_ctx *, int64_t, int64_t, int64_t, int, int) { return fail(MC_ERR_UNSUPPORTED, "mock"); }
int mc_scan_enqueue_many(mc_ctx *, const int64_t *, const int64_t *, const int64_t *, int, int, int) { return fail(MC_ERR_UNSUPPORTED, "mock"); }
int mc_scan_collect(mc_ctx *, int, int, mc_scan_result *) { return fail(MC_ERR_UNSUPPORTED, "mock"); }

static int64_t nearest_of(mc_ctx *c, const std::vector<int64_t> &rows, double *dist_out) {
	std::vector<uint8_t> packed(rows.size() * (size_t)c->nbins * c->tbytes);
	for (size_t i = 0; i < rows.size(); i++) memcpy(packed.data() + i * (size_t)c->nbins * c->tbytes, row(c, rows[i]), (size_t)c->nbins * c->tbytes);
	std::vector<double> mean((size_t)c->nbins);
	mco_mean(packed.data(), c->nbins, c->tbytes, (int)rows.size(), mean.data());
	int64_t best = -1; double bd = 0;
	for (size_t i = 0; i < rows.size(); i++) {
		const double d = mco_distance_d(row(c, rows[i]), c->nbins, c->tbytes, mean.data());
		if (best < 0 || d < bd) { best = rows[i]; bd = d; }
	}
	if (dist_out) *dist_out = bd;
	return best;
}
int mc_mean_nearest(mc_ctx *c, const int64_t *rows, int64_t m, int append, int64_t *nearest, double *dist) {
	if (!append) c->members.clear();
	c->members.insert(c->members.end(), rows, rows + m);
	*nearest = nearest_of(c, c->members, dist);
	return MC_OK;
}
int mc_accumulate_step(mc_ctx *c, int64_t center, int64_t lo, int64_t hi, int restart, mc_step_result *res, int64_t *rows_out, int64_t cap) {
	if (restart) { c->members.clear(); c->members.push_back(center); }
	std::vector<uint8_t> marks((size_t)(hi >= lo ? hi - lo + 1 : 0));
	mc_scan(c, center, lo, hi, &res->scan, marks.data());
	res->nearest_row = -1;
	if (res->scan.n_pos > 0) {
		if (rows_out && cap < res->scan.n_pos) return fail(MC_ERR_ARG, "marked_rows_out too small");
		int64_t w = 0;
		for (int64_t r = lo; r <= hi; r++)
			if (marks[(size_t)(r - lo)]) { c->members.push_back(r); if (rows_out) rows_out[w++] = r; }
		res->nearest_row = nearest_of(c, c->members, nullptr);
	}
	res->n_members = (int64_t)c->members.size();
	return MC_OK;
}
int mc_near_threshold_count(mc_ctx *c, int64_t *out, int reset) { *out = c->near; if (reset) c->near = 0; return MC_OK; }

// A literal restatement of the reference's length-binned container (bvec.cpp) on (row, length)
// entries -- vectors with erase, linear walks -- independent of host/bvec.hpp and of the device code.
namespace {
struct LitBvec {
	typedef std::pair<int64_t, uint64_t> Item;   // row, length
	std::vector<std::vector<Item>> data;
	std::vector<uint64_t> bounds;
	void index_of(uint64_t point, size_t *pfront, size_t *pback) const {   // bvec.cpp:123-149
		size_t low = bounds.size() - 1, high = 0;
		for (size_t i = 0; i < bounds.size(); i++) {
			const size_t prev = i > 0 ? bounds[i - 1] : 0;
			const size_t prev_index = i > 0 ? i - 1 : 0;
			if (point >= prev && point <= bounds[i]) { low = std::min(low, prev_index); high = std::max(high, prev_index); }
		}
		if (point >= bounds.back()) high = std::max(high, bounds.size() - 1);
		if (pfront) *pfront = low;
		if (pback) *pback = high;
	}
	void inner_index_of(uint64_t length, size_t &idx, size_t *pfront, size_t *pback) const {   // bvec.cpp:52-120
		if (data.at(idx).empty()) {
			if (pfront) for (size_t i = 0; i < data.size(); i++) if (!data[i].empty()) { idx = i; *pfront = 0; break; }
			if (pback) for (long i = (long)data.size() - 1; i >= 0; i--) if (!data[i].empty()) { idx = (size_t)i; *pback = 0; break; }
			return;
		}
		const std::vector<Item> &bin = data[idx];
		size_t front = 0, back = 0, low = 0, high = bin.size() - 1;
		while (low <= high) {
			const size_t mid = (low + high) / 2;
			const uint64_t d = bin[mid].second;
			if (d == length) { front = back = mid; break; }
			else if (length < d) high = mid;
			else low = mid + 1;
			if (low == high) { front = low; back = high; break; }
		}
		if (pfront) { for (long i = (long)front; i >= 0 && bin[(size_t)i].second == length; i--) front = (size_t)i; *pfront = front; }
		if (pback) { for (size_t i = back; i < bin.size() && bin[i].second == length; i++) back = i; *pback = back; }
	}
	long diff(size_t abin, size_t apos, size_t rbin, size_t rpos) const {   // bvec_iterator.h:61-76
		if (abin < rbin || (abin == rbin && apos < rpos)) return -diff(rbin, rpos, abin, apos);
		if (abin == rbin) return (long)(apos - rpos);
		long sum = (long)apos;
		sum += (long)(data.at(rbin).size() - rpos);
		for (size_t i = rbin + 1; i < abin; i++) sum += (long)data[i].size();
		return sum;
	}
	int64_t pop() {   // bvec.cpp:27-38
		for (auto &bin : data) if (!bin.empty()) { const int64_t r = bin[0].first; bin.erase(bin.begin()); return r; }
		return -1;
	}
};
}  // namespace

int mc_accumulate_run(mc_ctx *c, double sim, const uint64_t *bin_bounds, const int64_t *first_row, int64_t nb, int64_t *centers_out,
                      int64_t *offs_out, int64_t *members_out, mc_run_stats *st) {
	LitBvec bv;
	bv.bounds.assign(bin_bounds, bin_bounds + nb);
	bv.data.resize((size_t)nb);
	for (int64_t b = 0; b < nb; b++)
		for (int64_t r = first_row[b]; r < first_row[b + 1]; r++) bv.data[(size_t)b].push_back({r, c->len[(size_t)r]});
	c->alive.assign((size_t)c->n, 1);
	int64_t nc = 0, total = 0, scans = 0, evals = 0, steps = 0;
	const int64_t near0 = c->near;
	offs_out[0] = 0;
	int64_t last = bv.pop();
	if (last >= 0) c->alive[(size_t)last] = 0;
	while (last >= 0) {   // ClusterFactory.cpp:722-729 around accumulate (:637-714)
		std::vector<int64_t> current{last};
		int64_t next_seed = -1;
		for (bool is_min = false; !is_min;) {
			const uint64_t len = c->len[(size_t)last];
			size_t fbin = 0, fpos = 0, bbin = bv.data.size() - 1, bpos = bv.data[bbin].size() - 1;   // bvec.cpp:247-278
			bv.index_of((uint64_t)(len * sim), &fbin, nullptr);
			bv.index_of((uint64_t)(len / sim), nullptr, &bbin);
			bv.inner_index_of((uint64_t)(len * sim), fbin, &fpos, nullptr);
			bv.inner_index_of((uint64_t)(len / sim), bbin, nullptr, &bpos);
			mc_scan_result res;
			res.n_eval = 0; res.n_pos = 0; res.best_row = -1; res.best_f0 = -1;
			std::vector<uint8_t> marks;
			int64_t lo = 0, hi = -1;
			if (bv.diff(bbin, bpos, fbin, fpos) + 1 > 0) {
				lo = bv.data[fbin][fpos].first;
				hi = bv.data[bbin][bpos].first;
				marks.resize((size_t)(hi - lo + 1));
				mc_scan(c, last, lo, hi, &res, marks.data());
				scans++; evals += res.n_eval;
			}
			steps++;
			is_min = res.n_pos == 0;
			if (is_min) {
				if (res.best_row < 0) next_seed = bv.pop();
				else {
					next_seed = res.best_row;
					for (auto &bin : bv.data)
						for (size_t i = 0; i < bin.size(); i++) if (bin[i].first == next_seed) { bin.erase(bin.begin() + (long)i); break; }
				}
				if (next_seed >= 0) c->alive[(size_t)next_seed] = 0;
			} else {
				for (size_t b = fbin; b <= bbin && b < bv.data.size(); b++) {   // remove_available, bvec.cpp:290-317
					std::vector<LitBvec::Item> keep;
					for (const auto &it : bv.data[b]) {
						if (it.first >= lo && it.first <= hi && marks[(size_t)(it.first - lo)]) current.push_back(it.first);
						else keep.push_back(it);
					}
					bv.data[b].swap(keep);
				}
				last = nearest_of(c, current, nullptr);
			}
		}
		centers_out[nc] = last;
		for (int64_t r : current) members_out[total++] = r;
		offs_out[++nc] = total;
		last = next_seed;
	}
	if (st) { st->n_clusters = nc; st->n_scans = scans; st->n_evals = evals; st->n_near_threshold = c->near - near0; st->n_steps = steps; st->device_seconds = 0; st->n_compactions = 0; }
	return MC_OK;
}

int mc_clone_points(mc_ctx *dst, mc_ctx *src) { *dst = *src; return MC_OK; }
int mc_clone_sequences(mc_ctx *dst, mc_ctx *src) { *dst = *src; return MC_OK; }
int mc_comm_init(mc_ctx *, int, int, uint8_t *) { return MC_OK; }
int mc_comm_connect_local(mc_ctx *const *, int) { return MC_OK; }
int mc_accumulate_step_sharded(mc_ctx *const *ctxs, int world, int64_t center, int64_t lo, int64_t hi, int restart, mc_step_result *res, int64_t *rows_out, int64_t cap) {
	const int rc = mc_accumulate_step(ctxs[0], center, lo, hi, restart, res, rows_out, cap);
	for (int r = 1; r < world; r++) ctxs[r]->alive = ctxs[0]->alive;
	return rc;
}
int mc_reserve_permute(mc_ctx *) { return MC_OK; }
int mc_permute_rows(mc_ctx *c, const int64_t *old_of_new, int64_t count, int64_t n_alive) {
	const size_t rb = (size_t)c->nbins * c->tbytes;
	std::vector<uint8_t> h((size_t)count * rb);
	std::vector<uint64_t> l((size_t)count);
	for (int64_t i = 0; i < count; i++) {
		memcpy(h.data() + (size_t)i * rb, c->hist.data() + (size_t)old_of_new[i] * rb, rb);
		l[(size_t)i] = c->len[(size_t)old_of_new[i]];
	}
	memcpy(c->hist.data(), h.data(), h.size());
	for (int64_t i = 0; i < count; i++) { c->len[(size_t)i] = l[(size_t)i]; c->alive[(size_t)i] = i < n_alive; }
	c->members.clear();
	return MC_OK;
}
int mc_update_centers(mc_ctx *c, const int64_t *centers, int64_t nc, const int64_t *cand, int64_t, const int64_t *cb, const int64_t *ce, int64_t *next) {
#pragma omp parallel for schedule(dynamic)
	for (int64_t j = 0; j < nc; j++) {
		std::vector<int64_t> good;
		for (int64_t i = cb[j]; i < ce[j]; i++) {
			uint8_t fl;
			eval_one(c, cand[i], centers[j], nullptr, nullptr, &fl, nullptr);
			if (fl) good.push_back(cand[i]);
		}
		next[j] = good.empty() ? -1 : nearest_of(c, good, nullptr);
	}
	return MC_OK;
}
int mc_align_pairs(mc_ctx *c, const int32_t *a, const int32_t *b, int64_t m, int32_t *score, int32_t *alen, int32_t *matches) {
	mco_globalign_batch((const char *)c->digits.data(), c->offs.data(), a, b, (int)m, score, alen, matches);
	return MC_OK;
}
int mc_kmer_histograms_host(mc_ctx *, const uint8_t *, const int64_t *, int64_t, int, int, void *, uint64_t *) { return fail(MC_ERR_UNSUPPORTED, "mock"); }
int mc_scan_host(mc_ctx *, const void *, int, int, const uint64_t *, int64_t, const int64_t *, int, mc_scan_result *, uint8_t *) { return fail(MC_ERR_UNSUPPORTED, "mock"); }
}
