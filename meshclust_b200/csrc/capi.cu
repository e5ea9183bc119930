// The C-ABI of include/meshclust_b200.h: context, HBM residency, host<->device staging.
// Kernels live in kmer_hist.cu / pair_kernels.cu / center_mean.cu / nw_identity.cu.
#include <stdarg.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <cmath>

#include "mc_common.cuh"

// kernel launchers (defined in the other translation units)
int mc_upload_lut();
int mc_launch_encode(mc_ctx *ctx);
int mc_launch_validate(mc_ctx *ctx);
int mc_launch_kmer_hist(mc_ctx *ctx, int k, int tbytes);
int mc_launch_point_stats(mc_ctx *ctx, const uint64_t *lens_dev);
int mc_launch_point_stats_range(mc_ctx *ctx, int64_t row0, int64_t n, const uint64_t *lens_dev, cudaStream_t stream);
int mc_launch_alive_reset(mc_ctx *ctx);
int mc_launch_scan(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked, void *partials_dev, int *nparts_out);
int64_t mc_scan_max_blocks(mc_ctx *ctx);
int mc_launch_dist_keys(mc_ctx *ctx, const int32_t *center_rows_dev, int C, uint16_t *keys_dev);
int mc_launch_pair_list(mc_ctx *ctx, const int32_t *pa_dev, const int32_t *pb_dev, int64_t m, double *raw5_dev, uint64_t *dist_dev, double *sum_dev, double *f0_dev, uint8_t *flag_dev, double *feats_dev);
int mc_launch_mean_nearest(mc_ctx *ctx, const int64_t *new_rows_dev, int64_t m_new, unsigned long long *sum_dev, const int64_t *members_dev, int64_t m_all, uint8_t *tq_dev, unsigned long long *magc_dev, void *partials_dev, long long *out_row_dev, double *out_dist_dev);
size_t mc_acc_dev_bytes();
int mc_launch_accumulate_tail(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int restart, const void *partials_dev, int nparts, void *acc_dev, void *out_host_dev, int32_t *list_host_dev, unsigned long long seq, const unsigned int *err_dev);
int mc_launch_permute_rows(mc_ctx *ctx, const int32_t *old_of_new_dev, int64_t count, int64_t n_alive, void *hist_out, void *aux_out);
int mc_launch_update_centers(mc_ctx *ctx, const int64_t *center_rows_dev, int64_t ncenters, const int64_t *cand_rows_dev, const int64_t *cand_begin_dev, const int64_t *cand_end_dev, const int64_t *flag_off_dev, uint8_t *flags_dev, long long *next_rows_dev);
int mc_launch_nw(mc_ctx *ctx, const int32_t *pa_dev, const int32_t *pb_dev, int64_t m, int64_t max_len, int32_t *score_dev, int32_t *len_dev, int32_t *id_dev, void *scratch_a, void *scratch_b, int64_t scratch_stride, int64_t nwarps, int rows_per_lane, int team_warps);
int mc_nw_pick_rows(const int64_t *lb, int64_t m);

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void mc_set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char *mc_last_error(void) { return g_err; }
extern "C" const char *mc_version(void) { return "meshclust_b200 0.1 (sm_100a)"; }

extern "C" int mc_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int mc_ensure_scratch(mc_ctx *ctx, size_t bytes) {
	if (bytes <= ctx->scratch_bytes) return MC_OK;
	ctx->staged_raw = nullptr;   // (bytes parked by mc_stage_fasta_bytes do not survive a new buffer)
	if (ctx->d_scratch) { MC_CUDA(cudaStreamSynchronize(ctx->stream)); MC_CUDA(cudaFree(ctx->d_scratch)); ctx->d_scratch = nullptr; ctx->scratch_bytes = 0; }
	bytes = align_up(bytes + bytes / 4, 1 << 20);
	MC_CUDA(cudaMalloc(&ctx->d_scratch, bytes));
	ctx->scratch_bytes = bytes;
	return MC_OK;
}

// A caller that knows what is coming (the ingest, then Phase A with its staging copies) sizes the scratch buffer once:
// growing it later means cudaFree + cudaMalloc of gigabytes in the middle of the run (seen: 0.2 - 1 s, sporadically).
extern "C" int mc_reserve_scratch(mc_ctx *ctx, int64_t bytes) {
	MC_REQUIRE(ctx && bytes >= 0, MC_ERR_ARG, "mc_reserve_scratch: bad arguments");
	MC_CUDA(cudaSetDevice(ctx->device));
	if ((size_t)bytes <= ctx->scratch_bytes) return MC_OK;
	ctx->staged_raw = nullptr;
	if (ctx->d_scratch) { MC_CUDA(cudaStreamSynchronize(ctx->stream)); MC_CUDA(cudaFree(ctx->d_scratch)); ctx->d_scratch = nullptr; ctx->scratch_bytes = 0; }
	const size_t want = align_up((size_t)bytes, 1 << 20);
	if (cudaMalloc(&ctx->d_scratch, want) != cudaSuccess) { cudaGetLastError(); ctx->d_scratch = nullptr; return MC_OK; }   // a hint: later calls allocate what they need
	ctx->scratch_bytes = want;
	return MC_OK;
}

int mc_ensure_pinned(mc_ctx *ctx, size_t bytes) {
	if (bytes <= ctx->pinned_bytes) return MC_OK;
	if (ctx->h_pinned) { MC_CUDA(cudaStreamSynchronize(ctx->stream)); MC_CUDA(cudaFreeHost(ctx->h_pinned)); ctx->h_pinned = nullptr; ctx->pinned_bytes = 0; }
	bytes = align_up(bytes + bytes / 4, 1 << 16);
	MC_CUDA(cudaMallocHost(&ctx->h_pinned, bytes));
	ctx->pinned_bytes = bytes;
	return MC_OK;
}

// bump allocator over the scratch buffer
struct Carve {
	uint8_t *base;
	size_t off = 0;
	explicit Carve(void *b) : base((uint8_t *)b) {}
	template <class T>
	T *take(size_t count) {
		off = align_up(off, 256);
		T *p = reinterpret_cast<T *>(base + off);
		off += count * sizeof(T);
		return p;
	}
	static size_t need(std::initializer_list<size_t> sizes) {
		size_t t = 0;
		for (size_t s : sizes) t = align_up(t, 256) + s;
		return t + 256;
	}
};

// ---------------------------------------------------------------------------------------------
extern "C" int mc_ctx_create(mc_ctx **out, int device) {
	MC_REQUIRE(out != nullptr, MC_ERR_ARG, "mc_ctx_create: out is NULL");
	const bool dbg = getenv("MC_DEBUG_TIMING") != nullptr;
	auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
	double t_prev = now();
	auto lap = [&](const char *what) { if (dbg) { const double t = now(); fprintf(stderr, "[mc_ctx_create] %-28s %.3f s\n", what, t - t_prev); t_prev = t; } };
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		mc_set_error("no CUDA device available (%s); meshclust_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
		return MC_ERR_CUDA;
	}
	MC_REQUIRE(device >= 0 && device < ndev, MC_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
	lap("cudaGetDeviceCount");
	MC_CUDA(cudaSetDevice(device));
	MC_CUDA(cudaFree(0));
	lap("cudaSetDevice + context");
	mc_ctx *ctx = new mc_ctx();
	ctx->device = device;
	cudaDeviceProp prop;
	MC_CUDA(cudaGetDeviceProperties(&prop, device));
	// the per-scan record buffers (result slots, inboxes, tag-polled copies) hold MC_SCAN_PARTS entries per scan and
	// every kernel sizes its grid from num_sms: a part with more SMs uses MC_SCAN_PARTS of them for these kernels
	ctx->num_sms = std::min<int>(prop.multiProcessorCount, MC_SCAN_PARTS);
	lap("cudaGetDeviceProperties");
	MC_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	ctx->own_stream = ctx->stream;
	MC_CUDA(cudaMalloc(&ctx->d_ticket, 16 * sizeof(unsigned int)));
	MC_CUDA(cudaMemsetAsync(ctx->d_ticket, 0, 16 * sizeof(unsigned int), ctx->stream));
	MC_CUDA(cudaMalloc(&ctx->d_flags, 16 * sizeof(unsigned int)));
	MC_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 16 * sizeof(unsigned int), ctx->stream));
	MC_CUDA(cudaMalloc(&ctx->d_near, 64));
	MC_CUDA(cudaMemsetAsync(ctx->d_near, 0, 64, ctx->stream));
	ctx->model.near = ctx->d_near;
	lap("stream + small allocations");
	int rc = mc_upload_lut();
	if (rc) { delete ctx; return rc; }
	lap("constant LUT (module load)");
	rc = mc_ensure_pinned(ctx, 1 << 20);
	if (rc) { delete ctx; return rc; }
	lap("pinned staging buffer");
	ctx->model.valid = 0;
	*out = ctx;
	return MC_OK;
}

static void free_seq(mc_ctx *ctx) {
	cudaFree(ctx->d_seq); cudaFree(ctx->d_seq_off); cudaFree(ctx->d_segs); cudaFree(ctx->d_seg_off);
	ctx->d_seq = nullptr; ctx->d_seq_off = nullptr; ctx->d_segs = nullptr; ctx->d_seg_off = nullptr;
	ctx->have_seq = false;
}

static void free_hist(mc_ctx *ctx) {
	cudaFree(ctx->d_hist_tmp); cudaFree(ctx->d_aux_tmp);
	ctx->d_hist_tmp = nullptr; ctx->d_aux_tmp = nullptr; ctx->tmp_rows = 0;
	cudaFree(ctx->d_hist); cudaFree(ctx->d_aux);
	cudaFree(ctx->d_marks); cudaFree(ctx->d_members); cudaFree(ctx->d_sum);
	ctx->d_hist = nullptr; ctx->d_aux = nullptr;
	ctx->d_marks = nullptr; ctx->d_members = nullptr; ctx->d_sum = nullptr;
	ctx->hist_capacity = 0; ctx->aux_capacity = 0; ctx->members_cap = 0; ctx->members_n = 0; ctx->sum_bins = 0;
	ctx->have_hist = false;
}

extern "C" void mc_ctx_destroy(mc_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	mc_comm_destroy(ctx);
	free_seq(ctx);
	free_hist(ctx);
	cudaFree(ctx->d_scratch);
	cudaFree(ctx->d_scan_slots);
	cudaFree(ctx->d_scan_partials);
	cudaFreeHost(ctx->h_pinned);
	cudaFree(ctx->d_acc);
	if (ctx->h_step) cudaFreeHost(ctx->h_step);
	if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
	for (int i = 0; i < 8; i++) if (ctx->chunk_ev[i]) cudaEventDestroy(ctx->chunk_ev[i]);
	cudaFree(ctx->d_ticket);
	cudaFree(ctx->d_flags);
	cudaFree(ctx->d_near);
	cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

extern "C" void *mc_stream(mc_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int64_t mc_launch_count(mc_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int mc_set_stream(mc_ctx *ctx, void *stream) {
	MC_REQUIRE(ctx, MC_ERR_ARG, "ctx is NULL");
	MC_CUDA(cudaSetDevice(ctx->device));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
	return MC_OK;
}

extern "C" int mc_sync(mc_ctx *ctx) {
	MC_REQUIRE(ctx, MC_ERR_ARG, "ctx is NULL");
	MC_CUDA(cudaSetDevice(ctx->device));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// host helper: segments (Chromosome.cpp:162-258)
// ---------------------------------------------------------------------------------------------
extern "C" int mc_host_segments(const uint8_t *s, int64_t len, int32_t *segs, int max_segs) {
	// pass 1+2 fused: maximal non-N runs, merged when the gap start2 - end1 < 10, kept when >= 20 bp,
	// then cut at 1 Mbp.  A run that starts on the very last character is never closed by the
	// reference (removeN's else-if chain), so it is ignored here too.
	int nseg = 0;
	bool any_raw = false, have_cur = false;
	int64_t cs = 0, ce = 0;
	auto emit = [&](int64_t a, int64_t b) {
		const int64_t l = b - a + 1;
		if (l < 20) return;
		if (l > 1000000) {
			const int64_t frag = l / 1000000;
			for (int64_t h = 0; h < frag; h++) {
				const int64_t fs = a + h * 1000000, fe = (h == frag - 1) ? b : fs + 1000000 - 1;
				if (nseg < max_segs) { segs[2 * nseg] = (int32_t)fs; segs[2 * nseg + 1] = (int32_t)fe; }
				nseg++;
			}
		} else {
			if (nseg < max_segs) { segs[2 * nseg] = (int32_t)a; segs[2 * nseg + 1] = (int32_t)b; }
			nseg++;
		}
	};
	auto raw = [&](int64_t a, int64_t b) {
		any_raw = true;
		if (!have_cur) { cs = a; ce = b; have_cur = true; }
		else if (a - ce < 10) { ce = b; }
		else { emit(cs, ce); cs = a; ce = b; }
	};
	// fast path: no N at all (the usual case) -> the whole record is one run
	if (len > 0 && !memchr(s, 'N', (size_t)len) && !memchr(s, 'n', (size_t)len)) {
		// a one-character record never closes its run in the reference (see above)
		if (len == 1) return -1;
		emit(0, len - 1);
		return nseg;
	}
	int64_t start = -1;
	for (int64_t i = 0; i < len; i++) {
		const bool isn = (s[i] | 0x20) == 'n';
		if (!isn && start == -1) start = i;
		else if (isn && start != -1) { raw(start, i - 1); start = -1; }
		else if (i == len - 1 && !isn && start != -1) { raw(start, i); start = -1; }
	}
	if (!any_raw) return -1;
	emit(cs, ce);
	return nseg;
}

// ---------------------------------------------------------------------------------------------
// stage 0: sequences
// ---------------------------------------------------------------------------------------------
extern "C" int mc_load_sequences(mc_ctx *ctx, const uint8_t *letters, const int64_t *offsets, int64_t n,
                                 const int32_t *segs, const int64_t *seg_offsets) {
	MC_REQUIRE(ctx && letters && offsets && seg_offsets && n > 0, MC_ERR_ARG, "mc_load_sequences: bad arguments");
	MC_REQUIRE(n < (1LL << 31), MC_ERR_UNSUPPORTED, "more than 2^31 sequences");
	MC_CUDA(cudaSetDevice(ctx->device));
	free_seq(ctx);
	const int64_t total = offsets[n];
	const int64_t nseg = seg_offsets[n];
	MC_REQUIRE(total >= 0 && nseg >= 0 && (nseg == 0 || segs), MC_ERR_ARG, "mc_load_sequences: bad offsets");
	ctx->n = n; ctx->total_bases = total; ctx->nseg = nseg;
	MC_CUDA(cudaMalloc(&ctx->d_seq, (size_t)total + 64));
	MC_CUDA(cudaMalloc(&ctx->d_seq_off, (size_t)(n + 1) * sizeof(int64_t)));
	MC_CUDA(cudaMalloc(&ctx->d_seg_off, (size_t)(n + 1) * sizeof(int64_t)));
	MC_CUDA(cudaMalloc(&ctx->d_segs, (size_t)std::max<int64_t>(nseg, 1) * 2 * sizeof(int32_t)));
	MC_CUDA(cudaMemsetAsync(ctx->d_seq + total, 0, 64, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(ctx->d_seq, letters, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(ctx->d_seq_off, offsets, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(ctx->d_seg_off, seg_offsets, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
	if (nseg) MC_CUDA(cudaMemcpyAsync(ctx->d_segs, segs, (size_t)nseg * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(unsigned int), ctx->stream));
	// the letters stay letters: K1 counts them in one pass; the digit strings the aligner reads are made by
	// the in-place encode pass when they are first asked for (ensure_digits).  Invalid letters are reported now.
	ctx->digits_ready = false;
	int rc = mc_launch_validate(ctx);
	if (rc) return rc;
	unsigned int flags[4];
	MC_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->have_seq = true;
	ctx->h_seq_off.assign(offsets, offsets + n + 1);
	MC_REQUIRE(flags[0] == 0, MC_ERR_INPUT, "Invalid nucleotide in input (the reference throws InvalidInputException)");
	return MC_OK;
}

// FASTA ingest on the device: raw file bytes + per-record spans -> letters in row order + per-record flags
int mc_launch_ingest(mc_ctx *ctx, const uint8_t *raw_dev, const int64_t *span_begin_dev, const int64_t *span_end_dev, uint8_t *rec_flags_dev,
                     unsigned int *err_dev);

// The raw bytes can go up before the spans are known (the host derives the row order meanwhile): they are parked at
// the start of the scratch buffer, and the mc_ingest_fasta that follows with the same host buffer finds them there.
static size_t ingest_scratch_bytes(int64_t raw_bytes, int64_t n) {
	return Carve::need({(size_t)raw_bytes + 64, (size_t)n * 8, (size_t)n * 8, (size_t)n, 64});
}

extern "C" int mc_stage_fasta_bytes(mc_ctx *ctx, const uint8_t *raw, int64_t raw_bytes, int64_t n_records) {
	MC_REQUIRE(ctx && raw && raw_bytes >= 0 && n_records > 0, MC_ERR_ARG, "mc_stage_fasta_bytes: bad arguments");
	MC_CUDA(cudaSetDevice(ctx->device));
	ctx->staged_raw = nullptr;
	int rc = mc_ensure_scratch(ctx, ingest_scratch_bytes(raw_bytes, n_records));
	if (rc) return rc;
	Carve cv(ctx->d_scratch);
	uint8_t *d_raw = cv.take<uint8_t>((size_t)raw_bytes + 64);
	MC_CUDA(cudaMemcpyAsync(d_raw, raw, (size_t)raw_bytes, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemsetAsync(d_raw + raw_bytes, '\n', 64, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->staged_raw = raw;
	ctx->staged_raw_bytes = raw_bytes;
	ctx->staged_scratch = ctx->d_scratch;
	return MC_OK;
}

extern "C" int mc_ingest_fasta(mc_ctx *ctx, const uint8_t *raw, int64_t raw_bytes, const int64_t *span_begin, const int64_t *span_end,
                               const int64_t *offsets, int64_t n, uint8_t *rec_flags_out) {
	MC_REQUIRE(ctx && raw && span_begin && span_end && offsets && rec_flags_out && n > 0 && raw_bytes >= 0, MC_ERR_ARG, "mc_ingest_fasta: bad arguments");
	MC_REQUIRE(n < (1LL << 31), MC_ERR_UNSUPPORTED, "more than 2^31 sequences");
	for (int64_t i = 0; i < n; i++)
		MC_REQUIRE(span_begin[i] >= 0 && span_begin[i] <= span_end[i] && span_end[i] <= raw_bytes && offsets[i + 1] >= offsets[i] &&
		           offsets[i + 1] - offsets[i] <= span_end[i] - span_begin[i], MC_ERR_ARG, "mc_ingest_fasta: span / offsets of record %lld", (long long)i);
	MC_CUDA(cudaSetDevice(ctx->device));
	free_seq(ctx);
	const int64_t total = offsets[n];
	MC_REQUIRE(offsets[0] == 0 && total >= 0, MC_ERR_ARG, "mc_ingest_fasta: bad offsets");
	ctx->n = n; ctx->total_bases = total; ctx->nseg = 0;
	MC_CUDA(cudaMalloc(&ctx->d_seq, (size_t)total + 64));
	MC_CUDA(cudaMalloc(&ctx->d_seq_off, (size_t)(n + 1) * sizeof(int64_t)));
	MC_CUDA(cudaMemsetAsync(ctx->d_seq + total, 0, 64, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(ctx->d_seq_off, offsets, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
	// the raw bytes, the spans and the flags live in the scratch buffer for the length of this call
	const size_t need = ingest_scratch_bytes(raw_bytes, n);
	// (bytes staged by mc_stage_fasta_bytes are still there if the scratch buffer has not been replaced since)
	const bool staged = ctx->staged_raw == raw && ctx->staged_raw_bytes == raw_bytes && ctx->staged_scratch == ctx->d_scratch && need <= ctx->scratch_bytes;
	ctx->staged_raw = nullptr;
	int rc = mc_ensure_scratch(ctx, need);
	if (rc) return rc;
	Carve cv(ctx->d_scratch);
	uint8_t *d_raw = cv.take<uint8_t>((size_t)raw_bytes + 64);
	int64_t *d_sb = cv.take<int64_t>((size_t)n), *d_se = cv.take<int64_t>((size_t)n);
	uint8_t *d_fl = cv.take<uint8_t>((size_t)n);
	unsigned int *d_err = cv.take<unsigned int>(16);
	if (!staged) {
		MC_CUDA(cudaMemcpyAsync(d_raw, raw, (size_t)raw_bytes, cudaMemcpyHostToDevice, ctx->stream));
		MC_CUDA(cudaMemsetAsync(d_raw + raw_bytes, '\n', 64, ctx->stream));
	}
	MC_CUDA(cudaMemcpyAsync(d_sb, span_begin, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(d_se, span_end, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemsetAsync(d_err, 0, 64, ctx->stream));
	rc = mc_launch_ingest(ctx, d_raw, d_sb, d_se, d_fl, d_err);
	if (rc) return rc;
	unsigned int h_err = 0;
	MC_CUDA(cudaMemcpyAsync(rec_flags_out, d_fl, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->digits_ready = false;
	ctx->h_seq_off.assign(offsets, offsets + n + 1);
	MC_REQUIRE(h_err == 0, MC_ERR_INPUT, "mc_ingest_fasta: a record's span does not hold the number of letters its offsets announce");
	return MC_OK;
}

extern "C" int mc_load_segments(mc_ctx *ctx, const int32_t *segs, const int64_t *seg_offsets, int validate) {
	MC_REQUIRE(ctx && seg_offsets, MC_ERR_ARG, "mc_load_segments: bad arguments");
	MC_REQUIRE(ctx->d_seq && ctx->d_seq_off && !ctx->have_seq, MC_ERR_STATE, "mc_load_segments: call mc_ingest_fasta first");
	MC_CUDA(cudaSetDevice(ctx->device));
	const int64_t n = ctx->n, nseg = seg_offsets[n];
	MC_REQUIRE(nseg >= 0 && (nseg == 0 || segs), MC_ERR_ARG, "mc_load_segments: bad offsets");
	ctx->nseg = nseg;
	MC_CUDA(cudaMalloc(&ctx->d_seg_off, (size_t)(n + 1) * sizeof(int64_t)));
	MC_CUDA(cudaMalloc(&ctx->d_segs, (size_t)std::max<int64_t>(nseg, 1) * 2 * sizeof(int32_t)));
	MC_CUDA(cudaMemcpyAsync(ctx->d_seg_off, seg_offsets, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
	if (nseg) MC_CUDA(cudaMemcpyAsync(ctx->d_segs, segs, (size_t)nseg * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
	unsigned int flags[4] = {0, 0, 0, 0};
	MC_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(unsigned int), ctx->stream));
	if (validate) {
		const int rc = mc_launch_validate(ctx);
		if (rc) return rc;
		MC_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
	}
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->have_seq = true;
	MC_REQUIRE(flags[0] == 0, MC_ERR_INPUT, "Invalid nucleotide in input (the reference throws InvalidInputException)");
	return MC_OK;
}

extern "C" int mc_copy_letters(mc_ctx *ctx, uint8_t *out) {
	MC_REQUIRE(ctx && out && ctx->d_seq, MC_ERR_ARG, "mc_copy_letters: bad arguments");
	MC_CUDA(cudaSetDevice(ctx->device));
	MC_CUDA(cudaMemcpyAsync(out, ctx->d_seq, (size_t)ctx->total_bases, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	return MC_OK;
}

// ChromosomeOneDigit::encodeNucleotides (ChromosomeOneDigit.cpp:95-144) over the whole buffer, in place, once
static int ensure_digits(mc_ctx *ctx) {
	if (ctx->digits_ready) return MC_OK;
	MC_CUDA(cudaSetDevice(ctx->device));
	const int rc = mc_launch_encode(ctx);
	if (rc) return rc;
	ctx->digits_ready = true;
	return MC_OK;
}

extern "C" int mc_copy_digits(mc_ctx *ctx, uint8_t *out) {
	MC_REQUIRE(ctx && out, MC_ERR_ARG, "bad arguments");
	MC_REQUIRE(ctx->have_seq, MC_ERR_STATE, "mc_copy_digits: load sequences first");
	{
		const int rc = ensure_digits(ctx);
		if (rc) return rc;
	}
	MC_REQUIRE(ctx->have_seq, MC_ERR_STATE, "no sequences loaded");
	MC_CUDA(cudaSetDevice(ctx->device));
	MC_CUDA(cudaMemcpyAsync(out, ctx->d_seq, (size_t)ctx->total_bases, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// stage 1: histograms
// ---------------------------------------------------------------------------------------------
static int alloc_hist(mc_ctx *ctx, int64_t n, int k, int tbytes) {
	const int nbins = 1 << (2 * k);
	const size_t bytes = (size_t)n * nbins * tbytes + 256;   // +tail: 16-byte loads never leave the buffer
	if (bytes > ctx->hist_capacity) {
		if (ctx->d_hist) MC_CUDA(cudaFree(ctx->d_hist));
		ctx->d_hist = nullptr;
		MC_CUDA(cudaMalloc(&ctx->d_hist, bytes));
		ctx->hist_capacity = bytes;
	}
	if (n > ctx->aux_capacity) {
		cudaFree(ctx->d_aux); cudaFree(ctx->d_marks);
		MC_CUDA(cudaMalloc(&ctx->d_aux, ((size_t)n + 64) * sizeof(McRowAux)));   // +tail for tile-granular bulk copies
		MC_CUDA(cudaMalloc(&ctx->d_marks, (size_t)n + 64));
		ctx->aux_capacity = n;
	}
	if (nbins > ctx->sum_bins) {
		cudaFree(ctx->d_sum);
		// running sum (uint64 per bin) + truncated-mean row + its magnitude
		MC_CUDA(cudaMalloc(&ctx->d_sum, (size_t)nbins * 8 + (size_t)nbins * 2 + 64));
		ctx->sum_bins = nbins;
	}
	ctx->n = n; ctx->k = k; ctx->nbins = nbins; ctx->tbytes = tbytes;
	ctx->rows_permuted = false;
	MC_CUDA(cudaMemsetAsync(ctx->d_aux, 0, ((size_t)n + 64) * sizeof(McRowAux), ctx->stream));
	MC_CUDA(cudaMemsetAsync(ctx->d_marks, 0, (size_t)n, ctx->stream));
	ctx->members_n = 0;
	return MC_OK;
}

extern "C" int mc_build_histograms(mc_ctx *ctx, int k, int tbytes, int *tbytes_out, uint64_t *max_count_out) {
	MC_REQUIRE(ctx, MC_ERR_ARG, "ctx is NULL");
	MC_REQUIRE(ctx->have_seq, MC_ERR_STATE, "mc_build_histograms: load sequences first");
	MC_REQUIRE(k >= 1 && k <= 7, MC_ERR_UNSUPPORTED, "k=%d unsupported (1..7)", k);
	MC_REQUIRE(tbytes == 0 || tbytes == 1 || tbytes == 2, MC_ERR_ARG, "tbytes must be 0, 1 or 2");
	MC_CUDA(cudaSetDevice(ctx->device));
	int use = tbytes ? tbytes : 1;
	unsigned int flags[4] = {0, 0, 0, 0};
	for (;;) {
		int rc = alloc_hist(ctx, ctx->n, k, use);
		if (rc) return rc;
		MC_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(unsigned int), ctx->stream));
		static const bool dbg = getenv("MC_DEBUG_TIMING") != nullptr;
		cudaEvent_t e0 = nullptr, e1 = nullptr;
		if (dbg) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ctx->stream); }
		rc = mc_launch_kmer_hist(ctx, k, use);
		if (rc) return rc;
		if (dbg) cudaEventRecord(e1, ctx->stream);
		MC_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
		MC_CUDA(cudaStreamSynchronize(ctx->stream));
		if (dbg) {
			float ms = 0;
			cudaEventElapsedTime(&ms, e0, e1);
			const double by = (double)ctx->total_bases + (double)ctx->n * ctx->nbins * use;
			fprintf(stderr, "[mc_build_histograms] kmer_count_kernel k=%d, %d-byte bins: %.1f us on the device, %.0f GB/s of %.1f MB (letters read + histograms written)\n",
			        k, use, ms * 1e3, by / (ms * 1e-3) / 1e9, by / 1e6);
			cudaEventDestroy(e0); cudaEventDestroy(e1);
		}
		// Runner.cpp:75-89: the width is the smallest that holds the largest bin
		const int needed = flags[1] <= 0xffu ? 1 : (flags[1] <= 0xffffu ? 2 : 4);
		if (needed > 2) {
			mc_set_error("largest k-mer count %u needs 32-bit histograms; only 8/16-bit are on the GPU path", flags[1]);
			return MC_ERR_UNSUPPORTED;
		}
		if (tbytes == 0 && needed > use) { use = needed; continue; }
		MC_REQUIRE(needed <= use, MC_ERR_UNSUPPORTED, "largest k-mer count %u does not fit %d-byte bins", flags[1], use);
		break;
	}
	if (tbytes_out) *tbytes_out = use;
	if (max_count_out) *max_count_out = flags[1];
	ctx->have_hist = true;
	return MC_OK;
}

extern "C" int mc_load_histograms(mc_ctx *ctx, const void *hists, int tbytes, int k, const uint64_t *lens, int64_t n) {
	MC_REQUIRE(ctx && hists && lens && n > 0, MC_ERR_ARG, "mc_load_histograms: bad arguments");
	MC_REQUIRE(k >= 1 && k <= 8, MC_ERR_UNSUPPORTED, "k=%d unsupported (1..8)", k);
	MC_REQUIRE(tbytes == 1 || tbytes == 2, MC_ERR_ARG, "tbytes must be 1 or 2");
	MC_CUDA(cudaSetDevice(ctx->device));
	int rc = alloc_hist(ctx, n, k, tbytes);
	if (rc) return rc;
	const size_t bytes = (size_t)n * ctx->nbins * tbytes;
	MC_CUDA(cudaMemcpyAsync(ctx->d_hist, hists, bytes, cudaMemcpyHostToDevice, ctx->stream));
	rc = mc_ensure_scratch(ctx, (size_t)n * 8 + 256);
	if (rc) return rc;
	MC_CUDA(cudaMemcpyAsync(ctx->d_scratch, lens, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
	rc = mc_launch_point_stats(ctx, (const uint64_t *)ctx->d_scratch);
	if (rc) return rc;
	// stream-ordered: later calls on this context queue behind the copy; the host buffers may be
	// reused once the copies have been consumed
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->have_hist = true;
	return MC_OK;
}

extern "C" int mc_copy_histograms(mc_ctx *ctx, void *out) {
	MC_REQUIRE(ctx && out, MC_ERR_ARG, "bad arguments");
	MC_REQUIRE(ctx->have_hist, MC_ERR_STATE, "no histograms");
	MC_CUDA(cudaMemcpyAsync(out, ctx->d_hist, (size_t)ctx->n * ctx->nbins * ctx->tbytes, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	return MC_OK;
}

extern "C" int mc_copy_point_stats(mc_ctx *ctx, uint64_t *len, uint64_t *mag, uint64_t *sumsq) {
	MC_REQUIRE(ctx, MC_ERR_ARG, "ctx is NULL");
	MC_REQUIRE(ctx->have_hist, MC_ERR_STATE, "no histograms");
	const size_t b = (size_t)ctx->n * 8;
	(void)b;
	if (len) MC_CUDA(cudaMemcpy2DAsync(len, 8, &ctx->d_aux[0].len, sizeof(McRowAux), 8, (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
	if (mag) MC_CUDA(cudaMemcpy2DAsync(mag, 8, &ctx->d_aux[0].mag, sizeof(McRowAux), 8, (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
	if (sumsq) MC_CUDA(cudaMemcpy2DAsync(sumsq, 8, &ctx->d_aux[0].sq, sizeof(McRowAux), 8, (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// stage 2
// ---------------------------------------------------------------------------------------------
extern "C" int mc_set_model(mc_ctx *ctx, const double *mins, const double *maxs, const double *weights, int nfeat) {
	MC_REQUIRE(ctx && mins && maxs && weights, MC_ERR_ARG, "mc_set_model: bad arguments");
	MC_REQUIRE(nfeat == 3 || nfeat == 4, MC_ERR_ARG, "nfeat must be 3 or 4 (Trainer.cpp:603-646)");
	const int nlookup = nfeat >= 4 ? 5 : 4;
	for (int i = 0; i < 5; i++) {
		ctx->model.mins[i] = i < nlookup ? mins[i] : 0.0;
		ctx->model.maxs[i] = i < nlookup ? maxs[i] : 1.0;
	}
	for (int i = 0; i < 5; i++) ctx->model.w[i] = i <= nfeat ? weights[i] : 0.0;
	ctx->model.filt[0] = std::fabs(ctx->model.w[3] / (ctx->model.maxs[3] - ctx->model.mins[3]));
	ctx->model.filt[1] = nfeat >= 4 ? std::fabs(ctx->model.w[4] / (ctx->model.maxs[4] - ctx->model.mins[4])) : 0.0;
	// division-free normalisation: allowed per feature only when q + (a - b*q)*RN(1/b) reproduces
	// a / b bit for bit on a dense probe of the feature's value range (and on every special value)
	ctx->model.fast_div = 0;
	for (int j = 0; j < 5; j++) {
		const double lo = ctx->model.mins[j], b = ctx->model.maxs[j] - lo, y = 1.0 / b;
		ctx->model.rcp[j] = y;
		if (!(std::isfinite(b) && std::isfinite(y) && b != 0.0 && std::isfinite(lo)) || std::fabs(b) < 1e-290 || std::fabs(b) > 1e290) continue;
		bool ok = true;
		uint64_t rng = 0x9E3779B97F4A7C15ull + (uint64_t)j;
		for (int t = 0; t < 200000 && ok; t++) {
			rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
			const double u = (double)(rng >> 11) * (1.0 / 9007199254740992.0);   // [0,1)
			double a;
			switch (t & 3) {
			case 0: a = (u * 1.2 - 0.1) * b; break;                // inside the trained range
			case 1: a = (u * 40.0 - 20.0) * b; break;               // far outside it
			case 2: a = std::ldexp(u - 0.5, (int)(rng & 63) - 40) * b; break;   // many magnitudes
			default: a = std::floor(u * 65536.0) - lo; break;      // integer-valued raw features
			}
			const double q = a * y, r = std::fma(-b, q, a), fast = std::fma(r, y, q), ref = a / b;
			if (!(fast == ref) && !(fast != fast && ref != ref)) ok = false;
		}
		for (double a : {0.0, -0.0, b, -b, 1.0, -1.0, lo, -lo}) {
			const double q = a * y, r = std::fma(-b, q, a), fast = std::fma(r, y, q), ref = a / b;
			if (memcmp(&fast, &ref, 8) != 0) ok = false;
		}
		if (ok) ctx->model.fast_div |= 1 << j;
	}
	ctx->model.nfeat = nfeat;
	ctx->model.valid = 1;
	return MC_OK;
}

// every entry point makes its context's GPU current first: contexts of several GPUs may be driven
// from one host thread (cudaSetDevice is a no-op when the device is already current)
#define MC_NEED_HIST(ctx)                                                                             \
	do {                                                                                              \
		MC_REQUIRE((ctx) && (ctx)->have_hist, MC_ERR_STATE, "%s: histograms are not built", __func__); \
		MC_CUDA(cudaSetDevice((ctx)->device));                                                        \
	} while (0)
#define MC_NEED_MODEL(ctx) MC_REQUIRE((ctx)->model.valid, MC_ERR_STATE, "%s: mc_set_model has not been called", __func__)

static int check_rows32(mc_ctx *ctx, const int32_t *r, int64_t m) {
	for (int64_t i = 0; i < m; i++) MC_REQUIRE(r[i] >= 0 && r[i] < ctx->n, MC_ERR_ARG, "row %d out of range", (int)r[i]);
	return MC_OK;
}
static int check_rows64(mc_ctx *ctx, const int64_t *r, int64_t m) {
	for (int64_t i = 0; i < m; i++) MC_REQUIRE(r[i] >= 0 && r[i] < ctx->n, MC_ERR_ARG, "row %lld out of range", (long long)r[i]);
	return MC_OK;
}

extern "C" int mc_distance_keys(mc_ctx *ctx, const int32_t *center_rows, int C, uint16_t *keys_out) {
	MC_NEED_HIST(ctx);
	MC_REQUIRE(center_rows && keys_out && C > 0, MC_ERR_ARG, "mc_distance_keys: bad arguments");
	MC_CUDA(cudaSetDevice(ctx->device));
	int rc = check_rows32(ctx, center_rows, C);
	if (rc) return rc;
	// chunks of centers so the key buffer stays modest
	const int64_t n = ctx->n;
	int cchunk = (int)std::max<int64_t>(1, std::min<int64_t>(C, (512LL << 20) / (n * 2)));
	rc = mc_ensure_scratch(ctx, Carve::need({(size_t)C * 4, (size_t)cchunk * n * 2}));
	if (rc) return rc;
	Carve cv(ctx->d_scratch);
	int32_t *d_c = cv.take<int32_t>(C);
	uint16_t *d_k = cv.take<uint16_t>((size_t)cchunk * n);
	MC_CUDA(cudaMemcpyAsync(d_c, center_rows, (size_t)C * 4, cudaMemcpyHostToDevice, ctx->stream));
	for (int c0 = 0; c0 < C; c0 += cchunk) {
		const int cc = std::min(cchunk, C - c0);
		rc = mc_launch_dist_keys(ctx, d_c + c0, cc, d_k);
		if (rc) return rc;
		MC_CUDA(cudaMemcpyAsync(keys_out + (size_t)c0 * n, d_k, (size_t)cc * n * 2, cudaMemcpyDeviceToHost, ctx->stream));
		MC_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	return MC_OK;
}

static int pair_list_common(mc_ctx *ctx, const int32_t *a, const int32_t *b, int64_t m, double *raw5, uint64_t *dist,
                            double *sum, double *f0, uint8_t *flag, double *feats) {
	MC_CUDA(cudaSetDevice(ctx->device));
	int rc = check_rows32(ctx, a, m);
	if (rc) return rc;
	rc = check_rows32(ctx, b, m);
	if (rc) return rc;
	rc = mc_ensure_scratch(ctx, Carve::need({(size_t)m * 4, (size_t)m * 4, (size_t)m * 40, (size_t)m * 8, (size_t)m * 8, (size_t)m * 8, (size_t)m, (size_t)m * 32}));
	if (rc) return rc;
	Carve cv(ctx->d_scratch);
	int32_t *d_a = cv.take<int32_t>(m), *d_b = cv.take<int32_t>(m);
	double *d_raw = raw5 ? cv.take<double>(m * 5) : nullptr;
	uint64_t *d_dist = dist ? cv.take<uint64_t>(m) : nullptr;
	double *d_sum = sum ? cv.take<double>(m) : nullptr;
	double *d_f0 = f0 ? cv.take<double>(m) : nullptr;
	uint8_t *d_flag = flag ? cv.take<uint8_t>(m) : nullptr;
	double *d_feats = feats ? cv.take<double>(m * 4) : nullptr;
	MC_CUDA(cudaMemcpyAsync(d_a, a, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(d_b, b, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
	rc = mc_launch_pair_list(ctx, d_a, d_b, m, d_raw, d_dist, d_sum, d_f0, d_flag, d_feats);
	if (rc) return rc;
	if (raw5) MC_CUDA(cudaMemcpyAsync(raw5, d_raw, (size_t)m * 40, cudaMemcpyDeviceToHost, ctx->stream));
	if (dist) MC_CUDA(cudaMemcpyAsync(dist, d_dist, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
	if (sum) MC_CUDA(cudaMemcpyAsync(sum, d_sum, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
	if (f0) MC_CUDA(cudaMemcpyAsync(f0, d_f0, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
	if (flag) MC_CUDA(cudaMemcpyAsync(flag, d_flag, (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
	if (feats) MC_CUDA(cudaMemcpyAsync(feats, d_feats, (size_t)m * 32, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	return MC_OK;
}

extern "C" int mc_pair_features(mc_ctx *ctx, const int32_t *a, const int32_t *b, int64_t m, double *out5, uint64_t *dist_out) {
	MC_NEED_HIST(ctx);
	MC_REQUIRE(a && b && m >= 0, MC_ERR_ARG, "mc_pair_features: bad arguments");
	if (m == 0) return MC_OK;
	return pair_list_common(ctx, a, b, m, out5, dist_out, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int mc_pair_classify(mc_ctx *ctx, const int32_t *a, const int32_t *b, int64_t m, double *sum_out,
                                double *f0_out, uint8_t *flag_out, double *feats_out) {
	MC_NEED_HIST(ctx);
	MC_NEED_MODEL(ctx);
	MC_REQUIRE(a && b && m >= 0, MC_ERR_ARG, "mc_pair_classify: bad arguments");
	if (m == 0) return MC_OK;
	return pair_list_common(ctx, a, b, m, nullptr, nullptr, sum_out, f0_out, flag_out, feats_out);
}

// the fold the scan kernels leave to the reader: same rule as in the kernels (first maximum in
// row order wins, Trainer.cpp:99)
static void fold_partials(const mc_scan_result *p, int np, mc_scan_result *out) {
	mc_scan_result r;
	mc_scan_init(r);
	for (int i = 0; i < np; i++) mc_scan_merge(r, p[i]);
	*out = r;
}

static int ensure_scan_slots(mc_ctx *ctx) {
	if (ctx->d_scan_slots) return MC_OK;
	MC_CUDA(cudaSetDevice(ctx->device));
	MC_CUDA(cudaMalloc(&ctx->d_scan_slots, (size_t)MC_SCAN_SLOTS * MC_SCAN_PARTS * sizeof(mc_scan_result)));
	for (int i = 0; i < MC_SCAN_SLOTS; i++) ctx->slot_nparts[i] = 0;
	return MC_OK;
}

extern "C" int mc_alive_reset(mc_ctx *ctx) {
	MC_NEED_HIST(ctx);
	int rc = mc_launch_alive_reset(ctx);
	if (rc) return rc;
	MC_CUDA(cudaMemsetAsync(ctx->d_marks, 0, (size_t)ctx->n, ctx->stream));
	return MC_OK;
}

extern "C" int mc_alive_kill(mc_ctx *ctx, const int64_t *rows, int64_t m) {
	MC_NEED_HIST(ctx);
	MC_REQUIRE(rows || m == 0, MC_ERR_ARG, "bad arguments");
	int rc = check_rows64(ctx, rows, m);
	if (rc) return rc;
	for (int64_t i = 0; i < m; i++) MC_CUDA(cudaMemsetAsync(&ctx->d_aux[rows[i]].alive, 0, 4, ctx->stream));
	return MC_OK;
}

extern "C" int mc_scan(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, mc_scan_result *res, uint8_t *marks_out) {
	MC_NEED_HIST(ctx);
	MC_NEED_MODEL(ctx);
	MC_REQUIRE(res, MC_ERR_ARG, "mc_scan: res is NULL");
	MC_REQUIRE(center_row >= 0 && center_row < ctx->n, MC_ERR_ARG, "center row out of range");
	MC_REQUIRE(lo >= 0 && hi < ctx->n, MC_ERR_ARG, "scan range [%lld,%lld] out of range", (long long)lo, (long long)hi);
	MC_CUDA(cudaSetDevice(ctx->device));
	if (hi < lo) {   // empty bvec range (bvec_iterator.h:61-76 yields zero iterations)
		res->n_eval = 0; res->n_pos = 0; res->best_row = -1; res->best_f0 = -1.0;
		return MC_OK;
	}
	int rc = mc_ensure_scratch(ctx, Carve::need({(size_t)MC_SCAN_PARTS * 32}));
	if (rc) return rc;
	rc = mc_ensure_pinned(ctx, (size_t)MC_SCAN_PARTS * 32);
	if (rc) return rc;
	Carve cv(ctx->d_scratch);
	void *d_part = cv.take<uint8_t>((size_t)MC_SCAN_PARTS * 32);
	int nparts = 0;
	rc = mc_launch_scan(ctx, center_row, lo, hi, 1, d_part, &nparts);
	if (rc) return rc;
	MC_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_part, (size_t)nparts * sizeof(mc_scan_result), cudaMemcpyDeviceToHost, ctx->stream));
	if (marks_out) MC_CUDA(cudaMemcpyAsync(marks_out, ctx->d_marks + lo, (size_t)(hi - lo + 1), cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	fold_partials((const mc_scan_result *)ctx->h_pinned, nparts, res);
	return MC_OK;
}

extern "C" int mc_scan_enqueue(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked, int slot) {
	MC_NEED_HIST(ctx);
	MC_NEED_MODEL(ctx);
	MC_REQUIRE(slot >= 0 && slot < MC_SCAN_SLOTS, MC_ERR_ARG, "slot %d out of range", slot);
	MC_REQUIRE(center_row >= 0 && center_row < ctx->n, MC_ERR_ARG, "center row out of range");
	MC_REQUIRE(lo >= 0 && hi < ctx->n && lo <= hi, MC_ERR_ARG, "scan range [%lld,%lld] invalid", (long long)lo, (long long)hi);
	int rc = ensure_scan_slots(ctx);
	if (rc) return rc;
	return mc_launch_scan(ctx, center_row, lo, hi, remove_marked,
	                      (uint8_t *)ctx->d_scan_slots + (size_t)slot * MC_SCAN_PARTS * sizeof(mc_scan_result), &ctx->slot_nparts[slot]);
}

int mc_launch_scan_fold(mc_ctx *ctx, const void *slots_dev, const int *nparts_dev, int nslots, void *out_dev);

extern "C" int mc_scan_fold_dev(mc_ctx *ctx, int slot0, int nslots, mc_scan_result *out_dev) {
	MC_REQUIRE(ctx && out_dev, MC_ERR_ARG, "bad arguments");
	MC_REQUIRE(slot0 >= 0 && nslots > 0 && slot0 + nslots <= MC_SCAN_SLOTS, MC_ERR_ARG, "slot range invalid");
	MC_REQUIRE(ctx->d_scan_slots, MC_ERR_STATE, "nothing was enqueued");
	MC_CUDA(cudaSetDevice(ctx->device));
	int rc = mc_ensure_scratch(ctx, (size_t)MC_SCAN_SLOTS * sizeof(int) + 256);
	if (rc) return rc;
	int *d_np = (int *)ctx->d_scratch;
	MC_CUDA(cudaMemcpyAsync(d_np, ctx->slot_nparts + slot0, (size_t)nslots * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	return mc_launch_scan_fold(ctx, (uint8_t *)ctx->d_scan_slots + (size_t)slot0 * MC_SCAN_PARTS * sizeof(mc_scan_result), d_np, nslots, out_dev);
}

extern "C" int mc_scan_enqueue_many(mc_ctx *ctx, const int64_t *center_rows, const int64_t *lo, const int64_t *hi,
                                    int count, int remove_marked, int slot0) {
	MC_REQUIRE(center_rows && lo && hi && count > 0, MC_ERR_ARG, "mc_scan_enqueue_many: bad arguments");
	if (remove_marked == 0 && count > 1 && ctx && ctx->have_hist && ctx->tbytes * ctx->nbins >= 16 && !getenv("MC_SCAN_DIRECT") && !getenv("MC_SCAN_NO_BATCH")) {
		// scans that remove nothing are independent of each other: up to MC_SCAN_BATCH of them share one
		// launch (blockIdx.y = scan), so no launch latency sits between them
		MC_NEED_HIST(ctx);
		MC_NEED_MODEL(ctx);
		MC_REQUIRE(slot0 >= 0 && slot0 + count <= MC_SCAN_SLOTS, MC_ERR_ARG, "slot range invalid");
		int rc = ensure_scan_slots(ctx);
		if (rc) return rc;
		for (int i0 = 0; i0 < count; i0 += MC_SCAN_BATCH) {
			const int m = std::min(MC_SCAN_BATCH, count - i0);
			McScanReq req[MC_SCAN_BATCH];
			for (int i = 0; i < m; i++) {
				const int64_t c = center_rows[i0 + i], l = lo[i0 + i], h = hi[i0 + i];
				MC_REQUIRE(c >= 0 && c < ctx->n, MC_ERR_ARG, "center row out of range");
				MC_REQUIRE(l >= 0 && h < ctx->n && l <= h, MC_ERR_ARG, "scan range [%lld,%lld] invalid", (long long)l, (long long)h);
				req[i].lo = l; req[i].hi = h; req[i].center_row = c;
				req[i].partials_dev = (uint8_t *)ctx->d_scan_slots + (size_t)(slot0 + i0 + i) * MC_SCAN_PARTS * sizeof(mc_scan_result);
				req[i].marks_dev = nullptr;
				req[i].ll_partials_dev = nullptr;
				req[i].ll_tag = 0;
			}
			rc = mc_launch_scan_batch(ctx, req, m, 0, &ctx->slot_nparts[slot0 + i0], nullptr);
			if (rc == MC_ERR_UNSUPPORTED) break;   // shape only the direct-load kernel handles: one launch per scan below
			if (rc) return rc;
			if (i0 + m >= count) return MC_OK;
		}
	}
	for (int i = 0; i < count; i++) {
		const int rc = mc_scan_enqueue(ctx, center_rows[i], lo[i], hi[i], remove_marked & MC_SCAN_REMOVE, slot0 + i);
		if (rc) return rc;
	}
	return MC_OK;
}

extern "C" int mc_scan_collect(mc_ctx *ctx, int slot0, int nslots, mc_scan_result *res) {
	MC_REQUIRE(ctx && res, MC_ERR_ARG, "bad arguments");
	MC_REQUIRE(slot0 >= 0 && nslots > 0 && slot0 + nslots <= MC_SCAN_SLOTS, MC_ERR_ARG, "slot range invalid");
	MC_REQUIRE(ctx->d_scan_slots, MC_ERR_STATE, "nothing was enqueued");
	MC_CUDA(cudaSetDevice(ctx->device));
	const size_t slot_bytes = (size_t)MC_SCAN_PARTS * sizeof(mc_scan_result);
	int rc = mc_ensure_pinned(ctx, (size_t)nslots * slot_bytes);
	if (rc) return rc;
	MC_CUDA(cudaMemcpyAsync(ctx->h_pinned, (uint8_t *)ctx->d_scan_slots + (size_t)slot0 * slot_bytes, (size_t)nslots * slot_bytes,
	                        cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	for (int i = 0; i < nslots; i++)
		fold_partials((const mc_scan_result *)((uint8_t *)ctx->h_pinned + (size_t)i * slot_bytes), ctx->slot_nparts[slot0 + i], &res[i]);
	return MC_OK;
}

extern "C" int mc_near_threshold_count(mc_ctx *ctx, int64_t *count_out, int reset) {
	MC_REQUIRE(ctx && count_out, MC_ERR_ARG, "mc_near_threshold_count: bad arguments");
	MC_CUDA(cudaSetDevice(ctx->device));
	unsigned long long dev = 0;
	MC_CUDA(cudaMemcpyAsync(&dev, ctx->d_near, 8, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	*count_out = (int64_t)dev + ctx->near_host;
	if (reset) {
		MC_CUDA(cudaMemsetAsync(ctx->d_near, 0, 8, ctx->stream));
		ctx->near_host = 0;
	}
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// stage 3
// ---------------------------------------------------------------------------------------------
static int ensure_members(mc_ctx *ctx, int64_t cap) {
	if (cap <= ctx->members_cap) return MC_OK;
	int64_t *nm = nullptr;
	MC_CUDA(cudaMalloc(&nm, (size_t)cap * 8));
	if (ctx->members_n) MC_CUDA(cudaMemcpyAsync(nm, ctx->d_members, (size_t)ctx->members_n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	cudaFree(ctx->d_members);
	ctx->d_members = nm;
	ctx->members_cap = cap;
	return MC_OK;
}

extern "C" int mc_mean_nearest(mc_ctx *ctx, const int64_t *rows, int64_t m, int append, int64_t *nearest_row, double *nearest_dist) {
	MC_NEED_HIST(ctx);
	MC_REQUIRE(rows && m > 0 && nearest_row, MC_ERR_ARG, "mc_mean_nearest: bad arguments");
	MC_CUDA(cudaSetDevice(ctx->device));
	int rc = check_rows64(ctx, rows, m);
	if (rc) return rc;
	if (!append) ctx->members_n = 0;
	const int64_t total = ctx->members_n + m;
	if (total > ctx->members_cap) {
		rc = ensure_members(ctx, std::max<int64_t>(total * 2, 1024));
		if (rc) return rc;
	}
	unsigned long long *d_sum = reinterpret_cast<unsigned long long *>(ctx->d_sum);
	uint8_t *d_tq = reinterpret_cast<uint8_t *>(ctx->d_sum) + (size_t)ctx->sum_bins * 8;
	if (!append) MC_CUDA(cudaMemsetAsync(d_sum, 0, (size_t)ctx->nbins * 8, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(ctx->d_members + ctx->members_n, rows, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
	const size_t nblk = (size_t)ctx->num_sms * 4;
	rc = mc_ensure_scratch(ctx, Carve::need({nblk * 16, 64}));
	if (rc) return rc;
	Carve cv(ctx->d_scratch);
	void *d_part = cv.take<uint8_t>(nblk * 16);
	uint8_t *d_out = cv.take<uint8_t>(64);   // [0] magc, [8] row, [16] dist
	rc = mc_launch_mean_nearest(ctx, ctx->d_members + ctx->members_n, m, d_sum, ctx->d_members, total, d_tq,
	                            (unsigned long long *)d_out, d_part, (long long *)(d_out + 8), (double *)(d_out + 16));
	if (rc) return rc;
	ctx->members_n = total;
	MC_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_out, 24, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	*nearest_row = *reinterpret_cast<int64_t *>((uint8_t *)ctx->h_pinned + 8);
	if (nearest_dist) *nearest_dist = *reinterpret_cast<double *>((uint8_t *)ctx->h_pinned + 16);
	return MC_OK;
}

// The fused step publishes its result in host-mapped memory and writes a sequence word last; the
// host polls that word instead of synchronising with the stream (saves the completion interrupt /
// driver round trip of a few microseconds per step).  The stream is queried now and then so that a
// failed launch or a dead kernel is still reported.
static int wait_step(mc_ctx *ctx, unsigned long long seq) {
	volatile unsigned long long *word = reinterpret_cast<volatile unsigned long long *>((uint8_t *)ctx->h_step + 48);
	for (unsigned long spins = 0;; spins++) {
		if (*word == seq) return MC_OK;
		__builtin_ia32_pause();
		if ((spins & 0xfff) == 0xfff) {
			const cudaError_t q = cudaStreamQuery(ctx->stream);
			if (q == cudaSuccess) {   // the stream is idle: the word must be there (or the kernel never ran)
				if (*word == seq) return MC_OK;
				MC_REQUIRE(false, MC_ERR_CUDA, "fused step finished without publishing its result");
			}
			if (q != cudaErrorNotReady) MC_CUDA(q);
		}
	}
}

// device state of the fused step + its host-mapped result block: [0,48) mc_step_result, [56,60) error
// word of a sharded step, [64, ...) int32 marked rows
static int ensure_step_buffers(mc_ctx *ctx) {
	if (!ctx->d_acc) {
		MC_CUDA(cudaMalloc(&ctx->d_acc, mc_acc_dev_bytes()));
		MC_CUDA(cudaMemsetAsync(ctx->d_acc, 0, mc_acc_dev_bytes(), ctx->stream));
	}
	const size_t need = 64 + ((size_t)ctx->n + 16) * sizeof(int32_t);
	if (need > ctx->h_step_bytes) {
		if (ctx->h_step) { MC_CUDA(cudaStreamSynchronize(ctx->stream)); MC_CUDA(cudaFreeHost(ctx->h_step)); ctx->h_step = nullptr; ctx->h_step_bytes = 0; }
		MC_CUDA(cudaHostAlloc(&ctx->h_step, need, cudaHostAllocMapped));
		memset(ctx->h_step, 0, 64);   // sequence word starts at 0; step_seq counts from 1
		MC_CUDA(cudaHostGetDevicePointer(&ctx->h_step_dev, ctx->h_step, 0));
		ctx->h_step_bytes = need;
	}
	return MC_OK;
}

extern "C" int mc_accumulate_step(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int restart,
                                  mc_step_result *res, int64_t *marked_rows_out, int64_t cap) {
	MC_NEED_HIST(ctx);
	MC_NEED_MODEL(ctx);
	MC_REQUIRE(res, MC_ERR_ARG, "mc_accumulate_step: res is NULL");
	MC_REQUIRE(center_row >= 0 && center_row < ctx->n, MC_ERR_ARG, "center row out of range");
	MC_REQUIRE(hi < lo || (lo >= 0 && hi < ctx->n), MC_ERR_ARG, "scan range [%lld,%lld] out of range", (long long)lo, (long long)hi);
	MC_REQUIRE(restart || ctx->members_n > 0, MC_ERR_STATE, "mc_accumulate_step: no cluster has been started (restart = 0)");
	MC_CUDA(cudaSetDevice(ctx->device));
	int rc = ensure_members(ctx, ctx->n + 1);   // a cluster holds every row at most once
	if (rc) return rc;
	rc = ensure_step_buffers(ctx);
	if (rc) return rc;
	rc = mc_ensure_scratch(ctx, Carve::need({(size_t)MC_SCAN_PARTS * 32}));
	if (rc) return rc;
	Carve cv(ctx->d_scratch);
	void *d_part = cv.take<uint8_t>((size_t)MC_SCAN_PARTS * 32);
	int nparts = 0;
	if (hi >= lo) {
		rc = mc_launch_scan(ctx, center_row, lo, hi, 1, d_part, &nparts);
		if (rc) return rc;
	}
	const unsigned long long seq = ++ctx->step_seq;
	rc = mc_launch_accumulate_tail(ctx, center_row, lo, hi, restart, d_part, nparts, ctx->d_acc, ctx->h_step_dev,
	                               reinterpret_cast<int32_t *>((uint8_t *)ctx->h_step_dev + 64), seq, nullptr);
	if (rc) return rc;
	rc = wait_step(ctx, seq);
	if (rc) return rc;
	*res = *reinterpret_cast<const mc_step_result *>(ctx->h_step);
	ctx->members_n = res->n_members;
	if (marked_rows_out) {
		MC_REQUIRE(cap >= res->scan.n_pos, MC_ERR_ARG, "marked_rows_out holds %lld rows, %lld were marked", (long long)cap, (long long)res->scan.n_pos);
		const int32_t *src = reinterpret_cast<const int32_t *>((const uint8_t *)ctx->h_step + 64);
		for (int64_t i = 0; i < res->scan.n_pos; i++) marked_rows_out[i] = src[i];
	}
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// Phase A as one persistent kernel (phase_a.cu)
// ---------------------------------------------------------------------------------------------
int mc_pa_rows_per_tile(int tbytes, int nbins);
size_t mc_pa_exchange_bytes(int grid);
int mc_pa_max_grid();
int mc_pa_trace_slots();
size_t mc_pa_gsum_bytes(int nbins);
size_t mc_pa_range_bytes();
bool mc_pa_shape_supported(int tbytes, int nbins);
int mc_launch_pa_prepare(mc_ctx *ctx, const unsigned long long *bounds_dev, const int *row0_dev, int nb, double sim, void *range_tab_dev, unsigned int *err_dev);
int mc_launch_phase_a(mc_ctx *ctx, const unsigned long long *bounds_dev, const int *row0_dev, int nb, const void *range_tab_dev,
                      uint32_t *alive_bits_dev, unsigned long long *g_sum_dev, void *exch_dev,
                      unsigned long long *bar_dev, int *members_dev, int *cl_center_dev, int *cl_off_dev, long long *stats_dev,
                      unsigned long long *trace_dev, int trace_steps, int grid, int qmax, int *mcur_dev, double sim,
                      void *const *staging, const long long *staging_cap, long long compact_min, int compact_shift);

extern "C" int mc_accumulate_run(mc_ctx *ctx, double similarity, const uint64_t *bin_bounds, const int64_t *bin_first_row,
                                 int64_t nbins_bvec, int64_t *center_rows_out, int64_t *cluster_offsets_out,
                                 int64_t *member_rows_out, mc_run_stats *stats) {
	MC_NEED_HIST(ctx);
	MC_NEED_MODEL(ctx);
	MC_REQUIRE(bin_bounds && bin_first_row && nbins_bvec >= 1 && center_rows_out && cluster_offsets_out && member_rows_out,
	           MC_ERR_ARG, "mc_accumulate_run: bad arguments");
	MC_REQUIRE(similarity > 0 && similarity < 1, MC_ERR_ARG, "mc_accumulate_run: similarity must be between 0 and 1");
	MC_REQUIRE(mc_pa_shape_supported(ctx->tbytes, ctx->nbins), MC_ERR_UNSUPPORTED,
	           "mc_accumulate_run: histogram rows of %d bytes are not supported by the persistent kernel", ctx->tbytes * ctx->nbins);
	MC_REQUIRE(nbins_bvec <= 16384, MC_ERR_UNSUPPORTED, "mc_accumulate_run: %lld bvec bins do not fit shared memory", (long long)nbins_bvec);
	const int64_t n = ctx->n;
	const int nb = (int)nbins_bvec;
	MC_REQUIRE(bin_first_row[0] == 0 && bin_first_row[nb] == n, MC_ERR_ARG, "mc_accumulate_run: bin_first_row must run from 0 to the number of rows");
	std::vector<int> row0((size_t)nb + 1);
	for (int i = 0; i <= nb; i++) {
		MC_REQUIRE(i == 0 || bin_first_row[i] >= bin_first_row[i - 1], MC_ERR_ARG, "mc_accumulate_run: bin_first_row is not sorted");
		row0[(size_t)i] = (int)bin_first_row[i];
	}
	for (int i = 1; i < nb; i++) MC_REQUIRE(bin_bounds[i] >= bin_bounds[i - 1], MC_ERR_ARG, "mc_accumulate_run: bin_bounds is not sorted");
	const int rt = mc_pa_rows_per_tile(ctx->tbytes, ctx->nbins);
	const int64_t total_tiles = (n + rt - 1) / rt;
	int grid = (int)std::min<int64_t>(std::min(ctx->num_sms, mc_pa_max_grid()), std::max<int64_t>(1, total_tiles));
	if (getenv("MC_PA_GRID")) grid = std::max(1, std::min(grid, atoi(getenv("MC_PA_GRID"))));
	const int qmax = (int)((total_tiles + grid - 1) / grid) + 2;
	const int trace_steps = getenv("MC_PA_TRACE") ? atoi(getenv("MC_PA_TRACE")) : 0;
	const size_t words = (size_t)((n + 31) / 32) + 64;
	const size_t NB = (size_t)ctx->nbins;
	// staging for the row compactions: alive rows in order, ping-pong between two buffers (histograms, constants,
	// original row numbers, search records, bitmap).  A compaction happens once a fraction 2^-shift of the current
	// rows has left the bvec (default 1/8: the scans stream at most 8/7 of the rows they evaluate; the copies add up
	// to 7 n rows over a run, a few scans' worth), so the buffers hold 7/8 n and 49/64 n rows.  Without the memory
	// for it (or with MC_PA_NO_COMPACT) the run streams the original rows to the end.
	const size_t RBy = (size_t)ctx->tbytes * ctx->nbins;
	int compact_shift = getenv("MC_PA_COMPACT_SHIFT") ? atoi(getenv("MC_PA_COMPACT_SHIFT")) : 3;
	compact_shift = std::max(1, std::min(compact_shift, 6));
	const long long n1 = n - (n >> compact_shift);
	const long long cap[2] = {n1 + 64, n1 - (n1 >> compact_shift) + 64};
	std::vector<size_t> need_base = {(size_t)nb * 8, (size_t)(nb + 1) * 4, (size_t)n * mc_pa_range_bytes(), words * 4, mc_pa_gsum_bytes((int)NB),
	                                 mc_pa_exchange_bytes(grid), 64,
	                                 (size_t)n * 4, (size_t)n * 4, (size_t)(n + 1) * 4, 64, 64, (size_t)trace_steps * mc_pa_trace_slots() * 8 + 64, (size_t)n * 4};
	std::vector<size_t> need_st = need_base;
	for (int b = 0; b < 2; b++) {
		need_st.push_back((size_t)cap[b] * RBy + 256);
		need_st.push_back((size_t)cap[b] * sizeof(McRowAux) + 64);
		need_st.push_back((size_t)cap[b] * 4);
		need_st.push_back((size_t)cap[b] * mc_pa_range_bytes());
		need_st.push_back((size_t)(cap[b] / 32 + 2) * 4);
	}
	auto total_of = [](const std::vector<size_t> &v) { size_t t = 0; for (size_t x : v) t = align_up(t, 256) + x; return t + 256; };
	// (MC_PA_COMPACT_MIN: tests compact inputs of a few thousand rows)
	const long long compact_min = getenv("MC_PA_COMPACT_MIN") ? std::max(64LL, atoll(getenv("MC_PA_COMPACT_MIN"))) : 4096;
	bool compact = n >= 2 * compact_min && !getenv("MC_PA_NO_COMPACT");
	int rc = MC_OK;
	const bool dbg = getenv("MC_DEBUG_TIMING") != nullptr;
	auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
	double t_prev = now();
	auto lap = [&](const char *what) { if (dbg) { const double t = now(); fprintf(stderr, "[mc_accumulate_run] %-28s %.3f s\n", what, t - t_prev); t_prev = t; } };
	if (compact && mc_ensure_scratch(ctx, total_of(need_st)) != MC_OK) { compact = false; cudaGetLastError(); }
	if (!compact) {
		rc = mc_ensure_scratch(ctx, total_of(need_base));
		if (rc) return rc;
	}
	lap("scratch");
	Carve cv(ctx->d_scratch);
	unsigned long long *d_bounds = cv.take<unsigned long long>((size_t)nb);
	int *d_row0 = cv.take<int>((size_t)nb + 1);
	uint8_t *d_range = cv.take<uint8_t>((size_t)n * mc_pa_range_bytes());
	uint32_t *d_bits = cv.take<uint32_t>(words);
	unsigned long long *d_gsum = cv.take<unsigned long long>(mc_pa_gsum_bytes((int)NB) / 8);
	uint8_t *d_exch = cv.take<uint8_t>(mc_pa_exchange_bytes(grid));
	unsigned long long *d_bar = cv.take<unsigned long long>(8);
	int *d_members = cv.take<int>((size_t)n);
	int *d_center = cv.take<int>((size_t)n);
	int *d_off = cv.take<int>((size_t)n + 1);
	long long *d_stats = cv.take<long long>(8);
	unsigned int *d_err = cv.take<unsigned int>(16);
	const size_t TS = (size_t)mc_pa_trace_slots();
	unsigned long long *d_trace = cv.take<unsigned long long>((size_t)trace_steps * TS + 8);
	int *d_mcur = cv.take<int>((size_t)n);
	void *staging[10] = {};
	if (compact)
		for (int b = 0; b < 2; b++) {
			staging[b * 5 + 0] = cv.take<uint8_t>((size_t)cap[b] * RBy + 256);
			staging[b * 5 + 1] = cv.take<uint8_t>((size_t)cap[b] * sizeof(McRowAux) + 64);
			staging[b * 5 + 2] = cv.take<int>((size_t)cap[b]);
			staging[b * 5 + 3] = cv.take<uint8_t>((size_t)cap[b] * mc_pa_range_bytes());
			staging[b * 5 + 4] = cv.take<uint32_t>((size_t)(cap[b] / 32 + 2));
		}
	MC_CUDA(cudaMemcpyAsync(d_bounds, bin_bounds, (size_t)nb * 8, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(d_row0, row0.data(), (size_t)(nb + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
	// a fresh bvec: every row alive, the bits past the last row clear
	MC_CUDA(cudaMemsetAsync(d_bits, 0, words * 4, ctx->stream));
	MC_CUDA(cudaMemsetAsync(d_bits, 0xff, (size_t)(n / 32) * 4, ctx->stream));
	const uint32_t tail_word = (n % 32) ? ((1u << (n % 32)) - 1u) : 0u;
	if (n % 32) MC_CUDA(cudaMemcpyAsync(d_bits + n / 32, &tail_word, 4, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemsetAsync(d_gsum, 0, mc_pa_gsum_bytes((int)NB), ctx->stream));
	MC_CUDA(cudaMemsetAsync(d_bar, 0, 64, ctx->stream));
	MC_CUDA(cudaMemsetAsync(d_stats, 0, 64, ctx->stream));
	MC_CUDA(cudaMemsetAsync(d_err, 0, 64, ctx->stream));
	// tagged records: tag 0 = never written
	MC_CUDA(cudaMemsetAsync(d_exch, 0, mc_pa_exchange_bytes(grid), ctx->stream));
	if (trace_steps) MC_CUDA(cudaMemsetAsync(d_trace, 0, (size_t)trace_steps * TS * 8, ctx->stream));
	rc = mc_launch_pa_prepare(ctx, d_bounds, d_row0, nb, similarity, d_range, d_err);
	if (rc) return rc;
	unsigned int h_err = 0;
	MC_CUDA(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	MC_REQUIRE(h_err == 0, MC_ERR_ARG, "mc_accumulate_run: the rows of a bvec bin are not in non-decreasing length order");
	lap("uploads + search records");
	rc = mc_launch_phase_a(ctx, d_bounds, d_row0, nb, d_range, d_bits, d_gsum, d_exch, d_bar, d_members, d_center, d_off,
	                       d_stats, trace_steps ? d_trace : nullptr, trace_steps, grid, qmax, d_mcur, similarity, compact ? staging : nullptr, cap, compact_min, compact_shift);
	if (rc) return rc;
	long long h_stats[8];
	unsigned long long h_bar[2];
	MC_CUDA(cudaMemcpyAsync(h_stats, d_stats, 64, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(h_bar, d_bar, 16, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	lap("kernel");
	MC_REQUIRE(h_bar[1] == 0, MC_ERR_CUDA, "mc_accumulate_run: a grid-wide barrier timed out (code %llu)", h_bar[1]);
	const int64_t nc = h_stats[0];
	MC_REQUIRE(nc >= 0 && nc <= n, MC_ERR_CUDA, "mc_accumulate_run: the kernel reported %lld clusters", (long long)nc);
	std::vector<int> tmp((size_t)std::max<int64_t>(n + 1, 1));
	if (nc) {
		MC_CUDA(cudaMemcpyAsync(tmp.data(), d_center, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
		MC_CUDA(cudaStreamSynchronize(ctx->stream));
		for (int64_t i = 0; i < nc; i++) center_rows_out[i] = tmp[(size_t)i];
	}
	MC_CUDA(cudaMemcpyAsync(tmp.data(), d_off, (size_t)(nc + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	for (int64_t i = 0; i <= nc; i++) cluster_offsets_out[i] = tmp[(size_t)i];
	const int64_t total = cluster_offsets_out[nc];
	MC_REQUIRE(total == n || nc == 0, MC_ERR_CUDA, "mc_accumulate_run: %lld of %lld rows were assigned", (long long)total, (long long)n);
	if (total) {
		MC_CUDA(cudaMemcpyAsync(tmp.data(), d_members, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
		MC_CUDA(cudaStreamSynchronize(ctx->stream));
		for (int64_t i = 0; i < total; i++) member_rows_out[i] = tmp[(size_t)i];
	}
	ctx->near_host += h_stats[3];
	if (stats) {
		stats->n_clusters = nc;
		stats->n_scans = h_stats[1];
		stats->n_evals = h_stats[2];
		stats->n_near_threshold = h_stats[3];
		stats->n_steps = h_stats[4];
		stats->device_seconds = (double)h_stats[5] * 1e-9;
		stats->n_compactions = h_stats[6];
	}
	if (trace_steps) {
		std::vector<unsigned long long> tr((size_t)trace_steps * TS);
		MC_CUDA(cudaMemcpy(tr.data(), d_trace, tr.size() * 8, cudaMemcpyDeviceToHost));
		// CTA 0, averages over the traced steps in us: range of the window, scan, exchange of the summaries
		// (incl. waiting for the slowest CTA), tail, exchange of the candidates, rest
		double acc[32] = {0};
		long cnt[32] = {0};
		const long steps = (long)std::min<long long>(trace_steps, h_stats[4]);
		for (long s = 0; s < steps; s++) {
			const unsigned long long *t = tr.data() + (size_t)s * TS;
			auto add = [&](int k, unsigned long long a, unsigned long long b) { if (a && b && b >= a) { acc[k] += (double)(b - a) * 1e-3; cnt[k]++; } };
			add(0, t[0], t[1]); add(1, t[1], t[2]); add(2, t[2], t[3]);
			if (t[5]) { add(3, t[3], t[5]); add(4, t[5], t[6]); add(5, t[6], t[7]); }
			else add(6, t[3], t[7]);
			acc[7] += (double)(long long)t[8] * 1e-3; cnt[7]++;
			acc[8] += (double)(long long)t[9] * 1e-3; cnt[8]++;
			add(10, t[0], t[10]); add(11, t[10], t[11]); add(12, t[4], t[12]); add(13, t[12], t[13]);
			add(14, t[2], t[14]); add(15, t[14], t[4]); add(21, t[1], t[21]); add(22, t[21], t[14]);
			add(23, t[1], t[22]); add(24, t[22], t[23]); add(25, t[23], t[24]); add(26, t[24], t[25]);
			add(16, t[3], t[16]); add(17, t[16], t[17]); add(18, t[17], t[18]); add(19, t[18], t[19]); add(20, t[19], t[5]);
		}
		auto av = [&](int k) { return acc[k] / (double)std::max(1L, cnt[k]); };
		fprintf(stderr, "[mc_accumulate_run trace, %ld steps, us] range %.2f  scan %.2f  exchange1 %.2f (last CTA starts its scan %.2f and ends it %.2f after CTA 0) | "
		        "tail %.2f (x%ld)  exchange2 %.2f  rest %.2f | close %.2f (x%ld)\n",
		        steps, av(0), av(1), av(2), av(7), av(8), av(3), cnt[3], av(4), av(5), av(6), cnt[6]);
		fprintf(stderr, "[mc_accumulate_run trace, detail] range: search record loaded %.2f, position located +%.2f | exchange1: bin sums flushed %.2f, fold + fence +%.2f, "
		        "record stored +%.2f, all records folded +%.2f (last consumer warp done %.2f after the range, sums flushed +%.2f) | tail: mean %.2f, distances +%.2f, block barrier +%.2f, fold + fence +%.2f, record stored +%.2f\n",
		        av(10), av(11), av(14), av(15), av(12), av(13), av(21), av(22), av(16), av(17), av(18), av(19), av(20));
		fprintf(stderr, "[mc_accumulate_run trace, consumer warp 0] center in registers %.2f after the range, first tile landed +%.2f, reduced + decided +%.2f, rest of its tiles +%.2f\n",
		        av(23), av(24), av(25), av(26));
	}
	return MC_OK;
}

int mc_comm_scan_push(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked, int slot, int fence);
int mc_comm_combine_dev(mc_ctx *ctx, int slot, const void **rec_dev_out, unsigned int **err_dev_out);

extern "C" int mc_clone_points(mc_ctx *dst, mc_ctx *src) {
	MC_REQUIRE(dst && src && dst != src, MC_ERR_ARG, "mc_clone_points: bad arguments");
	MC_REQUIRE(src->have_hist, MC_ERR_STATE, "mc_clone_points: the source has no histograms");
	MC_CUDA(cudaSetDevice(src->device));
	MC_CUDA(cudaStreamSynchronize(src->stream));
	MC_CUDA(cudaSetDevice(dst->device));
	int rc = alloc_hist(dst, src->n, src->k, src->tbytes);
	if (rc) return rc;
	const size_t hb = (size_t)src->n * src->nbins * src->tbytes, ab = (size_t)src->n * sizeof(McRowAux);
	if (dst->device == src->device) {
		MC_CUDA(cudaMemcpyAsync(dst->d_hist, src->d_hist, hb, cudaMemcpyDeviceToDevice, dst->stream));
		MC_CUDA(cudaMemcpyAsync(dst->d_aux, src->d_aux, ab, cudaMemcpyDeviceToDevice, dst->stream));
	} else {
		MC_CUDA(cudaMemcpyPeerAsync(dst->d_hist, dst->device, src->d_hist, src->device, hb, dst->stream));
		MC_CUDA(cudaMemcpyPeerAsync(dst->d_aux, dst->device, src->d_aux, src->device, ab, dst->stream));
	}
	MC_CUDA(cudaStreamSynchronize(dst->stream));
	dst->model = src->model;
	dst->model.near = dst->d_near;   // every GPU counts into its own word
	dst->have_hist = true;
	return MC_OK;
}

// The sequences of `src` (letters or digit strings, offsets, segments) copied to `dst`, device to device: every GPU
// that aligns pairs (K4 is split by pairs) needs the strings of both partners of a pair.
extern "C" int mc_clone_sequences(mc_ctx *dst, mc_ctx *src) {
	MC_REQUIRE(dst && src && dst != src, MC_ERR_ARG, "mc_clone_sequences: bad arguments");
	MC_REQUIRE(src->have_seq, MC_ERR_STATE, "mc_clone_sequences: the source has no sequences");
	MC_CUDA(cudaSetDevice(src->device));
	MC_CUDA(cudaStreamSynchronize(src->stream));
	MC_CUDA(cudaSetDevice(dst->device));
	free_seq(dst);
	const int64_t n = src->n, total = src->total_bases, nseg = src->nseg;
	MC_REQUIRE(!dst->have_hist || dst->n == n, MC_ERR_STATE, "mc_clone_sequences: the destination holds histograms of %lld other rows", (long long)dst->n);
	dst->n = n; dst->total_bases = total; dst->nseg = nseg;
	MC_CUDA(cudaMalloc(&dst->d_seq, (size_t)total + 64));
	MC_CUDA(cudaMalloc(&dst->d_seq_off, (size_t)(n + 1) * sizeof(int64_t)));
	MC_CUDA(cudaMalloc(&dst->d_seg_off, (size_t)(n + 1) * sizeof(int64_t)));
	MC_CUDA(cudaMalloc(&dst->d_segs, (size_t)std::max<int64_t>(nseg, 1) * 2 * sizeof(int32_t)));
	auto copy = [&](void *d, const void *s_, size_t bytes) -> cudaError_t {
		if (bytes == 0) return cudaSuccess;
		return dst->device == src->device ? cudaMemcpyAsync(d, s_, bytes, cudaMemcpyDeviceToDevice, dst->stream)
		                                  : cudaMemcpyPeerAsync(d, dst->device, s_, src->device, bytes, dst->stream);
	};
	MC_CUDA(copy(dst->d_seq, src->d_seq, (size_t)total + 64));
	MC_CUDA(copy(dst->d_seq_off, src->d_seq_off, (size_t)(n + 1) * sizeof(int64_t)));
	MC_CUDA(copy(dst->d_seg_off, src->d_seg_off, (size_t)(n + 1) * sizeof(int64_t)));
	MC_CUDA(copy(dst->d_segs, src->d_segs, (size_t)nseg * 2 * sizeof(int32_t)));
	MC_CUDA(cudaStreamSynchronize(dst->stream));
	dst->h_seq_off = src->h_seq_off;
	dst->digits_ready = src->digits_ready;
	dst->rows_permuted = false;
	dst->have_seq = true;
	return MC_OK;
}

extern "C" int mc_accumulate_step_sharded(mc_ctx *const *ctxs, int world, int64_t center_row, int64_t lo, int64_t hi,
                                          int restart, mc_step_result *res, int64_t *marked_rows_out, int64_t cap) {
	MC_REQUIRE(ctxs && world >= 1 && world <= MC_MAX_PEERS && res, MC_ERR_ARG, "mc_accumulate_step_sharded: bad arguments");
	mc_ctx *root = ctxs[0];
	for (int r = 0; r < world; r++) {
		mc_ctx *c = ctxs[r];
		MC_REQUIRE(c && c->have_hist && c->model.valid, MC_ERR_STATE, "rank %d: histograms / model missing (mc_clone_points)", r);
		MC_REQUIRE(c->comm.world == world && c->comm.rank == r && c->comm.connected, MC_ERR_STATE, "rank %d: mc_comm_init + mc_comm_connect_local first", r);
		MC_REQUIRE(c->n == root->n && c->nbins == root->nbins && c->tbytes == root->tbytes, MC_ERR_STATE, "rank %d holds different points", r);
		MC_REQUIRE(!c->comm.slot_pending[0], MC_ERR_STATE, "rank %d: exchange slot 0 is busy", r);
		MC_REQUIRE(!c->comm.broken, MC_ERR_STATE, "rank %d: an earlier sharded step failed after some ranks had launched; the exchange is out of step (create the contexts again)", r);
	}
	MC_REQUIRE(center_row >= 0 && center_row < root->n, MC_ERR_ARG, "center row out of range");
	MC_REQUIRE(hi < lo || (lo >= 0 && hi < root->n), MC_ERR_ARG, "scan range [%lld,%lld] out of range", (long long)lo, (long long)hi);
	MC_REQUIRE(restart || root->members_n > 0, MC_ERR_STATE, "mc_accumulate_step_sharded: no cluster has been started (restart = 0)");
	MC_CUDA(cudaSetDevice(root->device));
	int rc = ensure_members(root, root->n + 1);
	if (rc) return rc;
	rc = ensure_step_buffers(root);
	if (rc) return rc;
	// every rank scans its tiles; marks land in rank 0's array, summaries in every inbox (the
	// system-scope fence in the kernel orders a CTA's marks before its record)
	static const bool dbg = getenv("MC_DEBUG_TIMING") != nullptr;
	static double t_launch = 0, t_scans = 0, t_tail = 0;
	static long n_steps = 0;
	auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
	const double ta = dbg ? now() : 0;
	for (int r = world - 1; r >= 0; r--) {
		mc_ctx *c = ctxs[r];
		c->comm.marks_target = r == 0 ? nullptr : root->d_marks;
		rc = mc_comm_scan_push(c, center_row, lo, hi, 1, 0, 1);
		c->comm.marks_target = nullptr;
		if (rc) {
			// ranks launched so far have advanced their epoch and pushed records, the others have not: later steps
			// would wait 4 s for records that never come.  Fail fast from now on instead.
			for (int q = 0; q < world; q++) ctxs[q]->comm.broken = true;
			return rc;
		}
	}
	double tb = 0, tc = 0;
	if (dbg) {   // phase times (serialising): launches, all scans finished, tail finished
		tb = now();
		for (int r = 0; r < world; r++) { cudaSetDevice(ctxs[r]->device); cudaStreamSynchronize(ctxs[r]->stream); }
		tc = now();
		cudaSetDevice(root->device);
	}
	const void *rec = nullptr;
	unsigned int *d_err = nullptr;
	rc = mc_comm_combine_dev(root, 0, &rec, &d_err);
	if (rc) return rc;
	const unsigned long long seq = ++root->step_seq;
	rc = mc_launch_accumulate_tail(root, center_row, lo, hi, restart, rec, 1, root->d_acc, root->h_step_dev,
	                               reinterpret_cast<int32_t *>((uint8_t *)root->h_step_dev + 64), seq, d_err);
	if (rc) return rc;
	rc = wait_step(root, seq);
	if (rc) return rc;
	unsigned int *h_err = reinterpret_cast<unsigned int *>((uint8_t *)root->h_step + 56);
	if (dbg) {
		t_launch += tb - ta; t_scans += tc - tb; t_tail += now() - tc;
		if (++n_steps % 500 == 0)
			fprintf(stderr, "[mc_accumulate_step_sharded x%d] %ld steps: launches %.1f us, scans done +%.1f us, combine+tail+sync +%.1f us (averages)\n",
			        world, n_steps, t_launch / n_steps * 1e6, t_scans / n_steps * 1e6, t_tail / n_steps * 1e6);
	}
	if (*h_err) {
		MC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned int), root->stream));
		mc_set_error("sharded step: a rank's records did not arrive (world %d)", world);
		return MC_ERR_CUDA;
	}
	*res = *reinterpret_cast<const mc_step_result *>(root->h_step);
	root->members_n = res->n_members;
	if (marked_rows_out) {
		MC_REQUIRE(cap >= res->scan.n_pos, MC_ERR_ARG, "marked_rows_out holds %lld rows, %lld were marked", (long long)cap, (long long)res->scan.n_pos);
		const int32_t *src = reinterpret_cast<const int32_t *>((const uint8_t *)root->h_step + 64);
		for (int64_t i = 0; i < res->scan.n_pos; i++) marked_rows_out[i] = src[i];
	}
	return MC_OK;
}

// staging buffers of mc_permute_rows for up to `count` rows (device copy of the rows, pinned + device
// copy of the permutation)
static int reserve_permute(mc_ctx *ctx, int64_t count) {
	const size_t rb = (size_t)ctx->nbins * ctx->tbytes;
	if ((size_t)count > ctx->tmp_rows) {
		MC_CUDA(cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_hist_tmp); cudaFree(ctx->d_aux_tmp);
		ctx->d_hist_tmp = nullptr; ctx->d_aux_tmp = nullptr; ctx->tmp_rows = 0;
		MC_CUDA(cudaMalloc(&ctx->d_hist_tmp, (size_t)count * rb + 256));
		MC_CUDA(cudaMalloc(&ctx->d_aux_tmp, ((size_t)count + 64) * sizeof(McRowAux)));
		ctx->tmp_rows = (size_t)count;
	}
	int rc = mc_ensure_scratch(ctx, (size_t)count * 4 + 256);
	if (rc) return rc;
	return mc_ensure_pinned(ctx, (size_t)count * 4);
}

// Allocate what mc_permute_rows will need for the whole point set now (e.g. while the host is busy
// with something else), so that no allocation happens between two scans later.
extern "C" int mc_reserve_permute(mc_ctx *ctx) {
	MC_NEED_HIST(ctx);
	return reserve_permute(ctx, ctx->n);
}

extern "C" int mc_permute_rows(mc_ctx *ctx, const int64_t *old_of_new, int64_t count, int64_t n_alive) {
	MC_NEED_HIST(ctx);
	MC_REQUIRE(old_of_new && count >= 0 && count <= ctx->n && n_alive >= 0 && n_alive <= count, MC_ERR_ARG, "mc_permute_rows: bad arguments");
	if (count == 0) return MC_OK;
	const bool dbg = getenv("MC_DEBUG_TIMING") != nullptr;
	auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
	double t_prev = now();
	auto lap = [&](const char *what) { if (dbg) { const double t = now(); fprintf(stderr, "[mc_permute_rows count=%lld] %-22s %.4f s\n", (long long)count, what, t - t_prev); t_prev = t; } };
	const size_t rb = (size_t)ctx->nbins * ctx->tbytes;
	int rc = reserve_permute(ctx, count);
	if (rc) return rc;
	// validate (a permutation of 0..count-1) while narrowing to 32 bits
	int32_t *h = (int32_t *)ctx->h_pinned;
	{
		std::vector<uint8_t> seen((size_t)count, 0);
		for (int64_t i = 0; i < count; i++) {
			const int64_t o = old_of_new[i];
			MC_REQUIRE(o >= 0 && o < count && !seen[(size_t)o], MC_ERR_ARG, "mc_permute_rows: old_of_new is not a permutation of 0..count-1 (entry %lld)", (long long)i);
			seen[(size_t)o] = 1;
			h[i] = (int32_t)o;
		}
	}
	lap("buffers + validation");
	int32_t *d_perm = (int32_t *)ctx->d_scratch;
	MC_CUDA(cudaMemcpyAsync(d_perm, h, (size_t)count * 4, cudaMemcpyHostToDevice, ctx->stream));
	rc = mc_launch_permute_rows(ctx, d_perm, count, n_alive, ctx->d_hist_tmp, ctx->d_aux_tmp);
	if (rc) return rc;
	MC_CUDA(cudaMemcpyAsync(ctx->d_hist, ctx->d_hist_tmp, (size_t)count * rb, cudaMemcpyDeviceToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(ctx->d_aux, ctx->d_aux_tmp, (size_t)count * sizeof(McRowAux), cudaMemcpyDeviceToDevice, ctx->stream));
	MC_CUDA(cudaMemsetAsync(ctx->d_marks, 0, (size_t)count, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));   // the pinned staging buffer is reused by later calls
	lap("upload + gather + copy");
	ctx->rows_permuted = true;
	ctx->members_n = 0;
	return MC_OK;
}

extern "C" int mc_update_centers(mc_ctx *ctx, const int64_t *center_rows, int64_t ncenters, const int64_t *cand_rows,
                                 int64_t ncand, const int64_t *cand_begin, const int64_t *cand_end, int64_t *next_rows) {
	MC_NEED_HIST(ctx);
	MC_NEED_MODEL(ctx);
	MC_REQUIRE(center_rows && cand_rows && cand_begin && cand_end && next_rows && ncenters > 0, MC_ERR_ARG, "mc_update_centers: bad arguments");
	MC_CUDA(cudaSetDevice(ctx->device));
	int rc = check_rows64(ctx, center_rows, ncenters);
	if (rc) return rc;
	rc = check_rows64(ctx, cand_rows, ncand);
	if (rc) return rc;
	std::vector<int64_t> flag_off((size_t)ncenters + 1, 0);
	for (int64_t c = 0; c < ncenters; c++) {
		MC_REQUIRE(cand_begin[c] >= 0 && cand_end[c] >= cand_begin[c] && cand_end[c] <= ncand, MC_ERR_ARG, "candidate range of center %lld is invalid", (long long)c);
		flag_off[c + 1] = flag_off[c] + (cand_end[c] - cand_begin[c]);
	}
	const size_t nflags = (size_t)flag_off[ncenters];
	rc = mc_ensure_scratch(ctx, Carve::need({(size_t)ncenters * 8, (size_t)ncand * 8, (size_t)ncenters * 8, (size_t)ncenters * 8, (size_t)(ncenters + 1) * 8, nflags + 16, (size_t)ncenters * 8}));
	if (rc) return rc;
	Carve cv(ctx->d_scratch);
	int64_t *d_cr = cv.take<int64_t>(ncenters), *d_cand = cv.take<int64_t>(std::max<int64_t>(ncand, 1));
	int64_t *d_cb = cv.take<int64_t>(ncenters), *d_ce = cv.take<int64_t>(ncenters), *d_fo = cv.take<int64_t>(ncenters + 1);
	uint8_t *d_fl = cv.take<uint8_t>(nflags + 16);
	long long *d_next = cv.take<long long>(ncenters);
	MC_CUDA(cudaMemcpyAsync(d_cr, center_rows, (size_t)ncenters * 8, cudaMemcpyHostToDevice, ctx->stream));
	if (ncand) MC_CUDA(cudaMemcpyAsync(d_cand, cand_rows, (size_t)ncand * 8, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(d_cb, cand_begin, (size_t)ncenters * 8, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(d_ce, cand_end, (size_t)ncenters * 8, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(d_fo, flag_off.data(), (size_t)(ncenters + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	rc = mc_launch_update_centers(ctx, d_cr, ncenters, d_cand, d_cb, d_ce, d_fo, d_fl, d_next);
	if (rc) return rc;
	MC_CUDA(cudaMemcpyAsync(next_rows, d_next, (size_t)ncenters * 8, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// stage 4
// ---------------------------------------------------------------------------------------------
extern "C" int mc_align_pairs(mc_ctx *ctx, const int32_t *a, const int32_t *b, int64_t m, int32_t *score, int32_t *alen, int32_t *matches) {
	MC_REQUIRE(ctx && ctx->have_seq, MC_ERR_STATE, "mc_align_pairs: load sequences first");
	MC_REQUIRE(a && b && score && alen && matches && m >= 0, MC_ERR_ARG, "mc_align_pairs: bad arguments");
	MC_REQUIRE(!ctx->rows_permuted, MC_ERR_STATE, "mc_align_pairs: rows were re-numbered by mc_permute_rows; the sequences still use the original rows");
	if (m == 0) return MC_OK;
	const bool dbg = getenv("MC_DEBUG_TIMING") != nullptr;
	auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
	double t_prev = now();
	auto lap = [&](const char *what) { if (dbg) { const double t = now(); fprintf(stderr, "[mc_align_pairs m=%lld] %-22s %.4f s\n", (long long)m, what, t - t_prev); t_prev = t; } };
	MC_CUDA(cudaSetDevice(ctx->device));
	int rc = check_rows32(ctx, a, m);
	if (rc) return rc;
	rc = check_rows32(ctx, b, m);
	if (rc) return rc;
	rc = ensure_digits(ctx);
	if (rc) return rc;
	// longest seq1 among the pairs decides the scratch line, the lengths of seq2 the strip height (the offsets are
	// kept on the host since mc_load_sequences)
	const std::vector<int64_t> &off = ctx->h_seq_off;
	MC_REQUIRE((int64_t)off.size() == ctx->n + 1, MC_ERR_STATE, "mc_align_pairs: sequence offsets missing");
	int64_t max_la = 0;
	std::vector<int64_t> lbs((size_t)m);
	for (int64_t i = 0; i < m; i++) {
		max_la = std::max(max_la, off[a[i] + 1] - off[a[i]]);
		lbs[(size_t)i] = off[b[i] + 1] - off[b[i]];
	}
	static const int force_rows = getenv("MC_NW_ROWS") ? atoi(getenv("MC_NW_ROWS")) : 0;
	const int rows_per_lane = force_rows ? force_rows : mc_nw_pick_rows(lbs.data(), m);
	const int64_t stride = align_up((size_t)max_la + 2, 32);
	// A batch too small to fill the GPU with one warp per pair, of pairs long enough to be cut into many strips, is
	// run by teams of four warps per pair (nw_identity.cu): the batch then takes a quarter of one pair's latency
	// instead of all of it.  MC_NW_TEAMS=1 / 4 / 6 / 0 forces teams (of 4 / 6 warps) / forbids them (tests run every batch both ways).
	int64_t sum_lb = 0;
	for (int64_t i = 0; i < m; i++) sum_lb += lbs[(size_t)i];
	const int force_teams = getenv("MC_NW_TEAMS") ? atoi(getenv("MC_NW_TEAMS")) : -1;   // (read per call: tests switch it)
	// (16 rows per lane, 168 registers: three CTAs of four warps or two of six per SM)
	const int64_t teams4 = (int64_t)ctx->num_sms * 3, teams6 = (int64_t)ctx->num_sms * 2;
	bool teams = force_teams >= 1 || (force_teams != 0 && m <= teams4 && sum_lb >= m * 8 * 512);
	if (max_la >= (1 << 24) || (force_rows && force_rows != 16)) teams = false;
	const int team_warps = !teams ? 1 : (force_teams == 4 || force_teams == 6 ? force_teams : (m <= teams6 ? 6 : 4));
	const int64_t resident_teams = team_warps == 6 ? teams6 : teams4;
	const int rows_launch = teams ? 16 : rows_per_lane;
	const int lines = teams ? team_warps + 1 : 2;
	int64_t nwarps = teams ? std::min<int64_t>(m, resident_teams) : std::min<int64_t>(m, (int64_t)ctx->num_sms * 32);
	// keep the scratch under ~4 GB
	while (nwarps > ctx->num_sms && nwarps * lines * stride * 24 > (4LL << 30)) nwarps /= 2;
	lap("offsets D2H");
	rc = mc_ensure_scratch(ctx, Carve::need({(size_t)m * 4, (size_t)m * 4, (size_t)m * 4, (size_t)m * 4, (size_t)m * 4, (size_t)nwarps * lines * stride * 16, (size_t)nwarps * lines * stride * 8}));
	if (rc) return rc;
	lap("scratch");
	Carve cv(ctx->d_scratch);
	int32_t *d_a = cv.take<int32_t>(m), *d_b = cv.take<int32_t>(m), *d_s = cv.take<int32_t>(m), *d_l = cv.take<int32_t>(m), *d_i = cv.take<int32_t>(m);
	void *d_sa = cv.take<uint8_t>((size_t)nwarps * lines * stride * 16);
	void *d_sb = cv.take<uint8_t>((size_t)nwarps * lines * stride * 8);
	MC_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(unsigned int), ctx->stream));
	MC_CUDA(cudaMemcpyAsync(d_a, a, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(d_b, b, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
	rc = mc_launch_nw(ctx, d_a, d_b, m, max_la, d_s, d_l, d_i, d_sa, d_sb, stride, nwarps, rows_launch, team_warps);
	if (rc) return rc;
	unsigned int flags[4];
	MC_CUDA(cudaMemcpyAsync(score, d_s, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(alen, d_l, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(matches, d_i, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	lap("kernel + copies");
	MC_REQUIRE(flags[2] == 0, MC_ERR_UNSUPPORTED, "a pair is longer than 65535 bases in total; not supported on the GPU path");
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// one-shot host-buffer entry points
// ---------------------------------------------------------------------------------------------
extern "C" int mc_kmer_histograms_host(mc_ctx *ctx, const uint8_t *letters, const int64_t *offsets, int64_t n, int k,
                                       int tbytes, void *hists_out, uint64_t *max_count_out) {
	MC_REQUIRE(ctx && letters && offsets && hists_out && n > 0, MC_ERR_ARG, "mc_kmer_histograms_host: bad arguments");
	std::vector<int32_t> segs;
	std::vector<int64_t> seg_off((size_t)n + 1, 0);
	int32_t buf[2 * 64];
	for (int64_t i = 0; i < n; i++) {
		const int64_t len = offsets[i + 1] - offsets[i];
		int ns = mc_host_segments(letters + offsets[i], len, buf, 64);
		MC_REQUIRE(ns >= 0, MC_ERR_INPUT, "sequence %lld has no non-N run (the reference throws at Chromosome.cpp:193)", (long long)i);
		if (ns > 64) {
			std::vector<int32_t> big((size_t)ns * 2);
			mc_host_segments(letters + offsets[i], len, big.data(), ns);
			segs.insert(segs.end(), big.begin(), big.end());
		} else {
			segs.insert(segs.end(), buf, buf + 2 * ns);
		}
		seg_off[i + 1] = seg_off[i] + ns;
	}
	int rc = mc_load_sequences(ctx, letters, offsets, n, segs.data(), seg_off.data());
	if (rc) return rc;
	int used = 0;
	rc = mc_build_histograms(ctx, k, tbytes, &used, max_count_out);
	if (rc) return rc;
	MC_REQUIRE(tbytes == 0 || used == tbytes, MC_ERR_STATE, "unexpected histogram width");
	return mc_copy_histograms(ctx, hists_out);
}

// get_close for several centers over HOST histograms, as a pipeline: the rows are uploaded in chunks
// on a copy stream while the compute stream derives the constants of the previous chunk and runs all
// centers against it in one launch (every center keeps its own mark array); marks return chunk by
// chunk.  What the call costs is the upload of the histograms over PCIe, not upload + compute.
extern "C" int mc_scan_host(mc_ctx *ctx, const void *hists, int tbytes, int k, const uint64_t *lens, int64_t n,
                            const int64_t *center_rows, int ncenters, mc_scan_result *res, uint8_t *marks_out) {
	MC_REQUIRE(ctx && hists && lens && center_rows && res && n > 0 && ncenters > 0, MC_ERR_ARG, "mc_scan_host: bad arguments");
	MC_NEED_MODEL(ctx);
	MC_REQUIRE(k >= 1 && k <= 8, MC_ERR_UNSUPPORTED, "k=%d unsupported (1..8)", k);
	MC_REQUIRE(tbytes == 1 || tbytes == 2, MC_ERR_ARG, "tbytes must be 1 or 2");
	MC_REQUIRE(ncenters <= MC_SCAN_SLOTS, MC_ERR_ARG, "at most %d centers per call", MC_SCAN_SLOTS);
	for (int c = 0; c < ncenters; c++) MC_REQUIRE(center_rows[c] >= 0 && center_rows[c] < n, MC_ERR_ARG, "row %lld out of range", (long long)center_rows[c]);
	MC_CUDA(cudaSetDevice(ctx->device));
	const int nbins = 1 << (2 * k);
	const size_t rb = (size_t)nbins * tbytes;
	constexpr int NCHUNK = 4;
	const bool staged_shape = rb >= 16 && (tbytes == 1 ? rb <= 4096 : rb <= 2048);   // shapes of the TMA-staged scan kernel
	const bool pipelined = staged_shape && ncenters <= MC_SCAN_BATCH && (int64_t)ncenters * NCHUNK <= MC_SCAN_SLOTS && n >= 4096 && !getenv("MC_SCAN_DIRECT");
	if (!pipelined) {
		// small inputs, many centers, or shapes only the direct-load kernel handles: upload, then scan by scan
		int rc = mc_load_histograms(ctx, hists, tbytes, k, lens, n);
		if (rc) return rc;
		rc = ensure_scan_slots(ctx);
		if (rc) return rc;
		const size_t slot_bytes = (size_t)MC_SCAN_PARTS * sizeof(mc_scan_result);
		for (int c = 0; c < ncenters; c++) {
			rc = mc_launch_scan(ctx, center_rows[c], 0, n - 1, 0, (uint8_t *)ctx->d_scan_slots + (size_t)c * slot_bytes, &ctx->slot_nparts[c]);
			if (rc) return rc;
			if (marks_out) MC_CUDA(cudaMemcpyAsync(marks_out + (size_t)c * n, ctx->d_marks, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
		}
		return mc_scan_collect(ctx, 0, ncenters, res);
	}
	// rows n .. n+ncenters-1 hold copies of the center rows, so that every chunk can be scanned as soon
	// as it has landed, wherever its centers live
	int rc = alloc_hist(ctx, n + ncenters, k, tbytes);
	if (rc) return rc;
	ctx->n = n;
	rc = ensure_scan_slots(ctx);
	if (rc) return rc;
	if (!ctx->copy_stream) {
		MC_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
		for (int i = 0; i < 8; i++) MC_CUDA(cudaEventCreateWithFlags(&ctx->chunk_ev[i], cudaEventDisableTiming));
	}
	rc = mc_ensure_scratch(ctx, Carve::need({(size_t)(n + ncenters) * 8, (size_t)ncenters * (size_t)n + 64}));
	if (rc) return rc;
	rc = mc_ensure_pinned(ctx, (size_t)ncenters * rb + (size_t)ncenters * 8 + (size_t)NCHUNK * ncenters * MC_SCAN_PARTS * sizeof(mc_scan_result));
	if (rc) return rc;
	Carve cv(ctx->d_scratch);
	uint64_t *d_lens = cv.take<uint64_t>((size_t)(n + ncenters));
	uint8_t *d_marks_multi = cv.take<uint8_t>((size_t)ncenters * (size_t)n + 64);
	uint8_t *d_hist = (uint8_t *)ctx->d_hist;
	// From here on copies out of / into the caller's buffers are in flight: EVERY failure leaves through fail(),
	// which waits for both streams (the caller may free its buffers as soon as the call returns) and drops the
	// half-uploaded histograms.
	ctx->have_hist = false;
	auto fail = [&](int code) {
		cudaStreamSynchronize(ctx->copy_stream);
		cudaStreamSynchronize(ctx->stream);
		ctx->have_hist = false;
		return code;
	};
#define SH_CUDA(call)                                                                                  \
	do {                                                                                               \
		const cudaError_t _e = (call);                                                                 \
		if (_e != cudaSuccess) {                                                                       \
			mc_set_error("mc_scan_host: %s failed: %s", #call, cudaGetErrorString(_e));               \
			return fail(MC_ERR_CUDA);                                                                  \
		}                                                                                              \
	} while (0)
	// the stream of the context may still be busy with earlier work on these buffers
	SH_CUDA(cudaEventRecord(ctx->chunk_ev[NCHUNK], ctx->stream));
	SH_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[NCHUNK], 0));
	// centers first (tiny): gathered on the host into pinned memory
	uint8_t *h_cent = (uint8_t *)ctx->h_pinned;
	uint64_t *h_clen = (uint64_t *)(h_cent + (size_t)ncenters * rb);
	mc_scan_result *h_part = (mc_scan_result *)(h_clen + ncenters);
	for (int c = 0; c < ncenters; c++) {
		memcpy(h_cent + (size_t)c * rb, (const uint8_t *)hists + (size_t)center_rows[c] * rb, rb);
		h_clen[c] = lens[center_rows[c]];
	}
	SH_CUDA(cudaMemcpyAsync(d_hist + (size_t)n * rb, h_cent, (size_t)ncenters * rb, cudaMemcpyHostToDevice, ctx->copy_stream));
	SH_CUDA(cudaMemcpyAsync(d_lens + n, h_clen, (size_t)ncenters * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
	const int64_t per = ((n + NCHUNK - 1) / NCHUNK + 1023) / 1024 * 1024;   // chunk boundaries on 1024 rows
	int nchunks = 0;
	for (int64_t r0 = 0; r0 < n; r0 += per, nchunks++) {
		const int64_t r1 = std::min(n, r0 + per);
		SH_CUDA(cudaMemcpyAsync(d_hist + (size_t)r0 * rb, (const uint8_t *)hists + (size_t)r0 * rb, (size_t)(r1 - r0) * rb, cudaMemcpyHostToDevice, ctx->copy_stream));
		SH_CUDA(cudaMemcpyAsync(d_lens + r0, lens + r0, (size_t)(r1 - r0) * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
		SH_CUDA(cudaEventRecord(ctx->chunk_ev[nchunks], ctx->copy_stream));
	}
	const size_t slot_bytes = (size_t)MC_SCAN_PARTS * sizeof(mc_scan_result);
	int chunk = 0;
	for (int64_t r0 = 0; r0 < n; r0 += per, chunk++) {
		const int64_t r1 = std::min(n, r0 + per);
		SH_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[chunk], 0));
		if (chunk == 0) {
			rc = mc_launch_point_stats_range(ctx, n, ncenters, d_lens + n, ctx->stream);
			if (rc) break;
		}
		rc = mc_launch_point_stats_range(ctx, r0, r1 - r0, d_lens + r0, ctx->stream);
		if (rc) break;
		McScanReq req[MC_SCAN_BATCH];
		for (int c = 0; c < ncenters; c++) {
			req[c].lo = r0; req[c].hi = r1 - 1; req[c].center_row = n + c;
			req[c].partials_dev = (uint8_t *)ctx->d_scan_slots + (size_t)(chunk * ncenters + c) * slot_bytes;
			req[c].marks_dev = d_marks_multi + (size_t)c * (size_t)n;
			req[c].ll_partials_dev = nullptr;
			req[c].ll_tag = 0;
		}
		// no programmatic overlap with the kernel in front: it has just written the constants these scans
		// read before their dependency wait
		ctx->pdl_enabled = false;
		rc = mc_launch_scan_batch(ctx, req, ncenters, 0, &ctx->slot_nparts[chunk * ncenters], nullptr);
		ctx->pdl_enabled = true;
		if (rc) break;
		if (marks_out && cudaMemcpy2DAsync(marks_out + r0, (size_t)n, d_marks_multi + r0, (size_t)n, (size_t)(r1 - r0), (size_t)ncenters, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
			mc_set_error("mc_scan_host: copying the marks back failed: %s", cudaGetErrorString(cudaGetLastError()));
			rc = MC_ERR_CUDA;
			break;
		}
	}
	if (rc) return fail(rc);
	SH_CUDA(cudaMemcpyAsync(h_part, ctx->d_scan_slots, (size_t)nchunks * ncenters * slot_bytes, cudaMemcpyDeviceToHost, ctx->stream));
	SH_CUDA(cudaStreamSynchronize(ctx->stream));
#undef SH_CUDA
	ctx->have_hist = true;
	// fold: per center, the CTA partials of all its chunks
	std::vector<mc_scan_result> tmp((size_t)nchunks * MC_SCAN_PARTS);
	for (int c = 0; c < ncenters; c++) {
		int np = 0;
		for (int ch = 0; ch < nchunks; ch++) {
			const int cnt = ctx->slot_nparts[ch * ncenters + c];
			memcpy(tmp.data() + np, (const uint8_t *)h_part + (size_t)(ch * ncenters + c) * slot_bytes, (size_t)cnt * sizeof(mc_scan_result));
			np += cnt;
		}
		fold_partials(tmp.data(), np, &res[c]);
	}
	return MC_OK;
}
