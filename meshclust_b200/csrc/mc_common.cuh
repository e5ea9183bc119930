// Shared definitions of the meshclust_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/meshclust_b200.h"

#define MC_NUM_SMS_FALLBACK 148
#define MC_FULL_MASK 0xffffffffu
#define MC_SCAN_PARTS 160   // per-scan partial records (>= SM count)
#define MC_SCAN_BATCH 16    // independent scans one launch can carry

struct McScanReq {          // one scan of a batch launch (host side)
	long long lo, hi, center_row;
	void *partials_dev;
	void *marks_dev;          // optional: a mark array of its own for this scan (scans of one launch that all keep their marks)
	void *ll_partials_dev;    // optional (burst path): MC_SCAN_PARTS x 8 {data, tag} words for this scan's CTA partials
	unsigned int ll_tag;
};

// ---------------------------------------------------------------------------------------------
// error plumbing (no exceptions across the C-ABI)
// ---------------------------------------------------------------------------------------------
void mc_set_error(const char *fmt, ...);

#define MC_CUDA(call)                                                                          \
	do {                                                                                       \
		cudaError_t _e = (call);                                                               \
		if (_e != cudaSuccess) {                                                               \
			mc_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
			return MC_ERR_CUDA;                                                                \
		}                                                                                      \
	} while (0)

#define MC_REQUIRE(cond, code, ...)   \
	do {                              \
		if (!(cond)) {                \
			mc_set_error(__VA_ARGS__); \
			return (code);            \
		}                             \
	} while (0)

// round(1/(1+exp(-sum))) == 1.0 (Trainer.cpp:95,102) holds exactly when fl(1+exp(-sum)) <= 2,
// i.e. exp(-sum) <= 1 + 2^-52 (ties-to-even), i.e. -sum < 1.5 * 2^-52 -- a STRICT bound: at
// -sum = 1.5 * 2^-52 the exponential lies just above the midpoint and rounds up to 1 + 2^-51, the
// sum to 2 + 2^-51 and the sigmoid below one half.  NaN is never similar.
#define MC_SIGMOID_SUM_THRESHOLD (-0x1.8p-52)
#define MC_IS_SIMILAR(sum) ((sum) > MC_SIGMOID_SUM_THRESHOLD)
// pairs this close to the decision threshold are the ones a different rounding of exp() could flip
// (north_star: "identical, except for pairs within that tolerance of the --id threshold, reported by count")
#define MC_NEAR_THRESHOLD 1e-9

// ---------------------------------------------------------------------------------------------
// the trained classifier as the kernels see it (passed by value as a kernel argument)
// lookup order [LD, INTERSECTION, MANHATTAN, PEARSON, KULCZYNSKI2]  (Feature.cpp:15-28)
// ---------------------------------------------------------------------------------------------
struct McModel {
	double mins[5];
	double maxs[5];
	double w[5];   // w[0] bias, w[1..nfeat]
	double rcp[5]; // correctly rounded 1 / (maxs - mins), for the division-free normalisation
	double filt[2]; // |w[3] * rcp[3]| and |w[4] * rcp[4]|: how far an error of PEARSON / KULCZYNSKI2 moves the GLM sum (mc_scan_decide)
	int fast_div;  // bit j set: (x - mins[j]) / (maxs[j] - mins[j]) may use rcp[j] (validated on the host)
	int nfeat;     // 3 or 4
	int valid;
	unsigned long long *near;   // device counter of decisions with |sum| < MC_NEAR_THRESHOLD (this context's GPU)
};

#ifdef __CUDACC__
__device__ __forceinline__ void mc_count_near(const McModel &m, double sum) {
	if (fabs(sum) < MC_NEAR_THRESHOLD) atomicAdd(m.near, 1ull);
}
#endif



// ---------------------------------------------------------------------------------------------
// THE scan summary and its ONE merge rule (Trainer.cpp:38-48,81,99): counts add up; the arg-max is
// the largest f0, and among equal f0 the first row in iteration order (the reference iterates
// serially with a strict >).  Every fold -- per lane, per warp, per CTA, per GPU, on the host --
// goes through mc_scan_merge.
// ---------------------------------------------------------------------------------------------
typedef mc_scan_result ScanPartial;

__host__ __device__ __forceinline__ void mc_scan_init(mc_scan_result &a) {
	a.n_eval = 0; a.n_pos = 0; a.best_row = -1; a.best_f0 = -1.0;
}

__host__ __device__ __forceinline__ void mc_scan_merge(mc_scan_result &a, const mc_scan_result &b) {
	a.n_eval += b.n_eval;
	a.n_pos += b.n_pos;
	if (b.best_row >= 0 && (b.best_f0 > a.best_f0 || (b.best_f0 == a.best_f0 && (a.best_row < 0 || b.best_row < a.best_row)))) {
		a.best_f0 = b.best_f0;
		a.best_row = b.best_row;
	}
}

#ifdef __CUDACC__
// butterfly over the 32 lanes of a warp: every lane ends with the warp's summary
__device__ __forceinline__ void mc_scan_warp_fold(mc_scan_result &v) {
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		mc_scan_result other;
		other.n_eval = __shfl_xor_sync(MC_FULL_MASK, v.n_eval, o);
		other.n_pos = __shfl_xor_sync(MC_FULL_MASK, v.n_pos, o);
		other.best_row = __shfl_xor_sync(MC_FULL_MASK, v.best_row, o);
		other.best_f0 = __shfl_xor_sync(MC_FULL_MASK, v.best_f0, o);
		mc_scan_merge(v, other);
	}
}
#endif

// per-point constants next to the histogram: one 32-byte record per row, so a tile of rows needs
// one bulk copy for the histograms and one for these
struct __align__(32) McRowAux {
	uint64_t len;     // sequence length in bases (all characters, Ns included)
	uint64_t mag;     // sum of bins (pseudo-counts included)  DivergencePoint.cpp:97-109
	uint64_t sq;      // sum of squared bins
	uint32_t alive;   // 1 while the row is still in the bvec (not yet assigned to a cluster)
	uint32_t pad;
};

// ---------------------------------------------------------------------------------------------
// multi-GPU exchange of scan summaries over NVLink peer memory.
// Every rank owns an "inbox" in its HBM; the scan kernel of rank r stores each CTA's 32-byte partial
// straight into the inbox of every rank as 8 low-latency words {u32 data, u32 epoch}: an 8-byte
// store is atomic, so a reader that sees the epoch in a word also sees its data and no fence or
// remote atomic is needed (the scheme NCCL calls LL).  Inbox address of a record:
//   ((parity * MC_XSLOTS + slot) * MC_MAX_PEERS + source rank) * MC_SCAN_PARTS + CTA) * 64
// parity = epoch & 1 double-buffers a slot: a rank can be at most one exchange ahead of a peer.
// ---------------------------------------------------------------------------------------------
#define MC_MAX_PEERS 8
#ifndef MC_XSLOTS
#define MC_XSLOTS 64
#endif
#define MC_LL_RECORD_BYTES 64
#define MC_INBOX_BYTES ((size_t)2 * MC_XSLOTS * MC_MAX_PEERS * MC_SCAN_PARTS * MC_LL_RECORD_BYTES)

struct McPeerPush {                            // by-value argument of the scan kernel
	unsigned long long inbox[MC_MAX_PEERS];    // every rank's inbox as this GPU addresses it
	int world;                                 // 0: single GPU, nothing is sent
	int rank;
	unsigned int epoch;                        // flag value of this exchange, never 0
	unsigned int fence;                        // 1: make earlier peer stores (marks) visible before the record
	unsigned long long slot_off;               // byte offset of (parity, slot) inside an inbox
	int tiles_only;                            // host-side only (launcher): 1 selects the variant that sends nothing
};

struct McComm {
	int world = 0, rank = 0;
	uint8_t *inbox = nullptr;                  // this rank's inbox (cudaMalloc, IPC-exportable)
	uint8_t *peer_inbox[MC_MAX_PEERS] = {};
	bool ipc_opened[MC_MAX_PEERS] = {};
	bool connected = false;
	bool broken = false;      // a sharded step failed after some ranks had launched: epochs are out of step, every later exchange would time out
	unsigned int slot_epoch[MC_XSLOTS] = {};
	unsigned char slot_pending[MC_XSLOTS] = {};   // 0 free, 1 scan enqueued, 2 combine enqueued, 3 burst scan enqueued (one folded record per rank)
	void *d_out = nullptr;                     // MC_XSLOTS combined records + error word
	void *h_out = nullptr;                     // pinned host copy of d_out
	cudaEvent_t done = nullptr;                // recorded behind the last combine
	uint8_t *marks_target = nullptr;           // sharded Phase A: the marks array of the rank that runs the tail
	cudaStream_t xstream = nullptr;            // exchange stream of the burst path (fold + send + combine)
	unsigned int *d_ll_partials = nullptr;     // per exchange slot: MC_SCAN_PARTS CTA partials as {data, tag} words (burst path)
	unsigned int slot_uses[MC_XSLOTS] = {};
	cudaEvent_t burst_done[4] = {};            // per bank of MC_SCAN_BATCH slots: the burst's summaries are on the host
};

// ---------------------------------------------------------------------------------------------
// the context
// ---------------------------------------------------------------------------------------------
struct mc_ctx {
	int device = 0;
	int num_sms = MC_NUM_SMS_FALLBACK;
	cudaStream_t stream = nullptr;
	cudaStream_t own_stream = nullptr;
	int64_t launches = 0;
	bool pdl_enabled = true;   // scans are launched with programmatic stream serialization unless a caller switches it off

	// sequences
	int64_t n = 0;            // rows
	int64_t total_bases = 0;
	uint8_t *d_seq = nullptr;      // letters; digits in place once an aligner asked for them
	bool digits_ready = false;
	int64_t *d_seq_off = nullptr;  // n+1
	std::vector<int64_t> h_seq_off; // the same on the host
	int32_t *d_segs = nullptr;     // 2*nseg
	int64_t *d_seg_off = nullptr;  // n+1
	int64_t nseg = 0;
	bool have_seq = false;
	const uint8_t *staged_raw = nullptr;   // mc_stage_fasta_bytes: host buffer whose bytes sit at the start of the scratch buffer
	int64_t staged_raw_bytes = 0;
	void *staged_scratch = nullptr;

	// histograms
	int k = 0;
	int nbins = 0;
	int tbytes = 0;
	void *d_hist = nullptr;
	size_t hist_capacity = 0;
	McRowAux *d_aux = nullptr;
	int64_t aux_capacity = 0;
	bool have_hist = false;

	// mc_scan_host: upload stream + one event per chunk of rows (created on first use)
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t chunk_ev[8] = {};

	// staging for mc_permute_rows (allocated on first use, sized like the histograms)
	void *d_hist_tmp = nullptr;
	McRowAux *d_aux_tmp = nullptr;
	size_t tmp_rows = 0;
	bool rows_permuted = false;   // sequences (d_seq_off) still use the original rows

	// marks of the last scan (the alive set lives in McRowAux::alive)
	uint8_t *d_marks = nullptr;

	// model
	McModel model{};

	// scratch
	void *d_scratch = nullptr;
	size_t scratch_bytes = 0;
	void *h_pinned = nullptr;
	size_t pinned_bytes = 0;
	unsigned int *d_ticket = nullptr;   // last-block-done counter + flags
	unsigned int *d_flags = nullptr;    // [0] invalid-input flag, [1] max count, ...
	unsigned long long *d_near = nullptr;   // near-threshold decisions counted by the kernels
	int64_t near_host = 0;                  // ... and by runs that report their count through their own result (mc_accumulate_run)

	// result slots of mc_scan_enqueue + per-launch block partials
	void *d_scan_slots = nullptr;      // MC_SCAN_SLOTS x MC_SCAN_PARTS partial records
	void *d_scan_partials = nullptr;   // block partials of the direct-load kernel
	int slot_nparts[MC_SCAN_SLOTS];

	// mean-shift member list (Phase A)
	int64_t *d_members = nullptr;
	int64_t members_cap = 0, members_n = 0;
	uint32_t *d_sum = nullptr;   // running per-bin sum of member histograms
	int64_t sum_bins = 0;

	McComm comm;

	// fused Phase-A step (mc_accumulate_step): device state + host-mapped result / marked-row list
	void *d_acc = nullptr;
	void *h_step = nullptr;      // cudaHostAlloc(Mapped): mc_step_result, then int32 rows
	void *h_step_dev = nullptr;  // device alias of h_step
	size_t h_step_bytes = 0;
	unsigned long long step_seq = 0;   // sequence number the fused step publishes behind its result
};

void mc_comm_destroy(mc_ctx *ctx);
int mc_launch_scan_batch(mc_ctx *ctx, const McScanReq *req, int count, int remove_marked, int *nparts_out, const struct McPeerPush *push);
int mc_ensure_scratch(mc_ctx *ctx, size_t bytes);
int mc_ensure_pinned(mc_ctx *ctx, size_t bytes);

// ---------------------------------------------------------------------------------------------
// the shared FP64 epilogue: raw features -> normalised -> combos -> GLM sum.
// Bit-faithful to Feature.cpp:207-340 + Feature.h:64-88 + Trainer.cpp:84-95 as compiled by the
// reference's flags: the file is built with -fmad=false and the one fused multiply-add the
// reference binary performs in the GLM sum is written as fma().
// S = sum min(p,q), D = sum p*q; (lp,mp,sp) = (length, mag, sum p^2) of the point, (lq,mq,sq) of
// the center; N = number of bins.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mc_raw_features(uint64_t S, uint64_t D, uint64_t lp, uint64_t mp,
                                                 uint64_t sp, uint64_t lq, uint64_t mq, uint64_t sq,
                                                 int N, bool need_kul, double c[5]) {
	c[0] = (double)(lp > lq ? lp - lq : lq - lp);
	c[1] = (double)(2 * S) / (double)(mp + mq);
	c[2] = (double)(int)(mp + mq - 2 * S);
	{
		const double dap = (double)mp / N, daq = (double)mq / N;
		const long long ap = (int)round(dap), aq = (int)round(daq);
		const long long np = (long long)sp - 2 * ap * (long long)mp + (long long)N * ap * ap;
		const long long nq = (long long)sq - 2 * aq * (long long)mq + (long long)N * aq * aq;
		const long long dot =
			(long long)D - aq * (long long)mp - ap * (long long)mq + (long long)N * ap * aq;
		const double prod = (double)(np * nq);
		c[3] = (double)dot / sqrt(prod > 0.5 ? prod : 0.5);
		if (need_kul) {
			const double coeff = N * (dap + daq) / (2 * dap * daq);
			c[4] = coeff * (double)S;
		} else {
			c[4] = 0.0;
		}
	}
}

__device__ __forceinline__ void mc_eval_model(const McModel &m, const double c_raw[5], double f[4],
                                               double &sum) {
	double c[5];
	// LD, MANHATTAN, PEARSON are distances (1 - v'), INTERSECTION and KULCZYNSKI2 similarities
	// (Feature.cpp:162-204); no clamping (Feature.cpp:42-51)
#pragma unroll
	for (int j = 0; j < 5; j++) {
		const double v = (c_raw[j] - m.mins[j]) / (m.maxs[j] - m.mins[j]);
		c[j] = (j == 1 || j == 4) ? v : 1 - v;
	}
	f[0] = (1.0 * c[0]) * c[1];
	f[1] = (1.0 * (c[0] * c[0])) * (c[2] * c[2]);
	f[2] = 1.0 * c[3];
	f[3] = (1.0 * (c[0] * c[0])) * (c[4] * c[4]);
	sum = m.w[0];
	sum = fma(m.w[1], f[0], sum);
	sum = fma(m.w[2], f[1], sum);
	sum = fma(m.w[3], f[2], sum);
	if (m.nfeat >= 4) sum = fma(m.w[4], f[3], sum);
}

// ---------------------------------------------------------------------------------------------
// The same epilogue, trimmed for the streaming scan kernel (bit-identical results):
//  * uint8 histograms with k <= 6 keep every integer quantity below 2^31, so the moments are
//    32-bit IMADs and the conversions single I2F instructions instead of 64-bit emulation;
//  * mag / N is a multiplication by the exact power of two 1/N;
//  * a quotient by a per-model constant b (the normalisation ranges) is q = a*y, r = a - b*q,
//    q' = q + r*y with y = RN(1/b): the correctly rounded a/b (Markstein), which mc_set_model
//    additionally verifies on the host against true division before setting the fast_div bit;
//  * KULCZYNSKI2 is skipped entirely for 3-feature models.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double mc_div_const(double a, double b, double y) {
	const double q = a * y;
	const double r = fma(-b, q, a);
	return fma(r, y, q);
}

template <int TB>
__device__ __forceinline__ void mc_scan_epilogue(const McModel &m, uint64_t S64, uint64_t D64, uint64_t lp,
                                                  uint64_t mp64, uint64_t sp64, uint64_t lq, uint64_t mq64,
                                                  uint64_t sq64, int N, double invN, double &f0, double &sum) {
	double c[5];
	if constexpr (TB == 1) {
		const uint32_t S = (uint32_t)S64, mp = (uint32_t)mp64, mq = (uint32_t)mq64;
		c[0] = (double)(lp > lq ? lp - lq : lq - lp);
		c[1] = (double)(2u * S) / (double)(mp + mq);
		c[2] = (double)(int)(mp + mq - 2u * S);
		const double dap = (double)mp * invN, daq = (double)mq * invN;
		const int ap = (int)round(dap), aq = (int)round(daq);
		const int np = (int)(uint32_t)sp64 - 2 * ap * (int)mp + N * ap * ap;
		const int nq = (int)(uint32_t)sq64 - 2 * aq * (int)mq + N * aq * aq;
		const int dot = (int)(uint32_t)D64 - aq * (int)mp - ap * (int)mq + N * ap * aq;
		const double prod = (double)((long long)np * (long long)nq);
		c[3] = (double)dot / sqrt(prod > 0.5 ? prod : 0.5);
		if (m.nfeat >= 4) {
			const double coeff = N * (dap + daq) / (2 * dap * daq);
			c[4] = coeff * (double)S;
		} else {
			c[4] = 0.0;
		}
	} else {
		mc_raw_features(S64, D64, lp, mp64, sp64, lq, mq64, sq64, N, m.nfeat >= 4, c);
	}
	double v[5];
#pragma unroll
	for (int j = 0; j < 5; j++) {
		if (j == 4 && m.nfeat < 4) { v[j] = 0.0; continue; }
		const double num = c[j] - m.mins[j], den = m.maxs[j] - m.mins[j];
		const double q = ((m.fast_div >> j) & 1) ? mc_div_const(num, den, m.rcp[j]) : num / den;
		v[j] = (j == 1 || j == 4) ? q : 1 - q;
	}
	const double f1 = (1.0 * (v[0] * v[0])) * (v[2] * v[2]);
	f0 = (1.0 * v[0]) * v[1];
	sum = m.w[0];
	sum = fma(m.w[1], f0, sum);
	sum = fma(m.w[2], f1, sum);
	sum = fma(m.w[3], 1.0 * v[3], sum);
	if (m.nfeat >= 4) sum = fma(m.w[4], (1.0 * (v[0] * v[0])) * (v[4] * v[4]), sum);
}

// ---------------------------------------------------------------------------------------------
// The decision of a scan without most of its FP64 work.  B200 issues FP64 at a fraction of the FP32
// rate, and a scan of short rows (k <= 4) is bound by this epilogue, not by HBM.  What the caller
// needs per row is f0 (bit-exact: it is the arg-max key), the DECISION sum > threshold, and whether the
// sum lies within MC_NEAR_THRESHOLD of it.  f0 and the MANHATTAN term are cheap (one true division).
// The expensive terms are PEARSON (integer -> double conversions, a square root and a division) and,
// for 4-feature models, KULCZYNSKI2 (two more divisions).  They enter the sum linearly through
// w3 * (1 - (c3 - min3) / range3) and w4 * (v0 * q4)^2, so an approximation of c3 / c4 with a known
// error bound gives an interval for the sum; when the interval does not contain the threshold
// (and stays 1e-9 away from it) the decision is the one the exact arithmetic would take, and the
// exact terms are never computed.  Only rows whose sum is within ~1e-4 of the threshold run the full
// epilogue (bit-identical to mc_scan_epilogue).
//   c3 = dot / sqrt(prod) in FP32: three conversions, two products and MUFU.RSQ (2 ulp): relative
//   error < 2^-21; |c3| <= 1 (Cauchy-Schwarz on the exact integer moments), so |c3 - c3f| < 2^-21.
//   The bound used is 2^-18 (8x slack); c4 likewise (relative 2^-18 of |c4|).
// flags: bit 0 = similar, bit 1 = within MC_NEAR_THRESHOLD of the threshold.
// ---------------------------------------------------------------------------------------------
template <int TB>
__device__ __forceinline__ unsigned mc_scan_decide(const McModel &m, uint64_t S64, uint64_t D64, uint64_t lp,
                                                   uint64_t mp64, uint64_t sp64, uint64_t lq, uint64_t mq64,
                                                   uint64_t sq64, int N, double invN, double &f0) {
	if constexpr (TB != 1) {
		double sum;
		mc_scan_epilogue<TB>(m, S64, D64, lp, mp64, sp64, lq, mq64, sq64, N, invN, f0, sum);
		return (MC_IS_SIMILAR(sum) ? 1u : 0u) | (fabs(sum) < MC_NEAR_THRESHOLD ? 2u : 0u);
	} else {
		const uint32_t S = (uint32_t)S64, mp = (uint32_t)mp64, mq = (uint32_t)mq64;
		auto norm = [&](int j, double c) {
			const double num = c - m.mins[j], den = m.maxs[j] - m.mins[j];
			return ((m.fast_div >> j) & 1) ? mc_div_const(num, den, m.rcp[j]) : num / den;
		};
		const double c0 = (double)(lp > lq ? lp - lq : lq - lp);
		const double c1 = (double)(2u * S) / (double)(mp + mq);
		const double c2 = (double)(int)(mp + mq - 2u * S);
		const double v0 = 1 - norm(0, c0), v1 = norm(1, c1), v2 = 1 - norm(2, c2);
		const double v00 = v0 * v0;
		const double f1 = v00 * (v2 * v2);
		f0 = v0 * v1;
		const double P = fma(m.w[2], f1, fma(m.w[1], f0, m.w[0]));
		// the exact integer moments of PEARSON (Feature.cpp:274-294)
		const double dap = (double)mp * invN, daq = (double)mq * invN;
		const int ap = (int)round(dap), aq = (int)round(daq);
		const int np = (int)(uint32_t)sp64 - 2 * ap * (int)mp + N * ap * ap;
		const int nq = (int)(uint32_t)sq64 - 2 * aq * (int)mq + N * aq * aq;
		const int dot = (int)(uint32_t)D64 - aq * (int)mp - ap * (int)mq + N * ap * aq;
		{
			const float prodf = fmaxf((float)np * (float)nq, 0.5f);
			const float c3f = (float)dot * rsqrtf(prodf);
			const double q3 = ((double)c3f - m.mins[3]) * m.rcp[3];
			double ra = m.w[3] * (1.0 - q3);
			double tol = m.filt[0] * 0x1p-18;
			if (m.nfeat >= 4) {
				const float dapf = (float)dap, daqf = (float)daq;
				const float c4f = ((float)N * (dapf + daqf) / (2.0f * dapf * daqf)) * (float)S;
				const double q4 = ((double)c4f - m.mins[4]) * m.rcp[4];
				const double e4 = m.filt[1] * fabs((double)c4f) * 0x1p-18;   // |w4| x the error of q4
				ra = fma(m.w[4], v00 * (q4 * q4), ra);
				tol += v00 * e4 * (2.0 * fabs(q4) + 1.0);
			}
			const double sa = P + ra;
			tol += MC_NEAR_THRESHOLD + 1e-12 * (fabs(P) + fabs(ra));
			if (fabs(sa) > tol) return sa > 0.0 ? 1u : 0u;   // false for NaN: falls through to the exact path
		}
		const double prod = (double)((long long)np * (long long)nq);
		const double c3 = (double)dot / sqrt(prod > 0.5 ? prod : 0.5);
		const double v3 = 1 - norm(3, c3);
		double sum = fma(m.w[3], v3, P);
		if (m.nfeat >= 4) {
			const double coeff = N * (dap + daq) / (2 * dap * daq);
			const double v4 = norm(4, coeff * (double)S);
			sum = fma(m.w[4], v00 * (v4 * v4), sum);
		}
		return (MC_IS_SIMILAR(sum) ? 1u : 0u) | (fabs(sum) < MC_NEAR_THRESHOLD ? 2u : 0u);
	}
}

// DivergencePoint::distance (DivergencePoint.cpp:68-81), with the fused 1 - f*f of the compiled
// reference
__device__ __forceinline__ uint64_t mc_distance_key(uint64_t S, uint64_t magsum) {
	const double frac = (double)(2 * S) / (double)magsum;
	return (uint64_t)(10000.0 * fma(-frac, frac, 1.0));
}

// DivergencePoint::distance_d (DivergencePoint.cpp:53-65) against a truncated mean whose bins sum
// to magc
__device__ __forceinline__ double mc_distance_d(uint64_t S, uint64_t magp, uint64_t magc) {
	const double frac = (double)(2 * S) / (double)(magp + magc);
	return 10000.0 * fma(-frac, frac, 1.0);
}
