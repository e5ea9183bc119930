// K2a, the headline kernel: Trainer::get_close (Trainer.cpp:34-114) + bvec::remove_available
// (bvec.cpp:290-317) for ONE center against every row of an inclusive row range (several such scans
// per launch when they are independent of each other, blockIdx.y = scan).
//
// HBM-bound streaming kernel, warp-specialised; the CTAs of a scan are persistent over its tiles:
//   * warp 0 = producer: one lane per ring slot feeds the shared-memory stages with 1-D TMA bulk
//     copies (cp.async.bulk ... mbarrier::complete_tx): per stage one copy of RT contiguous
//     histogram rows and one of their 32-byte McRowAux records.  One CTA per SM with ~216 KB in
//     flight for rows of 1 KB and more; two (single scan) or three (several scans per launch) smaller
//     CTAs per SM for rows up to 256 bytes, so that the next scan is resident while this one ends.
//   * the other warps = consumers: a warp takes a tile of RT rows, reduces them against the center
//     held in registers (VABSDIFF4 / IDP.4A on 16-byte LDS), transposes the partials so lane l owns
//     row l, and runs the FP64 feature + GLM epilogue on all 32 lanes; marks, alive flags, count and
//     arg-max follow.
//   * programmatic dependent launch: everything up to the epilogue of a warp's first two tiles runs
//     BEFORE griddepcontrol.wait -- only the alive flags (and the right to write) depend on the previous
//     scan of the stream; histograms and the constants len / mag / sum p^2 are immutable while scans run.
//   * every CTA leaves one partial (count, positives, arg-max); the <= 148-entry fold is done by the
//     reader of the result (host, the fused tail kernel, or the exchange stream), which keeps a
//     threadfence + atomic ticket + last-CTA pass out of a kernel whose whole body is a few microseconds.
//   * multi-GPU variants (PUSH): this rank's blocks of tiles only; CTA partials go to the peers' inboxes
//     from the kernel (direct) or are left as tag-polled copies for the exchange stream (burst).
// Dead rows are copied too; the host compacts the row arrays as Phase A consumes them (mc_permute_rows).
#include "pair_core.cuh"
#include "tma_utils.cuh"

// ---- geometry ----------------------------------------------------------------------------------
constexpr int TSCAN_MAX_CONSUMERS = 24;   // small rows: one tile per warp hides the FP64 epilogue latency of short scans
constexpr int TSCAN_MAX_STAGES = 32;
constexpr int TSCAN_SMEM_BUDGET = 216 * 1024;
// Narrow rows make short scans (C2 shape: 29 MB, ~4 us of HBM time per launch), where what matters is
// how soon the NEXT scan of the stream has its tiles in flight.  With half the shared memory and half
// the warps per CTA, two CTAs fit an SM: the next launch (programmatic dependent launch) becomes
// resident and prefetches while this one still runs its epilogue.
#ifndef MC_SCAN_SMALL_CTAS_PER_SM
#define MC_SCAN_SMALL_CTAS_PER_SM 2
#endif
// A launch that carries several independent scans never needs room for a NEXT launch: its own scans
// follow each other on the SMs, and what costs is the start and the end of every CTA.  Three smaller
// CTAs per SM leave two streaming while one starts or ends (measured on the C2 shape: 5.34 -> 4.95 us
// per scan; a single scan per launch is better off with two larger CTAs: 5.4 vs 7.0 us).
#ifndef MC_SCAN_BATCH_CTAS_PER_SM
#define MC_SCAN_BATCH_CTAS_PER_SM 3
#endif
constexpr int TSCAN_SMALL_ROW = 256;

// A stage holds one consumer tile: RT consecutive rows + their McRowAux records.  RT = 32 (one row
// per lane in the epilogue) while that fits 32 KB, fewer for very wide rows (lanes >= RT idle in
// the epilogue, which is cheap next to a multi-KB row).  Every consumer warp owns a private ring
// of D stages, so a warp never waits on a barrier more than one phase ahead of it.
template <int RB, int CPS>
struct TileCfg {
	static constexpr int RT = (RB * 32 <= 32 * 1024) ? 32 : ((32 * 1024) / RB > 0 ? (32 * 1024) / RB : 1);
	static constexpr int ROW_BYTES = RT * RB;
	static constexpr int AUX_BYTES = RT * 32;
	static constexpr int STAGE_BYTES = ((ROW_BYTES + AUX_BYTES + 127) / 128) * 128;
	static constexpr int CTAS_PER_SM = CPS;
	static constexpr int NS_RAW = (TSCAN_SMEM_BUDGET / CTAS_PER_SM) / STAGE_BYTES;
	static constexpr int NS_CAP = NS_RAW > TSCAN_MAX_STAGES ? TSCAN_MAX_STAGES : NS_RAW;
	// two CTAs per SM: 11 consumers + the producer = 384 threads, i.e. 80 registers per thread (12 consumers
	// would cap them at 72 and spill inside the tile loop)
	// three CTAs: 7 consumers (256 threads); four: 5 consumers (192 threads)
	static constexpr int MAXC = RB <= TSCAN_SMALL_ROW ? (CTAS_PER_SM == 2 ? 11 : (CTAS_PER_SM == 3 ? 7 : (CTAS_PER_SM == 4 ? 5 : TSCAN_MAX_CONSUMERS / CTAS_PER_SM))) : 16;
	static constexpr int NCW = NS_CAP > MAXC ? MAXC : NS_CAP;   // active consumer warps
	static constexpr int D = NS_CAP / NCW;                                                   // ring depth per consumer
	static constexpr int NS = NCW * D;
	// sharded scans: consecutive tiles one rank owns (a power of two, ~256 KB of rows)
	static constexpr int GROUP_TILES = (262144 / ROW_BYTES) >= 1 ? (262144 / ROW_BYTES) : 1;
	static_assert(NS_CAP >= 2, "row too wide for the staged scan");
};

#ifdef MC_SCAN_TRACE
// timelines of 8 consecutive launches (keyed by the result slot the launch writes to)
__device__ unsigned long long g_scan_trace[8 * 148 * 32 * 8];
#define TRACE(slot) do { if (lane == 0) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); g_scan_trace[((((unsigned long long)partials / (MC_SCAN_PARTS * 32)) & 7) * 148 * 32 + blockIdx.x * 32 + wib) * 8 + (slot)] = _t; } } while (0)
#else
#define TRACE(slot) do {} while (0)
#endif

// PUSH selects the sharded (multi-GPU) variants; the single-GPU instantiation (0) carries neither the
// extra arguments nor a branch.  All sharded variants evaluate only this rank's tiles of the range.
//   1  direct:     this scan's CTA partials go to the peers' inboxes at the very end of the kernel --
//                  lowest latency for a single scan whose result is awaited right away (the sharded
//                  Phase-A step).  A kernel that ends with remote stores completes only when NVLink
//                  has acknowledged them, which costs a stream of back-to-back scans ~5 us per launch
//                  (measured at 2 GPUs: 9.0 -> 13.8 us), hence:
//   2  tiles only: nothing is sent from the kernel; a second stream folds the CTA partials of a
//                  burst of scans and sends one record per scan and rank (peer_exchange.cu), so the
//                  scan stream never waits for the exchange
template <int PUSH>
struct PushArg { McPeerPush v; };
template <>
struct PushArg<0> {};

__device__ __forceinline__ void peer_store_words(const McPeerPush &push, int cta, int lane, unsigned int data) {
	// lane = 8 * (peer mod 4) + word: one 8-byte {data, epoch} store per word and peer
	const int w = lane & 7;
	for (int p = lane >> 3; p < push.world; p += 4) {
		const unsigned long long dst = push.inbox[p] + push.slot_off +
			((unsigned long long)push.rank * MC_SCAN_PARTS + cta) * MC_LL_RECORD_BYTES + (unsigned long long)w * 8;
		asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(data), "r"(push.epoch) : "memory");
	}
}

// One launch carries up to MC_SCAN_BATCH scans: blockIdx.y selects the scan.  Scans of one launch must be
// independent of each other (no removal of marked rows): they are the "several centers against the
// same generation of the alive set" form of mc_scan_enqueue_many / mc_scan_sharded_burst.  Their CTAs
// queue behind each other on the SMs, so consecutive scans overlap without any launch in between.
struct ScanDesc {
	long long lo, hi, center_row;
	ScanPartial *partials;
	uint8_t *marks;               // optional: this scan's own mark array (indexed by row) instead of the shared one
	unsigned int *ll_partials;    // optional (burst path): this scan's CTA partials as {data, tag} words
	unsigned int ll_tag;
};
struct ScanBatch {
	ScanDesc d[MC_SCAN_BATCH];
};

template <int TB, int RB, int PUSH, int CPS>
__global__ void __launch_bounds__(32 * (1 + TileCfg<RB, CPS>::NCW), CPS)
scan_tma_kernel(const uint8_t *__restrict__ hist, McRowAux *__restrict__ aux, uint8_t *__restrict__ marks,
                const __grid_constant__ ScanBatch batch, long long nrows_total, McModel model,
                int remove_marked, PushArg<PUSH> push_arg) {
	const long long lo = batch.d[blockIdx.y].lo, hi = batch.d[blockIdx.y].hi, center_row = batch.d[blockIdx.y].center_row;
	ScanPartial *__restrict__ partials = batch.d[blockIdx.y].partials;
	if (batch.d[blockIdx.y].marks) marks = batch.d[blockIdx.y].marks;
	using C = RowCfg<RB>;
	using T = TileCfg<RB, CPS>;
	constexpr int NB = RB / TB;
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) uint64_t full_bar[TSCAN_MAX_STAGES];
	__shared__ __align__(8) uint64_t empty_bar[TSCAN_MAX_STAGES];
	__shared__ ScanPartial warp_part[TSCAN_MAX_CONSUMERS];

	const int lane = threadIdx.x & 31;
	const int wib = threadIdx.x >> 5;
	TRACE(0);
	// programmatic dependent launch: let the next scan of the stream start its launch + prologue
	// while this one still runs; it blocks at griddepcontrol.wait below until this grid has
	// completed and flushed (alive flags / marks written here are read there)
	asm volatile("griddepcontrol.launch_dependents;");

	if (threadIdx.x == 0) {
		for (int s = 0; s < T::NS; s++) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	TRACE(1);

	// Tiles are RT rows on an ABSOLUTE grid (tile t = rows [t*RT, (t+1)*RT)), so which GPU owns a row
	// -- and with it the row's alive flag -- never depends on the range of a scan.  Ownership goes by
	// blocks of GT consecutive tiles (~256 KB of histogram rows): block b belongs to rank b mod world.
	// Any length window wider than a few blocks spreads over all GPUs, and every rank still streams
	// long contiguous runs (interleaving single 8 KB tiles cost 40 % of the bandwidth: 5.8 -> 10 us per
	// C2-shape scan).  A rank's tiles are dealt round-robin to its CTAs, then round-robin to the CTA's
	// consumer warps.  Rows of the first / last tile outside [lo, hi] are copied but never evaluated.
	int world = 1, rank = 0;
	if constexpr (PUSH != 0) { world = push_arg.v.world; rank = push_arg.v.rank; }
	constexpr long long GT = T::GROUP_TILES;
	// tiles of this rank below tile t
	auto owned_below = [&](long long t) {
		const long long period = GT * world, rem = t % period - (long long)rank * GT;
		return (t / period) * GT + (rem < 0 ? 0 : (rem > GT ? GT : rem));
	};
	const long long t0 = lo / T::RT, t1 = hi / T::RT;
	const long long base = owned_below(t0);
	const long long ntiles = hi >= lo ? owned_below(t1 + 1) - base : 0;
	// j-th tile of this rank inside the range -> absolute tile
	auto tile_of = [&](long long j) {
		const long long J = base + j;
		return (J / GT) * (GT * world) + (long long)rank * GT + J % GT;
	};
	const long long my_first = blockIdx.x;
	const long long nmine = my_first < ntiles ? (ntiles - my_first + gridDim.x - 1) / gridDim.x : 0;

	ScanPartial mine;
	mine.n_eval = 0; mine.n_pos = 0; mine.best_row = -1; mine.best_f0 = -1.0;

	if (wib == 0) {
		// ===================== producer =====================
		// lane l owns ring slot l (consumer l / D, ring position l % D) and issues every bulk copy
		// that lands there, in order.  Slots fill in parallel (issue cost spread over the warp) while
		// one lane never runs more than one phase ahead of its slot's barriers.
		// Everything the producer copies is immutable while scans run (histograms, and the constants
		// len / mag / sum p^2 of McRowAux; the alive word that travels along is ignored), so it never
		// waits for the previous scan of the stream: its tiles are in flight while that scan finishes.
		if (lane < T::NS) {
			const int w = lane / T::D;
			for (long long u = lane % T::D; ; u += T::D) {
				const long long jj = u * T::NCW + w;          // u-th tile of consumer w
				if (jj >= nmine) break;
				const long long round = u / T::D;
				if (round > 0) mbar_wait(&empty_bar[lane], (uint32_t)(round - 1) & 1);
				const long long r0 = tile_of(my_first + jj * gridDim.x) * T::RT;
				long long nr = nrows_total - r0;
				if (nr > T::RT) nr = T::RT;
				uint8_t *dst = smem + (size_t)lane * T::STAGE_BYTES;
				mbar_expect_tx(&full_bar[lane], (uint32_t)(nr * RB + nr * 32));
				tma_bulk_g2s(dst, hist + (size_t)r0 * RB, (uint32_t)(nr * RB), &full_bar[lane]);
				tma_bulk_g2s(dst + T::ROW_BYTES, aux + r0, (uint32_t)(nr * 32), &full_bar[lane]);
			}
		}
	} else if (wib - 1 < T::NCW) {
		// ===================== consumers =====================
		const int cw = wib - 1;
		const int g = lane / C::LPP, r = lane % C::LPP;
		const uint8_t *crow = hist + (size_t)center_row * RB;
		CenterRegs<RB> cen;
		cen.load(crow, r);
		const uint64_t lq = aux[center_row].len, mq = aux[center_row].mag, sq = aux[center_row].sq;

		// What depends on the previous scan of the stream is only WHICH rows are still alive (and the
		// right to write marks / alive flags).  The reductions and the FP64 feature + GLM epilogue of a
		// warp's first PRE tiles therefore run before griddepcontrol.wait -- while the previous scan is
		// still finishing -- and only their application (alive test, counts, arg-max, marks) waits.
		constexpr int PRE = 2;
		double pre_f0_0 = 0.0, pre_f0_1 = 0.0;
		unsigned pre_flag_0 = 0, pre_flag_1 = 0;
		int npre = 0;
		bool waited = false;
		auto row_of_tile = [&](long long uu) { return tile_of(my_first + ((long long)cw + uu * T::NCW) * gridDim.x) * T::RT + lane; };
		auto apply = [&](long long row, double f0, unsigned flag_in) {
			if (lane < T::RT && row >= lo && row <= hi) {
				unsigned flag = 0;
				if (__ldcg(&aux[row].alive)) {
					flag = flag_in & 1u;
					if (flag_in & 2u) atomicAdd(model.near, 1ull);
					mine.n_eval++;
					mine.n_pos += flag;
					if (f0 > mine.best_f0) { mine.best_f0 = f0; mine.best_row = row; }
					if (flag && remove_marked) aux[row].alive = 0;
				}
				marks[row] = (uint8_t)flag;
			}
		};
		auto dependency_wait = [&]() {
			asm volatile("griddepcontrol.wait;" ::: "memory");
			waited = true;
			if (npre > 0) apply(row_of_tile(0), pre_f0_0, pre_flag_0);
			if (npre > 1) apply(row_of_tile(1), pre_f0_1, pre_flag_1);
		};

		long long u = 0;
		for (long long jj = cw; jj < nmine; jj += T::NCW, u++) {
			const long long row_mine = row_of_tile(u);
			const bool have_row = lane < T::RT && row_mine >= lo && row_mine <= hi;
			const int slot = cw * T::D + (int)(u % T::D);
			if (u == 0) TRACE(2);
			mbar_wait(&full_bar[slot], (uint32_t)(u / T::D) & 1);
			if (u == 0) TRACE(3);
			const uint8_t *st = smem + (size_t)slot * T::STAGE_BYTES;
			McRowAux my_aux;
			my_aux.len = 0; my_aux.mag = 0; my_aux.sq = 0; my_aux.alive = 0; my_aux.pad = 0;
			if (have_row) my_aux = *reinterpret_cast<const McRowAux *>(st + T::ROW_BYTES + (size_t)lane * 32);
			// row p of the tile is reduced by lane group p / LPP in iteration p % LPP
			PairAcc<TB> part[C::LPP];
			// rows outside [lo, hi] (or past the end of the array, where the stage keeps stale bytes of an
			// older tile) are harmless: their lanes never reach the epilogue (have_row), so the
			// reduction itself is branch-free
#pragma unroll
			for (int it = 0; it < C::LPP; it++) {
				const int p = g * C::LPP + it;
				PairAcc<TB> acc;
				if (T::RT == 32 || p < T::RT) {
					const uint8_t *row = st + (size_t)p * RB;
#pragma unroll
					for (int c = 0; c < C::CH; c++) {
						const uint4 v = *reinterpret_cast<const uint4 *>(row + (size_t)(c * C::LPP + r) * 16);
						acc.add(v.x, cen.w[c][0]); acc.add(v.y, cen.w[c][1]);
						acc.add(v.z, cen.w[c][2]); acc.add(v.w, cen.w[c][3]);
					}
				}
				part[it] = acc;
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty_bar[slot]);   // the stage can be refilled during the epilogue
			const PairAcc<TB> tot = mc_transpose_reduce<C::LPP>(part, r);
			if (u == 0) TRACE(4);
			if (!waited && u >= PRE) dependency_wait();
			if (!waited) {
				// alive is not known yet: evaluate the row regardless, apply after the wait
				double f0 = 0.0;
				unsigned flag = 0;
				if (have_row) {
					flag = mc_scan_decide<TB>(model, tot.summin(my_aux.mag, mq), tot.dot(), my_aux.len, my_aux.mag, my_aux.sq, lq, mq, sq, NB, 1.0 / NB, f0);   // bit 1: near the threshold, counted when the row turns out to be alive
				}
				if (u == 0) { pre_f0_0 = f0; pre_flag_0 = flag; } else { pre_f0_1 = f0; pre_flag_1 = flag; }
				npre = (int)u + 1;
			} else if (have_row) {
				unsigned flag = 0;
				if (__ldcg(&aux[row_mine].alive)) {   // dead rows skip the epilogue
					double f0;
					const unsigned dec = mc_scan_decide<TB>(model, tot.summin(my_aux.mag, mq), tot.dot(), my_aux.len, my_aux.mag, my_aux.sq, lq, mq, sq, NB, 1.0 / NB, f0);
					flag = dec & 1u;
					if (dec & 2u) atomicAdd(model.near, 1ull);
					mine.n_eval++;
					mine.n_pos += flag;
					if (f0 > mine.best_f0) { mine.best_f0 = f0; mine.best_row = row_mine; }
					if (flag && remove_marked) aux[row_mine].alive = 0;
				}
				marks[row_mine] = (uint8_t)flag;
			}
		}
		if (!waited) dependency_wait();
		mc_scan_warp_fold(mine);
		if (lane == 0) warp_part[cw] = mine;
		TRACE(5);
	}
	__syncthreads();
	TRACE(6);
	// per-CTA partial; the (tiny) fold over <= 148 partials is left to whoever reads the result
	if (wib == 0) {
		asm volatile("griddepcontrol.wait;" ::: "memory");   // returns at once: the consumers of this CTA have passed it
		ScanPartial b;
		b.n_eval = 0; b.n_pos = 0; b.best_row = -1; b.best_f0 = -1.0;
		if (lane < T::NCW) b = warp_part[lane];
		mc_scan_warp_fold(b);
		if (lane == 0) partials[blockIdx.x] = b;
		if constexpr (PUSH == 2) {
			// burst path: the exchange stream is not ordered behind this launch by an event (which would sit
			// between two scan launches and undo their overlap) nor by a fence + counter (a microsecond at
			// the end of every CTA): the partial is written a second time as eight {data, tag} words, each
			// an atomic 8-byte store, and fold_send_kernel polls the tags
			unsigned int *ll = batch.d[blockIdx.y].ll_partials;
			if (ll) {
				unsigned long long f[4];
				f[0] = (unsigned long long)__shfl_sync(MC_FULL_MASK, b.n_eval, 0);
				f[1] = (unsigned long long)__shfl_sync(MC_FULL_MASK, b.n_pos, 0);
				f[2] = (unsigned long long)__shfl_sync(MC_FULL_MASK, b.best_row, 0);
				f[3] = (unsigned long long)__double_as_longlong(__shfl_sync(MC_FULL_MASK, b.best_f0, 0));
				if (lane < 8) {
					const unsigned long long fld = (lane >> 1) == 0 ? f[0] : ((lane >> 1) == 1 ? f[1] : ((lane >> 1) == 2 ? f[2] : f[3]));
					const unsigned int data = (lane & 1) ? (unsigned int)(fld >> 32) : (unsigned int)fld;
					unsigned int *dst = ll + ((size_t)blockIdx.x * 8 + lane) * 2;
					asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(data), "r"(batch.d[blockIdx.y].ll_tag) : "memory");
				}
			}
		}
		if constexpr (PUSH == 1) {
			// sharded scan: this CTA's partial goes straight into every rank's inbox over NVLink
			const McPeerPush &push = push_arg.v;
			if (push.fence) __threadfence_system();
			unsigned long long f[4];
			f[0] = (unsigned long long)__shfl_sync(MC_FULL_MASK, b.n_eval, 0);
			f[1] = (unsigned long long)__shfl_sync(MC_FULL_MASK, b.n_pos, 0);
			f[2] = (unsigned long long)__shfl_sync(MC_FULL_MASK, b.best_row, 0);
			f[3] = (unsigned long long)__double_as_longlong(__shfl_sync(MC_FULL_MASK, b.best_f0, 0));
			const int w = lane & 7;
			const unsigned long long fld = (w >> 1) == 0 ? f[0] : ((w >> 1) == 1 ? f[1] : ((w >> 1) == 2 ? f[2] : f[3]));
			peer_store_words(push, blockIdx.x, lane, (w & 1) ? (unsigned int)(fld >> 32) : (unsigned int)fld);
		}
		TRACE(7);
	}
}

#ifdef MC_SCAN_TRACE
extern "C" int mc_debug_scan_trace(unsigned long long *out) {
	return cudaMemcpyFromSymbol(out, g_scan_trace, sizeof(g_scan_trace)) == cudaSuccess ? 0 : -1;
}
#endif

// one warp per slot: fold its partial records (same rule as the host fold)
__global__ void scan_fold_kernel(const ScanPartial *__restrict__ slots, const int *__restrict__ nparts, int nslots,
                                 ScanPartial *__restrict__ out) {
	const int lane = threadIdx.x & 31;
	const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (s >= nslots) return;
	const ScanPartial *p = slots + (size_t)s * MC_SCAN_PARTS;
	ScanPartial b;
	b.n_eval = 0; b.n_pos = 0; b.best_row = -1; b.best_f0 = -1.0;
	for (int i = lane; i < nparts[s]; i += 32) mc_scan_merge(b, p[i]);
	mc_scan_warp_fold(b);
	if (lane == 0) out[s] = b;
}

int mc_launch_scan_fold(mc_ctx *ctx, const void *slots_dev, const int *nparts_dev, int nslots, void *out_dev) {
	const int threads = 128;
	scan_fold_kernel<<<(nslots * 32 + threads - 1) / threads, threads, 0, ctx->stream>>>((const ScanPartial *)slots_dev, nparts_dev, nslots, (ScanPartial *)out_dev);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

// rows narrower than 16 bytes (k = 1) and rows too wide for two stages keep the direct-load kernel
int mc_launch_scan_direct(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked,
                          void *partials_dev, int *nparts_out);

template <int TB, int RB, int PUSH, int CPS>
static int launch_tma_cps(mc_ctx *ctx, const McScanReq *req, int count, int remove_marked, int *nparts_out, const McPeerPush *push) {
	using T = TileCfg<RB, CPS>;
	const size_t smem = (size_t)T::NS * T::STAGE_BYTES;
	static bool attr_set[64] = {};   // function attributes are per device
	if (!attr_set[ctx->device & 63]) {
		MC_CUDA(cudaFuncSetAttribute(scan_tma_kernel<TB, RB, PUSH, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attr_set[ctx->device & 63] = true;
	}
	int64_t blocks = ctx->num_sms;
	if (count == 1 && PUSH == 0) {   // a short single scan does not need every SM
		const int64_t ntiles = req[0].hi >= req[0].lo ? req[0].hi / T::RT - req[0].lo / T::RT + 1 : 0;   // absolute tiles touched by [lo, hi]
		if (blocks > ntiles) blocks = ntiles;
		if (blocks < 1) blocks = 1;
	}
	// sharded variants: every rank always leaves num_sms records per scan, so a reader knows how many to expect
	PushArg<PUSH> pa{};
	uint8_t *marks = ctx->d_marks;
	if constexpr (PUSH != 0) {
		pa.v = *push;
		if (ctx->comm.marks_target) marks = ctx->comm.marks_target;   // marks go to the rank that compacts them
	}
	ScanBatch batch{};
	for (int i = 0; i < count; i++) {
		batch.d[i].lo = req[i].lo; batch.d[i].hi = req[i].hi; batch.d[i].center_row = req[i].center_row;
		batch.d[i].partials = (ScanPartial *)req[i].partials_dev;
		batch.d[i].marks = (uint8_t *)req[i].marks_dev;
		batch.d[i].ll_partials = (unsigned int *)req[i].ll_partials_dev;
		batch.d[i].ll_tag = req[i].ll_tag;
		nparts_out[i] = (int)blocks;
	}
	static const bool no_pdl = getenv("MC_NO_PDL") != nullptr;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3((unsigned)blocks, (unsigned)count);
	cfg.blockDim = dim3(32 * (1 + T::NCW));
	cfg.dynamicSmemBytes = smem;
	cfg.stream = ctx->stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = (no_pdl || !ctx->pdl_enabled) ? 0 : 1;
	MC_CUDA(cudaLaunchKernelEx(&cfg, scan_tma_kernel<TB, RB, PUSH, CPS>, (const uint8_t *)ctx->d_hist, ctx->d_aux, marks,
	                           batch, (long long)ctx->n, ctx->model, remove_marked, pa));
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

// rows up to 256 bytes: two CTAs per SM for a single scan per launch (room for the next launch of the
// stream), three for a launch that carries several scans; wider rows: one
template <int TB, int RB, int PUSH>
static int launch_tma_impl(mc_ctx *ctx, const McScanReq *req, int count, int remove_marked, int *nparts_out, const McPeerPush *push) {
	if constexpr (RB <= TSCAN_SMALL_ROW) {
		if (count > 1 && PUSH != 1) return launch_tma_cps<TB, RB, PUSH, MC_SCAN_BATCH_CTAS_PER_SM>(ctx, req, count, remove_marked, nparts_out, push);
		return launch_tma_cps<TB, RB, PUSH, MC_SCAN_SMALL_CTAS_PER_SM>(ctx, req, count, remove_marked, nparts_out, push);
	} else {
		return launch_tma_cps<TB, RB, PUSH, 1>(ctx, req, count, remove_marked, nparts_out, push);
	}
}

template <int TB, int RB>
static int launch_tma(mc_ctx *ctx, const McScanReq *req, int count, int remove_marked, int *nparts_out, const McPeerPush *push) {
	if (push && push->tiles_only) return launch_tma_impl<TB, RB, 2>(ctx, req, count, remove_marked, nparts_out, push);
	if (push) return launch_tma_impl<TB, RB, 1>(ctx, req, count, remove_marked, nparts_out, push);
	return launch_tma_impl<TB, RB, 0>(ctx, req, count, remove_marked, nparts_out, nullptr);
}

// rows narrower than 16 bytes (k = 1) and rows too wide for two stages keep the direct-load kernel
int mc_launch_scan_direct(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked,
                          void *partials_dev, int *nparts_out);

// `count` scans (1 <= count <= MC_SCAN_BATCH) in ONE launch; count > 1 requires scans that are
// independent of each other.  Every req[i].partials_dev must hold MC_SCAN_PARTS entries of 32 bytes;
// nparts_out[i] says how many were written.  push != NULL: sharded variants (this rank's tiles only).
// Returns MC_ERR_UNSUPPORTED for shapes only the direct-load kernel handles when count > 1 or push.
int mc_launch_scan_batch(mc_ctx *ctx, const McScanReq *req, int count, int remove_marked, int *nparts_out, const McPeerPush *push) {
	static const bool legacy = getenv("MC_SCAN_DIRECT") != nullptr;
	MC_REQUIRE(count >= 1 && count <= MC_SCAN_BATCH, MC_ERR_ARG, "scan batch of %d", count);
	const int rb = ctx->tbytes * ctx->nbins;
	if (!legacy) {
		if (ctx->tbytes == 1) {
			switch (rb) {
			case 16: return launch_tma<1, 16>(ctx, req, count, remove_marked, nparts_out, push);
			case 64: return launch_tma<1, 64>(ctx, req, count, remove_marked, nparts_out, push);
			case 256: return launch_tma<1, 256>(ctx, req, count, remove_marked, nparts_out, push);
			case 1024: return launch_tma<1, 1024>(ctx, req, count, remove_marked, nparts_out, push);
			case 4096: return launch_tma<1, 4096>(ctx, req, count, remove_marked, nparts_out, push);
			default: break;
			}
		} else {
			switch (rb) {
			case 32: return launch_tma<2, 32>(ctx, req, count, remove_marked, nparts_out, push);
			case 128: return launch_tma<2, 128>(ctx, req, count, remove_marked, nparts_out, push);
			case 512: return launch_tma<2, 512>(ctx, req, count, remove_marked, nparts_out, push);
			case 2048: return launch_tma<2, 2048>(ctx, req, count, remove_marked, nparts_out, push);
			default: break;
			}
		}
	}
	MC_REQUIRE(!push, MC_ERR_UNSUPPORTED, "sharded scans need the staged scan kernel (16-byte rows and wider, no MC_SCAN_DIRECT)");
	MC_REQUIRE(count == 1, MC_ERR_UNSUPPORTED, "scan batches need the staged scan kernel");
	return mc_launch_scan_direct(ctx, req[0].center_row, req[0].lo, req[0].hi, remove_marked, req[0].partials_dev, nparts_out);
}

int mc_launch_scan_push(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked,
                        void *partials_dev, int *nparts_out, const McPeerPush *push) {
	McScanReq r;
	r.lo = lo; r.hi = hi; r.center_row = center_row; r.partials_dev = partials_dev; r.marks_dev = nullptr; r.ll_partials_dev = nullptr; r.ll_tag = 0;
	return mc_launch_scan_batch(ctx, &r, 1, remove_marked, nparts_out, push);
}

int mc_launch_scan(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked,
                   void *partials_dev, int *nparts_out) {
	return mc_launch_scan_push(ctx, center_row, lo, hi, remove_marked, partials_dev, nparts_out, nullptr);
}
