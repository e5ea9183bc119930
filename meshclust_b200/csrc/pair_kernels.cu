// Stage 2: point-vs-center feature evaluation + GLM decision (K2a scan, K2b keys / pair lists).
// Replaces Trainer::get_close / filter / merge loop bodies and the DivergencePoint::distance
// calls of Trainer::split's sort comparators.  HBM-bound byte work: one pass over the histogram
// rows with 16-byte streaming loads, SIMD-in-word integer reductions, FP64 epilogue on all lanes.
#include "pair_core.cuh"
#include "tma_utils.cuh"

// ---------------------------------------------------------------------------------------------
// per-point constants
// ---------------------------------------------------------------------------------------------
template <int TB>
__global__ void point_stats_kernel(const uint8_t *__restrict__ hist, int nbins, int64_t row0, int64_t n,
                                   const uint64_t *__restrict__ lens, McRowAux *__restrict__ aux) {
	// rows [row0, row0 + n); lens[i] belongs to row row0 + i
	const int lane = threadIdx.x & 31;
	const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
	const int rb = nbins * TB;
	for (int64_t i = warp; i < n; i += nwarps) {
		const int64_t row = row0 + i;
		unsigned long long m = 0, s = 0;
		const uint8_t *p = hist + (size_t)row * rb;
		if (rb >= 16) {
			// 16-byte loads; uint8: sum and sum of squares with two dot products per word
			for (int c = lane; c < rb / 16; c += 32) {
				const uint4 v = *reinterpret_cast<const uint4 *>(p + (size_t)c * 16);
				const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
				for (int j = 0; j < 4; j++) {
					if (TB == 1) {
						m += __dp4a(w[j], 0x01010101u, 0u);
						s += __dp4a(w[j], w[j], 0u);
					} else {
						const unsigned long long a = w[j] & 0xffffu, b = w[j] >> 16;
						m += a + b;
						s += a * a + b * b;
					}
				}
			}
		} else if (TB == 1) {
			for (int b = lane; b < nbins; b += 32) { unsigned v = p[b]; m += v; s += v * v; }
		} else {
			const uint16_t *q = reinterpret_cast<const uint16_t *>(p);
			for (int b = lane; b < nbins; b += 32) { unsigned long long v = q[b]; m += v; s += v * v; }
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			m += __shfl_xor_sync(MC_FULL_MASK, m, o);
			s += __shfl_xor_sync(MC_FULL_MASK, s, o);
		}
		if (lane == 0) {
			McRowAux a;
			a.len = lens[i]; a.mag = m; a.sq = s; a.alive = 1; a.pad = 0;
			aux[row] = a;
		}
	}
}

// constants of rows [row0, row0 + n) from their histograms and lens_dev[0 .. n)
int mc_launch_point_stats_range(mc_ctx *ctx, int64_t row0, int64_t n, const uint64_t *lens_dev, cudaStream_t stream) {
	const int threads = 256;
	int64_t blocks = (n * 32 + threads - 1) / threads;
	if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
	if (blocks < 1) blocks = 1;
	if (ctx->tbytes == 1)
		point_stats_kernel<1><<<(int)blocks, threads, 0, stream>>>((const uint8_t *)ctx->d_hist, ctx->nbins, row0, n, lens_dev, ctx->d_aux);
	else
		point_stats_kernel<2><<<(int)blocks, threads, 0, stream>>>((const uint8_t *)ctx->d_hist, ctx->nbins, row0, n, lens_dev, ctx->d_aux);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

int mc_launch_point_stats(mc_ctx *ctx, const uint64_t *lens_dev) {
	return mc_launch_point_stats_range(ctx, 0, ctx->n, lens_dev, ctx->stream);
}

__global__ void alive_reset_kernel(McRowAux *__restrict__ aux, long long n) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) aux[i].alive = 1;
}

int mc_launch_alive_reset(mc_ctx *ctx) {
	int64_t blocks = (ctx->n + 255) / 256;
	if (blocks > (int64_t)ctx->num_sms * 8) blocks = (int64_t)ctx->num_sms * 8;
	alive_reset_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(ctx->d_aux, ctx->n);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// K2a: scan = Trainer::get_close + bvec::remove_available
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;

template <int TB, int RB>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const uint8_t *__restrict__ hist, McRowAux *__restrict__ aux,
            uint8_t *__restrict__ marks, long long lo, long long hi, long long center_row,
            McModel model, int remove_marked, ScanPartial *__restrict__ partials,
            unsigned int *__restrict__ ticket, ScanPartial *__restrict__ result) {
	using C = RowCfg<RB>;
	constexpr int NB = RB / TB;
	extern __shared__ __align__(16) uint32_t cen_smem[];
	__shared__ ScanPartial warp_part[SCAN_THREADS / 32];
	__shared__ bool is_last;

	const int lane = threadIdx.x & 31;
	const int wib = threadIdx.x >> 5;
	const int g = lane / C::LPP, r = lane % C::LPP;

	const uint8_t *crow = hist + (size_t)center_row * RB;
	CenterRegs<RB> cen;
	cen.load(crow, r);
	if constexpr (!C::CENTER_IN_REGS) {
		for (int i = threadIdx.x; i < RB / 16; i += blockDim.x)
			reinterpret_cast<uint4 *>(cen_smem)[i] = __ldg(reinterpret_cast<const uint4 *>(crow) + i);
		__syncthreads();
	}
	const uint64_t lq = aux[center_row].len, mq = aux[center_row].mag, sq = aux[center_row].sq;

	ScanPartial mine;
	mine.n_eval = 0; mine.n_pos = 0; mine.best_row = -1; mine.best_f0 = -1.0;

	const long long warps_total = (long long)gridDim.x * (SCAN_THREADS / 32);
	const long long warp_id = (long long)blockIdx.x * (SCAN_THREADS / 32) + wib;
	for (long long batch = lo + warp_id * 32; batch <= hi; batch += warps_total * 32) {
		const long long row_mine = batch + lane;
		const unsigned alive_mine = (row_mine <= hi) ? aux[row_mine].alive : 0u;
		const unsigned alive_bits = __ballot_sync(MC_FULL_MASK, alive_mine != 0);
		if (alive_bits == 0) {
			if (row_mine <= hi) marks[row_mine] = 0;
			continue;
		}
		PairAcc<TB> part[C::LPP];
#pragma unroll
		for (int it = 0; it < C::LPP; it++) {
			const int pidx = g * C::LPP + it;
			// dead / out-of-range rows read the (cache-hot) center row instead: loads stay
			// unconditional so they can all be issued up front
			const bool a = (alive_bits >> pidx) & 1u;
			const uint8_t *row = a ? hist + (size_t)(batch + pidx) * RB : crow;
			part[it] = mc_row_partial<TB, RB>(row, r, cen, cen_smem);
		}
		const PairAcc<TB> tot = mc_transpose_reduce<C::LPP>(part, r);
		if (row_mine <= hi) {
			unsigned flag = 0;
			if (alive_mine) {
				const uint64_t lp = aux[row_mine].len, mp = aux[row_mine].mag, sp = aux[row_mine].sq;
				const uint64_t S = tot.summin(mp, mq);
				double c[5], f[4], sum;
				mc_raw_features(S, tot.dot(), lp, mp, sp, lq, mq, sq, NB, model.nfeat >= 4, c);
				mc_eval_model(model, c, f, sum);
				flag = MC_IS_SIMILAR(sum) ? 1u : 0u;
				mc_count_near(model, sum);
				mine.n_eval++;
				mine.n_pos += flag;
				if (f[0] > mine.best_f0) { mine.best_f0 = f[0]; mine.best_row = row_mine; }
				if (flag && remove_marked) aux[row_mine].alive = 0;
			}
			marks[row_mine] = (uint8_t)flag;
		}
	}

	// warp -> block -> grid reduction, deterministic (merge rule is order independent)
	mc_scan_warp_fold(mine);
	if (lane == 0) warp_part[wib] = mine;
	__syncthreads();
	if (threadIdx.x == 0) {
		ScanPartial b = warp_part[0];
		for (int w = 1; w < SCAN_THREADS / 32; w++) mc_scan_merge(b, warp_part[w]);
		partials[blockIdx.x] = b;
		__threadfence();
		const unsigned t = atomicAdd(ticket, 1u);
		is_last = (t == gridDim.x - 1);
	}
	__syncthreads();
	if (is_last) {
		// the last block folds every block's partial (tiny: <= a few thousand entries)
		__threadfence();
		ScanPartial b;
		b.n_eval = 0; b.n_pos = 0; b.best_row = -1; b.best_f0 = -1.0;
		for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
			ScanPartial p;   // L2 loads: the partials were written by other SMs
			p.n_eval = __ldcg(&partials[i].n_eval);
			p.n_pos = __ldcg(&partials[i].n_pos);
			p.best_row = __ldcg(&partials[i].best_row);
			p.best_f0 = __ldcg(&partials[i].best_f0);
			mc_scan_merge(b, p);
		}
		mc_scan_warp_fold(b);
		if (lane == 0) warp_part[wib] = b;
		__syncthreads();
		if (threadIdx.x == 0) {
			ScanPartial t = warp_part[0];
			for (int w = 1; w < SCAN_THREADS / 32; w++) mc_scan_merge(t, warp_part[w]);
			*result = t;
			*ticket = 0;   // re-arm for the next launch on this stream
		}
	}
}

int64_t mc_scan_max_blocks(mc_ctx *ctx);

// direct-load variant (rows narrower than 16 bytes, very wide rows, MC_SCAN_DIRECT=1): folds on the
// device, so it reports a single partial in partials_dev[0]; block partials go to a context buffer
int mc_launch_scan_direct(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked,
                          void *partials_dev, int *nparts_out) {
	if (!ctx->d_scan_partials) MC_CUDA(cudaMalloc(&ctx->d_scan_partials, (size_t)mc_scan_max_blocks(ctx) * 32));
	void *result_dev = partials_dev;
	partials_dev = ctx->d_scan_partials;
	*nparts_out = 1;
	const int64_t rows = hi - lo + 1;
	int64_t blocks = (rows + SCAN_THREADS - 1) / SCAN_THREADS;
	const int64_t cap = (int64_t)ctx->num_sms * 8;
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
#define SCAN_CASE(TBv, RBv)                                                                            \
	{                                                                                                  \
		const size_t smem = RowCfg<RBv>::CENTER_IN_REGS ? 0 : (size_t)RBv;                             \
		if (smem > 48 * 1024)                                                                          \
			MC_CUDA(cudaFuncSetAttribute(scan_kernel<TBv, RBv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
		scan_kernel<TBv, RBv><<<(int)blocks, SCAN_THREADS, smem, ctx->stream>>>(                       \
			(const uint8_t *)ctx->d_hist, ctx->d_aux, ctx->d_marks, lo, hi, center_row,         \
			ctx->model, remove_marked, (ScanPartial *)partials_dev, ctx->d_ticket, (ScanPartial *)result_dev); \
	}
	MC_DISPATCH_ROW(ctx->tbytes, ctx->nbins, SCAN_CASE);
#undef SCAN_CASE
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

int64_t mc_scan_max_blocks(mc_ctx *ctx) { return (int64_t)ctx->num_sms * 8; }

// ---------------------------------------------------------------------------------------------
// K2b: distance keys of every row against C centers (Trainer::split sort keys)
// grid.y = center index; same 32-rows-per-warp tile as the scan, integer epilogue only.
// ---------------------------------------------------------------------------------------------
template <int TB, int RB>
__global__ void __launch_bounds__(SCAN_THREADS)
dist_keys_kernel(const uint8_t *__restrict__ hist, const McRowAux *__restrict__ aux, long long n,
                 const int32_t *__restrict__ center_rows, uint16_t *__restrict__ keys) {
	using C = RowCfg<RB>;
	extern __shared__ __align__(16) uint32_t cen_smem[];
	const int lane = threadIdx.x & 31;
	const int wib = threadIdx.x >> 5;
	const int g = lane / C::LPP, r = lane % C::LPP;
	const long long center_row = center_rows[blockIdx.y];
	const uint8_t *crow = hist + (size_t)center_row * RB;
	CenterRegs<RB> cen;
	cen.load(crow, r);
	if constexpr (!C::CENTER_IN_REGS) {
		for (int i = threadIdx.x; i < RB / 16; i += blockDim.x)
			reinterpret_cast<uint4 *>(cen_smem)[i] = __ldg(reinterpret_cast<const uint4 *>(crow) + i);
		__syncthreads();
	}
	const uint64_t mq = aux[center_row].mag;
	uint16_t *out = keys + (size_t)blockIdx.y * n;
	const long long warps_total = (long long)gridDim.x * (SCAN_THREADS / 32);
	const long long warp_id = (long long)blockIdx.x * (SCAN_THREADS / 32) + wib;
	for (long long batch = warp_id * 32; batch < n; batch += warps_total * 32) {
		PairAcc<TB> part[C::LPP];
#pragma unroll
		for (int it = 0; it < C::LPP; it++) {
			const long long row = batch + g * C::LPP + it;
			const uint8_t *p = row < n ? hist + (size_t)row * RB : crow;
			part[it] = mc_row_partial<TB, RB>(p, r, cen, cen_smem);
		}
		const PairAcc<TB> tot = mc_transpose_reduce<C::LPP>(part, r);
		const long long row_mine = batch + lane;
		if (row_mine < n) {
			const uint64_t mp = aux[row_mine].mag;
			out[row_mine] = (uint16_t)mc_distance_key(tot.summin(mp, mq), mp + mq);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K2b as a tile: C centers x all rows with every row read ONCE.  The centers of a group (all C of them
// when C * row bytes fit shared memory: 150 KB for 150 centers at k = 5) are staged in shared memory with
// 1-D bulk copies (cp.async.bulk, one per center); a warp keeps one row in registers (lane l owns the 16-byte
// chunks l, l + 32, ... of it: a warp-wide 16-byte load of a center from shared memory is then 512
// consecutive bytes, free of bank conflicts), runs it against 32 centers at a time -- lane-private partial sums of
// |p - q| per center -- and a transposing reduction (31 shuffles for 32 centers) leaves the total of center c
// in lane c, which computes the key.  DRAM traffic = n * 4^k (+ keys), not C times that.
// Sum min(p, q) = (mag_p + mag_q - sum |p - q|) / 2 exactly.
// ---------------------------------------------------------------------------------------------
struct AbsAcc {
	uint32_t a;
	__device__ __forceinline__ void shfl_add_from(const AbsAcc &send, int mask) { a += __shfl_xor_sync(MC_FULL_MASK, send.a, mask); }
};

template <int TB>
__device__ __forceinline__ uint32_t absdiff_acc(uint32_t p, uint32_t q, uint32_t acc) {
	if constexpr (TB == 1) {
		asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(p), "r"(q));
		return acc;
	} else {
		// two 16-bit bins per word: |a - b| = max - min
		const uint32_t mx = __vmaxu2(p, q), mn = __vminu2(p, q), d = mx - mn;   // no borrow across the halves: max >= min in each
		return acc + (d & 0xffffu) + (d >> 16);
	}
}

constexpr int TILE_THREADS = 512;

template <int TB, int RB>
__global__ void __launch_bounds__(TILE_THREADS)
dist_keys_tile_kernel(const uint8_t *__restrict__ hist, const McRowAux *__restrict__ aux, long long n,
                      const int32_t *__restrict__ center_rows, int C, int group, uint16_t *__restrict__ keys) {
	constexpr int W = RB / 128;   // 32-bit words of a row per lane
	static_assert(W >= 1, "the tile kernel wants rows of 128 bytes and more");
	extern __shared__ __align__(128) uint8_t tsm[];
	__shared__ __align__(8) uint64_t bar;
	uint8_t *cen = tsm;                                                      // group x RB
	unsigned long long *cmag = reinterpret_cast<unsigned long long *>(tsm + (size_t)group * RB);   // group
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const long long nwarps = (long long)gridDim.x * (TILE_THREADS / 32);
	const long long warp = (long long)blockIdx.x * (TILE_THREADS / 32) + wib;
	if (threadIdx.x == 0) {
		mbar_init(&bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	uint32_t phase = 0;
	for (int g0 = 0; g0 < C; g0 += group) {
		const int gn = min(group, C - g0);
		const int gpad = (gn + 31) & ~31;   // whole blocks of 32 centers: the padding repeats the group's first center
		if (threadIdx.x == 0) {
			mbar_expect_tx(&bar, (uint32_t)gpad * RB);
			for (int c = 0; c < gpad; c++) {
				const long long crow = center_rows[g0 + (c < gn ? c : 0)];
				tma_bulk_g2s(cen + (size_t)c * RB, hist + (size_t)crow * RB, RB, &bar);
			}
		}
		for (int c = threadIdx.x; c < gpad; c += TILE_THREADS) cmag[c] = aux[center_rows[g0 + (c < gn ? c : 0)]].mag;
		mbar_wait(&bar, phase);
		phase ^= 1;
		__syncthreads();
		for (long long row = warp; row < n; row += nwarps) {
			uint32_t rw[W];
			// W >= 4: chunk j of this lane = bytes (j * 32 + lane) * 16 of the row; narrower rows: W consecutive words
			const uint32_t *src = reinterpret_cast<const uint32_t *>(hist + (size_t)row * RB) + (W >= 4 ? lane * 4 : lane * W);
			if constexpr (W >= 4) {
#pragma unroll
				for (int w = 0; w < W; w += 4) {
					const uint4 v = mc_ld_stream16(src + w * 32);
					rw[w] = v.x; rw[w + 1] = v.y; rw[w + 2] = v.z; rw[w + 3] = v.w;
				}
			} else {
#pragma unroll
				for (int w = 0; w < W; w++) rw[w] = __ldg(src + w);
			}
			const unsigned long long mp = aux[row].mag;
			for (int sub = 0; sub < gpad; sub += 32) {
				AbsAcc acc[32];
#pragma unroll
				for (int c = 0; c < 32; c++) {
					const uint32_t *q = reinterpret_cast<const uint32_t *>(cen + (size_t)(sub + c) * RB) + (W >= 4 ? lane * 4 : lane * W);
					uint32_t a = 0;
					if constexpr (W >= 4) {
#pragma unroll
						for (int w = 0; w < W; w += 4) {
							const uint4 v = *reinterpret_cast<const uint4 *>(q + w * 32);
							a = absdiff_acc<TB>(rw[w], v.x, a); a = absdiff_acc<TB>(rw[w + 1], v.y, a);
							a = absdiff_acc<TB>(rw[w + 2], v.z, a); a = absdiff_acc<TB>(rw[w + 3], v.w, a);
						}
					} else {
#pragma unroll
						for (int w = 0; w < W; w++) a = absdiff_acc<TB>(rw[w], q[w], a);
					}
					acc[c].a = a;
				}
				const AbsAcc tot = mc_transpose_reduce<32>(acc, lane);   // lane c: center sub + c
				const int c = sub + lane;
				if (c < gn) {
					const unsigned long long mq = cmag[c];
					const unsigned long long S = (mp + mq - (unsigned long long)tot.a) >> 1;
					keys[(size_t)(g0 + c) * n + row] = (uint16_t)mc_distance_key(S, mp + mq);
				}
			}
		}
		__syncthreads();   // every warp is done with this group's centers before the next group lands
	}
}

template <int TB, int RB>
static int launch_keys_tile(mc_ctx *ctx, const int32_t *center_rows_dev, int C, uint16_t *keys_dev) {
	int dev_smem = 0;
	MC_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
	// centers per group: whole blocks of 32, as many as shared memory holds (row + 8 bytes each)
	int group = (int)(((size_t)dev_smem - 2048) / ((size_t)RB + 8)) & ~31;
	const int cpad = (C + 31) & ~31;
	if (group > cpad) group = cpad;
	MC_REQUIRE(group >= 32, MC_ERR_UNSUPPORTED, "rows of %d bytes: 32 centers do not fit shared memory", RB);
	const size_t smem = (size_t)group * ((size_t)RB + 8);
	MC_CUDA(cudaFuncSetAttribute(dist_keys_tile_kernel<TB, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 1;
	MC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dist_keys_tile_kernel<TB, RB>, TILE_THREADS, smem));
	if (per_sm < 1) per_sm = 1;
	int64_t blocks = (int64_t)ctx->num_sms * per_sm;
	const int64_t need = (ctx->n + TILE_THREADS / 32 - 1) / (TILE_THREADS / 32);
	if (blocks > need) blocks = need;
	if (blocks < 1) blocks = 1;
	dist_keys_tile_kernel<TB, RB><<<(unsigned)blocks, TILE_THREADS, smem, ctx->stream>>>((const uint8_t *)ctx->d_hist, ctx->d_aux, ctx->n, center_rows_dev, C, group, keys_dev);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

int mc_launch_dist_keys(mc_ctx *ctx, const int32_t *center_rows_dev, int C, uint16_t *keys_dev) {
	// several centers against rows of 128 bytes and more: the tile kernel (every row read once)
	static const bool no_tile = getenv("MC_KEYS_NO_TILE") != nullptr;
	const int rbytes = ctx->tbytes * ctx->nbins;
	if (!no_tile && C >= 4 && rbytes >= 128 && rbytes <= 4096) {
		if (ctx->tbytes == 1) {
			switch (rbytes) {
			case 256: return launch_keys_tile<1, 256>(ctx, center_rows_dev, C, keys_dev);
			case 1024: return launch_keys_tile<1, 1024>(ctx, center_rows_dev, C, keys_dev);
			case 4096: return launch_keys_tile<1, 4096>(ctx, center_rows_dev, C, keys_dev);
			default: break;
			}
		} else {
			switch (rbytes) {
			case 128: return launch_keys_tile<2, 128>(ctx, center_rows_dev, C, keys_dev);
			case 512: return launch_keys_tile<2, 512>(ctx, center_rows_dev, C, keys_dev);
			case 2048: return launch_keys_tile<2, 2048>(ctx, center_rows_dev, C, keys_dev);
			default: break;
			}
		}
	}

	int64_t blocks = (ctx->n + SCAN_THREADS - 1) / SCAN_THREADS;
	const int64_t cap = (int64_t)ctx->num_sms * 4;
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
	dim3 grid((unsigned)blocks, (unsigned)C);
#define KEYS_CASE(TBv, RBv)                                                                            \
	{                                                                                                  \
		const size_t smem = RowCfg<RBv>::CENTER_IN_REGS ? 0 : (size_t)RBv;                             \
		if (smem > 48 * 1024)                                                                          \
			MC_CUDA(cudaFuncSetAttribute(dist_keys_kernel<TBv, RBv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
		dist_keys_kernel<TBv, RBv><<<grid, SCAN_THREADS, smem, ctx->stream>>>(                          \
			(const uint8_t *)ctx->d_hist, ctx->d_aux, ctx->n, center_rows_dev, keys_dev);              \
	}
	MC_DISPATCH_ROW(ctx->tbytes, ctx->nbins, KEYS_CASE);
#undef KEYS_CASE
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// pair lists (training feature matrix, normalisation bounds, merge candidates): one warp per pair
// ---------------------------------------------------------------------------------------------
template <int TB>
__global__ void pair_list_kernel(const uint8_t *__restrict__ hist, const McRowAux *__restrict__ aux, int nbins,
                                 const int32_t *__restrict__ pa, const int32_t *__restrict__ pb,
                                 long long m, McModel model, double *__restrict__ raw5,
                                 unsigned long long *__restrict__ dist, double *__restrict__ sum_out,
                                 double *__restrict__ f0_out, uint8_t *__restrict__ flag_out,
                                 double *__restrict__ feats_out) {
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	const int rb = nbins * TB;
	for (long long i = warp; i < m; i += nwarps) {
		const long long a = pa[i], b = pb[i];
		const uint8_t *p = hist + (size_t)a * rb, *q = hist + (size_t)b * rb;
		PairAcc<TB> acc;
		if (rb >= 16) {
			for (int c = lane; c < rb / 16; c += 32) {
				const uint4 x = __ldg(reinterpret_cast<const uint4 *>(p) + c);
				const uint4 y = __ldg(reinterpret_cast<const uint4 *>(q) + c);
				acc.add(x.x, y.x); acc.add(x.y, y.y); acc.add(x.z, y.z); acc.add(x.w, y.w);
			}
		} else {
			for (int c = lane; c < rb / 4; c += 32)
				acc.add(__ldg(reinterpret_cast<const uint32_t *>(p) + c), __ldg(reinterpret_cast<const uint32_t *>(q) + c));
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) acc.shfl_add_from(acc, o);
		if (lane == 0) {
			const uint64_t lp = aux[a].len, mp = aux[a].mag, sp = aux[a].sq;
			const uint64_t lq = aux[b].len, mq = aux[b].mag, sq = aux[b].sq;
			const uint64_t S = acc.summin(mp, mq);
			double c[5];
			mc_raw_features(S, acc.dot(), lp, mp, sp, lq, mq, sq, nbins, true, c);
			if (raw5) {
#pragma unroll
				for (int j = 0; j < 5; j++) raw5[i * 5 + j] = c[j];
			}
			if (dist) dist[i] = mc_distance_key(S, mp + mq);
			if (sum_out || f0_out || flag_out || feats_out) {
				double f[4], sum;
				mc_eval_model(model, c, f, sum);
				if (sum_out) sum_out[i] = sum;
				if (f0_out) f0_out[i] = f[0];
				if (flag_out) { flag_out[i] = MC_IS_SIMILAR(sum) ? 1 : 0; mc_count_near(model, sum); }
				if (feats_out) {
#pragma unroll
					for (int j = 0; j < 4; j++) feats_out[i * 4 + j] = f[j];
				}
			}
		}
	}
}

int mc_launch_pair_list(mc_ctx *ctx, const int32_t *pa_dev, const int32_t *pb_dev, int64_t m,
                        double *raw5_dev, uint64_t *dist_dev, double *sum_dev, double *f0_dev,
                        uint8_t *flag_dev, double *feats_dev) {
	const int threads = 256;
	int64_t blocks = (m * 32 + threads - 1) / threads;
	if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
	if (blocks < 1) blocks = 1;
	const McRowAux *aux = ctx->d_aux;
	if (ctx->tbytes == 1)
		pair_list_kernel<1><<<(int)blocks, threads, 0, ctx->stream>>>((const uint8_t *)ctx->d_hist, aux, ctx->nbins, pa_dev, pb_dev, m, ctx->model, raw5_dev, (unsigned long long *)dist_dev, sum_dev, f0_dev, flag_dev, feats_dev);
	else
		pair_list_kernel<2><<<(int)blocks, threads, 0, ctx->stream>>>((const uint8_t *)ctx->d_hist, aux, ctx->nbins, pa_dev, pb_dev, m, ctx->model, raw5_dev, (unsigned long long *)dist_dev, sum_dev, f0_dev, flag_dev, feats_dev);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// mc_permute_rows: gather rows (histogram + constants) into a staging buffer, one warp per row
// ---------------------------------------------------------------------------------------------
__global__ void permute_rows_kernel(const uint8_t *__restrict__ hist, const McRowAux *__restrict__ aux,
                                    const int32_t *__restrict__ old_of_new, long long count, long long n_alive, int rb,
                                    uint8_t *__restrict__ hist_out, McRowAux *__restrict__ aux_out) {
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	for (long long i = warp; i < count; i += nwarps) {
		const long long o = old_of_new[i];
		const uint8_t *src = hist + (size_t)o * rb;
		uint8_t *dst = hist_out + (size_t)i * rb;
		if (rb >= 16) {
			for (int c = lane; c < rb / 16; c += 32) reinterpret_cast<uint4 *>(dst)[c] = reinterpret_cast<const uint4 *>(src)[c];
		} else {
			for (int c = lane; c < rb / 4; c += 32) reinterpret_cast<uint32_t *>(dst)[c] = reinterpret_cast<const uint32_t *>(src)[c];
		}
		if (lane == 0) {
			McRowAux a = aux[o];
			a.alive = i < n_alive ? 1u : 0u;
			aux_out[i] = a;
		}
	}
}

int mc_launch_permute_rows(mc_ctx *ctx, const int32_t *old_of_new_dev, int64_t count, int64_t n_alive, void *hist_out, void *aux_out) {
	const int threads = 256;
	int64_t blocks = (count * 32 + threads - 1) / threads;
	if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
	if (blocks < 1) blocks = 1;
	permute_rows_kernel<<<(int)blocks, threads, 0, ctx->stream>>>((const uint8_t *)ctx->d_hist, ctx->d_aux, old_of_new_dev, count, n_alive,
	                                                              ctx->nbins * ctx->tbytes, (uint8_t *)hist_out, (McRowAux *)aux_out);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}
