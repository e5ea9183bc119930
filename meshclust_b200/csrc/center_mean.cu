// Stage 3: mean-shift center accumulate / update (K3a center_sum, K3b nearest_to_mean).
// Replaces get_mean (ClusterFactory.cpp:382-425) and mean_shift_update (ClusterFactory.cpp:289-380)
// with Trainer::filter / Trainer::closest (Trainer.cpp:334-365) folded in.
//
// The mean never has to exist as doubles: DivergencePoint::distance_d truncates it to the bin
// type inside the min and accumulates its magnitude through a uint64 (DivergencePoint.cpp:53-65),
// so only tq_i = floor(sum_i / count) (integer division; exact) and
// magc = sum_i tq_i are needed.
#include "pair_core.cuh"

// generic-address pair reduction of one row against a "center" row (global or shared memory)
template <int TB>
__device__ __forceinline__ PairAcc<TB> warp_pair_reduce(const uint8_t *p, const uint8_t *q, int rb, int lane) {
	PairAcc<TB> acc;
	if (rb >= 16) {
		for (int c = lane; c < rb / 16; c += 32) {
			const uint4 x = *(reinterpret_cast<const uint4 *>(p) + c);
			const uint4 y = *(reinterpret_cast<const uint4 *>(q) + c);
			acc.add(x.x, y.x); acc.add(x.y, y.y); acc.add(x.z, y.z); acc.add(x.w, y.w);
		}
	} else {
		for (int c = lane; c < rb / 4; c += 32)
			acc.add(*(reinterpret_cast<const uint32_t *>(p) + c), *(reinterpret_cast<const uint32_t *>(q) + c));
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) acc.shfl_add_from(acc, o);
	return acc;
}

// ---------------------------------------------------------------------------------------------
// Phase A: running sum of member histograms
// ---------------------------------------------------------------------------------------------
template <int TB>
__global__ void sum_rows_kernel(const uint8_t *__restrict__ hist, int nbins, const int64_t *__restrict__ rows,
                                long long m, unsigned long long *__restrict__ sum) {
	// blockIdx.y strides over rows, threads over bins
	for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nbins; b += gridDim.x * blockDim.x) {
		unsigned long long acc = 0;
		for (long long i = blockIdx.y; i < m; i += gridDim.y) {
			const size_t off = (size_t)rows[i] * nbins + b;
			acc += TB == 1 ? (unsigned long long)hist[off] : (unsigned long long)reinterpret_cast<const uint16_t *>(hist)[off];
		}
		if (acc) atomicAdd(&sum[b], acc);
	}
}

// tq row (bin type) = floor(sum / count), magc = sum tq
template <int TB>
__global__ void trunc_mean_kernel(const unsigned long long *__restrict__ sum, int nbins, long long count,
                                  uint8_t *__restrict__ tq, unsigned long long *__restrict__ magc) {
	__shared__ unsigned long long red[32];
	unsigned long long local = 0;
	for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
		const unsigned long long v = sum[b] / (unsigned long long)count;
		if (TB == 1) tq[b] = (uint8_t)v; else reinterpret_cast<uint16_t *>(tq)[b] = (uint16_t)v;
		local += v;
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(MC_FULL_MASK, local, o);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long t = 0;
		for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
		*magc = t;
	}
}

struct NearPartial {
	long long pos;   // position in the member list (first minimum wins)
	double dist;
};

__device__ __forceinline__ void near_merge(NearPartial &a, const NearPartial &b) {
	if (b.pos >= 0 && (a.pos < 0 || b.dist < a.dist || (b.dist == a.dist && b.pos < a.pos))) a = b;
}

template <int TB>
__global__ void nearest_kernel(const uint8_t *__restrict__ hist, const McRowAux *__restrict__ aux, int nbins,
                               const int64_t *__restrict__ rows, long long m, const uint8_t *__restrict__ tq,
                               const unsigned long long *__restrict__ magc_p, NearPartial *__restrict__ partials,
                               unsigned int *__restrict__ ticket, long long *__restrict__ out_row,
                               double *__restrict__ out_dist) {
	__shared__ NearPartial wp[8];
	__shared__ bool is_last;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int rb = nbins * TB;
	const unsigned long long magc = *magc_p;
	NearPartial best; best.pos = -1; best.dist = 0;
	const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
	for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + wib; i < m; i += nwarps) {
		const long long row = rows[i];
		const PairAcc<TB> acc = warp_pair_reduce<TB>(hist + (size_t)row * rb, tq, rb, lane);
		const uint64_t mp = aux[row].mag;
		const double d = mc_distance_d(acc.summin(mp, magc), mp, magc);
		NearPartial c; c.pos = i; c.dist = d;
		near_merge(best, c);   // NaN never replaces (comparisons false), like the reference's `<`
	}
	if (lane == 0) wp[wib] = best;
	__syncthreads();
	if (threadIdx.x == 0) {
		NearPartial b = wp[0];
		for (int w = 1; w < (int)(blockDim.x >> 5); w++) near_merge(b, wp[w]);
		partials[blockIdx.x] = b;
		__threadfence();
		is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (is_last && threadIdx.x == 0) {
		__threadfence();
		NearPartial b; b.pos = -1; b.dist = 0;
		for (int i = 0; i < (int)gridDim.x; i++) {
			NearPartial p;   // L2 loads: the partials were written by other SMs
			p.pos = __ldcg(&partials[i].pos);
			p.dist = __ldcg(&partials[i].dist);
			near_merge(b, p);
		}
		*out_row = b.pos >= 0 ? rows[b.pos] : -1;
		*out_dist = b.dist;
		*ticket = 0;
	}
}

int mc_launch_mean_nearest(mc_ctx *ctx, const int64_t *new_rows_dev, int64_t m_new, unsigned long long *sum_dev,
                           const int64_t *members_dev, int64_t m_all, uint8_t *tq_dev, unsigned long long *magc_dev,
                           void *partials_dev, long long *out_row_dev, double *out_dist_dev) {
	const int nbins = ctx->nbins;
	if (m_new > 0) {
		const int threads = nbins >= 256 ? 256 : (nbins < 32 ? 32 : nbins);
		dim3 grid((nbins + threads - 1) / threads, (unsigned)(m_new < 64 ? m_new : 64));
		if (ctx->tbytes == 1)
			sum_rows_kernel<1><<<grid, threads, 0, ctx->stream>>>((const uint8_t *)ctx->d_hist, nbins, new_rows_dev, m_new, sum_dev);
		else
			sum_rows_kernel<2><<<grid, threads, 0, ctx->stream>>>((const uint8_t *)ctx->d_hist, nbins, new_rows_dev, m_new, sum_dev);
		ctx->launches++;
	}
	if (ctx->tbytes == 1) trunc_mean_kernel<1><<<1, 256, 0, ctx->stream>>>(sum_dev, nbins, m_all, tq_dev, magc_dev);
	else trunc_mean_kernel<2><<<1, 256, 0, ctx->stream>>>(sum_dev, nbins, m_all, tq_dev, magc_dev);
	ctx->launches++;
	int64_t blocks = (m_all + 7) / 8;
	if (blocks > (int64_t)ctx->num_sms * 4) blocks = (int64_t)ctx->num_sms * 4;
	if (blocks < 1) blocks = 1;
	if (ctx->tbytes == 1)
		nearest_kernel<1><<<(int)blocks, 256, 0, ctx->stream>>>((const uint8_t *)ctx->d_hist, ctx->d_aux, nbins, members_dev, m_all, tq_dev, magc_dev, (NearPartial *)partials_dev, ctx->d_ticket + 1, out_row_dev, out_dist_dev);
	else
		nearest_kernel<2><<<(int)blocks, 256, 0, ctx->stream>>>((const uint8_t *)ctx->d_hist, ctx->d_aux, nbins, members_dev, m_all, tq_dev, magc_dev, (NearPartial *)partials_dev, ctx->d_ticket + 1, out_row_dev, out_dist_dev);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// Phase B: one CTA per center -- filter, mean of survivors, nearest survivor
// ---------------------------------------------------------------------------------------------
constexpr int UPD_THREADS = 256;

template <int TB>
__global__ void __launch_bounds__(UPD_THREADS)
update_centers_kernel(const uint8_t *__restrict__ hist, const McRowAux *__restrict__ aux, int nbins,
                      const int64_t *__restrict__ center_rows, const int64_t *__restrict__ cand_rows,
                      const int64_t *__restrict__ cand_begin, const int64_t *__restrict__ cand_end,
                      const int64_t *__restrict__ flag_off, uint8_t *__restrict__ flags, McModel model,
                      long long *__restrict__ next_rows) {
	extern __shared__ __align__(16) uint8_t smem_raw[];
	uint32_t *sums = reinterpret_cast<uint32_t *>(smem_raw);            // nbins
	uint8_t *tq = smem_raw + (size_t)nbins * 4;                          // nbins * TB (16-byte aligned)
	__shared__ unsigned long long s_magc;
	__shared__ unsigned int s_count;
	__shared__ NearPartial wp[UPD_THREADS / 32];

	const int c = blockIdx.x;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int rb = nbins * TB;
	const long long crow = center_rows[c];
	const long long cb = cand_begin[c], ncand = cand_end[c] - cb;
	uint8_t *fl = flags + flag_off[c];
	for (int b = threadIdx.x; b < nbins; b += blockDim.x) sums[b] = 0;
	if (threadIdx.x == 0) { s_count = 0; s_magc = 0; }
	__syncthreads();

	// pass 1: Trainer::filter -- keep candidates classified similar to the center
	const uint64_t lq = aux[crow].len, mq = aux[crow].mag, sq = aux[crow].sq;
	unsigned int kept = 0;
	for (long long i = wib; i < ncand; i += UPD_THREADS / 32) {
		const long long row = cand_rows[cb + i];
		const PairAcc<TB> acc = warp_pair_reduce<TB>(hist + (size_t)row * rb, hist + (size_t)crow * rb, rb, lane);
		if (lane == 0) {
			const uint64_t lp = aux[row].len, mp = aux[row].mag, sp = aux[row].sq;
			double cc[5], f[4], sum;
			mc_raw_features(acc.summin(mp, mq), acc.dot(), lp, mp, sp, lq, mq, sq, nbins, model.nfeat >= 4, cc);
			mc_eval_model(model, cc, f, sum);
			const bool keep = sum >= MC_SIGMOID_SUM_THRESHOLD;
			fl[i] = keep ? 1 : 0;
			kept += keep;
		}
	}
	if (lane == 0 && kept) atomicAdd(&s_count, kept);
	__syncthreads();
	const unsigned int count = s_count;
	if (count == 0) {
		if (threadIdx.x == 0) next_rows[c] = -1;
		return;
	}

	// pass 2: integer sum of the survivors' histograms (ClusterFactory.cpp:316-335)
	{
		const int words = rb / 4 > 0 ? rb / 4 : 1;
		const int lanes_per_row = words < UPD_THREADS ? words : UPD_THREADS;
		const int rows_par = UPD_THREADS / lanes_per_row;
		const int sub = threadIdx.x / lanes_per_row, t = threadIdx.x % lanes_per_row;
		if (sub < rows_par) {
			for (int w = t; w < words; w += lanes_per_row) {
				uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
				for (long long i = sub; i < ncand; i += rows_par) {
					if (!fl[i]) continue;
					const uint32_t v = *(reinterpret_cast<const uint32_t *>(hist + (size_t)cand_rows[cb + i] * rb) + w);
					if (TB == 1) { a0 += v & 0xff; a1 += (v >> 8) & 0xff; a2 += (v >> 16) & 0xff; a3 += v >> 24; }
					else { a0 += v & 0xffff; a1 += v >> 16; }
				}
				if (TB == 1) {
					atomicAdd(&sums[w * 4 + 0], a0); atomicAdd(&sums[w * 4 + 1], a1);
					atomicAdd(&sums[w * 4 + 2], a2); atomicAdd(&sums[w * 4 + 3], a3);
				} else {
					atomicAdd(&sums[w * 2 + 0], a0); atomicAdd(&sums[w * 2 + 1], a1);
				}
			}
		}
	}
	__syncthreads();
	{
		unsigned long long local = 0;
		for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
			const uint32_t v = sums[b] / count;
			if (TB == 1) tq[b] = (uint8_t)v; else reinterpret_cast<uint16_t *>(tq)[b] = (uint16_t)v;
			local += v;
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(MC_FULL_MASK, local, o);
		if (lane == 0 && local) atomicAdd(&s_magc, local);
	}
	__syncthreads();
	const unsigned long long magc = s_magc;

	// pass 3: Trainer::closest -- first survivor with the smallest distance_d to the mean
	NearPartial best; best.pos = -1; best.dist = 0;
	for (long long i = wib; i < ncand; i += UPD_THREADS / 32) {
		if (!fl[i]) continue;
		const long long row = cand_rows[cb + i];
		const PairAcc<TB> acc = warp_pair_reduce<TB>(hist + (size_t)row * rb, tq, rb, lane);
		const uint64_t mp = aux[row].mag;
		NearPartial cnd; cnd.pos = i; cnd.dist = mc_distance_d(acc.summin(mp, magc), mp, magc);
		near_merge(best, cnd);
	}
	if (lane == 0) wp[wib] = best;
	__syncthreads();
	if (threadIdx.x == 0) {
		NearPartial b = wp[0];
		for (int w = 1; w < UPD_THREADS / 32; w++) near_merge(b, wp[w]);
		next_rows[c] = b.pos >= 0 ? cand_rows[cb + b.pos] : -1;
	}
}

int mc_launch_update_centers(mc_ctx *ctx, const int64_t *center_rows_dev, int64_t ncenters,
                             const int64_t *cand_rows_dev, const int64_t *cand_begin_dev,
                             const int64_t *cand_end_dev, const int64_t *flag_off_dev, uint8_t *flags_dev,
                             long long *next_rows_dev) {
	const size_t smem = (size_t)ctx->nbins * 4 + (size_t)ctx->nbins * ctx->tbytes + 16;
	MC_REQUIRE(smem <= 200 * 1024, MC_ERR_UNSUPPORTED, "k too large for the update kernel");
	const McRowAux *aux = ctx->d_aux;
	if (ctx->tbytes == 1) {
		if (smem > 48 * 1024) MC_CUDA(cudaFuncSetAttribute(update_centers_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		update_centers_kernel<1><<<(unsigned)ncenters, UPD_THREADS, smem, ctx->stream>>>((const uint8_t *)ctx->d_hist, aux, ctx->nbins, center_rows_dev, cand_rows_dev, cand_begin_dev, cand_end_dev, flag_off_dev, flags_dev, ctx->model, next_rows_dev);
	} else {
		if (smem > 48 * 1024) MC_CUDA(cudaFuncSetAttribute(update_centers_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		update_centers_kernel<2><<<(unsigned)ncenters, UPD_THREADS, smem, ctx->stream>>>((const uint8_t *)ctx->d_hist, aux, ctx->nbins, center_rows_dev, cand_rows_dev, cand_begin_dev, cand_end_dev, flag_off_dev, flags_dev, ctx->model, next_rows_dev);
	}
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}
