// Stage 3: mean-shift center accumulate / update (K3a center_sum, K3b nearest_to_mean).
// Replaces get_mean (ClusterFactory.cpp:382-425) and mean_shift_update (ClusterFactory.cpp:289-380)
// with Trainer::filter / Trainer::closest (Trainer.cpp:334-365) folded in.
//
// The mean never has to exist as doubles: DivergencePoint::distance_d truncates it to the bin
// type inside the min and accumulates its magnitude through a uint64 (DivergencePoint.cpp:53-65),
// so only tq_i = floor(sum_i / count) (integer division; exact) and
// magc = sum_i tq_i are needed.
#include <cooperative_groups.h>

#include "pair_core.cuh"

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
		const PairAcc<TB> acc = mc_warp_pair_reduce<TB>(hist + (size_t)row * rb, tq, rb, lane);
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
// Phase A, fused: everything accumulate() does between two get_close calls, in ONE cooperative
// launch behind the scan kernel (ClusterFactory.cpp:676-692):
//   fold the scan's per-CTA partials -> ordered compaction of the marked rows (what
//   bvec::remove_available appends to `current`, bvec.cpp:302-316) into the device-resident member
//   list and into a host-mapped list -> running bin sums += their histograms -> truncated mean ->
//   nearest member (get_mean's argmin, first minimum in `current` order).
// The grid is one CTA per SM; phases are separated by grid-wide barriers.  When nothing was marked
// (is_min: once per cluster) every CTA leaves right after the fold.
// ---------------------------------------------------------------------------------------------
struct AccDev {
	long long members_n;                     // |current|
	unsigned int counts[MC_SCAN_PARTS];      // marked rows per CTA chunk
	NearPartial near[MC_SCAN_PARTS];
};

constexpr int ACC_THREADS = 256;

// The result block lives in host-mapped pinned memory: [0,48) mc_step_result, [48,56) sequence word,
// [56,60) error word of a sharded step.  The host does not synchronise with the stream; it polls the
// sequence word, which is written last, behind a system-scope fence.
__device__ __forceinline__ void step_publish(mc_step_result *out_host, unsigned long long seq, const unsigned int *err_dev) {
	volatile unsigned int *err = reinterpret_cast<volatile unsigned int *>(reinterpret_cast<uint8_t *>(out_host) + 56);
	*err = err_dev ? __ldcg(err_dev) : 0u;
	__threadfence_system();
	*reinterpret_cast<volatile unsigned long long *>(reinterpret_cast<uint8_t *>(out_host) + 48) = seq;
}

template <int TB>
__global__ void __launch_bounds__(ACC_THREADS)
accumulate_tail_kernel(const uint8_t *__restrict__ hist, const McRowAux *__restrict__ aux,
                       const uint8_t *__restrict__ marks, int nbins, long long lo, long long hi,
                       long long center_row, int restart, const mc_scan_result *__restrict__ partials,
                       int nparts, long long *__restrict__ members, unsigned long long *__restrict__ sum,
                       AccDev *__restrict__ acc, mc_step_result *__restrict__ out_host,
                       int32_t *__restrict__ list_host, unsigned long long seq,
                       const unsigned int *__restrict__ err_dev) {
	namespace cg = cooperative_groups;
	cg::grid_group grid = cg::this_grid();
	extern __shared__ __align__(16) uint8_t tq[];   // truncated mean, nbins * TB bytes
	__shared__ mc_scan_result s_scan;
	__shared__ unsigned int s_warp[ACC_THREADS / 32];
	__shared__ unsigned long long s_red[ACC_THREADS / 32];
	__shared__ unsigned int s_base;
	__shared__ unsigned long long s_magc;
	__shared__ NearPartial s_near[ACC_THREADS / 32];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int rb = nbins * TB;
	const int G = gridDim.x, b = blockIdx.x;

	// ---- fold (every CTA, redundantly: <= 160 records out of L2)
	if (wib == 0) {
		mc_scan_result r;
		r.n_eval = 0; r.n_pos = 0; r.best_row = -1; r.best_f0 = -1.0;
		for (int i = lane; i < nparts; i += 32) {
			mc_scan_result p;
			p.n_eval = __ldcg(&partials[i].n_eval); p.n_pos = __ldcg(&partials[i].n_pos);
			p.best_row = __ldcg(&partials[i].best_row); p.best_f0 = __ldcg(&partials[i].best_f0);
			mc_scan_merge(r, p);
		}
		mc_scan_warp_fold(r);
		if (lane == 0) s_scan = r;
	}
	const long long m0 = restart ? 1 : acc->members_n;   // read before anyone may rewrite it (end of the kernel)
	__syncthreads();
	const mc_scan_result scan = s_scan;
	const long long n_new = scan.n_pos;
	if (b == 0 && restart) {
		// `current` = {center}: member 0 and the running sums start from its histogram
		if (threadIdx.x == 0) members[0] = center_row;
		for (int i = threadIdx.x; i < nbins; i += ACC_THREADS)
			sum[i] = TB == 1 ? (unsigned long long)hist[(size_t)center_row * rb + i]
			                 : (unsigned long long)reinterpret_cast<const uint16_t *>(hist + (size_t)center_row * rb)[i];
	}
	if (n_new == 0) {
		if (b == 0 && threadIdx.x == 0) {
			out_host->scan = scan;
			out_host->nearest_row = -1;
			out_host->n_members = m0;
			acc->members_n = m0;
			step_publish(out_host, seq, err_dev);
		}
		return;   // uniform across the grid: nobody reaches a barrier
	}

	// ---- ordered compaction of the marks of [lo, hi]: count, barrier, scatter
	const long long nrows = hi - lo + 1;
	long long chunk = (nrows + G - 1) / G;
	const long long per = (chunk + ACC_THREADS - 1) / ACC_THREADS;
	chunk = per * ACC_THREADS;
	const long long t0 = lo + (long long)b * chunk + (long long)threadIdx.x * per;
	long long t1 = t0 + per;
	if (t1 > hi + 1) t1 = hi + 1;
	unsigned int mine = 0;
	for (long long r = t0; r < t1; r++) mine += marks[r];
	unsigned int incl = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned int v = __shfl_up_sync(MC_FULL_MASK, incl, o);
		if (lane >= o) incl += v;
	}
	if (lane == 31) s_warp[wib] = incl;
	__syncthreads();
	unsigned int warp_base = 0, block_total = 0;
	for (int w = 0; w < ACC_THREADS / 32; w++) {
		if (w < wib) warp_base += s_warp[w];
		block_total += s_warp[w];
	}
	const unsigned int excl = warp_base + incl - mine;
	if (threadIdx.x == 0) acc->counts[b] = block_total;
	grid.sync();
	if (wib == 0) {
		unsigned int v = 0;
		for (int i = lane; i < b; i += 32) v += __ldcg(&acc->counts[i]);
#pragma unroll
		for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(MC_FULL_MASK, v, o);
		if (lane == 0) s_base = v;
	}
	__syncthreads();
	if (mine) {
		long long at = (long long)s_base + excl;
		for (long long r = t0; r < t1; r++)
			if (marks[r]) {
				members[m0 + at] = r;
				list_host[at] = (int32_t)r;
				at++;
			}
		__threadfence_system();   // the list is read by the host as soon as the sequence word appears
	}
	grid.sync();

	// ---- running sums += the new members' histograms (exact integers)
	{
		const int words = rb / 4;
		for (int w0 = 0; w0 < words; w0 += ACC_THREADS) {
			const int w = w0 + threadIdx.x;
			uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
			if (w < words) {
				for (long long i = b; i < n_new; i += G) {
					const long long row = __ldcg(&members[m0 + i]);
					const uint32_t v = *(reinterpret_cast<const uint32_t *>(hist + (size_t)row * rb) + w);
					if (TB == 1) { a0 += v & 0xff; a1 += (v >> 8) & 0xff; a2 += (v >> 16) & 0xff; a3 += v >> 24; }
					else { a0 += v & 0xffff; a1 += v >> 16; }
				}
				if (TB == 1) {
					if (a0) atomicAdd(&sum[w * 4 + 0], (unsigned long long)a0);
					if (a1) atomicAdd(&sum[w * 4 + 1], (unsigned long long)a1);
					if (a2) atomicAdd(&sum[w * 4 + 2], (unsigned long long)a2);
					if (a3) atomicAdd(&sum[w * 4 + 3], (unsigned long long)a3);
				} else {
					if (a0) atomicAdd(&sum[w * 2 + 0], (unsigned long long)a0);
					if (a1) atomicAdd(&sum[w * 2 + 1], (unsigned long long)a1);
				}
			}
		}
	}
	grid.sync();

	// ---- truncated mean (per CTA, in shared memory) and its magnitude
	const long long m_all = m0 + n_new;
	{
		unsigned long long local = 0;
		for (int i = threadIdx.x; i < nbins; i += ACC_THREADS) {
			const unsigned long long v = __ldcg(&sum[i]) / (unsigned long long)m_all;
			if (TB == 1) tq[i] = (uint8_t)v; else reinterpret_cast<uint16_t *>(tq)[i] = (uint16_t)v;
			local += v;
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(MC_FULL_MASK, local, o);
		if (lane == 0) s_red[wib] = local;
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned long long t = 0;
			for (int w = 0; w < ACC_THREADS / 32; w++) t += s_red[w];
			s_magc = t;
		}
		__syncthreads();
	}
	const unsigned long long magc = s_magc;

	// ---- nearest member: first minimum of distance_d in `current` order
	NearPartial best; best.pos = -1; best.dist = 0;
	for (long long i = (long long)b * (ACC_THREADS / 32) + wib; i < m_all; i += (long long)G * (ACC_THREADS / 32)) {
		const long long row = __ldcg(&members[i]);
		const PairAcc<TB> pa = mc_warp_pair_reduce<TB>(hist + (size_t)row * rb, tq, rb, lane);
		const uint64_t mp = aux[row].mag;
		NearPartial cnd; cnd.pos = i; cnd.dist = mc_distance_d(pa.summin(mp, magc), mp, magc);
		near_merge(best, cnd);
	}
	if (lane == 0) s_near[wib] = best;
	__syncthreads();
	if (threadIdx.x == 0) {
		NearPartial r = s_near[0];
		for (int w = 1; w < ACC_THREADS / 32; w++) near_merge(r, s_near[w]);
		acc->near[b] = r;
	}
	grid.sync();
	if (b == 0 && wib == 0) {
		NearPartial r; r.pos = -1; r.dist = 0;
		for (int i = lane; i < G; i += 32) {
			NearPartial p;
			p.pos = __ldcg(&acc->near[i].pos);
			p.dist = __ldcg(&acc->near[i].dist);
			near_merge(r, p);
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			NearPartial other;
			other.pos = __shfl_xor_sync(MC_FULL_MASK, r.pos, o);
			other.dist = __shfl_xor_sync(MC_FULL_MASK, r.dist, o);
			near_merge(r, other);
		}
		if (lane == 0) {
			out_host->scan = scan;
			out_host->nearest_row = r.pos >= 0 ? __ldcg(&members[r.pos]) : -1;
			out_host->n_members = m_all;
			acc->members_n = m_all;
			step_publish(out_host, seq, err_dev);
		}
	}
}

size_t mc_acc_dev_bytes() { return sizeof(AccDev); }

int mc_launch_accumulate_tail(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int restart,
                              const void *partials_dev, int nparts, void *acc_dev, void *out_host_dev,
                              int32_t *list_host_dev, unsigned long long seq, const unsigned int *err_dev) {
	const int nbins = ctx->nbins;
	const size_t smem = (size_t)nbins * ctx->tbytes;
	MC_REQUIRE(smem <= 96 * 1024, MC_ERR_UNSUPPORTED, "k too large for the accumulate kernel");
	const uint8_t *hist = (const uint8_t *)ctx->d_hist;
	const McRowAux *aux = ctx->d_aux;
	const uint8_t *marks = ctx->d_marks;
	long long lo_ = lo, hi_ = hi, cr = center_row;
	long long *members = (long long *)ctx->d_members;
	unsigned long long *sum = (unsigned long long *)ctx->d_sum;
	int nb = nbins, rs = restart, np = nparts;
	void *args[] = {&hist, &aux, &marks, &nb, &lo_, &hi_, &cr, &rs, &partials_dev, &np, &members, &sum, &acc_dev, &out_host_dev, &list_host_dev, &seq, &err_dev};
	int grid = ctx->num_sms < MC_SCAN_PARTS ? ctx->num_sms : MC_SCAN_PARTS;
	const void *fn = ctx->tbytes == 1 ? (const void *)accumulate_tail_kernel<1> : (const void *)accumulate_tail_kernel<2>;
	if (smem > 48 * 1024) MC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	MC_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(ACC_THREADS), args, smem, ctx->stream));
	ctx->launches++;
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
		const PairAcc<TB> acc = mc_warp_pair_reduce<TB>(hist + (size_t)row * rb, hist + (size_t)crow * rb, rb, lane);
		if (lane == 0) {
			const uint64_t lp = aux[row].len, mp = aux[row].mag, sp = aux[row].sq;
			double cc[5], f[4], sum;
			mc_raw_features(acc.summin(mp, mq), acc.dot(), lp, mp, sp, lq, mq, sq, nbins, model.nfeat >= 4, cc);
			mc_eval_model(model, cc, f, sum);
			const bool keep = MC_IS_SIMILAR(sum);
			mc_count_near(model, sum);
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
		const PairAcc<TB> acc = mc_warp_pair_reduce<TB>(hist + (size_t)row * rb, tq, rb, lane);
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
