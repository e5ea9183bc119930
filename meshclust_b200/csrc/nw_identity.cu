// Stage 4: GlobAlignE global alignment with path length / match bookkeeping (K4).
// Replaces utility/GlobAlignE.cpp:123-305 as called by Trainer::align (Trainer.cpp:15-31) and
// Feature::align (Feature.cpp:222-243), parameters (match 1, mismatch -1, open 2, continue 1).
//
// Integer, compute-bound.  One warp per pair, anti-diagonal wavefront: the rows j of seq2 are cut
// into strips of 32, lane l owns row j0+1+l and at step t fills column i = t-l+1 of seq1.  The
// row above arrives by __shfl_up from lane l-1 (which filled the same column one step earlier);
// lane 31's row is parked in a global scratch line for lane 0 of the next strip.
// Each DP state carries (score, len, id); len and id travel packed as (len << 16) | id so a path
// copy is one move and "+1 column [+1 match]" one add.  Requires la + lb <= 65535.
//
// Tie-breaking is the reference's: U and L prefer "open from M" on ties (:178-193,:258-273),
// M prefers M, then L, then U (:201-241), the final pick prefers M, L, U (:278-291), and the
// "minus infinity" is the finite, data-dependent value of :125-135 that takes part in sums.
#include "mc_common.cuh"

struct NwCell {
	int m, u, l;          // scores
	uint32_t pm, pu, pl;  // packed (len << 16) | id
};

constexpr int NW_OPEN_EXT = 3;   // gapOpen + gapContinue
constexpr int NW_EXT = 1;        // gapContinue
constexpr int NW_OPEN = 2;
constexpr uint32_t NW_LEN1 = 0x10000u;

__global__ void __launch_bounds__(128)
nw_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
          const int32_t *__restrict__ pa, const int32_t *__restrict__ pb, long long npairs,
          int4 *__restrict__ scratch_a, int2 *__restrict__ scratch_b, long long scratch_stride,
          int32_t *__restrict__ score_out, int32_t *__restrict__ len_out, int32_t *__restrict__ id_out,
          unsigned int *__restrict__ flags) {
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	// two scratch lines per warp (strip parity) so a strip never overwrites what it still reads
	int4 *sa0 = scratch_a + warp * 2 * scratch_stride;   // per column: m, u, l, pm
	int2 *sb0 = scratch_b + warp * 2 * scratch_stride;   //             pu, pl

	for (long long pr = warp; pr < npairs; pr += nwarps) {
		const long long ia = pa[pr], ib = pb[pr];
		const uint8_t *A = seq + seq_off[ia];
		const uint8_t *B = seq + seq_off[ib];
		const int la = (int)(seq_off[ia + 1] - seq_off[ia]);
		const int lb = (int)(seq_off[ib + 1] - seq_off[ib]);
		const int shorter = min(la, lb), diff = abs(la - lb);
		const int ninf = (diff >= 1 ? -NW_OPEN - diff * NW_EXT : 0) - shorter - 1;

		if (la + lb > 65535) {   // packed len/id would overflow; reported to the host
			if (lane == 0) { atomicOr(&flags[2], 1u); score_out[pr] = 0; len_out[pr] = 0; id_out[pr] = 0; }
			continue;
		}
		if (la == 0 || lb == 0) {
			// no DP cell: the answer is the boundary itself (see the init row / column-0 reset)
			if (lane == 0) {
				int sc, ln;
				if (la == 0 && lb == 0) { sc = 0; ln = 0; }
				else if (lb == 0) { sc = -NW_OPEN - la * NW_EXT; ln = la; }   // L[la] of the init row
				else { sc = ninf; ln = lb; }                                   // M[0] after lb rows
				score_out[pr] = sc; len_out[pr] = ln; id_out[pr] = 0;
			}
			continue;
		}

		int res_sc = 0; uint32_t res_p = 0;
		const int nstrips = (lb + 31) / 32;
		for (int strip = 0; strip < nstrips; strip++) {
			const int j = strip * 32 + 1 + lane;          // this lane's row (1-based)
			const int4 *sa_in = sa0 + (strip & 1) * scratch_stride;
			const int2 *sb_in = sb0 + (strip & 1) * scratch_stride;
			int4 *sa_out = sa0 + ((strip + 1) & 1) * scratch_stride;
			int2 *sb_out = sb0 + ((strip + 1) & 1) * scratch_stride;
			const int bj = (j <= lb) ? B[j - 1] : 0xfe;   // 0xfe never equals a base
			// state of (row j, column i-1): starts at column 0 = {M,L} = ninf with len j
			int m_left = ninf, l_left = ninf;
			uint32_t pm_left = (uint32_t)j << 16, pl_left = (uint32_t)j << 16;
			// diagonal (row j-1, column i-1): starts at column 0 of the row above
			int m_d = (j == 1) ? 0 : ninf, l_d = ninf, u_d = -NW_OPEN - (j - 1) * NW_EXT;
			uint32_t pm_d = (uint32_t)(j - 1) << 16, pl_d = pm_d, pu_d = pm_d;
			// what this lane publishes to the lane below: its last cell
			NwCell mine; mine.m = 0; mine.u = 0; mine.l = 0; mine.pm = 0; mine.pu = 0; mine.pl = 0;
			uint32_t achunk = 0, achar = 0;

			const int nsteps = la + 31;
			for (int t = 0; t < nsteps; t++) {
				if ((t & 31) == 0) achunk = (t + lane < la) ? A[t + lane] : 0xff;
				// base of column t+1 enters at lane 0 and moves one lane down per step
				const uint32_t a_in = __shfl_sync(MC_FULL_MASK, achunk, t & 31);
				const uint32_t a_dn = __shfl_up_sync(MC_FULL_MASK, achar, 1);
				achar = lane == 0 ? a_in : a_dn;

				NwCell up;
				up.m = __shfl_up_sync(MC_FULL_MASK, mine.m, 1);
				up.u = __shfl_up_sync(MC_FULL_MASK, mine.u, 1);
				up.l = __shfl_up_sync(MC_FULL_MASK, mine.l, 1);
				up.pm = __shfl_up_sync(MC_FULL_MASK, mine.pm, 1);
				up.pu = __shfl_up_sync(MC_FULL_MASK, mine.pu, 1);
				up.pl = __shfl_up_sync(MC_FULL_MASK, mine.pl, 1);
				const int i = t - lane + 1;               // column (1-based)
				if (lane == 0 && i <= la) {
					if (strip == 0) {
						// init row (GlobAlignE.cpp:137-160)
						up.m = ninf; up.u = ninf; up.l = -NW_OPEN - i * NW_EXT;
						up.pm = up.pu = up.pl = (uint32_t)i << 16;
					} else {
						const int4 x = __ldcg(&sa_in[i]);
						const int2 y = __ldcg(&sb_in[i]);
						up.m = x.x; up.u = x.y; up.l = x.z; up.pm = (uint32_t)x.w;
						up.pu = (uint32_t)y.x; up.pl = (uint32_t)y.y;
					}
				}
				if (i >= 1 && i <= la) {
					// vertical gap
					const int ub = up.m - NW_OPEN_EXT, uc = up.u - NW_EXT;
					const bool ubeg = ub >= uc;
					const int u = ubeg ? ub : uc;
					const uint32_t pu = (ubeg ? up.pm : up.pu) + NW_LEN1;
					// diagonal
					const bool eq = achar == (uint32_t)bj;
					const int s = eq ? 1 : -1;
					int best = m_d; uint32_t pb_ = pm_d;
					if (l_d > best) { best = l_d; pb_ = pl_d; }
					if (u_d > best) { best = u_d; pb_ = pu_d; }
					const int m = best + s;
					const uint32_t pm = pb_ + NW_LEN1 + (eq ? 1u : 0u);
					// horizontal gap on the current row
					const int lb_ = m_left - NW_OPEN_EXT, lc = l_left - NW_EXT;
					const bool lbeg = lb_ >= lc;
					const int l = lbeg ? lb_ : lc;
					const uint32_t pl = (lbeg ? pm_left : pl_left) + NW_LEN1;
					// roll
					m_d = up.m; l_d = up.l; u_d = up.u; pm_d = up.pm; pl_d = up.pl; pu_d = up.pu;
					m_left = m; l_left = l; pm_left = pm; pl_left = pl;
					mine.m = m; mine.u = u; mine.l = l; mine.pm = pm; mine.pu = pu; mine.pl = pl;
					if (lane == 31 && strip + 1 < nstrips) {
						__stcg(&sa_out[i], make_int4(m, u, l, (int)pm));
						__stcg(&sb_out[i], make_int2((int)pu, (int)pl));
					}
					if (j == lb && i == la) {
						// GlobAlignE.cpp:278-291: tie order M, L, U
						int bs = m; uint32_t bp = pm;
						if (l > bs) { bs = l; bp = pl; }
						if (u > bs) { bs = u; bp = pu; }
						res_sc = bs; res_p = bp;
					}
				}
			}
			__syncwarp();
		}
		// the lane that owned row lb holds the result
		const int owner = (lb - 1) & 31;
		res_sc = __shfl_sync(MC_FULL_MASK, res_sc, owner);
		res_p = __shfl_sync(MC_FULL_MASK, res_p, owner);
		if (lane == 0) {
			score_out[pr] = res_sc;
			len_out[pr] = (int32_t)(res_p >> 16);
			id_out[pr] = (int32_t)(res_p & 0xffffu);
		}
	}
}

// scratch: two lines of scratch_stride >= max_len + 1 columns per resident warp
int mc_launch_nw(mc_ctx *ctx, const int32_t *pa_dev, const int32_t *pb_dev, int64_t m, int64_t max_len,
                 int32_t *score_dev, int32_t *len_dev, int32_t *id_dev, void *scratch_a, void *scratch_b,
                 int64_t scratch_stride, int64_t nwarps) {
	const int threads = 128;
	const int64_t blocks = (nwarps * 32 + threads - 1) / threads;
	(void)max_len;
	nw_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(ctx->d_seq, ctx->d_seq_off, pa_dev, pb_dev, m, (int4 *)scratch_a, (int2 *)scratch_b, scratch_stride, score_dev, len_dev, id_dev, ctx->d_flags);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}
