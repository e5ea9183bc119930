// Stage 4: GlobAlignE global alignment with path length / match bookkeeping (K4).
// Replaces utility/GlobAlignE.cpp:123-305 as called by Trainer::align (Trainer.cpp:15-31) and
// Feature::align (Feature.cpp:222-243), parameters (match 1, mismatch -1, open 2, continue 1).
//
// Integer, compute-bound.  One warp per pair, anti-diagonal wavefront with register blocking:
// the rows j of seq2 are cut into strips of 32*R rows, lane l owns R consecutive rows and at step t
// fills column i = t-l+1 of seq1 for all of them (top row first: the vertical dependency stays
// inside the lane).  Only the lane's LAST row travels to the lane below (__shfl_up, 6 words per
// step for 32*R cells); lane 31's last row is parked in a global scratch line for lane 0 of the
// next strip, which reads it back in coalesced 32-column chunks fetched one chunk ahead.
// Each DP state carries (score, diag, id): every transition adds one column to the alignment, so its
// length is i + j - (diagonal moves) and only the diagonal moves need counting; diag and id travel packed
// as (diag << 16) | id: a gap move copies the word, a diagonal move adds (1 << 16) + match.  Both
// counts are at most min(la, lb): requires min(la, lb) <= 65535.
//
// Tie-breaking is the reference's: U and L prefer "open from M" on ties (:178-193,:258-273),
// M prefers M, then L, then U (:201-241), the final pick prefers M, L, U (:278-291), and the
// "minus infinity" is the finite, data-dependent value of :125-135 that takes part in sums.
#include "mc_common.cuh"

struct NwCell {
	int m, u, l;          // scores
	uint32_t pm, pu, pl;  // packed (diagonal moves << 16) | id
};

constexpr int NW_OPEN_EXT = 3;   // gapOpen + gapContinue
constexpr int NW_EXT = 1;        // gapContinue
constexpr int NW_OPEN = 2;
constexpr uint32_t NW_LEN1 = 0x10000u;
// rows per lane: a template parameter.  One wavefront step costs ~125 instructions of bookkeeping (14 shuffles,
// chunk hand-over, boundary store) next to ~26 per cell, so 4 rows per lane spend half of the issue slots
// outside the recurrence; 16 rows per lane (strips of 512 rows, ~160 registers) leave 8 per cell.  Short
// sequences keep the narrow strips (rows past the end of seq2 are wasted work): the launcher picks per batch.

__device__ __forceinline__ NwCell nw_shfl_up(const NwCell &c) {
	NwCell r;
	r.m = __shfl_up_sync(MC_FULL_MASK, c.m, 1);
	r.u = __shfl_up_sync(MC_FULL_MASK, c.u, 1);
	r.l = __shfl_up_sync(MC_FULL_MASK, c.l, 1);
	r.pm = __shfl_up_sync(MC_FULL_MASK, c.pm, 1);
	r.pu = __shfl_up_sync(MC_FULL_MASK, c.pu, 1);
	r.pl = __shfl_up_sync(MC_FULL_MASK, c.pl, 1);
	return r;
}

__device__ __forceinline__ NwCell nw_shfl_from(const NwCell &c, int src) {
	NwCell r;
	r.m = __shfl_sync(MC_FULL_MASK, c.m, src);
	r.u = __shfl_sync(MC_FULL_MASK, c.u, src);
	r.l = __shfl_sync(MC_FULL_MASK, c.l, src);
	r.pm = __shfl_sync(MC_FULL_MASK, c.pm, src);
	r.pu = __shfl_sync(MC_FULL_MASK, c.pu, src);
	r.pl = __shfl_sync(MC_FULL_MASK, c.pl, src);
	return r;
}

// W = warps that share a pair (a "team" = one CTA of 128 threads when W = 4).  One warp per pair (W = 1) fills the GPU
// only with thousands of pairs; the alignment-driven binary search of Trainer::split offers a few hundred long pairs
// per round, and a round then lasts as long as ONE warp needs for ONE pair (0.125 s for 10 kb x 10 kb).  With a team
// the strips of a pair form a pipeline: strip s is taken by warp (running strip number) mod W and may read column
// chunk c of the boundary row as soon as strip s - 1 has stored it -- the boundary lines already live in global
// memory, one line per strip in flight (W + 1 lines per team, line = running strip number mod (W + 1): the line a
// strip writes was last read W strips ago, by the same warp), and what was missing is only flow control: a word per
// line in shared memory, {running strip number, columns stored}, published by the writer every 32 columns behind a
// fence and polled by the reader before it fetches a chunk.  Running strip numbers continue across the pairs of a
// team, so warps drift into the next pair without a barrier.  The arithmetic is untouched.
template <int R, int W>
__global__ void __launch_bounds__(W == 1 ? 128 : 32 * W, W == 6 ? 2 : (R >= 16 ? 3 : 1))
nw_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
          const int32_t *__restrict__ pa, const int32_t *__restrict__ pb, long long npairs,
          int4 *__restrict__ scratch_a, int2 *__restrict__ scratch_b, long long scratch_stride,
          int32_t *__restrict__ score_out, int32_t *__restrict__ len_out, int32_t *__restrict__ id_out,
          unsigned int *__restrict__ flags) {
	constexpr int LINES = W == 1 ? 2 : W + 1;
	__shared__ unsigned long long s_prog[W == 1 ? 1 : LINES];   // per boundary line: (running strip number + 1) << 24 | columns stored
	const int lane = threadIdx.x & 31;
	const int wteam = W == 1 ? 0 : (int)(threadIdx.x >> 5);      // this warp's place in its team
	const long long warp = W == 1 ? (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) : (long long)blockIdx.x;   // pair owner: warp or team
	const long long nwarps = W == 1 ? (((long long)gridDim.x * blockDim.x) >> 5) : (long long)gridDim.x;
	// boundary lines of this warp / team (W = 1: two, by strip parity, so a strip never overwrites what it still reads)
	int4 *sa0 = scratch_a + warp * LINES * scratch_stride;   // per column: m, u, l, pm
	int2 *sb0 = scratch_b + warp * LINES * scratch_stride;   //             pu, pl
	unsigned long long gbase = 0;   // strips of the pairs this team has been through (identical in all its warps)
	if (W > 1) {
		if (threadIdx.x < LINES) s_prog[threadIdx.x] = 0;
		__syncthreads();
	}

	for (long long pr = warp; pr < npairs; pr += nwarps) {
		const long long ia = pa[pr], ib = pb[pr];
		const uint8_t *A = seq + seq_off[ia];
		const uint8_t *B = seq + seq_off[ib];
		const int la = (int)(seq_off[ia + 1] - seq_off[ia]);
		const int lb = (int)(seq_off[ib + 1] - seq_off[ib]);
		const int shorter = min(la, lb), diff = abs(la - lb);
		const int ninf = (diff >= 1 ? -NW_OPEN - diff * NW_EXT : 0) - shorter - 1;

		if (shorter > 65535) {   // the packed counters would overflow; reported to the host
			if (lane == 0 && wteam == 0) { atomicOr(&flags[2], 1u); score_out[pr] = 0; len_out[pr] = 0; id_out[pr] = 0; }
			continue;
		}
		if (la == 0 || lb == 0) {
			// no DP cell: the answer is the boundary itself (see the init row / column-0 reset)
			if (lane == 0 && wteam == 0) {
				int sc, ln;
				if (la == 0 && lb == 0) { sc = 0; ln = 0; }
				else if (lb == 0) { sc = -NW_OPEN - la * NW_EXT; ln = la; }   // L[la] of the init row
				else { sc = ninf; ln = lb; }                                   // M[0] after lb rows
				score_out[pr] = sc; len_out[pr] = ln; id_out[pr] = 0;
			}
			continue;
		}

		int res_sc = 0; uint32_t res_p = 0;
		const int rows_per_strip = 32 * R;
		const int nstrips = (lb + rows_per_strip - 1) / rows_per_strip;
		bool mine_last = W == 1;   // this warp took the pair's last strip (it holds the result)
		for (int strip = 0; strip < nstrips; strip++) {
			const unsigned long long g = gbase + (unsigned long long)strip;   // running strip number of the team
			if (W > 1 && (int)(g % W) != wteam) continue;
			if (W > 1) mine_last = strip == nstrips - 1;
			const int j0 = strip * rows_per_strip + lane * R + 1;   // first row of this lane (1-based)
			const int line_in = W == 1 ? (strip & 1) : (int)(g % LINES), line_out = W == 1 ? ((strip + 1) & 1) : (int)((g + 1) % LINES);
			const int4 *sa_in = sa0 + line_in * scratch_stride;
			const int2 *sb_in = sb0 + line_in * scratch_stride;
			int4 *sa_out = sa0 + line_out * scratch_stride;
			int2 *sb_out = sb0 + line_out * scratch_stride;
			// flow control of the team: columns 1..need of line_in have been stored by the strip before this one
			auto wait_cols = [&](int need) {
				if (W == 1) return;
				volatile unsigned long long *p = &s_prog[line_in];
				for (;;) {
					const unsigned long long v = *p;
					if ((v >> 24) == g && (long long)(v & 0xffffffull) >= (long long)need) break;   // (tag g = strip g - 1, stored as number + 1)
				}
				__threadfence();
			};

			int bj[R];
			// per row: the cell just to the left (row j, column i-1).  The diagonal of row r+1 is that
			// same cell before this step overwrites it, so only the top row of the lane needs a separate
			// diagonal (row j0-1, column i-1); the U fields of `left`, which the horizontal recurrence
			// never reads, carry the synthesised U of column 0 for that purpose.
			NwCell left[R];    // (row r, column i-1)
			NwCell diag0;      // (row j0-1, column i-1): what the lane's first row adds its match score to
#pragma unroll
			for (int r = 0; r < R; r++) {
				const int j = j0 + r;
				bj[r] = (j <= lb) ? B[j - 1] : 0xfe;   // 0xfe never equals a base
				// column 0 of row j: M = L = ninf with len j; U[0] as the row below synthesises it
				// (GlobAlignE.cpp:164-170,250-256)
				left[r].m = ninf; left[r].l = ninf; left[r].u = -NW_OPEN - j * NW_EXT;
				left[r].pm = left[r].pl = left[r].pu = 0;   // j vertical moves, no diagonal
			}
			// column 0 of row j0-1
			diag0.m = (j0 == 1) ? 0 : ninf;
			diag0.l = ninf;
			diag0.u = -NW_OPEN - (j0 - 1) * NW_EXT;
			diag0.pm = diag0.pl = diag0.pu = 0;
			NwCell mine;   // last row's newest cell, published to the lane below
			mine.m = 0; mine.u = 0; mine.l = 0; mine.pm = 0; mine.pu = 0; mine.pl = 0;
			uint32_t achunk = 0, achar = 0;
			NwCell bchunk, bnext;   // boundary row of the previous strip, 32 columns per lane-chunk
			bchunk.m = bchunk.u = bchunk.l = 0; bchunk.pm = bchunk.pu = bchunk.pl = 0;
			bnext = bchunk;
			if (strip > 0) {
				// prefetch columns 1..32 (chunk 0); chunk c holds columns 32c+1 .. 32c+32
				wait_cols(la < 32 ? la : 32);
				const int col = 1 + lane;
				if (col <= la) {
					const int4 x = __ldcg(&sa_in[col]);
					const int2 y = __ldcg(&sb_in[col]);
					bnext.m = x.x; bnext.u = x.y; bnext.l = x.z; bnext.pm = (uint32_t)x.w;
					bnext.pu = (uint32_t)y.x; bnext.pl = (uint32_t)y.y;
				}
			}

			const int nsteps = la + 31;
			for (int t = 0; t < nsteps; t++) {
				if ((t & 31) == 0) {
					achunk = (t + lane < la) ? A[t + lane] : 0xff;
					if (strip > 0) {
						bchunk = bnext;
						if (t + 33 <= la) wait_cols(la < t + 64 ? la : t + 64);
						const int col = t + 32 + 1 + lane;   // next chunk, one chunk ahead
						if (col <= la) {
							const int4 x = __ldcg(&sa_in[col]);
							const int2 y = __ldcg(&sb_in[col]);
							bnext.m = x.x; bnext.u = x.y; bnext.l = x.z; bnext.pm = (uint32_t)x.w;
							bnext.pu = (uint32_t)y.x; bnext.pl = (uint32_t)y.y;
						}
					}
				}
				// base of column t+1 enters at lane 0 and moves one lane down per step
				const uint32_t a_in = __shfl_sync(MC_FULL_MASK, achunk, t & 31);
				const uint32_t a_dn = __shfl_up_sync(MC_FULL_MASK, achar, 1);
				achar = lane == 0 ? a_in : a_dn;

				// the row above this lane's first row, same column
				NwCell up = nw_shfl_up(mine);
				const int i = t - lane + 1;               // column (1-based)
				if (strip > 0) {
					const NwCell b = nw_shfl_from(bchunk, t & 31);   // lane 0's column is t+1
					if (lane == 0) up = b;
				} else if (lane == 0) {
					// init row (GlobAlignE.cpp:137-160)
					up.m = ninf; up.u = ninf; up.l = -NW_OPEN - i * NW_EXT;
					up.pm = up.pu = up.pl = 0;
				}
				if (i >= 1 && i <= la) {
					NwCell dg = diag0;   // diagonal of the row being filled
					diag0 = up;          // (row j0-1, column i) is the top row's diagonal at the next column
#pragma unroll
					for (int r = 0; r < R; r++) {
						// vertical gap: from (row-1, i)
						// (DPX: VIMNMX with the predicate "first operand won" / three-input maximum)
						NwCell cur;
						bool ubeg;
						cur.u = __vibmax_s32(up.m - NW_OPEN_EXT, up.u - NW_EXT, &ubeg);   // ties open from M
						cur.pu = ubeg ? up.pm : up.pu;
						// diagonal: from (row-1, i-1), tie order M, L, U
						const int best = __vimax3_s32(dg.m, dg.l, dg.u);
						const uint32_t pbst = dg.m == best ? dg.pm : (dg.l == best ? dg.pl : dg.pu);
						cur.m = best - 1;
						cur.pm = pbst + NW_LEN1;
						if (achar == (uint32_t)bj[r]) { cur.m += 2; cur.pm += 1u; }   // two predicated adds
						// horizontal gap on the current row: from (row, i-1)
						bool hbeg;
						cur.l = __vibmax_s32(left[r].m - NW_OPEN_EXT, left[r].l - NW_EXT, &hbeg);
						cur.pl = hbeg ? left[r].pm : left[r].pl;
						// roll: (row, i-1) is the diagonal of the row below; this cell the left of (row, i+1)
						dg = left[r];
						left[r] = cur;
						up = cur;   // the next row of this lane sits right below
					}
					if (i == la) {
#pragma unroll
						for (int r = 0; r < R; r++)
							if (j0 + r == lb) {
								// GlobAlignE.cpp:278-291: tie order M, L, U
								int bs = left[r].m; uint32_t bp = left[r].pm;
								if (left[r].l > bs) { bs = left[r].l; bp = left[r].pl; }
								if (left[r].u > bs) { bs = left[r].u; bp = left[r].pu; }
								res_sc = bs; res_p = bp;
							}
					}
					mine = up;      // == last row's cell
					if (lane == 31 && strip + 1 < nstrips) {
						__stcg(&sa_out[i], make_int4(mine.m, mine.u, mine.l, (int)mine.pm));
						__stcg(&sb_out[i], make_int2((int)mine.pu, (int)mine.pl));
						if (W > 1 && ((i & 31) == 0 || i == la)) {
							__threadfence();
							*(volatile unsigned long long *)&s_prog[line_out] = ((g + 1) << 24) | (unsigned long long)i;
						}
					}
				}
			}
			__syncwarp();
		}
		gbase += (unsigned long long)nstrips;
		// the lane that owned row lb holds the result (in the warp that took the last strip)
		const int owner = ((lb - 1) % (32 * R)) / R;
		res_sc = __shfl_sync(MC_FULL_MASK, res_sc, owner);
		res_p = __shfl_sync(MC_FULL_MASK, res_p, owner);
		if (lane == 0 && mine_last) {
			score_out[pr] = res_sc;
			len_out[pr] = la + lb - (int32_t)(res_p >> 16);
			id_out[pr] = (int32_t)(res_p & 0xffffu);
		}
	}
}

// scratch: boundary lines of scratch_stride >= max_len + 1 columns: two per resident warp (team_warps = 1) or
// team_warps + 1 per resident team.  nwarps = resident warps / teams.
// rows_per_lane: 4, 8 or 16 (strips of 128 / 256 / 512 rows of seq2); team_warps: 1, or 4 / 6 (16 rows per lane only:
// three CTAs of 128 threads or two of 192 per SM)
int mc_launch_nw(mc_ctx *ctx, const int32_t *pa_dev, const int32_t *pb_dev, int64_t m, int64_t max_len,
                 int32_t *score_dev, int32_t *len_dev, int32_t *id_dev, void *scratch_a, void *scratch_b,
                 int64_t scratch_stride, int64_t nwarps, int rows_per_lane, int team_warps) {
	const int threads = team_warps > 1 ? 32 * team_warps : 128;
	const int64_t blocks = team_warps > 1 ? nwarps : (nwarps * 32 + threads - 1) / threads;
	(void)max_len;
#define NW_LAUNCH(RR, WW) nw_kernel<RR, WW><<<(unsigned)blocks, threads, 0, ctx->stream>>>(ctx->d_seq, ctx->d_seq_off, pa_dev, pb_dev, m, (int4 *)scratch_a, (int2 *)scratch_b, scratch_stride, score_dev, len_dev, id_dev, ctx->d_flags)
	if (team_warps > 1) {
		MC_REQUIRE((team_warps == 4 || team_warps == 6) && rows_per_lane == 16, MC_ERR_ARG, "mc_launch_nw: teams are four or six warps of 16 rows per lane");
		if (team_warps == 6) NW_LAUNCH(16, 6);
		else NW_LAUNCH(16, 4);
	} else {
		switch (rows_per_lane) {
		case 16: NW_LAUNCH(16, 1); break;
		case 8: NW_LAUNCH(8, 1); break;
		default: NW_LAUNCH(4, 1); break;
		}
	}
#undef NW_LAUNCH
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

// the strip height that wastes the fewest issue slots on a batch: cost per real cell ~ (rows a strip pads the
// pair to) x (recurrence + per-step bookkeeping / rows per lane)
int mc_nw_pick_rows(const int64_t *lb, int64_t m) {
	int best = 4;
	double best_cost = 0;
	for (int r : {4, 8, 16}) {
		double padded = 0, real = 0;
		const int64_t strip = 32 * r;
		for (int64_t i = 0; i < m; i++) {
			padded += (double)((lb[i] + strip - 1) / strip * strip);
			real += (double)lb[i];
		}
		const double cost = (real > 0 ? padded / real : 1.0) * (26.0 + 125.0 / r);
		if (r == 4 || cost < best_cost) { best = r; best_cost = cost; }
	}
	return best;
}
