// Stage 0/1: letters -> digit strings -> dense 4^k k-mer count vectors (K1).
// Replaces ChromosomeOneDigit::encodeNucleotides (ChromosomeOneDigit.cpp:95-144) and
// fill_table / KmerHashTable::wholesaleIncrement (ClusterFactory.h:40-55,
// KmerHashTable.cpp:133-223) with the pseudo-count of ClusterFactory.cpp:995.
//
// One warp per sequence.  Bases are read with aligned 16-byte loads (the sequence start is
// aligned down, leading bytes masked), each lane rolls the base-4 index over its 16 k-mer starts
// and increments a per-warp shared-memory table with ATOMS; the table is then written once,
// narrowed to the output width, together with mag = sum, sum of squares and the running maximum.
#include "mc_common.cuh"

// code LUT: 0..3 digit, 4 = 'N', 0xff = invalid   (ChromosomeOneDigit.cpp:59-85)
__device__ __constant__ uint8_t c_code_lut[256];

static uint8_t h_code_lut[256];
static bool h_lut_ready = false;

static void build_lut() {
	for (int i = 0; i < 256; i++) h_code_lut[i] = 0xff;
	auto set = [](const char *s, uint8_t v) {
		for (; *s; s++) {
			h_code_lut[(unsigned char)*s] = v;
			h_code_lut[(unsigned char)(*s | 0x20)] = v;   // toUpperCase, Chromosome.cpp:153-157
		}
	};
	set("AMV", 0);
	set("CYH", 1);
	set("GRSX", 2);
	set("TKWBD", 3);
	set("N", 4);
	h_lut_ready = true;
}

int mc_upload_lut() {
	if (!h_lut_ready) build_lut();
	MC_CUDA(cudaMemcpyToSymbol(c_code_lut, h_code_lut, 256));
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// encode: one warp per sequence, in place.  Outside segments N stays the byte 'N' and every
// other letter becomes its digit; inside segments N becomes C (1).  With no segment at all the
// string is only upper-cased (encodeNucleotides touches nothing when segNum == 0).
// flags[0] is set when an invalid letter is met (the reference throws InvalidInputException).
// ---------------------------------------------------------------------------------------------
__global__ void encode_kernel(uint8_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                              const int32_t *__restrict__ segs, const int64_t *__restrict__ seg_off,
                              long long n, unsigned int *__restrict__ flags) {
	__shared__ uint8_t lut[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = c_code_lut[i];
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	for (long long s = warp; s < n; s += nwarps) {
		const long long b0 = seq_off[s], b1 = seq_off[s + 1];
		const long long g0 = seg_off[s], g1 = seg_off[s + 1];
		const bool has_seg = g1 > g0;
		bool bad = false;
		const long long a0 = b0 & ~15LL;
		for (long long base = a0 + lane * 16; base < b1; base += 32 * 16) {
			uint4 v = *reinterpret_cast<const uint4 *>(seq + base);
			uint32_t w[4] = {v.x, v.y, v.z, v.w};
			bool touched = false;
#pragma unroll
			for (int j = 0; j < 16; j++) {
				const long long pos = base + j;
				if (pos < b0 || pos >= b1) continue;
				const uint32_t ch = (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
				uint32_t out;
				if (!has_seg) {
					out = (ch >= 'a' && ch <= 'z') ? ch - 32 : ch;
				} else {
					const uint32_t code = lut[ch];
					if (code == 4) {
						// N: digit 1 inside a segment, byte 'N' outside
						const int rel = (int)(pos - b0);
						bool inside = false;
						for (long long gi = g0; gi < g1; gi++) {
							if (rel >= segs[2 * gi] && rel <= segs[2 * gi + 1]) { inside = true; break; }
						}
						out = inside ? 1u : (uint32_t)'N';
					} else if (code == 0xffu) {
						bad = true;
						out = ch;
					} else {
						out = code;
					}
				}
				w[j >> 2] = (w[j >> 2] & ~(0xffu << ((j & 3) * 8))) | (out << ((j & 3) * 8));
				touched = true;
			}
			if (touched) {
				// neighbouring sequences share the first / last 16-byte word: write bytes there
				if (base >= b0 && base + 16 <= b1) {
					*reinterpret_cast<uint4 *>(seq + base) = make_uint4(w[0], w[1], w[2], w[3]);
				} else {
#pragma unroll
					for (int j = 0; j < 16; j++) {
						const long long pos = base + j;
						if (pos >= b0 && pos < b1) seq[pos] = (uint8_t)((w[j >> 2] >> ((j & 3) * 8)) & 0xffu);
					}
				}
			}
		}
		if (bad) atomicOr(&flags[0], 1u);
	}
}

int mc_launch_encode(mc_ctx *ctx) {
	const int threads = 256;
	int64_t blocks = (ctx->n * 32 + threads - 1) / threads;
	if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
	if (blocks < 1) blocks = 1;
	encode_kernel<<<(int)blocks, threads, 0, ctx->stream>>>(ctx->d_seq, ctx->d_seq_off, ctx->d_segs, ctx->d_seg_off, ctx->n, ctx->d_flags);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// histogram: WPB warps per block, each with a private 4^k-entry uint32 table in shared memory
// ---------------------------------------------------------------------------------------------
template <int TB>
__global__ void kmer_hist_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                                 const int32_t *__restrict__ segs, const int64_t *__restrict__ seg_off,
                                 long long n, int k, uint8_t *__restrict__ hist,
                                 McRowAux *__restrict__ aux_out, unsigned int *__restrict__ flags) {
	extern __shared__ uint32_t tables[];
	const int nbins = 1 << (2 * k);
	const uint32_t mask = (uint32_t)nbins - 1u;
	const int lane = threadIdx.x & 31;
	const int wib = threadIdx.x >> 5;
	uint32_t *tab = tables + (size_t)wib * nbins;
	const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
	const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
	unsigned int local_max = 0;
	for (long long s = warp; s < n; s += nwarps) {
		for (int i = lane; i < nbins; i += 32) tab[i] = 0;
		__syncwarp();
		const long long b0 = seq_off[s], b1 = seq_off[s + 1];
		const long long g0 = seg_off[s], g1 = seg_off[s + 1];
		for (long long gi = g0; gi < g1; gi++) {
			// k-mer starts [ss, last] inside this segment (absolute byte positions)
			const long long ss = b0 + segs[2 * gi];
			const long long last = b0 + (long long)segs[2 * gi + 1] - k + 1;
			const long long a0 = ss & ~15LL;
			for (long long base = a0 + lane * 16; base <= last; base += 32 * 16) {
				// 32-byte window: own 16 bytes + the following 16 (k-1 <= 15 look-ahead)
				const uint4 v0 = *reinterpret_cast<const uint4 *>(seq + base);
				uint4 v1 = make_uint4(0, 0, 0, 0);
				if (base + 16 < b1) v1 = *reinterpret_cast<const uint4 *>(seq + base + 16);
				const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
				uint32_t idx = 0;
#pragma unroll
				for (int u = 0; u < 31; u++) {
					uint32_t d = (w[u >> 2] >> ((u & 3) * 8)) & 0xffu;
					d = (d == (uint32_t)'N') ? 1u : d;   // N swallowed by a merged segment counts as C
					idx = ((idx << 2) | (d & 3u)) & mask;
					const int t = u - (k - 1);           // start (within the window) of the k-mer ending at u
					if (t >= 0 && t < 16) {
						const long long pos = base + t;
						if (pos >= ss && pos <= last) atomicAdd(&tab[idx], 1u);
					}
				}
			}
		}
		__syncwarp();
		// write the row: count + 1 (pseudo-count), narrowed to TB bytes
		unsigned long long m = 0, q = 0;
		if (TB == 1) {
			uint8_t *row = hist + (size_t)s * nbins;
			if (nbins >= 128) {
				for (int i = lane * 4; i < nbins; i += 128) {
					const uint32_t c0 = tab[i] + 1, c1 = tab[i + 1] + 1, c2 = tab[i + 2] + 1, c3 = tab[i + 3] + 1;
					local_max = max(max(local_max, c0), max(c1, max(c2, c3)));
					m += c0 + c1 + c2 + c3;
					q += (unsigned long long)c0 * c0 + (unsigned long long)c1 * c1 + (unsigned long long)c2 * c2 + (unsigned long long)c3 * c3;
					*reinterpret_cast<uint32_t *>(row + i) = (c0 & 0xff) | ((c1 & 0xff) << 8) | ((c2 & 0xff) << 16) | ((c3 & 0xff) << 24);
				}
			} else {
				for (int i = lane; i < nbins; i += 32) {
					const uint32_t c = tab[i] + 1;
					local_max = max(local_max, c);
					m += c; q += (unsigned long long)c * c;
					row[i] = (uint8_t)c;
				}
			}
		} else {
			uint16_t *row = reinterpret_cast<uint16_t *>(hist) + (size_t)s * nbins;
			for (int i = lane; i < nbins; i += 32) {
				const uint32_t c = tab[i] + 1;
				local_max = max(local_max, c);
				m += c; q += (unsigned long long)c * c;
				row[i] = (uint16_t)c;
			}
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			m += __shfl_xor_sync(MC_FULL_MASK, m, o);
			q += __shfl_xor_sync(MC_FULL_MASK, q, o);
		}
		if (lane == 0) {
			McRowAux a;
			a.len = (uint64_t)(b1 - b0);   // ClusterFactory.cpp:1007 set_length(base.length())
			a.mag = m;
			a.sq = q;
			a.alive = 1;
			a.pad = 0;
			aux_out[s] = a;
		}
		__syncwarp();
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) local_max = max(local_max, __shfl_xor_sync(MC_FULL_MASK, local_max, o));
	if (lane == 0 && local_max) atomicMax(&flags[1], local_max);
}

int mc_launch_kmer_hist(mc_ctx *ctx, int k, int tbytes) {
	const int nbins = 1 << (2 * k);
	const size_t per_warp = (size_t)nbins * 4;
	int wpb = 8;
	while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
	MC_REQUIRE(per_warp * wpb <= 200 * 1024, MC_ERR_UNSUPPORTED, "k=%d needs a %zu-byte table per sequence; k <= 7 is supported", k, per_warp);
	const size_t smem = per_warp * wpb;
	const int threads = wpb * 32;
	int64_t blocks = (ctx->n + wpb - 1) / wpb;
	const int64_t cap = (int64_t)ctx->num_sms * 16;
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
	if (tbytes == 1) {
		if (smem > 48 * 1024) MC_CUDA(cudaFuncSetAttribute(kmer_hist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		kmer_hist_kernel<1><<<(int)blocks, threads, smem, ctx->stream>>>(ctx->d_seq, ctx->d_seq_off, ctx->d_segs, ctx->d_seg_off, ctx->n, k, (uint8_t *)ctx->d_hist, ctx->d_aux, ctx->d_flags);
	} else {
		if (smem > 48 * 1024) MC_CUDA(cudaFuncSetAttribute(kmer_hist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		kmer_hist_kernel<2><<<(int)blocks, threads, smem, ctx->stream>>>(ctx->d_seq, ctx->d_seq_off, ctx->d_segs, ctx->d_seg_off, ctx->n, k, (uint8_t *)ctx->d_hist, ctx->d_aux, ctx->d_flags);
	}
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}
