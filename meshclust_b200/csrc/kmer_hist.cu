// Stage 0/1: letters -> digit strings -> dense 4^k k-mer count vectors (K1).
// Replaces ChromosomeOneDigit::encodeNucleotides (ChromosomeOneDigit.cpp:95-144) and
// fill_table / KmerHashTable::wholesaleIncrement (ClusterFactory.h:40-55,
// KmerHashTable.cpp:133-223) with the pseudo-count of ClusterFactory.cpp:995.
//
// One warp per sequence, one pass over the LETTERS (kmer_count_kernel below); the table is written once,
// narrowed to the output width, together with mag = sum, sum of squares and the running maximum.
// The in-place encode to digit strings (what the aligner reads) runs lazily, mc_ensure_digits.
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
// K1, one pass: letters -> 2-bit codes in registers -> k-mer counts.  One warp per sequence.
// A lane takes 16 letters (one aligned 16-byte load) and packs them into one 32-bit word, first base
// in the two most significant bits; with the word of the next 16 letters (the lane above, or lane 0 of
// the next round for lane 31) the k-mer that starts at letter j is the top 2k bits of the 64-bit
// window shifted left by 2j: a funnel shift and a shift per k-mer, then one shared-memory reduction
// into the warp's table.  Plain ACGT / acgt words are coded with three logic operations per four
// letters and checked with one byte permute against the letters the codes stand for; any other word
// (IUPAC codes, N inside a merged segment -> C, digit strings of an already encoded buffer) goes
// through the 256-entry table.  The digit strings the aligner wants are NOT written here: the in-place
// encode pass runs when the first alignment (or mc_copy_digits) asks for them.
// Bytes: L letters read + 4^k * sizeof(T) written per sequence.
// ---------------------------------------------------------------------------------------------
// codes of four letters by arithmetic; fast = all four are A, C, G or T in either case
__device__ __forceinline__ uint32_t k1_codes4_check(uint32_t w, bool &fast) {
	uint32_t x = (w >> 1) & 0x03030303u;
	x ^= (x >> 1) & 0x01010101u;
	const uint32_t s = x | (x >> 4);
	const uint32_t sel = (s & 0xffu) | ((s >> 8) & 0xff00u);
	fast = __byte_perm(0x54474341u, 0u, sel) == (w & 0xdfdfdfdfu);
	return x;
}

// the 2-bit codes of four letters (one per byte, first letter in byte 0) by arithmetic, valid for A, C, G, T in either
// case: A 0x41 C 0x43 G 0x47 T 0x54, bits 2..1 are 00 01 11 10 -> code = x ^ (x >> 1).  `diff` collects, over the
// words of a chunk, the bits in which a word differs from the letters its codes stand for (picked from "ACGT" by one
// byte permute, selector = code per nibble): zero <=> every letter was plain.
__device__ __forceinline__ uint32_t k1_codes4_acc(uint32_t w, uint32_t &diff) {
	uint32_t x = (w >> 1) & 0x03030303u;
	x ^= (x >> 1) & 0x01010101u;
	const uint32_t s = x | (x >> 4);
	const uint32_t sel = (s & 0x00ffu) | ((s >> 8) & 0xff00u);
	diff |= __byte_perm(0x54474341u, 0u, sel) ^ (w & 0xdfdfdfdfu);
	return x;
}

// the same through the 256-entry table (IUPAC codes, N inside a merged segment -> C, digit strings of an encoded buffer)
__device__ __forceinline__ uint32_t k1_codes4_lut(uint32_t w, const uint8_t *lut) {
	return (uint32_t)(lut[w & 0xffu] & 3u) | ((uint32_t)(lut[(w >> 8) & 0xffu] & 3u) << 8) |
	       ((uint32_t)(lut[(w >> 16) & 0xffu] & 3u) << 16) | ((uint32_t)(lut[w >> 24] & 3u) << 24);
}

// sixteen letters -> one 32-bit word of 2-bit codes, first letter in the two most significant bits.  Four codes
// (bits 1..0 of the bytes of x) are gathered into the top byte of x * 0x40100401: the multiplier places code i at
// bit 30 - 2 i, and every partial product is a 2-bit field of its own (no carries); three byte permutes collect the
// four top bytes.
__device__ __forceinline__ uint32_t k1_pack16(const uint4 v, const uint8_t *lut) {
	uint32_t diff = 0;
	uint32_t a = k1_codes4_acc(v.x, diff), b = k1_codes4_acc(v.y, diff), c = k1_codes4_acc(v.z, diff), d = k1_codes4_acc(v.w, diff);
	if (diff) {   // rare: some letter of the chunk is not a plain A, C, G or T
		a = k1_codes4_lut(v.x, lut); b = k1_codes4_lut(v.y, lut); c = k1_codes4_lut(v.z, lut); d = k1_codes4_lut(v.w, lut);
	}
	constexpr uint32_t M = 0x40100401u;
	const uint32_t lo = __byte_perm(d * M, c * M, 0x0073u);   // {top byte of d, top byte of c}
	const uint32_t hi = __byte_perm(b * M, a * M, 0x0073u);
	return __byte_perm(lo, hi, 0x5410u);
}

// K = k as a compile-time constant: every shift of the inner loop is an immediate
template <int TB, int K>
__global__ void __launch_bounds__(256)
kmer_count_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                  const int32_t *__restrict__ segs, const int64_t *__restrict__ seg_off,
                  long long n, uint8_t *__restrict__ hist,
                  McRowAux *__restrict__ aux_out, unsigned int *__restrict__ flags) {
	extern __shared__ uint32_t tables_raw[];
	__shared__ uint8_t lut[256];
	// N swallowed by a merged segment counts as C (ChromosomeOneDigit.cpp:59-85); digits stay what they are
	for (int i = threadIdx.x; i < 256; i += blockDim.x) {
		const uint8_t c = c_code_lut[i];
		lut[i] = i < 4 ? (uint8_t)i : (c == 4 ? 1 : c);
	}
	constexpr int nbins = 1 << (2 * K);
	constexpr uint32_t TABLE_BYTES = (uint32_t)nbins * 4u;
	constexpr uint32_t IDX_MASK = (uint32_t)(nbins - 1) << 2;
	const int lane = threadIdx.x & 31;
	const int wib = threadIdx.x >> 5;
	// the warp's table starts at a multiple of its size in the shared window: address = table | (k-mer << 2), one LOP3
	const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(tables_raw);
	const uint32_t base_s = (raw_s + TABLE_BYTES - 1u) & ~(TABLE_BYTES - 1u);
	uint32_t *tab = tables_raw + (base_s - raw_s) / 4u + (size_t)wib * nbins;
	const uint32_t tab_s = base_s + (uint32_t)wib * TABLE_BYTES;
	for (int i = lane; i < nbins; i += 32) tab[i] = 0;   // (every row write-out leaves the table zeroed again)
	__syncthreads();
	const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
	const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
	unsigned int local_max = 0;
	for (long long s = warp; s < n; s += nwarps) {
		const long long b0 = seq_off[s], b1 = seq_off[s + 1];
		const long long g0 = seg_off[s], g1 = seg_off[s + 1];
		const uint8_t *sbase = seq + (b0 & ~15LL);         // 16-byte aligned; positions below are relative to it
		const int off0 = (int)(b0 & 15LL);
		const int len = (int)(b1 - b0);
		for (long long gi = g0; gi < g1; gi++) {
			// k-mer starts [ss, last] inside this segment
			const int ss = off0 + segs[2 * gi];
			const int last = off0 + segs[2 * gi + 1] - K + 1;
			if (last < ss) continue;
			const int c_first = ss >> 4, c_last = last >> 4;   // chunks (16 letters) with a k-mer start
			const int end = off0 + len;                        // letters of this sequence end here
			auto chunk_word = [&](int c) -> uint32_t {
				if (16 * c >= end) return 0u;                  // (the buffer has a 64-byte tail: a chunk that starts inside is whole)
				return k1_pack16(*reinterpret_cast<const uint4 *>(sbase + 16 * c), lut);
			};
			// Every chunk of [c_first, c_last] counts ALL sixteen of its starts, without a predicate; the starts of
			// the first chunk that lie before the segment and those of the last chunk behind `last` are taken back
			// afterwards (one lane per start): the same words give the same k-mers, +1 - 1 cancels exactly.
			uint32_t hp = 0, hn = 0, tp = 0, tn = 0;
			uint32_t pn = chunk_word(c_first + lane);
			for (int c0 = c_first; c0 <= c_last; c0 += 32) {
				const uint32_t pc = pn;
				pn = chunk_word(c0 + 32 + lane);   // next round's word, also lane 31's neighbour
				uint32_t nb = __shfl_down_sync(MC_FULL_MASK, pc, 1);
				const uint32_t n0 = __shfl_sync(MC_FULL_MASK, pn, 0);
				if (lane == 31) nb = n0;
				if (c0 == c_first) { hp = __shfl_sync(MC_FULL_MASK, pc, 0); hn = __shfl_sync(MC_FULL_MASK, nb, 0); }
				if (c_last - c0 < 32) { tp = __shfl_sync(MC_FULL_MASK, pc, c_last - c0); tn = __shfl_sync(MC_FULL_MASK, nb, c_last - c0); }
				if (c0 + lane <= c_last) {
#pragma unroll
					for (int j = 0; j < 16; j++) {
						// (k-mer << 2) = bits of the 64-bit window pc:nb shifted right by 62 - 2j - 2K
						const int sh = 62 - 2 * j - 2 * K;
						const uint32_t t = sh >= 32 ? (pc >> (sh - 32)) : __funnelshift_r(nb, pc, sh);
						const uint32_t addr = (t & IDX_MASK) | tab_s;
						asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
					}
				}
			}
			{
				const int h = ss - 16 * c_first, tl = last - 16 * c_last;   // starts j < h of the first, j > tl of the last chunk
				const int j = lane & 15;
				const bool head = lane < 16;
				const uint32_t p = head ? hp : tp, q = head ? hn : tn;
				const uint32_t kmer = __funnelshift_l(q, p, 2 * j) >> (32 - 2 * K);
				if (head ? j < h : j > tl) {
					const uint32_t addr = (kmer << 2) | tab_s;
					asm volatile("red.shared.add.u32 [%0], 0xffffffff;" ::"r"(addr) : "memory");
				}
			}
		}
		__syncwarp();
		// write the row: count + 1 (pseudo-count), narrowed to TB bytes; the table is left zeroed
		unsigned long long m = 0, q = 0;
		if (TB == 1) {
			uint8_t *row = hist + (size_t)s * nbins;
			if (nbins >= 128) {
				for (int i = lane * 4; i < nbins; i += 128) {
					const uint4 t = *reinterpret_cast<const uint4 *>(tab + i);
					*reinterpret_cast<uint4 *>(tab + i) = make_uint4(0u, 0u, 0u, 0u);
					const uint32_t c0 = t.x + 1, c1 = t.y + 1, c2 = t.z + 1, c3 = t.w + 1;
					local_max = max(max(local_max, c0), max(c1, max(c2, c3)));
					m += c0 + c1 + c2 + c3;
					q += (unsigned long long)c0 * c0 + (unsigned long long)c1 * c1 + (unsigned long long)c2 * c2 + (unsigned long long)c3 * c3;
					*reinterpret_cast<uint32_t *>(row + i) = (c0 & 0xff) | ((c1 & 0xff) << 8) | ((c2 & 0xff) << 16) | ((c3 & 0xff) << 24);
				}
			} else {
				for (int i = lane; i < nbins; i += 32) {
					const uint32_t c = tab[i] + 1;
					tab[i] = 0;
					local_max = max(local_max, c);
					m += c; q += (unsigned long long)c * c;
					row[i] = (uint8_t)c;
				}
			}
		} else {
			uint16_t *row = reinterpret_cast<uint16_t *>(hist) + (size_t)s * nbins;
			for (int i = lane; i < nbins; i += 32) {
				const uint32_t c = tab[i] + 1;
				tab[i] = 0;
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

// ---------------------------------------------------------------------------------------------
// validation at load time: every letter must be one the reference encodes (ChromosomeOneDigit.cpp:59-85);
// flags[0] is set otherwise (the reference throws InvalidInputException).  Read-only.
// ---------------------------------------------------------------------------------------------
__global__ void validate_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                                const int64_t *__restrict__ seg_off, long long n, unsigned int *__restrict__ flags) {
	__shared__ uint8_t lut[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = c_code_lut[i];
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	bool bad = false;
	for (long long s = warp; s < n; s += nwarps) {
		if (seg_off[s + 1] == seg_off[s]) continue;   // no segment: encodeNucleotides touches nothing (and checks nothing)
		const long long b0 = seq_off[s], b1 = seq_off[s + 1];
		for (long long base = (b0 & ~15LL) + lane * 16; base < b1; base += 32 * 16) {
			const uint4 x = *reinterpret_cast<const uint4 *>(seq + base);
			const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
			for (int q = 0; q < 4; q++) {
				// words of plain ACGT / acgt inside the sequence pass on the permute check of K1; the table sees the rest
				bool fast;
				k1_codes4_check(w[q], fast);
				if (fast && base + 4 * q >= b0 && base + 4 * q + 3 < b1) continue;
#pragma unroll
				for (int j = 4 * q; j < 4 * q + 4; j++) {
					const uint32_t ch = (w[q] >> ((j & 3) * 8)) & 0xffu;
					if (base + j >= b0 && base + j < b1 && lut[ch] == 0xffu) bad = true;
				}
			}
		}
	}
	if (__any_sync(MC_FULL_MASK, bad) && lane == 0) atomicOr(&flags[0], 1u);
}

int mc_launch_validate(mc_ctx *ctx) {
	const int threads = 256;
	int64_t blocks = (ctx->n * 32 + threads - 1) / threads;
	if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
	if (blocks < 1) blocks = 1;
	validate_kernel<<<(int)blocks, threads, 0, ctx->stream>>>(ctx->d_seq, ctx->d_seq_off, ctx->d_seg_off, ctx->n, ctx->d_flags);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

template <int TB, int K>
static int k1_launch(mc_ctx *ctx) {
	constexpr int nbins = 1 << (2 * K);
	constexpr size_t per_warp = (size_t)nbins * 4;
	int wpb = 8;
	while (wpb > 1 && per_warp * (wpb + 1) > 200 * 1024) wpb >>= 1;
	const size_t smem = per_warp * (wpb + 1);   // + one table of slack: the tables start at a multiple of their size
	const int threads = wpb * 32;
	int64_t blocks = (ctx->n + wpb - 1) / wpb;
	const int64_t cap = (int64_t)ctx->num_sms * 16;
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
	if (smem > 48 * 1024) MC_CUDA(cudaFuncSetAttribute(kmer_count_kernel<TB, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	kmer_count_kernel<TB, K><<<(int)blocks, threads, smem, ctx->stream>>>(ctx->d_seq, ctx->d_seq_off, ctx->d_segs, ctx->d_seg_off, ctx->n, (uint8_t *)ctx->d_hist, ctx->d_aux, ctx->d_flags);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}

int mc_launch_kmer_hist(mc_ctx *ctx, int k, int tbytes) {
	MC_REQUIRE(k >= 1 && k <= 7 && (tbytes == 1 || tbytes == 2), MC_ERR_UNSUPPORTED, "k=%d with %d-byte bins is not supported (k 1..7, 1 or 2 bytes)", k, tbytes);
#define K1_CASE(KK) case KK: return tbytes == 1 ? k1_launch<1, KK>(ctx) : k1_launch<2, KK>(ctx);
	switch (k) {
		K1_CASE(1) K1_CASE(2) K1_CASE(3) K1_CASE(4) K1_CASE(5) K1_CASE(6) K1_CASE(7)
	default: break;
	}
#undef K1_CASE
	return MC_ERR_UNSUPPORTED;
}
