// FASTA ingest on the device: the sequence lines of every record, as they stand in the file,
// become the record's letters -- ChromListMaker::makeChromOneDigitList (nonltr/ChromListMaker.cpp:92-120) appends
// every non-header line verbatim to the sequence (Chromosome::appendToSequence, Chromosome.cpp:73-82) -- and the
// scan for N that decides the segments (Chromosome::removeN, Chromosome.cpp:162-184) is reduced to one flag per
// record: records without N are one run, the usual case; only the others go through the host's segment code.
//
// The host indexes the file (where the header lines are, how many letters every record holds) and hands over
// the raw bytes plus, per record, the byte span of its sequence lines and the place of its letters in the letter
// buffer (the caller's row order, so the rows come out already permuted).  One warp per record: 512 raw bytes per
// round (one aligned 16-byte load per lane), line feeds squeezed out with a warp-wide prefix sum of the kept
// counts, letters stored at their final place.  Bytes: span read + letters written, once.
#include "mc_common.cuh"

// record flags
constexpr unsigned MC_REC_HAS_N = 1u;       // an 'N' or 'n': segments must be derived from the letters
constexpr unsigned MC_REC_NOT_PLAIN = 2u;   // a letter other than A, C, G, T, N in either case: needs validation

__global__ void __launch_bounds__(256)
ingest_compact_kernel(const uint8_t *__restrict__ raw, const int64_t *__restrict__ span_begin, const int64_t *__restrict__ span_end,
                      const int64_t *__restrict__ seq_off, long long n, uint8_t *__restrict__ seq, uint8_t *__restrict__ rec_flags,
                      unsigned int *__restrict__ err) {
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	for (long long s = warp; s < n; s += nwarps) {
		const long long b = span_begin[s], e = span_end[s];
		uint8_t *out = seq + seq_off[s];
		const long long want = seq_off[s + 1] - seq_off[s];
		long long written = 0;
		unsigned fl = 0;
		for (long long rb = b & ~15LL; rb < e; rb += 32 * 16) {   // (warp-uniform trip count: the prefix sums are warp-wide)
			const long long base = rb + lane * 16;
			uint4 v = make_uint4(0u, 0u, 0u, 0u);
			if (base < e) v = *reinterpret_cast<const uint4 *>(raw + base);   // (the raw buffer has a 64-byte tail)
			const uint32_t w[4] = {v.x, v.y, v.z, v.w};
			unsigned km = 0;   // bit j: byte j of this lane is a letter of the record
#pragma unroll
			for (int j = 0; j < 16; j++) {
				const uint32_t ch = (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
				const long long pos = base + j;
				if (pos >= b && pos < e && ch != '\n') {
					km |= 1u << j;
					const uint32_t u = ch & 0xdfu;
					if (u == 'N') fl |= MC_REC_HAS_N;
					else if (u != 'A' && u != 'C' && u != 'G' && u != 'T') fl |= MC_REC_NOT_PLAIN;
				}
			}
			const int cnt = __popc(km);
			int incl = cnt;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const int t = __shfl_up_sync(MC_FULL_MASK, incl, o);
				if (lane >= o) incl += t;
			}
			const int tot = __shfl_sync(MC_FULL_MASK, incl, 31);
			const long long at = written + incl - cnt;
			if (at + cnt <= want) {
#pragma unroll
				for (int j = 0; j < 16; j++)
					if ((km >> j) & 1u) out[at + __popc(km & ((1u << j) - 1u))] = (uint8_t)((w[j >> 2] >> ((j & 3) * 8)) & 0xffu);
			}
			written += tot;
		}
		fl = __reduce_or_sync(MC_FULL_MASK, fl);
		if (lane == 0) {
			rec_flags[s] = (uint8_t)fl;
			if (written != want) atomicExch(err, 1u);   // the index and the bytes disagree
		}
	}
}

int mc_launch_ingest(mc_ctx *ctx, const uint8_t *raw_dev, const int64_t *span_begin_dev, const int64_t *span_end_dev, uint8_t *rec_flags_dev,
                     unsigned int *err_dev) {
	const int threads = 256;
	int64_t blocks = (ctx->n * 32 + threads - 1) / threads;
	if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
	if (blocks < 1) blocks = 1;
	ingest_compact_kernel<<<(int)blocks, threads, 0, ctx->stream>>>(raw_dev, span_begin_dev, span_end_dev, ctx->d_seq_off, ctx->n, ctx->d_seq,
	                                                                 rec_flags_dev, err_dev);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	return MC_OK;
}
