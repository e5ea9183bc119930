// Byte/halfword SIMD reductions over k-mer histograms + the warp-level "32 rows per warp" tile
// shared by the scan, distance-key and mean-shift kernels.
//
// The two reductions every live feature derives from:
//   S = sum_i min(p_i, q_i)      D = sum_i p_i * q_i
// uint8 bins:  sum|p-q| with VABSDIFF4.U8.ACC (1 instr / 4 bins), S = (mag_p + mag_q - sum|p-q|)/2
//              exactly;  D with IDP.4A.U8.U8 (1 instr / 4 bins).
// uint16 bins: VIMNMX.U16x2 + IDP.2A for S, two IMADs for D (64-bit accumulator).
#pragma once
#include "mc_common.cuh"

// --- streaming 16-byte load, no L1 allocation (each row is read once per scan) ---------------
__device__ __forceinline__ uint4 mc_ld_stream16(const void *p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
	             : "l"(p));
	return r;
}

template <int TB>
struct PairAcc;

template <>
struct PairAcc<1> {
	uint32_t a = 0;   // sum |p-q|
	uint32_t d = 0;   // sum p*q   (fits 32 bits up to k = 8)
	__device__ __forceinline__ void add(uint32_t p, uint32_t q) {
		asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a) : "r"(p), "r"(q));
		d = __dp4a(p, q, d);
	}
	__device__ __forceinline__ void shfl_add_from(const PairAcc &send, int mask) {
		a += __shfl_xor_sync(MC_FULL_MASK, send.a, mask);
		d += __shfl_xor_sync(MC_FULL_MASK, send.d, mask);
	}
	// S from the absolute-difference sum
	__device__ __forceinline__ uint64_t summin(uint64_t mp, uint64_t mq) const {
		return (mp + mq - (uint64_t)a) >> 1;
	}
	__device__ __forceinline__ uint64_t dot() const { return d; }
};

template <>
struct PairAcc<2> {
	uint32_t a = 0;             // sum min(p,q) directly
	unsigned long long d = 0;   // sum p*q
	__device__ __forceinline__ void add(uint32_t p, uint32_t q) {
		a = __dp2a_lo(__vminu2(p, q), 0x0101u, a);
		d += (unsigned long long)((p & 0xffffu) * (q & 0xffffu));
		d += (unsigned long long)((p >> 16) * (q >> 16));
	}
	__device__ __forceinline__ void shfl_add_from(const PairAcc &send, int mask) {
		a += __shfl_xor_sync(MC_FULL_MASK, send.a, mask);
		d += __shfl_xor_sync(MC_FULL_MASK, send.d, mask);
	}
	__device__ __forceinline__ uint64_t summin(uint64_t, uint64_t) const { return a; }
	__device__ __forceinline__ uint64_t dot() const { return d; }
};

// --- row geometry -----------------------------------------------------------------------------
// RB = bytes per histogram row.  A row is covered by LPP lanes, each owning CH chunks of CHUNK
// bytes (chunk c of lane r sits at byte (c*LPP + r)*CHUNK, so one group reads contiguous memory).
// A warp works on 32 rows at a time: G = 32/LPP groups, LPP iterations; in iteration `it` group
// g reduces row g*LPP + it, and the transposing reduction below leaves row `lane` in lane `lane`,
// so all 32 lanes run the FP64 epilogue on distinct rows.
template <int RB>
struct RowCfg {
	static constexpr int CHUNK = RB >= 16 ? 16 : RB;
	static constexpr int WORDS = CHUNK / 4;
	static constexpr int LPP = RB >= 512 ? 32 : (RB >= 16 ? RB / 16 : 1);
	static constexpr int G = 32 / LPP;
	static constexpr int CH = RB / (CHUNK * LPP);
	static constexpr bool CENTER_IN_REGS = CH <= 8;
};

template <int WORDS>
__device__ __forceinline__ void mc_load_chunk(const uint8_t *p, uint32_t (&w)[WORDS]) {
	if constexpr (WORDS == 4) {
		const uint4 v = mc_ld_stream16(p);
		w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
	} else if constexpr (WORDS == 2) {
		const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
		w[0] = v.x; w[1] = v.y;
	} else {
		w[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
	}
}

// transposing warp reduction: lane (g, r) enters with v[it], it in [0,LPP), its partial for row
// g*LPP + it; leaves with the group total of row g*LPP + r, i.e. row `lane`.  LPP-1 shuffles per
// accumulator field instead of LPP*log2(LPP).
template <int LPP, class Acc>
__device__ __forceinline__ Acc mc_transpose_reduce(Acc (&v)[LPP], int r) {
#pragma unroll
	for (int m = LPP / 2; m >= 1; m >>= 1) {
		const bool upper = (r & m) != 0;
#pragma unroll
		for (int i = 0; i < m; i++) {
			const Acc send = upper ? v[i] : v[i + m];
			Acc keep = upper ? v[i + m] : v[i];
			keep.shfl_add_from(send, m);
			v[i] = keep;
		}
	}
	return v[0];
}

// center row -> registers (CH*WORDS words per lane) or shared memory for very wide rows
template <int RB>
struct CenterRegs {
	using C = RowCfg<RB>;
	uint32_t w[C::CENTER_IN_REGS ? C::CH : 1][C::WORDS];
	__device__ __forceinline__ void load(const uint8_t *row, int r) {
		if constexpr (C::CENTER_IN_REGS) {
#pragma unroll
			for (int c = 0; c < C::CH; c++) mc_load_chunk<C::WORDS>(row + (size_t)(c * C::LPP + r) * C::CHUNK, w[c]);
		}
	}
};

// reduce one row against the center: returns this lane's partial
template <int TB, int RB>
__device__ __forceinline__ PairAcc<TB> mc_row_partial(const uint8_t *row, int r, const CenterRegs<RB> &cen,
                                                       const uint32_t *cen_smem) {
	using C = RowCfg<RB>;
	PairAcc<TB> acc;
	if constexpr (C::CENTER_IN_REGS) {
		uint32_t w[C::CH][C::WORDS];
#pragma unroll
		for (int c = 0; c < C::CH; c++) mc_load_chunk<C::WORDS>(row + (size_t)(c * C::LPP + r) * C::CHUNK, w[c]);
#pragma unroll
		for (int c = 0; c < C::CH; c++)
#pragma unroll
			for (int j = 0; j < C::WORDS; j++) acc.add(w[c][j], cen.w[c][j]);
	} else {
#pragma unroll 4
		for (int c = 0; c < C::CH; c++) {
			uint32_t w[C::WORDS];
			const int chunk = c * C::LPP + r;
			mc_load_chunk<C::WORDS>(row + (size_t)chunk * C::CHUNK, w);
			const uint4 q = *reinterpret_cast<const uint4 *>(cen_smem + chunk * 4);
			acc.add(w[0], q.x); acc.add(w[1], q.y); acc.add(w[2], q.z); acc.add(w[3], q.w);
		}
	}
	return acc;
}

// generic-address pair reduction of one row against a "center" row (global or shared memory)
template <int TB>
__device__ __forceinline__ PairAcc<TB> mc_warp_pair_reduce(const uint8_t *p, const uint8_t *q, int rb, int lane) {
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

// dispatch a functor over the supported (tbytes, row bytes) pairs
#define MC_DISPATCH_ROW(tbytes, nbins, FN)                                   \
	do {                                                                     \
		const int _rb = (tbytes) * (nbins);                                  \
		if ((tbytes) == 1) {                                                 \
			switch (_rb) {                                                   \
			case 4: FN(1, 4); break;                                         \
			case 16: FN(1, 16); break;                                       \
			case 64: FN(1, 64); break;                                       \
			case 256: FN(1, 256); break;                                     \
			case 1024: FN(1, 1024); break;                                   \
			case 4096: FN(1, 4096); break;                                   \
			case 16384: FN(1, 16384); break;                                 \
			case 65536: FN(1, 65536); break;                                 \
			default: MC_REQUIRE(false, MC_ERR_UNSUPPORTED, "unsupported k"); \
			}                                                                \
		} else {                                                             \
			switch (_rb) {                                                   \
			case 8: FN(2, 8); break;                                         \
			case 32: FN(2, 32); break;                                       \
			case 128: FN(2, 128); break;                                     \
			case 512: FN(2, 512); break;                                     \
			case 2048: FN(2, 2048); break;                                   \
			case 8192: FN(2, 8192); break;                                   \
			case 32768: FN(2, 32768); break;                                 \
			default: MC_REQUIRE(false, MC_ERR_UNSUPPORTED, "unsupported k"); \
			}                                                                \
		}                                                                    \
	} while (0)
