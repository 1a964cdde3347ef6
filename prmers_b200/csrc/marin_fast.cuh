// marin_fast.cuh -- specialised kernels of the three-pass squaring / multiplication for sm_100a.
//
// Same mathematics, same tables and the same register images as the generic phase functions of
// marin_kernels.cuh (every value is canonical, so any exact evaluation order gives identical bits);
// what changes is how the work is laid out on the SM:
//   * every stride is a compile-time constant: shared memory is addressed with LDS / STS and immediate
//     offsets (the generic path pays a 64-bit generic-pointer add per element: 17 % of its instructions);
//   * phases that only exchange data inside a block of 512 lane elements are WARP-LOCAL: the warp that
//     wrote the block reads it back after a __syncwarp(), so a contiguous pass has two block-wide
//     barriers instead of seven and a strided inverse+carry+forward pass three instead of nine; warps
//     drift apart and mix their ALU-heavy and multiply-heavy stretches instead of hitting them together;
//   * the carry of a tile row, its publication, the wait for the row before it in the previous tile, the
//     absorption and the write-back are ONE thread's job (no block barrier around the hand-off: a late
//     neighbour stalls one warp, not the CTA);
//   * digits are re-weighted with a 32 x 64-bit product (a digit is < 2^32), 11 instead of 18 instructions.
// Replaces, by function, kernels/marin.cl:960-1107 (mid-stage forward/backward passes), :1245-1625 (last
// stage forward + pointwise + backward) and :1693-1728 + :2196-2216 (unweight, carry, reweight, fix-up).
//
// Covered shapes: contiguous pass S = 4096 = 16^3; strided pass R = 256 T, T in {1, 2, 4}, four pair columns
// per tile.  Everything else (5 * 2^k, second-level strided passes, small transforms) runs the generic kernels.
#pragma once
#include "marin_kernels.cuh"

#if defined(__CUDACC__)
namespace marin { namespace fast {

// ------------------------------------------------------------------ arithmetic helpers

// d * w mod P for a 32-bit d and canonical w; canonical result, identical to gl::mul(d, w).
// d w = t2 2^64 + T, T = (t1:t0); T + t2 (2^32 - 1) = T - Nb with Nb = P - t2 (2^32 - 1) in (2^32, P]:
// T + (t2 + 1)(2^32 - 1) carries out exactly when T >= Nb, and the canonical result is that sum when it
// carries, that sum minus (2^32 - 1) (i.e. plus P) when it does not.
__device__ __forceinline__ u64 mul_small(u32 d, u64 w)
{
	u32 w0 = (u32)w, w1 = (u32)(w >> 32), t0, t1, t2, r0, r1;
	asm("{\n\t.reg .u32 u, m;\n\t"
		"mul.lo.u32 %0, %5, %6;\n\t"
		"mul.hi.u32 %1, %5, %6;\n\t"
		"mad.lo.cc.u32 %1, %5, %7, %1;\n\t"
		"madc.hi.u32 %2, %5, %7, 0;\n\t"
		"add.u32 u, %2, 1;\n\t"
		"mad.lo.cc.u32 %3, u, 0xffffffff, %0;\n\t"
		"madc.hi.cc.u32 %4, u, 0xffffffff, %1;\n\t"
		"madc.lo.u32 m, %8, %8, 0xffffffff;\n\t"  // m = carry ? 0 : 2^32 - 1
		"sub.cc.u32 %3, %3, m;\n\t"
		"subc.u32 %4, %4, 0;\n\t"
		"}" : "=&r"(t0), "=&r"(t1), "=&r"(t2), "=&r"(r0), "=&r"(r1) : "r"(d), "r"(w0), "r"(w1), "r"(GL_ZERO));
	return ((u64)r1 << 32) | r0;
}

// ------------------------------------------------------------------ radix-16 stages with compile-time strides
// Forward (DIF): inputs natural, output k in slot sw16(k), then times the stage twiddle tw[(k - 1) TWS] (k >= 1).
// Inverse: input q (times the inverse stage twiddle) in slot sw16(q), outputs natural.  Unnormalised.

template <int STW> __device__ __forceinline__ void ld16_nat(u64 *a, const u64 *p)
{
	#pragma unroll
	for (int q = 0; q < 16; ++q) a[q] = p[q * STW];
}
template <int STW> __device__ __forceinline__ void ld16_sw(u64 *a, const u64 *p)
{
	#pragma unroll
	for (int q = 0; q < 16; ++q) a[sw16(q)] = p[q * STW];
}
template <int STW> __device__ __forceinline__ void st16_nat(u64 *p, const u64 *a)
{
	#pragma unroll
	for (int k = 0; k < 16; ++k) p[k * STW] = a[k];
}
template <int STW> __device__ __forceinline__ void st16_sw(u64 *p, const u64 *a)
{
	#pragma unroll
	for (int k = 0; k < 16; ++k) p[k * STW] = a[sw16(k)];
}
template <int TWS> __device__ __forceinline__ void tw16(u64 *a, const u64 *tw)
{
	#pragma unroll
	for (int k = 1; k < 16; ++k) a[sw16(k)] = gl::mul(a[sw16(k)], tw[(k - 1) * TWS]);
}

// ================================================================== contiguous pass, S = 4096 = 16 x 16 x 16
// Shared layout: pair i at pair slot i + (i >> 4) (one pair of padding per 16), a thread works on ONE lane of the
// pairs (even or odd digit): adjacent threads take the two lanes of the same pairs.
//   F0  s = 256  global -> shared          (block-wide barrier after it)
//   per 256-pair block, owned by one warp: F1 (s = 16), F2 (s = 1), pointwise, I2, I1   (__syncwarp between)
//   I0  s = 256  shared -> global          (block-wide barrier before it)

static constexpr u32 MID_S = 4096, MID_THREADS = 256;
static constexpr u32 MID_SMEM = (MID_S + MID_S / 16 + 1) * 16;
// FIVE: rows of S = 5120 = 5 x 4 x 16 x 16 pairs (n = 5 * 2^21: the radix-5 stage of a 5 * 2^k transform lives in the
// contiguous pass, so the strided passes are the power-of-two kernels below): a radix-5 and a radix-4 stage across the
// row, then twenty warp-local blocks of 256 pairs; 320 threads, two CTAs per SM.
static constexpr u32 MID5_S = 5120, MID5_THREADS = 320;
static constexpr u32 MID5_SMEM = (MID5_S + MID5_S / 16 + 1) * 16;
// SHAPE 2, 3: rows of S = 2048 = 2 x 4 x 16 x 16 and S = 1024 = 4 x 16 x 16 pairs (n = 2^22 and 2^21 with 1024 rows, so
// that the strided passes are the specialised R = 1024 kernels): the radix-2 and radix-4 stages across the row are done
// by one thread on eight (four) elements 256 apart, straight from / to global memory, then one warp-local block per warp.
// SHAPE 4, 5: S = 2560 = 5 x 2 x 16 x 16 and S = 1280 = 5 x 16 x 16 (n = 5 * 2^20 and 5 * 2^19 with 1024 rows): the radix-5
// stage (and the radix-2 stage after it) in one thread on elements 256 apart, one warp-local block per warp.
// SHAPE 6: rows of S = 8192 = 2 x 16 x 16 x 16 pairs (n = 2^24 with 1024 rows: the transform stays on the three-pass path and
// on the specialised strided kernels instead of paying two more passes over memory): one thread takes the sixteen elements
// j + 256 q of BOTH halves of the row, does the radix-2 stage between them and the radix-16 stage of each half, straight
// from / to global memory; then 32 warp-local blocks over 16 warps.  136 KB of shared memory: one CTA of 512 threads per SM.
// SHAPE 7: rows of S = 10240 = 5 x 2 x 4 x 16 x 16 pairs (n = 5 * 2^22 with 1024 rows, same reason): the radix-5 stage across
// the row from / to global memory, the radix-2 and radix-4 stages as one eight-element step in shared memory, then 40
// warp-local blocks over 20 warps.  170 KB of shared memory: one CTA of 640 threads per SM.
static constexpr u32 MID_SHAPE_4096 = 0, MID_SHAPE_5120 = 1, MID_SHAPE_2048 = 2, MID_SHAPE_1024 = 3, MID_SHAPE_2560 = 4, MID_SHAPE_1280 = 5, MID_SHAPE_8192 = 6, MID_SHAPE_10240 = 7;
__host__ __device__ constexpr u32 mid_shape_S(int shape) { return shape == 0 ? 4096u : shape == 1 ? 5120u : shape == 2 ? 2048u : shape == 3 ? 1024u : shape == 4 ? 2560u : shape == 5 ? 1280u : shape == 6 ? 8192u : 10240u; }
__host__ __device__ constexpr u32 mid_shape_threads(int shape) { return shape == 0 ? 256u : shape == 1 ? 320u : shape == 2 ? 256u : shape == 3 ? 128u : shape == 4 ? 320u : shape == 5 ? 160u : shape == 6 ? 512u : 640u; }
__host__ __device__ constexpr u32 mid_shape_smem(int shape) { return (mid_shape_S(shape) + mid_shape_S(shape) / 16 + 1) * 16; }
__host__ __device__ constexpr u32 mid_shape_minblocks(int shape) { return shape == 0 ? 3u : shape == 1 ? 2u : shape == 2 ? 3u : shape == 3 ? 6u : shape == 4 ? 2u : shape == 5 ? 4u : 1u; }

template <int MODE, int SHAPE> __global__ void __launch_bounds__(mid_shape_threads(SHAPE), mid_shape_minblocks(SHAPE)) k_mid_fast(const __grid_constant__ MidArgs A)
{
	if (MARIN_PDL_MODE & 1) MARIN_PDL_TRIGGER();
	MARIN_PDL_WAIT();
	extern __shared__ __align__(16) unsigned char smraw[];
	pair_t *sm = reinterpret_cast<pair_t *>(smraw);
	const Tables &t = A.t;
	constexpr bool FIVE = SHAPE == 1, POW16 = SHAPE == 0, OCT = SHAPE == 2, QUAD = SHAPE == 3, TEN = SHAPE == 4, FIVE1 = SHAPE == 5, DBL = SHAPE == 6, FIVE2 = SHAPE == 7;
	constexpr u32 S = mid_shape_S(SHAPE), NTHR = mid_shape_threads(SHAPE), NW = NTHR / 32;
	constexpr u32 NIT = S / 256 / NW;                // warp-local blocks per warp
	constexpr u32 SB = FIVE2 ? 3 : (FIVE || OCT || TEN || DBL) ? 2 : 1; // index of the s = 16 stage in the stage list
	const u32 tid = threadIdx.x, lane = tid & 1, h = tid >> 1, wl = tid & 31, warp = tid >> 5;
	const size_t rowbase = (size_t)blockIdx.x * S;
	u64 *smw = reinterpret_cast<u64 *>(sm) + lane;
	const u64 *tw0 = t.twS + A.st[0].toff, *tw1 = t.twS + A.st[SB].toff;
	const u64 *twi0 = t.twSi + A.st[0].toff, *twi1 = t.twSi + A.st[SB].toff;

	if (POW16)
	{
		// ---- F0: elements j + 256 q of the row, straight from global memory
		const u64 *g = reinterpret_cast<const u64 *>(A.x + rowbase) + lane;
		#pragma unroll 1
		for (u32 it = 0; it < 2; ++it)
		{
			const u32 j = h + 128 * it;
			u64 a[16];
			ld16_nat<512>(a, g + 2 * j);
			dft16<false>(a);
			tw16<256>(a, tw0 + j);
			st16_sw<544>(smw + 2 * (j + (j >> 4)), a);
		}
	}
	else if (FIVE2)
	{
		// ---- F0: radix 5, elements j + 2048 q from global memory; F1 + F2: radix 2 (s = 1024) and radix 4 (s = 256) on the
		//      eight elements 256 apart inside each fifth of the row
		const u64 *g = reinterpret_cast<const u64 *>(A.x + rowbase) + lane;
		u64 nx[5];
		#pragma unroll
		for (int q = 0; q < 5; ++q) nx[q] = g[2 * (h + 2048 * q)];
		#pragma unroll 1
		for (u32 j = h; j < 2048; j += NTHR / 2)
		{
			u64 a[5];
			#pragma unroll
			for (int q = 0; q < 5; ++q) a[q] = nx[q];
			if (j + NTHR / 2 < 2048)
			{
				#pragma unroll
				for (int q = 0; q < 5; ++q) nx[q] = g[2 * (j + NTHR / 2 + 2048 * q)];
			}
			dft5(a, t.w5);
			#pragma unroll
			for (int k = 1; k < 5; ++k) a[k] = gl::mul(a[k], tw0[(k - 1) * 2048 + j]);
			u64 *o = smw + 2 * (j + (j >> 4));
			#pragma unroll
			for (int k = 0; k < 5; ++k) o[k * 4352] = a[k];
		}
		__syncthreads();
		const u64 *twa = t.twS + A.st[1].toff, *twq = t.twS + A.st[2].toff;
		#pragma unroll 1
		for (u32 rb = h; rb < 1280; rb += NTHR / 2)
		{
			const u32 j = rb & 255, base = (rb >> 8) * 2048 + j;
			u64 *o = smw + 2 * (base + (base >> 4));
			u64 a[8];
			#pragma unroll
			for (int q = 0; q < 8; ++q) a[q] = o[q * 544];
			#pragma unroll
			for (int q = 0; q < 4; ++q) { addsub(a[q], a[q + 4]); a[q + 4] = gl::mul(a[q + 4], twa[j + 256 * q]); }
			#pragma unroll
			for (int hf = 0; hf < 2; ++hf)
			{
				u64 *x = a + 4 * hf;
				dft4<false>(x[0], x[1], x[2], x[3]);
				o[hf * 2176] = x[0]; o[hf * 2176 + 544] = gl::mul(x[1], twq[j]);
				o[hf * 2176 + 1088] = gl::mul(x[2], twq[256 + j]); o[hf * 2176 + 1632] = gl::mul(x[3], twq[512 + j]);
			}
		}
	}
	else if (DBL)
	{
		// ---- F0 (radix 2, s = 4096) between the halves of the row and F1 (radix 16, s = 256) inside each, from global memory
		const u64 *g = reinterpret_cast<const u64 *>(A.x + rowbase) + lane;
		const u64 *twh = t.twS + A.st[1].toff;
		const u32 j = h;
		u64 a[16], b[16];
		ld16_nat<512>(a, g + 2 * j);
		ld16_nat<512>(b, g + 2 * (j + 4096));
		#pragma unroll
		for (int q = 0; q < 16; ++q) { addsub(a[q], b[q]); b[q] = gl::mul(b[q], tw0[j + 256 * q]); }
		dft16<false>(a);
		tw16<256>(a, twh + j);
		st16_sw<544>(smw + 2 * (j + (j >> 4)), a);
		dft16<false>(b);
		tw16<256>(b, twh + j);
		st16_sw<544>(smw + 2 * (4352 + j + (j >> 4)), b);
	}
	else if (OCT)
	{
		// ---- F0 (radix 2, s = 1024) and F1 (radix 4, s = 256) on the eight elements j + 256 q, from global memory
		const u64 *g = reinterpret_cast<const u64 *>(A.x + rowbase) + lane;
		const u64 *twq = t.twS + A.st[1].toff;
		#pragma unroll 1
		for (u32 it = 0; it < 2; ++it)
		{
			const u32 j = h + 128 * it;
			u64 a[8];
			#pragma unroll
			for (int q = 0; q < 8; ++q) a[q] = g[2 * (j + 256 * q)];
			#pragma unroll
			for (int q = 0; q < 4; ++q) { addsub(a[q], a[q + 4]); a[q + 4] = gl::mul(a[q + 4], tw0[j + 256 * q]); }
			u64 *o = smw + 2 * (j + (j >> 4));
			#pragma unroll
			for (int hf = 0; hf < 2; ++hf)
			{
				u64 *x = a + 4 * hf;
				dft4<false>(x[0], x[1], x[2], x[3]);
				o[hf * 2176] = x[0]; o[hf * 2176 + 544] = gl::mul(x[1], twq[j]);
				o[hf * 2176 + 1088] = gl::mul(x[2], twq[256 + j]); o[hf * 2176 + 1632] = gl::mul(x[3], twq[512 + j]);
			}
		}
	}
	else if (TEN)
	{
		// ---- F0 (radix 5, s = 512) on j + 512 q and j + 256 + 512 q, then F1 (radix 2, s = 256), from global memory
		const u64 *g = reinterpret_cast<const u64 *>(A.x + rowbase) + lane;
		const u64 *twq = t.twS + A.st[1].toff;
		#pragma unroll 1
		for (u32 j = h; j < 256; j += NTHR / 2)
		{
			u64 a[5], b[5];
			#pragma unroll
			for (int q = 0; q < 5; ++q) { a[q] = g[2 * (j + 512 * q)]; b[q] = g[2 * (j + 256 + 512 * q)]; }
			dft5(a, t.w5); dft5(b, t.w5);
			#pragma unroll
			for (int k = 1; k < 5; ++k) { a[k] = gl::mul(a[k], tw0[(k - 1) * 512 + j]); b[k] = gl::mul(b[k], tw0[(k - 1) * 512 + j + 256]); }
			const u64 w = twq[j];
			u64 *o = smw + 2 * (j + (j >> 4));
			#pragma unroll
			for (int k = 0; k < 5; ++k) { addsub(a[k], b[k]); o[k * 1088] = a[k]; o[k * 1088 + 544] = gl::mul(b[k], w); }
		}
	}
	else if (FIVE1)
	{
		// ---- F0 (radix 5, s = 256) on the elements j + 256 q, from global memory
		const u64 *g = reinterpret_cast<const u64 *>(A.x + rowbase) + lane;
		#pragma unroll 1
		for (u32 j = h; j < 256; j += NTHR / 2)
		{
			u64 a[5];
			#pragma unroll
			for (int q = 0; q < 5; ++q) a[q] = g[2 * (j + 256 * q)];
			dft5(a, t.w5);
			u64 *o = smw + 2 * (j + (j >> 4));
			o[0] = a[0];
			#pragma unroll
			for (int k = 1; k < 5; ++k) o[k * 544] = gl::mul(a[k], tw0[(k - 1) * 256 + j]);
		}
	}
	else if (QUAD)
	{
		// ---- F0 (radix 4, s = 256) on the elements j + 256 q, from global memory
		const u64 *g = reinterpret_cast<const u64 *>(A.x + rowbase) + lane;
		#pragma unroll 1
		for (u32 it = 0; it < 4; ++it)
		{
			const u32 j = h + 64 * it;
			u64 a0 = g[2 * j], a1 = g[2 * (j + 256)], a2 = g[2 * (j + 512)], a3 = g[2 * (j + 768)];
			dft4<false>(a0, a1, a2, a3);
			u64 *o = smw + 2 * (j + (j >> 4));
			o[0] = a0; o[544] = gl::mul(a1, tw0[j]); o[1088] = gl::mul(a2, tw0[256 + j]); o[1632] = gl::mul(a3, tw0[512 + j]);
		}
	}
	else
	{
		// ---- F0: radix 5, elements j + 1024 q from global memory; F1: radix 4 (s = 256) inside each fifth of the row
		const u64 *g = reinterpret_cast<const u64 *>(A.x + rowbase) + lane;
		u64 nx[5];                                           // the next butterfly's inputs are requested before this one is computed
		#pragma unroll
		for (int q = 0; q < 5; ++q) nx[q] = g[2 * (h + 1024 * q)];
		#pragma unroll 1
		for (u32 j = h; j < 1024; j += NTHR / 2)
		{
			u64 a[5];
			#pragma unroll
			for (int q = 0; q < 5; ++q) a[q] = nx[q];
			if (j + NTHR / 2 < 1024)
			{
				#pragma unroll
				for (int q = 0; q < 5; ++q) nx[q] = g[2 * (j + NTHR / 2 + 1024 * q)];
			}
			dft5(a, t.w5);
			#pragma unroll
			for (int k = 1; k < 5; ++k) a[k] = gl::mul(a[k], tw0[(k - 1) * 1024 + j]);
			u64 *o = smw + 2 * (j + (j >> 4));
			#pragma unroll
			for (int k = 0; k < 5; ++k) o[k * 2176] = a[k];
		}
		__syncthreads();
		const u64 *twq = t.twS + A.st[1].toff;
		#pragma unroll 1
		for (u32 rb = h; rb < 1280; rb += NTHR / 2)
		{
			const u32 j = rb & 255, base = (rb >> 8) * 1024 + j;
			u64 *o = smw + 2 * (base + (base >> 4));
			u64 a0 = o[0], a1 = o[544], a2 = o[1088], a3 = o[1632];
			dft4<false>(a0, a1, a2, a3);
			o[0] = a0; o[544] = gl::mul(a1, twq[j]); o[1088] = gl::mul(a2, twq[256 + j]); o[1632] = gl::mul(a3, twq[512 + j]);
		}
	}
	__syncthreads();

	// ---- warp-local part: blocks of 256 pairs
	const u32 jl = wl >> 1; // butterfly of this thread inside a block (F1: column j, F2: group of 16)
	#pragma unroll 1
	for (u32 it = 0; it < NIT; ++it)
	{
		const u32 Bk = warp + NW * it;
		u64 *blk = smw + 2 * (272 * Bk);
		{	// s = 16
			u64 a[16];
			ld16_nat<34>(a, blk + 2 * jl);
			dft16<false>(a);
			tw16<16>(a, tw1 + jl);
			st16_sw<34>(blk + 2 * jl, a);
		}
		__syncwarp();
		{	// s = 1, no twiddles
			u64 a[16];
			ld16_nat<2>(a, blk + 2 * (17 * jl));
			dft16<false>(a);
			st16_sw<2>(blk + 2 * (17 * jl), a);
		}
		__syncwarp();
		if (MODE == MODE_FWD)
		{
			// multiplicand: the forward transform in position order, copied out 16 bytes per thread
			pair_t *dst = A.dst + rowbase + 256 * Bk;
			#pragma unroll
			for (u32 u = 0; u < 8; ++u) { const u32 i = wl + 32 * u; dst[i] = sm[272 * Bk + i + (i >> 4)]; }
			continue;
		}
		{	// pointwise square / multiply modulo (t^2 - u), u = w_m^freq(position): one pair per thread
			const u64 rr = t.rowroot[blockIdx.x];
			#pragma unroll 2
			for (u32 u = 0; u < 8; ++u)
			{
				const u32 i = wl + 32 * u, ig = 256 * Bk + i;
				const u64 uu = gl::mul(rr, t.colroot[ig]);
				pair_t &v = sm[272 * Bk + i + (i >> 4)];
				v = (MODE == MODE_SQR) ? sqr22(v, uu) : mul22(v, A.y[rowbase + ig], uu);
			}
		}
		__syncwarp();
		{	// inverse, s = 1
			u64 a[16];
			ld16_sw<2>(a, blk + 2 * (17 * jl));
			dft16<true>(a);
			st16_nat<2>(blk + 2 * (17 * jl), a);
		}
		__syncwarp();
		{	// inverse, s = 16
			u64 a[16];
			ld16_sw<34>(a, blk + 2 * jl);
			tw16<16>(a, twi1 + jl);
			dft16<true>(a);
			st16_nat<34>(blk + 2 * jl, a);
		}
	}
	if (!(MARIN_PDL_MODE & 1)) MARIN_PDL_TRIGGER();   // late trigger: the last stage below is what is left
	if (MODE == MODE_FWD) return;
	__syncthreads();

	if (POW16)
	{
		// ---- I0: s = 256, back to global memory (in place)
		u64 *g = reinterpret_cast<u64 *>(A.x + rowbase) + lane;
		#pragma unroll 1
		for (u32 it = 0; it < 2; ++it)
		{
			const u32 j = h + 128 * it;
			u64 a[16];
			ld16_sw<544>(a, smw + 2 * (j + (j >> 4)));
			tw16<256>(a, twi0 + j);
			dft16<true>(a);
			st16_nat<512>(g + 2 * j, a);
		}
	}
	else if (FIVE2)
	{
		const u64 *twa = t.twSi + A.st[1].toff, *twq = t.twSi + A.st[2].toff;
		#pragma unroll 1
		for (u32 rb = h; rb < 1280; rb += NTHR / 2)
		{
			const u32 j = rb & 255, base = (rb >> 8) * 2048 + j;
			u64 *o = smw + 2 * (base + (base >> 4));
			u64 a[8];
			#pragma unroll
			for (int hf = 0; hf < 2; ++hf)
			{
				u64 *x = a + 4 * hf;
				x[0] = o[hf * 2176]; x[1] = gl::mul(o[hf * 2176 + 544], twq[j]);
				x[2] = gl::mul(o[hf * 2176 + 1088], twq[256 + j]); x[3] = gl::mul(o[hf * 2176 + 1632], twq[512 + j]);
				dft4<true>(x[0], x[1], x[2], x[3]);
			}
			#pragma unroll
			for (int q = 0; q < 4; ++q) { a[q + 4] = gl::mul(a[q + 4], twa[j + 256 * q]); addsub(a[q], a[q + 4]); }
			#pragma unroll
			for (int q = 0; q < 8; ++q) o[q * 544] = a[q];
		}
		__syncthreads();
		u64 *g = reinterpret_cast<u64 *>(A.x + rowbase) + lane;
		#pragma unroll 1
		for (u32 j = h; j < 2048; j += NTHR / 2)
		{
			const u64 *o = smw + 2 * (j + (j >> 4));
			u64 a[5];
			a[0] = o[0];
			#pragma unroll
			for (int k = 1; k < 5; ++k) a[k] = gl::mul(o[k * 4352], twi0[(k - 1) * 2048 + j]);
			dft5(a, t.w5i);
			#pragma unroll
			for (int q = 0; q < 5; ++q) g[2 * (j + 2048 * q)] = a[q];
		}
	}
	else if (DBL)
	{
		// ---- I1 (radix 16, s = 256) inside each half and I0 (radix 2, s = 4096) between them, back to global memory (in place)
		u64 *g = reinterpret_cast<u64 *>(A.x + rowbase) + lane;
		const u64 *twh = t.twSi + A.st[1].toff;
		const u32 j = h;
		u64 a[16], b[16];
		ld16_sw<544>(a, smw + 2 * (j + (j >> 4)));
		tw16<256>(a, twh + j);
		dft16<true>(a);
		ld16_sw<544>(b, smw + 2 * (4352 + j + (j >> 4)));
		tw16<256>(b, twh + j);
		dft16<true>(b);
		#pragma unroll
		for (int q = 0; q < 16; ++q) { b[q] = gl::mul(b[q], twi0[j + 256 * q]); addsub(a[q], b[q]); }
		st16_nat<512>(g + 2 * j, a);
		st16_nat<512>(g + 2 * (j + 4096), b);
	}
	else if (OCT)
	{
		// ---- I1 (radix 4, s = 256) and I0 (radix 2, s = 1024), back to global memory (in place)
		u64 *g = reinterpret_cast<u64 *>(A.x + rowbase) + lane;
		const u64 *twq = t.twSi + A.st[1].toff;
		#pragma unroll 1
		for (u32 it = 0; it < 2; ++it)
		{
			const u32 j = h + 128 * it;
			const u64 *o = smw + 2 * (j + (j >> 4));
			u64 a[8];
			#pragma unroll
			for (int hf = 0; hf < 2; ++hf)
			{
				u64 *x = a + 4 * hf;
				x[0] = o[hf * 2176]; x[1] = gl::mul(o[hf * 2176 + 544], twq[j]);
				x[2] = gl::mul(o[hf * 2176 + 1088], twq[256 + j]); x[3] = gl::mul(o[hf * 2176 + 1632], twq[512 + j]);
				dft4<true>(x[0], x[1], x[2], x[3]);
			}
			#pragma unroll
			for (int q = 0; q < 4; ++q) { a[q + 4] = gl::mul(a[q + 4], twi0[j + 256 * q]); addsub(a[q], a[q + 4]); }
			#pragma unroll
			for (int q = 0; q < 8; ++q) g[2 * (j + 256 * q)] = a[q];
		}
	}
	else if (TEN)
	{
		u64 *g = reinterpret_cast<u64 *>(A.x + rowbase) + lane;
		const u64 *twq = t.twSi + A.st[1].toff;
		#pragma unroll 1
		for (u32 j = h; j < 256; j += NTHR / 2)
		{
			const u64 w = twq[j];
			const u64 *o = smw + 2 * (j + (j >> 4));
			u64 a[5], b[5];
			#pragma unroll
			for (int k = 0; k < 5; ++k) { a[k] = o[k * 1088]; b[k] = gl::mul(o[k * 1088 + 544], w); addsub(a[k], b[k]); }
			#pragma unroll
			for (int k = 1; k < 5; ++k) { a[k] = gl::mul(a[k], twi0[(k - 1) * 512 + j]); b[k] = gl::mul(b[k], twi0[(k - 1) * 512 + j + 256]); }
			dft5(a, t.w5i); dft5(b, t.w5i);
			#pragma unroll
			for (int q = 0; q < 5; ++q) { g[2 * (j + 512 * q)] = a[q]; g[2 * (j + 256 + 512 * q)] = b[q]; }
		}
	}
	else if (FIVE1)
	{
		u64 *g = reinterpret_cast<u64 *>(A.x + rowbase) + lane;
		#pragma unroll 1
		for (u32 j = h; j < 256; j += NTHR / 2)
		{
			const u64 *o = smw + 2 * (j + (j >> 4));
			u64 a[5];
			a[0] = o[0];
			#pragma unroll
			for (int k = 1; k < 5; ++k) a[k] = gl::mul(o[k * 544], twi0[(k - 1) * 256 + j]);
			dft5(a, t.w5i);
			#pragma unroll
			for (int q = 0; q < 5; ++q) g[2 * (j + 256 * q)] = a[q];
		}
	}
	else if (QUAD)
	{
		u64 *g = reinterpret_cast<u64 *>(A.x + rowbase) + lane;
		#pragma unroll 1
		for (u32 it = 0; it < 4; ++it)
		{
			const u32 j = h + 64 * it;
			const u64 *o = smw + 2 * (j + (j >> 4));
			u64 a0 = o[0], a1 = gl::mul(o[544], twi0[j]), a2 = gl::mul(o[1088], twi0[256 + j]), a3 = gl::mul(o[1632], twi0[512 + j]);
			dft4<true>(a0, a1, a2, a3);
			g[2 * j] = a0; g[2 * (j + 256)] = a1; g[2 * (j + 512)] = a2; g[2 * (j + 768)] = a3;
		}
	}
	else
	{
		const u64 *twq = t.twSi + A.st[1].toff;
		#pragma unroll 1
		for (u32 rb = h; rb < 1280; rb += NTHR / 2)
		{
			const u32 j = rb & 255, base = (rb >> 8) * 1024 + j;
			u64 *o = smw + 2 * (base + (base >> 4));
			u64 a0 = o[0], a1 = gl::mul(o[544], twq[j]), a2 = gl::mul(o[1088], twq[256 + j]), a3 = gl::mul(o[1632], twq[512 + j]);
			dft4<true>(a0, a1, a2, a3);
			o[0] = a0; o[544] = a1; o[1088] = a2; o[1632] = a3;
		}
		__syncthreads();
		u64 *g = reinterpret_cast<u64 *>(A.x + rowbase) + lane;
		#pragma unroll 1
		for (u32 j = h; j < 1024; j += NTHR / 2)
		{
			const u64 *o = smw + 2 * (j + (j >> 4));
			u64 a[5];
			a[0] = o[0];
			#pragma unroll
			for (int k = 1; k < 5; ++k) a[k] = gl::mul(o[k * 2176], twi0[(k - 1) * 1024 + j]);
			dft5(a, t.w5i);
			#pragma unroll
			for (int q = 0; q < 5; ++q) g[2 * (j + 1024 * q)] = a[q];
		}
	}
}

// ================================================================== strided passes, R = 256 T rows, 4 pair columns
// Forward stage list (marin_plan.hpp, factor_stages): [16 (s = 16 T), 16 (s = T), T (s = 1)]; the last one is next to
// global memory and carries the boundary twiddle w_m^(K(row) column).  Tile = R rows x 4 pairs in shared memory with a
// pitch of 5 pairs.  Blocks of 16 T rows (512 lane elements for T = 4) are warp-local for the two small stages.

static constexpr u32 STR_THREADS = 256;

// Tile layout of the R = 1024 kernels: pair (row, c) of the 1024 x 4 tile sits at pair slot
//   4 row + (row >> 2) + (row >> 6) + c
// (one pair of padding after every 4 rows, one more after every 64).  Every access pattern of the pass is then free of
// bank conflicts: the 64-bit accesses of a radix-16 stage (a warp = 4 adjacent rows x 4 columns x 2 lanes = 32
// consecutive words), the 128-bit accesses of the radix-4 stage on pairs (rows 4 apart: slots 17 apart), of the carry
// (rows r + 64 b: slots 273 b apart, 273 = 1 mod 8) and of the coalesced re-weighting.  With a plain pitch of 5 pairs the
// radix-16 stages ran at 3 wavefronts per access instead of 2 and the carry at 8 instead of 4 (ncu: 16.2 M shared
// wavefronts per launch against 8.6 M ideal).  Rows r0 + 4 q and r0 + 64 q are still equally spaced: 17 and 273 pairs.
__device__ __forceinline__ constexpr u32 tile_pair(u32 row, u32 c) { return 4 * row + (row >> 2) + (row >> 6) + c; }
static constexpr u32 ST_SMALL = 2 * 17, ST_BIG = 2 * 273;      // words between rows 4 apart / 64 apart

template <int T> struct SC
{
	static constexpr u32 R = 256 * T;
	static constexpr u32 S0 = 16 * T;          // stride (rows) of the first radix-16 stage
	static constexpr u32 NB = 16;              // blocks of 16 T rows per tile
	static constexpr u32 SMEM = (4 * R + R / 4 + R / 64) * 16;
};

__device__ __forceinline__ u64 wait_slot_fast(const StridedArgs &A, size_t idx)
{
	u32 spins = 0;
	for (;;)
	{
		const pair_t v = marin_ld_slot(&A.slots[idx]);
		if (v.y == (u64)A.epoch) return v.x;
		if (++spins > SPIN_LIMIT) { MARIN_ST_VOLATILE_U32(&A.flags[A.tiles], 1u); return ~0ull; } // poisoned: the error word makes every later read-back fail
		if (spins > 64) MARIN_NANOSLEEP(100);
	}
}

// streaming loads: data read exactly once per launch (tile input, weight streams) must not evict the twiddle tables
// from L1 (no allocation there) (no allocation there)
__device__ __forceinline__ pair_t ld_stream(const pair_t *p)
{ pair_t r; asm volatile("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p)); return r; }
__device__ __forceinline__ u64 ld_stream(const u64 *p)
{ u64 r; asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p)); return r; }

// Boundary twiddles of a pair butterfly: rows base + q have frequency indices K0 + q Kstep, so the four factors are
// f0 g^q with f0 = w_m^(+-K0 colscale) (one product of two table entries) and g = w_m^(+-Kstep colscale), a per-thread
// constant (a thread keeps its column): two table loads per butterfly instead of eight, same number of products.
struct BoundaryPow { u64 g1, g2, g3; };
__device__ __forceinline__ BoundaryPow boundary_pow(const Tables &t, u64 colscale, u32 Kstep, bool inv)
{
	BoundaryPow b;
	b.g1 = boundary_tw(t, (u64)Kstep * colscale, inv);
	b.g2 = gl::mul(b.g1, b.g1);
	b.g3 = gl::mul(b.g2, b.g1);
	return b;
}

// ---- the small stages of one block of 16 T rows (warp-local), T = 4: radix 16 with s = 4, then radix 4 on whole pairs

// forward direction: shared -> radix-16 (s = 4) -> shared -> radix-4 on pairs, boundary twiddle -> global
__device__ __forceinline__ void fwd_small_T4(const StridedArgs &A, pair_t *sm, pair_t *gdst, u32 Bk, u32 c0, u32 wl, const BoundaryPow &bp)
{
	const Tables &t = A.t;
	{
		const u32 lane = wl & 1, cc = (wl >> 1) & 3, j = wl >> 3;
		u64 *p = reinterpret_cast<u64 *>(sm) + 2 * tile_pair(64 * Bk + j, cc) + lane;
		u64 a[16];
		ld16_nat<ST_SMALL>(a, p);
		dft16<false>(a);
		tw16<4>(a, A.tw + A.st[1].toff + j);
		st16_sw<ST_SMALL>(p, a);
	}
	__syncwarp();
	const u32 c = wl & 3;
	const size_t gcol = (size_t)c0 + c;
	const u64 colscale = (u64)(c0 + c) * A.escale;
	#pragma unroll
	for (u32 i = 0; i < 2; ++i)
	{
		const u32 base = 64 * Bk + 4 * ((wl + 32 * i) >> 2);
		const u64 f0 = boundary_tw(t, (u64)A.rowK[base] * colscale, false);
		u64 ax[4], ay[4];
		#pragma unroll
		for (u32 q = 0; q < 4; ++q) { const pair_t v = sm[tile_pair(base + q, c)]; ax[q] = v.x; ay[q] = v.y; }
		dft4<false>(ax[0], ax[1], ax[2], ax[3]); dft4<false>(ay[0], ay[1], ay[2], ay[3]);
		const u64 f1 = gl::mul(f0, bp.g1), f2 = gl::mul(f0, bp.g2), f3 = gl::mul(f0, bp.g3);
		pair_t o;
		o.x = gl::mul(ax[0], f0); o.y = gl::mul(ay[0], f0); gdst[gcol + (size_t)(base + 0) * A.stride] = o;
		o.x = gl::mul(ax[1], f1); o.y = gl::mul(ay[1], f1); gdst[gcol + (size_t)(base + 1) * A.stride] = o;
		o.x = gl::mul(ax[2], f2); o.y = gl::mul(ay[2], f2); gdst[gcol + (size_t)(base + 2) * A.stride] = o;
		o.x = gl::mul(ax[3], f3); o.y = gl::mul(ay[3], f3); gdst[gcol + (size_t)(base + 3) * A.stride] = o;
	}
}

// ---- the big radix-16 stage (s = 64 rows for T = 4): group G of four columns j = 4 G .. 4 G + 3, rows j + 64 q
// forward: FROMG = read the rows from global memory (first pass of an operation) instead of the tile
template <bool FROMG> __device__ __forceinline__ void fwd_big_T4(const StridedArgs &A, pair_t *sm, u32 G, u32 c0, u32 wl)
{
	const u32 lane = wl & 1, c = (wl >> 1) & 3, j = 4 * G + (wl >> 3);
	u64 *p = reinterpret_cast<u64 *>(sm) + 2 * tile_pair(j, c) + lane;
	u64 a[16];
	if (FROMG)
	{
		const u64 *g = reinterpret_cast<const u64 *>(A.src + (size_t)c0 + c + (size_t)j * A.stride) + lane;
		const size_t gs = 2 * (size_t)64 * A.stride;
		#pragma unroll
		for (int q = 0; q < 16; ++q) a[q] = g[q * gs];
	}
	else ld16_nat<ST_BIG>(a, p);
	dft16<false>(a);
	tw16<64>(a, A.tw + A.st[0].toff + j);
	st16_sw<ST_BIG>(p, a);
}

// second half, after EVERY run of the tile's warp has been published (a tile never holds back a carry its successor
// needs while it waits for its own predecessor; tile 0 waits for the last tile: end-around): wait for the run before
// each row in the previous tile, absorb, re-weight (coalesced; to the register unless the chain keeps the parked
// transform as the state, back into the tile when the forward pass follows).  The weights are requested before the wait.
// bits4: the width bits of the first four digits of the thread's two rows (4 + 4 bits), kept from the carry
template <bool FWD_FOLLOWS> __device__ __forceinline__ void absorb_reweight_T4(const StridedArgs &A, pair_t *sm, u32 G, u32 g, u32 c0, u32 wl, u32 bits4)
{
	const Tables &t = A.t;
	constexpr u32 R = 1024;
	const pair_t *wt = reinterpret_cast<const pair_t *>(t.w_tile) + (size_t)g * R * TILE_C;
	pair_t w[8];
	#pragma unroll
	for (u32 u = 0; u < 8; ++u)
	{
		const u32 ri = (wl >> 2) + 8 * u, row = 4 * G + (ri & 3) + 64 * (ri >> 2);
		w[u] = ld_stream(wt + (size_t)row * TILE_C + (wl & 3));
	}
	const u32 rows[2] = {4 * G + (wl & 3) + 64 * (wl >> 2), 4 * G + (wl & 3) + 64 * (wl >> 2) + 512};
	// both hand-off slots of the thread are requested before either is examined
	const size_t si[2] = {slot_before(A, g, rows[0]), slot_before(A, g, rows[1])};
	const pair_t s0 = marin_ld_slot(&A.slots[si[0]]), s1 = marin_ld_slot(&A.slots[si[1]]);
	u64 cins[2];
	cins[0] = (s0.y == (u64)A.epoch) ? s0.x : wait_slot_fast(A, si[0]);
	cins[1] = (s1.y == (u64)A.epoch) ? s1.x : wait_slot_fast(A, si[1]);
	#pragma unroll
	for (u32 u = 0; u < 2; ++u)
	{
		const u32 row = rows[u];
		const u64 cin = cins[u];
		if (cin == 0) continue;
		const pair_t p0 = sm[tile_pair(row, 0)], p1 = sm[tile_pair(row, 1)];
		u64 d[4] = {p0.x, p0.y, p1.x, p1.y};
		carry_absorb(d, 4, t.wbase, (bits4 >> (4 * u)) & 15u, cin);
		pair_t q0, q1; q0.x = d[0]; q0.y = d[1]; q1.x = d[2]; q1.y = d[3];
		sm[tile_pair(row, 0)] = q0; sm[tile_pair(row, 1)] = q1;
	}
	__syncwarp();
	#pragma unroll
	for (u32 u = 0; u < 8; ++u)
	{
		const u32 ri = (wl >> 2) + 8 * u, c = wl & 3, row = 4 * G + (ri & 3) + 64 * (ri >> 2);
		const pair_t d = sm[tile_pair(row, c)];
		pair_t o; o.x = mul_small((u32)d.x, w[u].x); o.y = mul_small((u32)d.y, w[u].y);
		if (!A.skip_dst) A.dst[(size_t)row * A.stride + c0 + c] = o;
		if (A.dst2) A.dst2[(size_t)row * A.stride + c0 + c] = o;
		if (FWD_FOLLOWS) sm[tile_pair(row, c)] = o;
	}
}

// ---- radix-16 stages with run-time strides: ONE copy of the inverse and ONE of the forward butterfly code serves both
// stages of a pass.  Warps run different phases at the same time here (that is the point of the warp-local phases), so
// the code they cover together must fit the instruction cache: with every stage unrolled separately the kernel was
// 170 KB and 21 % of its stall samples were "no instruction"; 15 address adds per butterfly are the cheaper price.
__device__ __forceinline__ void ld16_sw_rt(u64 *a, const u64 *p, u32 stw)
{
	#pragma unroll
	for (int q = 0; q < 16; ++q) a[sw16(q)] = p[q * stw];
}
__device__ __forceinline__ void ld16_nat_rt(u64 *a, const u64 *p, u32 stw)
{
	#pragma unroll
	for (int q = 0; q < 16; ++q) a[q] = p[q * stw];
}
__device__ __forceinline__ void st16_nat_rt(u64 *p, const u64 *a, u32 stw)
{
	#pragma unroll
	for (int k = 0; k < 16; ++k) p[k * stw] = a[k];
}
__device__ __forceinline__ void st16_sw_rt(u64 *p, const u64 *a, u32 stw)
{
	#pragma unroll
	for (int k = 0; k < 16; ++k) p[k * stw] = a[sw16(k)];
}
__device__ __forceinline__ void tw16_rt(u64 *a, const u64 *tw, u32 tws)
{
	#pragma unroll
	for (int k = 1; k < 16; ++k) a[sw16(k)] = gl::mul(a[sw16(k)], tw[(k - 1) * tws]);
}

// carry and publish the 64 rows of group G (one thread per row, two rows per thread: two independent serial chains,
// written out side by side so that their latencies overlap)
__device__ __forceinline__ void carry_publish_T4(const StridedArgs &A, pair_t *sm, u32 G, u32 g, u32 c0, u32 wl, u32 bits0, u32 bits1)
{
	const Tables &t = A.t;
	const u32 row0 = 4 * G + (wl & 3) + 64 * (wl >> 2), row1 = row0 + 64 * 8;
	u64 v0[RUN], v1[RUN];
	#pragma unroll
	for (u32 c = 0; c < TILE_C; ++c)
	{
		const pair_t p0 = sm[tile_pair(row0, c)], p1 = sm[tile_pair(row1, c)];
		v0[2 * c] = p0.x; v0[2 * c + 1] = p0.y; v1[2 * c] = p1.x; v1[2 * c + 1] = p1.y;
	}
	const u64 cout0 = carry_run(v0, RUN, t.wbase, bits0, A.a, 0);
	const u64 cout1 = carry_run(v1, RUN, t.wbase, bits1, A.a, 0);
	marin_st_slot(&A.slots[(size_t)g * 1024 + row0], cout0, A.epoch);
	marin_st_slot(&A.slots[(size_t)g * 1024 + row1], cout1, A.epoch);
	#pragma unroll
	for (u32 c = 0; c < TILE_C; ++c)
	{
		pair_t p0, p1; p0.x = v0[2 * c]; p0.y = v0[2 * c + 1]; p1.x = v1[2 * c]; p1.y = v1[2 * c + 1];
		sm[tile_pair(row0, c)] = p0; sm[tile_pair(row1, c)] = p1;
	}
}

// ---- bulk asynchronous copy (the TMA engine's 1-D form, cp.async.bulk; UBLKCP in SASS) with an mbarrier: one thread
// arms the barrier with the byte count and issues the copies, the data lands in shared memory without passing through
// any thread's registers, and whoever needs it waits on the barrier's phase
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, u32 count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, u32 bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, u32 bytes, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, u32 parity)
{
	asm volatile("{\n\t.reg .pred p;\n\tMBAR_WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra MBAR_DONE_%=;\n\tbra MBAR_WAIT_%=;\n\tMBAR_DONE_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// inverse strided pass with carry (level 0), optionally followed by the forward strided pass of the next squaring.
// Shared memory: the tile, then the stage twiddles of the direction at work (1020 words) -- L1 is a stream of
// tile inputs and weights here and a twiddle that misses costs an L2 round trip in the middle of a butterfly.
static constexpr u32 STR_TW = 15 * 64 + 15 * 4;               // stage twiddles of one direction, R = 1024
// ONE twiddle buffer: the inverse tables first, the forward tables copied over them once every warp is through the inverse
// stages.  Two resident CTAs then need 2 x 76.2 KB, which fits the 164 KB shared-memory configuration of the SM and leaves
// 64 KB of L1 for the boundary-twiddle and width tables; with both tables resident (2 x 84.2 KB) the SM was configured
// for 196 KB and 32 KB of L1 (forcing the 228 KB configuration costs this kernel another 17 us: profiles/r2/pdl_ab.log).
static constexpr u32 INV_SMEM_T4 = SC<4>::SMEM + STR_TW * 8;

// development aid (-DMARIN_PHASE_TIMING, tools/phase_clocks.py): SM cycle counter at the phase boundaries of every warp of
// four tiles in the middle of the grid
#ifdef MARIN_PHASE_TIMING
__device__ long long g_phase_clk[4 * 8 * 32];
#define MARIN_STAMP(i) do { if (g >= 600 && g < 604 && wl == 0) g_phase_clk[((g - 600) * 8 + warp) * 32 + (i)] = clock64(); } while (0)
#else
#define MARIN_STAMP(i) do { } while (0)
#endif

template <bool FWD_FOLLOWS> __global__ void __launch_bounds__(STR_THREADS, 2) k_inv_fast_T4(const __grid_constant__ StridedArgs A)
{
#ifdef MARIN_PHASE_TIMING
	const long long t_entry = clock64();
#endif
	extern __shared__ __align__(16) unsigned char smraw[];
	pair_t *sm = reinterpret_cast<pair_t *>(smraw);
	u64 *twsm = reinterpret_cast<u64 *>(smraw + SC<4>::SMEM);
	__shared__ u32 s_tile;
	__shared__ __align__(8) unsigned long long s_bar;
	if (MARIN_PDL_MODE & 2) MARIN_PDL_TRIGGER();
	if (threadIdx.x == 0)
	{
		// the stage twiddles (constant tables: nothing a predecessor wrote) come in by bulk asynchronous copy while the
		// contiguous pass before this kernel drains and the tile's first loads are in flight
		mbar_init(&s_bar, 1);
		mbar_expect_tx(&s_bar, STR_TW * 8);
		bulk_g2s(twsm, A.twi, STR_TW * 8, &s_bar);
	}
	MARIN_PDL_WAIT();
	// decoupled look-back needs "the tile I wait for has started": tile numbers come from an atomic ticket (taken after
	// the wait: the carry pass of the operation before this one hands out tickets from the same counter)
	if (threadIdx.x == 0) s_tile = (u32)(atomicAdd(A.ticket, 1ull) - A.ticket_base);
	__syncthreads();
	const Tables &t = A.t;
	const u32 g = s_tile, c0 = g * TILE_C, wl = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u32 c = wl & 3;                                      // column of this thread in the pair-wise radix-4 stages
	const u64 colscale = (u64)(c0 + c) * A.escale;
	const u32 lane = wl & 1, cc = (wl >> 1) & 3, jq = wl >> 3; // radix-16 stages: lane of the pair, column, butterfly of four
	u64 *smw = reinterpret_cast<u64 *>(sm) + 2 * cc + lane;
	const size_t gcol = (size_t)c0 + c;
	u32 keep = 0;                                             // width bits of the first four digits of the thread's four carry rows
#ifdef MARIN_PHASE_TIMING
	if (g >= 600 && g < 604 && wl == 0) g_phase_clk[((g - 600) * 8 + warp) * 32 + 0] = t_entry;
#endif
	MARIN_STAMP(1);
	{
		// ---- inverse: [radix 4 on pairs from global, radix 16 (s = 4)] per block of 64 rows, block barrier,
		//      [radix 16 (s = 64) + unweight, carry, publish] per group of 4 columns.
		// ONE register array serves as the queue of prefetched boundary inputs (two butterflies = 8 pairs, requested two
		// butterflies ahead of their use) in the first two phases and as the 16 unweighting factors in the last two.
		u64 pf[16];
		#pragma unroll
		for (u32 i = 0; i < 2; ++i)
			#pragma unroll
			for (u32 q = 0; q < 4; ++q)
			{
				const pair_t x = ld_stream(A.src + gcol + (size_t)(64 * warp + 4 * ((wl + 32 * i) >> 2) + q) * A.stride);
				pf[8 * i + 2 * q] = x.x; pf[8 * i + 2 * q + 1] = x.y;
			}
		// boundary twiddles: f0 of the four butterflies (two random table reads each) formed up front, in use order
		u64 fq[4];
		#pragma unroll
		for (u32 b = 0; b < 4; ++b)
			fq[b] = boundary_tw(t, (u64)A.rowK[64 * (warp + 8 * (b >> 1)) + 4 * ((wl + 32 * (b & 1)) >> 2)] * colscale, true);
		const BoundaryPow bp = boundary_pow(t, colscale, 256, true);
		u32 raw0 = 0, raw1 = 0;                                  // words of the width-bit table for the thread's two carry rows
		auto finish_group = [&](u32 blk)
		{
			const u32 r0 = 4 * blk + (wl & 3) + 64 * (wl >> 2);
			const u32 bits0 = (raw0 >> ((2 * (r0 * A.stride + c0)) & 31)) & 0xffu, bits1 = (raw1 >> ((2 * ((r0 + 512) * A.stride + c0)) & 31)) & 0xffu;
			keep = (keep << 8) | (bits0 & 15u) | ((bits1 & 15u) << 4);
			__syncwarp();
			carry_publish_T4(A, sm, blk, g, c0, wl, bits0, bits1);
		};
		#pragma unroll 1
		for (u32 ph = 0; ph < 4; ++ph)
		{
			const u32 blk = warp + 8 * (ph & 1);
			const bool big = ph >= 2;
			MARIN_STAMP(2 + 3 * ph);
			if (ph == 2) __syncthreads();
			MARIN_STAMP(3 + 3 * ph);
			if (!big)
			{
				#pragma unroll 1
				for (u32 i = 0; i < 2; ++i)
				{
					const u32 base = 64 * blk + 4 * ((wl + 32 * i) >> 2);
					const u64 f0 = fq[0];
					const u64 f1 = gl::mul(f0, bp.g1), f2 = gl::mul(f0, bp.g2), f3 = gl::mul(f0, bp.g3);
					u64 ax[4], ay[4];
					ax[0] = gl::mul(pf[0], f0); ay[0] = gl::mul(pf[1], f0);
					ax[1] = gl::mul(pf[2], f1); ay[1] = gl::mul(pf[3], f1);
					ax[2] = gl::mul(pf[4], f2); ay[2] = gl::mul(pf[5], f2);
					ax[3] = gl::mul(pf[6], f3); ay[3] = gl::mul(pf[7], f3);
					dft4<true>(ax[0], ax[1], ax[2], ax[3]); dft4<true>(ay[0], ay[1], ay[2], ay[3]);
					#pragma unroll
					for (u32 q = 0; q < 4; ++q) { pair_t o; o.x = ax[q]; o.y = ay[q]; sm[tile_pair(base + q, c)] = o; }
					fq[0] = fq[1]; fq[1] = fq[2]; fq[2] = fq[3];
					#pragma unroll
					for (u32 q = 0; q < 8; ++q) pf[q] = pf[8 + q];
					if (ph == 0)
					{
						const u32 nbase = 64 * (warp + 8) + 4 * ((wl + 32 * i) >> 2);
						#pragma unroll
						for (u32 q = 0; q < 4; ++q)
						{
							const pair_t x = ld_stream(A.src + gcol + (size_t)(nbase + q) * A.stride);
							pf[8 + 2 * q] = x.x; pf[8 + 2 * q + 1] = x.y;
						}
					}
				}
				__syncwarp();
			}
			else
			{
				// the 16 unweighting factors (a stream from HBM) and the width bits of this thread's two carry rows are
				// requested before the butterfly starts
				const u64 *wp = t.winvs_tile + 2 * (((size_t)g * 1024 + 4 * blk + jq) * TILE_C + cc) + lane;
				#pragma unroll
				for (int q = 0; q < 16; ++q) pf[q] = ld_stream(wp + (size_t)q * (64 * TILE_C * 2));
				// (the words only: the shift that extracts the eight bits waits until they are needed, so that the loads fly)
				const u32 r0 = 4 * blk + (wl & 3) + 64 * (wl >> 2);
				raw0 = t.wbits[(2 * (r0 * A.stride + c0)) >> 5];
				raw1 = t.wbits[(2 * ((r0 + 512) * A.stride + c0)) >> 5];
			}
			const u32 row0 = big ? 4 * blk + jq : 64 * blk + jq;
			const u32 stw = big ? ST_BIG : ST_SMALL, tws = big ? 64u : 4u;
			const u64 *tw = twsm + (big ? 4 * blk : 960u) + jq;
			u64 *p = smw + 2 * tile_pair(row0, 0);
			u64 a[16];
			ld16_sw_rt(a, p, stw);
			if (ph == 0) mbar_wait(&s_bar, 0);                   // the twiddles have landed (long ago: this is the first use)
			tw16_rt(a, tw, tws);
			dft16<true>(a);
			if (big)
			{
				if (!FWD_FOLLOWS && A.add_src) // mul_add: the other register joins while both are weighted (times m: pf carries the 1/m)
				{
					const u64 *ap = reinterpret_cast<const u64 *>(A.add_src) + 2 * ((size_t)row0 * A.stride + c0 + cc) + lane;
					const size_t as = 2 * (size_t)64 * A.stride;
					#pragma unroll
					for (int q = 0; q < 16; ++q) a[q] = gl::add(a[q], gl::mul(ap[q * as], (u64)t.m));
				}
				#pragma unroll
				for (int q = 0; q < 16; ++q) a[q] = gl::mul(a[q], pf[q]);
			}
			st16_nat_rt(p, a, stw);
			MARIN_STAMP(4 + 3 * ph);
			if (ph == 2) finish_group(blk);
		}
		finish_group(warp + 8);
	}
	MARIN_STAMP(14);
	#pragma unroll 1
	for (u32 it = 0; it < 2; ++it)
	{
		absorb_reweight_T4<FWD_FOLLOWS>(A, sm, warp + 8 * it, g, c0, wl, (keep >> (8 - 8 * it)) & 0xffu);
		MARIN_STAMP(15 + it);
	}
	if (!(MARIN_PDL_MODE & 2)) MARIN_PDL_TRIGGER();   // late trigger: the carries are out, the forward pass of the tile is left
	if (!FWD_FOLLOWS) return;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		// every warp is through with the inverse twiddles: the forward ones land in the same buffer while the first
		// forward butterflies run (they need them only after their 16-point transform)
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		mbar_expect_tx(&s_bar, STR_TW * 8);
		bulk_g2s(twsm, A.tw, STR_TW * 8, &s_bar);
	}
	MARIN_STAMP(17);
	// ---- forward pass of the next squaring: radix 16 (s = 64) per group, block barrier, [radix 16 (s = 4), radix 4 on
	//      pairs + boundary twiddle -> global] per block
	u64 fq[4];
	#pragma unroll
	for (u32 b = 0; b < 4; ++b)
		fq[b] = boundary_tw(t, (u64)A.rowK[64 * (warp + 8 * (b >> 1)) + 4 * ((wl + 32 * (b & 1)) >> 2)] * colscale, false);
	const BoundaryPow bpf = boundary_pow(t, colscale, 256, false);
	#pragma unroll 1
	for (u32 ph = 0; ph < 4; ++ph)
	{
		const u32 blk = warp + 8 * (ph & 1);
		const bool big = ph < 2;
		MARIN_STAMP(18 + 2 * ph);
		if (ph == 2) __syncthreads();
		MARIN_STAMP(19 + 2 * ph);
		const u32 row0 = big ? 4 * blk + jq : 64 * blk + jq;
		const u32 stw = big ? ST_BIG : ST_SMALL, tws = big ? 64u : 4u;
		const u64 *tw = twsm + (big ? 4 * blk : 960u) + jq;
		u64 *p = smw + 2 * tile_pair(row0, 0);
		u64 a[16];
		ld16_nat_rt(a, p, stw);
		dft16<false>(a);
		if (ph == 0) mbar_wait(&s_bar, 1);                   // the forward twiddles have landed
		tw16_rt(a, tw, tws);
		st16_sw_rt(p, a, stw);
		if (!big)
		{
			__syncwarp();
			#pragma unroll 1
			for (u32 i = 0; i < 2; ++i)
			{
				const u32 base = 64 * blk + 4 * ((wl + 32 * i) >> 2);
				const u64 f0 = fq[0];
				fq[0] = fq[1]; fq[1] = fq[2]; fq[2] = fq[3];
				u64 ax[4], ay[4];
				#pragma unroll
				for (u32 q = 0; q < 4; ++q) { const pair_t pv = sm[tile_pair(base + q, c)]; ax[q] = pv.x; ay[q] = pv.y; }
				dft4<false>(ax[0], ax[1], ax[2], ax[3]); dft4<false>(ay[0], ay[1], ay[2], ay[3]);
				const u64 f1 = gl::mul(f0, bpf.g1), f2 = gl::mul(f0, bpf.g2), f3 = gl::mul(f0, bpf.g3);
				pair_t o;
				o.x = gl::mul(ax[0], f0); o.y = gl::mul(ay[0], f0); A.fwd_dst[gcol + (size_t)(base + 0) * A.stride] = o;
				o.x = gl::mul(ax[1], f1); o.y = gl::mul(ay[1], f1); A.fwd_dst[gcol + (size_t)(base + 1) * A.stride] = o;
				o.x = gl::mul(ax[2], f2); o.y = gl::mul(ay[2], f2); A.fwd_dst[gcol + (size_t)(base + 2) * A.stride] = o;
				o.x = gl::mul(ax[3], f3); o.y = gl::mul(ay[3], f3); A.fwd_dst[gcol + (size_t)(base + 3) * A.stride] = o;
			}
		}
	}
	MARIN_STAMP(26);
}

// forward strided pass from a register in global memory
__global__ void __launch_bounds__(STR_THREADS, 2) k_fwd_fast_T4(const __grid_constant__ StridedArgs A)
{
	if (MARIN_PDL_MODE & 2) MARIN_PDL_TRIGGER();
	MARIN_PDL_WAIT();
	extern __shared__ __align__(16) unsigned char smraw[];
	pair_t *sm = reinterpret_cast<pair_t *>(smraw);
	const u32 g = blockIdx.x, c0 = g * TILE_C, wl = threadIdx.x & 31, warp = threadIdx.x >> 5;
	#pragma unroll 1
	for (u32 it = 0; it < 2; ++it) fwd_big_T4<true>(A, sm, warp + 8 * it, c0, wl);
	__syncthreads();
	const BoundaryPow bpf = boundary_pow(A.t, (u64)(c0 + (wl & 3)) * A.escale, 256, false);
	#pragma unroll 1
	for (u32 it = 0; it < 2; ++it) fwd_small_T4(A, sm, A.dst, warp + 8 * it, c0, wl, bpf);
}

// ================================================================== second-level strided pass with four rows
// n >= 2^25: m = 1024 x 4 x S.  The first-level pass (1024 rows, stride 4 S) is the specialised kernel above; what is left
// of the second level is ONE radix-4 butterfly per pair column of every block of 4 S pairs, with the boundary twiddle
// w_m^(+-K(row) column escale) -- a streaming kernel (coalesced 16-byte accesses, one thread per column and both lanes),
// not a tile kernel.  Same operation order as the generic stage (marin_kernels.cuh, run_stage_tile / bfly_core).
static constexpr u32 L1_THREADS = 256;
template <bool INV> __global__ void __launch_bounds__(L1_THREADS) k_level1_r4(const __grid_constant__ StridedArgs A)
{
	MARIN_PDL_WAIT();
	const Tables &t = A.t;
	const u32 col = blockIdx.x * L1_THREADS + threadIdx.x;
	if (col < A.stride)
	{
		const size_t base = (size_t)blockIdx.y * A.Lb + col;
		const u64 colscale = (u64)col * A.escale;
		const u32 K0 = A.rowK[0];
		pair_t v[4];
		#pragma unroll
		for (u32 q = 0; q < 4; ++q) v[q] = A.src[base + (size_t)q * A.stride];
		u64 ax[4], ay[4];
		if (INV)
		{
			#pragma unroll
			for (u32 q = 0; q < 4; ++q)
			{
				const u64 f = boundary_tw(t, (u64)(K0 + q) * colscale, true);
				ax[q] = gl::mul(v[q].x, f); ay[q] = gl::mul(v[q].y, f);
			}
			dft4<true>(ax[0], ax[1], ax[2], ax[3]); dft4<true>(ay[0], ay[1], ay[2], ay[3]);
		}
		else
		{
			#pragma unroll
			for (u32 q = 0; q < 4; ++q) { ax[q] = v[q].x; ay[q] = v[q].y; }
			dft4<false>(ax[0], ax[1], ax[2], ax[3]); dft4<false>(ay[0], ay[1], ay[2], ay[3]);
			#pragma unroll
			for (u32 k = 0; k < 4; ++k)
			{
				const u64 f = boundary_tw(t, (u64)(K0 + k) * colscale, false);
				ax[k] = gl::mul(ax[k], f); ay[k] = gl::mul(ay[k], f);
			}
		}
		#pragma unroll
		for (u32 k = 0; k < 4; ++k) { pair_t o; o.x = ax[k]; o.y = ay[k]; A.dst[base + (size_t)k * A.stride] = o; }
	}
	MARIN_PDL_TRIGGER();
}

// The same for eight rows (n = 5 * 2^25 = 2 x 1024 x 8 x 10240: the size of the reference's MM31 example): stage list [4 (s = 2), 2 (s = 1)],
// both stages in the registers of the thread that owns the column; the radix-2 stage carries the boundary twiddle.
template <bool INV> __global__ void __launch_bounds__(L1_THREADS) k_level1_r8(const __grid_constant__ StridedArgs A)
{
	MARIN_PDL_WAIT();
	const Tables &t = A.t;
	const u32 col = blockIdx.x * L1_THREADS + threadIdx.x;
	if (col < A.stride)
	{
		const size_t base = (size_t)blockIdx.y * A.Lb + col;
		const u64 colscale = (u64)col * A.escale;
		const u64 *tws = (INV ? A.twi : A.tw) + A.st[0].toff;      // tab[(k - 1) 2 + j] = w_8^(+-j k)
		u64 x[8], y[8];
		#pragma unroll
		for (u32 q = 0; q < 8; ++q) { const pair_t v = A.src[base + (size_t)q * A.stride]; x[q] = v.x; y[q] = v.y; }
		if (!INV)
		{
			#pragma unroll
			for (u32 j = 0; j < 2; ++j)
			{
				dft4<false>(x[j], x[j + 2], x[j + 4], x[j + 6]); dft4<false>(y[j], y[j + 2], y[j + 4], y[j + 6]);
				#pragma unroll
				for (u32 k = 1; k < 4; ++k) { const u64 w = tws[(k - 1) * 2 + j]; x[j + 2 * k] = gl::mul(x[j + 2 * k], w); y[j + 2 * k] = gl::mul(y[j + 2 * k], w); }
			}
			#pragma unroll
			for (u32 b = 0; b < 4; ++b)
			{
				addsub(x[2 * b], x[2 * b + 1]); addsub(y[2 * b], y[2 * b + 1]);
				const u32 K0 = A.rowK[2 * b];
				#pragma unroll
				for (u32 k = 0; k < 2; ++k)
				{
					const u64 f = boundary_tw(t, (u64)(K0 + 4 * k) * colscale, false);
					x[2 * b + k] = gl::mul(x[2 * b + k], f); y[2 * b + k] = gl::mul(y[2 * b + k], f);
				}
			}
		}
		else
		{
			#pragma unroll
			for (u32 b = 0; b < 4; ++b)
			{
				const u32 K0 = A.rowK[2 * b];
				#pragma unroll
				for (u32 k = 0; k < 2; ++k)
				{
					const u64 f = boundary_tw(t, (u64)(K0 + 4 * k) * colscale, true);
					x[2 * b + k] = gl::mul(x[2 * b + k], f); y[2 * b + k] = gl::mul(y[2 * b + k], f);
				}
				addsub(x[2 * b], x[2 * b + 1]); addsub(y[2 * b], y[2 * b + 1]);
			}
			#pragma unroll
			for (u32 j = 0; j < 2; ++j)
			{
				#pragma unroll
				for (u32 k = 1; k < 4; ++k) { const u64 w = tws[(k - 1) * 2 + j]; x[j + 2 * k] = gl::mul(x[j + 2 * k], w); y[j + 2 * k] = gl::mul(y[j + 2 * k], w); }
				dft4<true>(x[j], x[j + 2], x[j + 4], x[j + 6]); dft4<true>(y[j], y[j + 2], y[j + 4], y[j + 6]);
			}
		}
		#pragma unroll
		for (u32 k = 0; k < 8; ++k) { pair_t o; o.x = x[k]; o.y = y[k]; A.dst[base + (size_t)k * A.stride] = o; }
	}
	MARIN_PDL_TRIGGER();
}

// Sixteen rows (n = 5 * 2^26): one radix-16 stage; one lane of the pair column per thread.
template <bool INV> __global__ void __launch_bounds__(L1_THREADS) k_level1_r16(const __grid_constant__ StridedArgs A)
{
	MARIN_PDL_WAIT();
	const Tables &t = A.t;
	const u32 e = blockIdx.x * L1_THREADS + threadIdx.x, col = e >> 1, lane = e & 1;
	if (col < A.stride)
	{
		const size_t base = 2 * ((size_t)blockIdx.y * A.Lb + col) + lane, rs = 2 * (size_t)A.stride;
		const u64 *src = reinterpret_cast<const u64 *>(A.src); u64 *dst = reinterpret_cast<u64 *>(A.dst);
		const u64 colscale = (u64)col * A.escale;
		const u32 K0 = A.rowK[0];
		u64 a[16];
		if (INV)
		{
			#pragma unroll
			for (u32 q = 0; q < 16; ++q) a[sw16(q)] = gl::mul(src[base + q * rs], boundary_tw(t, (u64)(K0 + q) * colscale, true));
			dft16<true>(a);
			#pragma unroll
			for (u32 k = 0; k < 16; ++k) dst[base + k * rs] = a[k];
		}
		else
		{
			#pragma unroll
			for (u32 q = 0; q < 16; ++q) a[q] = src[base + q * rs];
			dft16<false>(a);
			#pragma unroll
			for (u32 k = 0; k < 16; ++k) dst[base + k * rs] = gl::mul(a[sw16(k)], boundary_tw(t, (u64)(K0 + k) * colscale, false));
		}
	}
	MARIN_PDL_TRIGGER();
}

}} // namespace marin::fast
#endif
