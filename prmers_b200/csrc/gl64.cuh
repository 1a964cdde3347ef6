// gl64.cuh -- arithmetic modulo the Goldilocks prime P = 2^64 - 2^32 + 1 for sm_100a.
//
// Replaces the device field helpers of the reference (kernels/marin.cl:111-148, host twins
// include/marin/arith.h:24-72).  Every function takes and returns CANONICAL values (< P) unless
// its name says otherwise; the device versions are carry-chain PTX on 32-bit halves (IADD3 / IMAD.WIDE
// in SASS, no tensor cores: this is modular integer arithmetic), the host versions are portable C
// with identical results so that the kernels can be emulated on a CPU by the tests.
//
// Identities used: 2^64 = 2^32 - 1, 2^96 = -1, 2^192 = 1 (mod P); 2^48 is a 4th root of unity.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

typedef unsigned long long u64;
typedef unsigned int u32;

#if defined(__CUDACC__)
// 0, 1 and -1 that ptxas cannot fold: a madc with them stays an IMAD / IMAD.X, i.e. the add moves from the ALU
// pipe (IADD3, the bottleneck of this engine) to the FMA pipe, which otherwise idles.
__constant__ u32 GL_ONE = 1u, GL_MONE = 0xffffffffu, GL_ZERO = 0u;
#endif

namespace gl {

static constexpr u64 P = 0xffffffff00000001ull;
static constexpr u32 EPS = 0xffffffffu; // 2^64 mod P

HD u64 neg(u64 a) { return a ? P - a : 0; }

// ---- P - b for b <= P (result in [0, P]); used as the subtrahend that turns a subtraction into an addition
HD u64 pminus(u64 b)
{
#if defined(__CUDA_ARCH__)
	// one ALU + one FMA-pipe instruction
	u32 b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1;
	asm("sub.cc.u32 %0, 1, %2;\n\t"
		"madc.lo.u32 %1, %3, %4, 0xfffffffe;\n\t"  // 2^32 - 1 - b1 - borrow
		: "=&r"(r0), "=&r"(r1) : "r"(b0), "r"(b1), "r"(GL_MONE));
	return ((u64)r1 << 32) | r0;
#else
	return P - b;
#endif
}

// ---- a - b, canonical inputs, canonical output
HD u64 sub(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
	u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), d0, d1, m;
	// 3 ALU + 2 FMA-pipe instructions: the two carry consumers that produce no carry themselves are written as
	// madc, which ptxas keeps on the FMA pipe as IMAD.X (the integer ALU pipe is what bounds this engine; measured
	// with tools/micro/bfly_bench.cu: 35 -> 51 G butterflies/s with this, pminus and reduce128 below).  After a
	// sub.cc chain the flag that madc reads is the hardware carry (1 = no borrow), hence the 2^32 - 1 addends.
	asm("sub.cc.u32 %0, %3, %5;\n\t"
		"subc.cc.u32 %1, %4, %6;\n\t"
		"madc.lo.u32 %2, %7, %7, 0xffffffff;\n\t" // m = borrow ? 2^32 - 1 : 0
		"sub.cc.u32 %0, %0, %2;\n\t"              // d -= borrow ? 2^32 - 1 : 0   (i.e. d += P)
		"madc.lo.u32 %1, %1, %8, 0xffffffff;\n\t"
		: "=&r"(d0), "=&r"(d1), "=&r"(m) : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(GL_ZERO), "r"(GL_ONE));
	return ((u64)d1 << 32) | d0;
#else
	const u64 d = a - b;
	return (a < b) ? d - EPS : d;
#endif
}

// ---- a + b, canonical inputs, canonical output: a - (P - b), 4 ALU + 3 FMA-pipe instructions
HD u64 add(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
	return sub(a, pminus(b));
#else
	const u64 s = a + b;
	return (s < a || s >= P) ? s + EPS : s;
#endif
}

// ---- (hi, lo) mod P for ANY 128-bit value; canonical output.
// X = lo + h0 (2^32 - 1) - h1; add 2^32 - 1 first so that the sum never goes negative
// (Craig-Wood's trick, the one marin.cl:115-142 also relies on), then take it off again.
HD u64 reduce128(u64 lo, u64 hi)
{
#if defined(__CUDA_ARCH__)
	u32 l0 = (u32)lo, l1 = (u32)(lo >> 32), h0 = (u32)hi, h1 = (u32)(hi >> 32), r0, r1;
	// the five instructions that consume a carry without producing one go to the FMA pipe (IMAD / IMAD.X)
	asm("{\n\t"
		".reg .u32 d, s0, s1, nc, c;\n\t"
		"mad.lo.u32 d, %5, %7, 0xffffffff;\n\t"    // d = 2^32 - 1 - h1
		"sub.cc.u32 s0, d, %4;\n\t"
		"madc.lo.u32 s1, %4, %8, 0xffffffff;\n\t"  // s1 = h0 - borrow
		"add.cc.u32 %0, s0, %2;\n\t"
		"addc.cc.u32 %1, s1, %3;\n\t"
		"madc.lo.u32 nc, %6, %6, 0xffffffff;\n\t"  // nc = carry ? 0 : 2^32 - 1
		"sub.cc.u32 %0, %0, nc;\n\t"
		"subc.cc.u32 %1, %1, 0;\n\t"
		"madc.lo.u32 c, %6, %6, 0xffffffff;\n\t"   // c = borrow ? 2^32 - 1 : 0
		"sub.cc.u32 %0, %0, c;\n\t"
		"madc.lo.u32 %1, %1, %8, 0xffffffff;\n\t"
		"}"
		: "=&r"(r0), "=&r"(r1) : "r"(l0), "r"(l1), "r"(h0), "r"(h1), "r"(GL_ZERO), "r"(GL_MONE), "r"(GL_ONE));
	return ((u64)r1 << 32) | r0;
#else
	const u32 h0 = (u32)hi, h1 = (u32)(hi >> 32);
	const u64 s = (((u64)h0 << 32) | (u32)(EPS - h1)) - h0;
	u64 r = s + lo;
	if (r >= s)
	{
		const u64 c = (r < EPS) ? EPS : 0;
		r -= EPS; r -= c;
	}
	return r;
#endif
}

// ---- a * b mod P; inputs may be any u64 (lazy sums are fine), output canonical
HD u64 mul(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
	// 64 x 64 -> 128 as a 32-bit mad chain (IMAD.WIDE with carry in SASS; cheaper than mul.lo.u64 +
	// mul.hi.u64, which ptxas expands twice), then the reduction above.
	u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1, r2, r3;
	asm("mul.lo.u32 %0, %4, %6;\n\t"
		"mul.hi.u32 %1, %4, %6;\n\t"
		"mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
		"madc.hi.u32 %2, %4, %7, 0;\n\t"
		"mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
		"madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
		"addc.u32 %3, 0, 0;\n\t"
		"mad.lo.cc.u32 %2, %5, %7, %2;\n\t"
		"madc.hi.u32 %3, %5, %7, %3;\n\t"
		: "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
	return reduce128(((u64)r1 << 32) | r0, ((u64)r3 << 32) | r2);
#else
	const unsigned __int128 t = (unsigned __int128)a * b;
	return reduce128((u64)t, (u64)(t >> 64));
#endif
}

HD u64 sqr(u64 a) { return mul(a, a); }

// ---- a * 2^K mod P for a compile-time K in [0, 192); canonical in, canonical out
template <int K> HD u64 mul_pow2(u64 a)
{
	static_assert(K >= 0 && K < 192, "2 has order 192");
	if constexpr (K == 0) return a;
	else if constexpr (K >= 96) return neg(mul_pow2<K - 96>(a)); // 2^96 = -1
	else if constexpr (K < 64) return reduce128(a << K, a >> (64 - K));
	else
	{
		// 64 <= K < 96, J = K - 64: a 2^J = v2 2^64 + v1 2^32 + v0, and 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32:
		// a 2^K = v0 (2^32 - 1) - (v2 2^32 + v1)
		constexpr int J = K - 64;
		const u64 v0 = (u32)(a << J);
		const u64 hi = (J == 0) ? (a >> 32) : (a >> ((32 - J) & 63));   // < 2^(32 + J) <= 2^63 < P
		return sub((v0 << 32) - v0, hi);
	}
}

HD u64 pow(u64 a, u64 e)
{
	u64 r = 1;
	for (; e != 0; e >>= 1) { if (e & 1) r = mul(r, a); a = mul(a, a); }
	return r;
}

HD u64 inv(u64 a) { return pow(a, P - 2); }
HD u64 half(u64 a) { return (a & 1) ? (a >> 1) + ((P + 1) >> 1) : (a >> 1); }

// the transform's own generator: 7^21, chosen so that the 16th root of unity is 2^12
// (7 is the reference's generator, arith.h:72; any generator gives the same canonical results)
static constexpr u64 GENERATOR_BASE = 7, GENERATOR_EXP = 21;
HD u64 root_of_unity(u64 order) { return pow(pow(GENERATOR_BASE, GENERATOR_EXP), (P - 1) / order); }

} // namespace gl
