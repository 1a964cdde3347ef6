// marin_plan.hpp -- host-side planning for the B200 Marin engine: transform size, pass/stage
// decomposition, launch shapes and the small constant tables.  Pure C++ (no CUDA) so the kernel
// emulator in tests/emu shares it with the engine.
//
// Replaces ibdwt::transform_size / roots (include/marin/ibdwt.h:17-43,79-108) and the per-n kernel
// selection tables of engine_gpu_flat (include/marin/engine_gpu.h:108-129,551-691,1568-1630).
#pragma once
#include <cstdlib>
#include <string>
#include <vector>
#include <cstring>
#include "marin_kernels.cuh"

namespace marin {

// Smallest n in {2^k, 5 2^k}, k <= 26, with 2 (w + 1) + log2 n < 64, w = floor(p / n): the digits
// of a product must not overflow P.  Same rule (and the same 2.4 for log2 5) as ibdwt.h:17-43, so
// that residues, digit widths and checkpoints line up with the reference digit for digit.
inline size_t transform_size(uint32_t exponent)
{
	uint32_t w = 0, l2 = 1, l5 = 2;
	do { ++l2; w = exponent >> l2; } while ((w + 1) * 2 + l2 >= 64);
	do { ++l5; w = exponent / (5u << l5); } while ((w + 1) * 2 + (l5 + 2.4) >= 64);
	const size_t bad = (size_t)-1;
	const size_t n2 = (l2 <= 26) ? ((size_t)1 << l2) : bad, n5 = (l5 <= 26) ? ((size_t)5 << l5) : bad;
	return n2 < n5 ? n2 : n5;
}

static constexpr u32 SMALL_MAX_PAIRS = 1024;  // one CTA does the whole operation up to here; beyond, the grid-wide passes win
                                              // (B200, us per squaring: n = 4096 one CTA 36 / three passes 24, n = 10240 79 / 33, n = 16384 119 / 30)
static constexpr u32 MID_S = 4096;            // contiguous-pass block for large transforms (64 KiB + pad)
static constexpr u32 MID_S5 = 5120;           // 5 * 2^k transforms with 1024 rows: the radix-5 stage moves into the contiguous pass
                                              // (85 KiB rows), and the strided passes are the specialised R = 1024 kernels
static constexpr u32 MAX_S = 10240;           // longest row of the contiguous pass (2 x 5120 pairs: 170 KiB of shared memory, one CTA per SM)
static constexpr u32 MAX_R = 1280;            // strided tile rows: 1280 * 5 pairs (4 + 1 of padding) * 16 B = 100 KiB, two CTAs per SM

struct Plan
{
	u32 p = 0, n = 0, m = 0;
	bool small = false;
	u32 S = 0, R = 1;            // m = R * R2 * S (R2 = 1 unless a second strided pass is needed)
	u32 R2 = 1;
	u32 nstS = 0, nstR = 0, nstR2 = 0;
	Stage stS[MAX_STAGES], stR[MAX_STAGES], stR2[MAX_STAGES];
	u32 thrMid = 0, thrStr = 0, thrStr2 = 0;  // threads per CTA
	u32 D = 0;                   // digits per carry run of the linear / small kernels
	u32 wbase = 0;
	u32 a_max = 1;               // largest multiplier `a` the carry chain supports without overflow
};

#ifndef MARIN_SMALL_R4
#define MARIN_SMALL_R4 1024 // pairs: measured on B200, n = 512: 18.4 -> 14.3 us, n = 1024: 26.6 -> 17.4, n = 2048: 24.6 -> 22.6, n = 4096: 36.9 -> 41.8 (not used there)
#endif
#ifndef MARIN_R_SMALL_FIRST
#define MARIN_R_SMALL_FIRST 0
#endif

static_assert(MARIN_SMALL_R4 >= SMALL_MAX_PAIRS, "k_program is compiled without the radix-16 stage: single-CTA engines must get radix-4 lists");
inline u32 factor_stages(u32 len, Stage *st, bool sixteen_last, bool single_cta = false, bool force_r4 = false)
{
	// radices: optional 5 first, then 2, 4 and 16s (or 16s first) with product len
	u32 r[MAX_STAGES], k = 0, rest = len;
	if (rest % 5 == 0) { r[k++] = 5; rest /= 5; }
	u32 l2 = 0; while ((1u << l2) < rest) ++l2;
	u32 n16 = l2 / 4, rem = l2 % 4;
	// tiny transforms are pure instruction latency: a radix-16 butterfly is ~900 dependent-ish instructions in a
	// handful of threads, radix-4 stages are shorter chains on four times the threads
	if ((single_cta && len <= MARIN_SMALL_R4) || force_r4) { n16 = 0; rem = l2; }
	u32 small_r[MAX_STAGES], ns = 0;
	if (rem & 1) small_r[ns++] = 2;
	for (u32 i = 0; i < rem / 2; ++i) small_r[ns++] = 4;
	if (sixteen_last) { for (u32 i = 0; i < ns; ++i) r[k++] = small_r[i]; for (u32 i = 0; i < n16; ++i) r[k++] = 16; }
	else { for (u32 i = 0; i < n16; ++i) r[k++] = 16; for (u32 i = ns; i > 0; --i) r[k++] = small_r[i - 1]; }
	u32 L = len;
	u32 off = 0;
	for (u32 i = 0; i < k; ++i)
	{
		st[i].radix = r[i]; st[i].L = L; st[i].s = L / r[i]; st[i].toff = off;
		st[i].ls = 0; while ((1u << st[i].ls) < st[i].s) ++st[i].ls; // the radix-5 stage, if any, comes first: s is always a power of two
		if (st[i].s > 1) off += (r[i] - 1) * st[i].s;
		L /= r[i];
	}
	return k;
}

// Row length S of the contiguous pass for the transform sizes whose three-pass split was measured on a B200
// (tools/shape_sweep.py, profiles/r2/shape_sweep.log): the rows the specialised contiguous kernels cover
// (S = 256 x {4, 8, 16, 32} and 256 x {5, 10, 20, 40}) with 1024 strided rows from n = 2^21 / 5 * 2^19 up, so that both strided
// passes are the specialised R = 1024 kernels as well; 256 strided rows for the two sizes below.  0 = no entry.
inline u32 measured_row_length(u32 m)
{
	switch (m)
	{
		case 1u << 25: return 8192;        // n = 2^26       with 1024 x 4 rows (measured_second_level)
		case 1u << 24: return 4096;        // n = 2^25       with 1024 x 4 rows
		case 1u << 23: return 8192;        // n = 2^24       671 us (five passes, R = 64 x 32, S = 4096: 1084)
		case 1u << 22: return 4096;        // n = 2^23       300 us
		case 1u << 21: return 2048;        // n = 2^22       166 us (S = 4096, R = 512: 230)
		case 1u << 20: return 1024;        // n = 2^21        92 us (95)
		case 1u << 19: return 2048;        // n = 2^20        51 us (66)
		case 1u << 18: return 1024;        // n = 2^19        39 us (49)
		case 1u << 15: return 1024;        // n = 2^16        33 us (S = 512: 42)
		case 1u << 14: return 2048;        // n = 2^15        25 us (S = 256: 35)
		case 1u << 12: return 1024;        // n = 2^13        22 us (S = 512: 31)
		case 5u << 25: return 10240;       // n = 5 * 2^26   with 1024 x 16 rows
		case 5u << 24: return 10240;       // n = 5 * 2^25   with 1024 x 8 rows
		case 5u << 23: return 10240;       // n = 5 * 2^24   with 1024 x 4 rows
		case 5u << 22: return MID_S5;      // n = 5 * 2^23   with 1024 x 4 rows
		case 5u << 21: return 10240;       // n = 5 * 2^22   830 us (five passes, R = 80 x 32, S = 4096: 1702)
		case 5u << 20: return MID_S5;      // n = 5 * 2^21   432 us
		case 5u << 19: return 2560;        // n = 5 * 2^20   224 us (S = 4096, R = 640: 323)
		case 5u << 18: return 1280;        // n = 5 * 2^19   129 us (157)
		case 5u << 17: return 2560;        // n = 5 * 2^18    77 us (105)
		case 5u << 16: return 1280;        // n = 5 * 2^17    45 us (55)
		case 5u << 13: return 2048;        // n = 5 * 2^14    29 us (S = 512: 43)
		default: return 0;
	}
}

// Rows of the second strided pass for the sizes that need one and were measured: m = 1024 x 4 x S keeps the first level
// on the specialised strided kernels and leaves a single radix-4 stage for the second (a streaming kernel,
// marin_fast.cuh: k_level1_r4).  0 = the generic rule below.
inline u32 measured_second_level(u32 m)
{
	switch (m)
	{
		case 1u << 24: case 1u << 25: case 5u << 22: case 5u << 23: return 4;
		case 5u << 24: return 8;
		case 5u << 25: return 16;
		default: return 0;
	}
}

inline u32 round_threads(u32 t, u32 cap) { t = (t + 31) / 32 * 32; if (t < 32) t = 32; return t > cap ? cap : t; }

inline bool make_plan(u32 p, Plan &pl, std::string &err, u32 force_S = 0, u32 force_R2 = 0)
{
	if (p < 3) { err = "exponent must be >= 3"; return false; }
	const size_t n = transform_size(p);
	if (n == (size_t)-1) { err = "exponent too large for the Goldilocks IBDWT (n would exceed 5*2^26)"; return false; }
	pl.p = p; pl.n = (u32)n; pl.m = pl.n / 2; pl.wbase = p / pl.n;
	const u32 m = pl.m;
	if (m <= SMALL_MAX_PAIRS && !force_S)
	{
		pl.small = true; pl.S = m; pl.R = 1;
		pl.nstS = factor_stages(m, pl.stS, true, true);
		pl.nstR = 0;
		pl.thrMid = round_threads((m <= MARIN_SMALL_R4) ? m / 2 : m / 8, 512); // one butterfly per thread and lane in the widest stage
	}
	else
	{
		pl.small = false;
		const u32 five = (m % 5 == 0) ? 5 : 1;
		const u32 measured_S = force_S ? 0u : measured_row_length(m);
		if (force_S) { pl.S = force_S; pl.R = m / force_S; }
		else if (measured_S) { pl.S = measured_S; pl.R = m / measured_S; }
		else if (m / five >= ((five == 5) ? 16u : 64u) * MID_S) { pl.S = MID_S; pl.R = m / MID_S; }
		else if (five == 5) { pl.R = 80; pl.S = m / pl.R; }
		else { pl.R = (m > 8192) ? 64 : 8; pl.S = m / pl.R; } // measured at n = 4096 and n = 16384: few long rows win (latency, not throughput)
		if (pl.R * pl.S != m || pl.S % TILE_C != 0 || ((pl.S & (pl.S - 1)) != 0 && (pl.S % 5 != 0 || ((pl.S / 5) & (pl.S / 5 - 1)) != 0)) || pl.R < 2) { err = "bad forced split"; return false; }
		// more rows than one tile can hold: split them over two strided passes, R = RA * RB, the radix-5
		// factor (if any) staying in the first
		if (pl.R > MAX_R || force_R2)
		{
			const u32 Rall = pl.R, ramin = (five == 5) ? 80u : 64u;
			const u32 measured_rb = (force_R2 || force_S) ? 0u : measured_second_level(m);
			u32 rb = force_R2 ? force_R2 : (measured_rb ? measured_rb : 128);
			if (!force_R2 && !measured_rb) while (rb > 2 && Rall / rb < ramin) rb /= 2;
			if (rb < 2 || Rall % rb != 0 || (rb & (rb - 1)) != 0 || Rall / rb < 2) { err = "bad second-level split"; return false; }
			pl.R2 = rb; pl.R = Rall / rb;
		}
		if (pl.R > MAX_R || pl.S > MAX_S) { err = "transform size n = " + std::to_string(n) + " is outside the supported decompositions"; return false; }
		// Latency-bound sizes (profiles/r2/lowlat_sweep.log): a pass there is a handful of CTAs walking through dependent
		// phases, and a radix-16 butterfly is ~1100 instructions in one thread; radix-4 stages put four times the threads
		// on a quarter of the chain.  Strided passes up to n = 2^18 (n = 65536: 32.8 -> 26.6 us, n = 5 * 2^11: 33.3 -> 24.6),
		// contiguous passes up to n = 5 * 2^13 unless the row is one of the specialised shapes (n = 4096: 24.6 -> 20.5).
		const bool fastS = pl.S == 1024 || pl.S == 2048 || pl.S == 4096 || pl.S == 8192 || pl.S == 1280 || pl.S == 2560 || pl.S == MID_S5 || pl.S == 10240;
		const bool r4S = !force_S && !fastS && n <= 40960, r4R = !force_S && n <= 262144;
		pl.nstS = factor_stages(pl.S, pl.stS, true, false, r4S);
		pl.nstR = factor_stages(pl.R, pl.stR, MARIN_R_SMALL_FIRST != 0, false, r4R);
		pl.thrMid = round_threads(r4S ? pl.S / 2 : pl.S / 8, 512);
		const u32 bf = pl.R * TILE_C * 2 / (r4R ? 4 : 16); // butterflies x lanes per stage
		pl.thrStr = (bf <= 1024) ? round_threads(bf, 1024) : round_threads(bf / 2, 1024);
		if (pl.R2 > 1)
		{
			pl.nstR2 = factor_stages(pl.R2, pl.stR2, false);
			const u32 bf2 = pl.R2 * TILE_C * 2 / 16;
			pl.thrStr2 = round_threads(bf2 ? bf2 : 1, 1024);
		}
	}
	pl.D = (n % 16 == 0 && n >= 256) ? 16 : ((n % 8 == 0 && n >= 16) ? 8 : 4);
	// a * conv < 2^(64 + w) keeps the carry out of a run below 2^64, and the carry left after three
	// digits must stay below 2^(w - 1): a <= 2^min(61 - w - log2 n, 3 w - 3 - log2 n)
	u32 l2n = 0; while ((1ull << l2n) < n) ++l2n;
	const int w = (int)pl.wbase;
	int e = 61 - w - (int)l2n; const int e2 = 3 * w - 3 - (int)l2n; if (e2 < e) e = e2;
	if (e < 0) e = 0; if (e > 31) e = 31;
	pl.a_max = 1u << e;
	return true;
}

// frequency index of the element that ends at position pos after the DIF stages st[0..nst)
inline u32 freq_of_pos(const Stage *st, u32 nst, u32 pos)
{
	u32 K = 0, mult = 1, rem = pos;
	for (u32 i = 0; i < nst; ++i) { const u32 k = rem / st[i].s; rem %= st[i].s; K += k * mult; mult *= st[i].radix; }
	return K;
}

struct HostTables
{
	std::vector<u32> wbits, rowK, rowK2;
	std::vector<u64> twS, twSi, twR, twRi, twR2, twR2i, bl_lo, bl_hi, bl_loi, bl_hii, rowroot, colroot;
	u64 inv_m = 0, nr2 = 0, nr2_inv = 0, w5[5], w5i[5];
	u32 bl_h = 0;
};

inline void powers(std::vector<u64> &v, u64 root, size_t cnt)
{
	v.resize(cnt);
	u64 x = 1;
	for (size_t i = 0; i < cnt; ++i) { v[i] = x; x = gl::mul(x, root); }
}

// stage twiddles of a transform of length len, one block per stage with s > 1: block[(k - 1) s + j] = root^((len / L) j k)
inline void stage_twiddles(std::vector<u64> &v, const Stage *st, u32 nst, u32 len, u64 root)
{
	std::vector<u64> pw;
	powers(pw, root, len);
	v.clear();
	for (u32 i = 0; i < nst; ++i)
	{
		if (st[i].s <= 1) continue;
		if (v.size() != st[i].toff) v.resize(st[i].toff);
		const u64 scale = len / st[i].L;
		for (u32 k = 1; k < st[i].radix; ++k)
			for (u32 j = 0; j < st[i].s; ++j) v.push_back(pw[(size_t)((scale * j * k) % len)]);
	}
	if (v.empty()) v.push_back(1);
}

inline void build_tables(const Plan &pl, HostTables &h)
{
	const u32 n = pl.n, m = pl.m, p = pl.p;
	// digit widths: width[j] = ceil(p (j+1) / n) - ceil(p j / n)  (ibdwt.h:122-133), as one bit each
	h.wbits.assign((n + 31) / 32, 0);
	u32 prev = 0;
	for (u32 j = 1; j <= n; ++j)
	{
		const u64 qj = (u64)p * j;
		const u32 c = (u32)((qj - 1) / n + 1);
		if (c - prev == pl.wbase + 1) h.wbits[(j - 1) >> 5] |= 1u << ((j - 1) & 31);
		prev = c;
	}
	// n-th root of two: 554^((P-1)/192/n)  (ibdwt.h:116; 554^((P-1)/192) = 2)
	h.nr2 = gl::pow(554, (gl::P - 1) / 192 / n);
	h.nr2_inv = gl::inv(h.nr2);
	h.inv_m = gl::inv((u64)m);
	stage_twiddles(h.twS, pl.stS, pl.nstS, pl.S, gl::root_of_unity(pl.S));
	stage_twiddles(h.twSi, pl.stS, pl.nstS, pl.S, gl::inv(gl::root_of_unity(pl.S)));
	stage_twiddles(h.twR, pl.stR, pl.nstR, pl.R, gl::root_of_unity(pl.R));
	stage_twiddles(h.twRi, pl.stR, pl.nstR, pl.R, gl::inv(gl::root_of_unity(pl.R)));
	h.bl_h = 11;
	const u64 rm = gl::root_of_unity(m), rmi = gl::inv(rm);
	powers(h.bl_lo, rm, (size_t)1 << h.bl_h);
	powers(h.bl_hi, gl::pow(rm, (u64)1 << h.bl_h), (m >> h.bl_h) + 1);
	powers(h.bl_loi, rmi, (size_t)1 << h.bl_h);
	powers(h.bl_hii, gl::pow(rmi, (u64)1 << h.bl_h), (m >> h.bl_h) + 1);
	// frequency index of each row after the strided pass(es): block rA * R2 + rB holds K_A(rA) + R * K_B(rB)
	h.rowK.resize(pl.R); h.rowK2.resize(pl.R2); h.rowroot.resize((size_t)pl.R * pl.R2);
	for (u32 r = 0; r < pl.R; ++r) h.rowK[r] = pl.nstR ? freq_of_pos(pl.stR, pl.nstR, r) : 0;
	for (u32 r = 0; r < pl.R2; ++r) h.rowK2[r] = pl.nstR2 ? freq_of_pos(pl.stR2, pl.nstR2, r) : 0;
	for (u32 ra = 0; ra < pl.R; ++ra)
		for (u32 rb = 0; rb < pl.R2; ++rb)
			h.rowroot[(size_t)ra * pl.R2 + rb] = gl::pow(rm, (u64)h.rowK[ra] + (u64)pl.R * h.rowK2[rb]);
	if (pl.R2 > 1)
	{
		stage_twiddles(h.twR2, pl.stR2, pl.nstR2, pl.R2, gl::root_of_unity(pl.R2));
		stage_twiddles(h.twR2i, pl.stR2, pl.nstR2, pl.R2, gl::inv(gl::root_of_unity(pl.R2)));
	}
	h.colroot.resize(pl.S);
	const u64 rS = gl::root_of_unity(pl.S);
	for (u32 s = 0; s < pl.S; ++s) h.colroot[s] = gl::pow(rS, freq_of_pos(pl.stS, pl.nstS, s));
	// radix-5 constants
	const u64 w = gl::root_of_unity(5), w2 = gl::mul(w, w), w3 = gl::mul(w2, w), w4 = gl::mul(w3, w);
	const u64 c1 = gl::half(gl::add(w, w4)), c2 = gl::half(gl::add(w2, w3)), s1 = gl::half(gl::sub(w, w4)), s2 = gl::half(gl::sub(w2, w3));
	h.w5[0] = gl::half(gl::add(c1, c2)); h.w5[1] = gl::half(gl::sub(c1, c2));          // u = -1/4, v
	h.w5[2] = s1; h.w5[3] = gl::sub(s2, s1); h.w5[4] = gl::add(s1, s2);
	h.w5i[0] = h.w5[0]; h.w5i[1] = h.w5[1];
	for (int i = 2; i < 5; ++i) h.w5i[i] = gl::neg(h.w5[i]);
}

inline void fill_mid_args(const Plan &pl, MidArgs &A)
{
	A.S = pl.S; A.nst = pl.nstS;
	for (u32 i = 0; i < pl.nstS; ++i) A.st[i] = pl.stS[i];
	A.use_rowroot = pl.small ? 0 : 1;
	A.small = pl.small ? 1 : 0;
	A.D = pl.D; A.a = 1;
}

// level 0: the (first) strided pass over the whole register; level 1: the second strided pass, run
// independently on each of the R blocks of R2 * S pairs the first one leaves
inline void fill_strided_args(const Plan &pl, StridedArgs &A, int level = 0)
{
	if (level == 0)
	{
		A.R = pl.R; A.nst = pl.nstR;
		for (u32 i = 0; i < pl.nstR; ++i) A.st[i] = pl.stR[i];
		A.Lb = pl.m; A.stride = pl.R2 * pl.S; A.escale = 1;
	}
	else
	{
		A.R = pl.R2; A.nst = pl.nstR2;
		for (u32 i = 0; i < pl.nstR2; ++i) A.st[i] = pl.stR2[i];
		A.Lb = pl.R2 * pl.S; A.stride = pl.S; A.escale = pl.R;
	}
	A.tiles = A.stride / TILE_C; A.a = 1; A.with_carry = 0;
}

} // namespace marin
