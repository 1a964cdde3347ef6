// marin_kernels.cuh -- device code of the B200 Marin engine (sm_100a), written from scratch.
//
// What it replaces, by function (not by name): the NTT pass kernels, pointwise square/multiply and
// carry kernels of kernels/marin.cl:475-2365, re-derived for Blackwell:
//   * a register is n weighted IBDWT digits (u64, natural order), viewed as m = n/2 pairs; the
//     transform is a length-m NTT applied lane-wise to pairs, the last radix-2 level being replaced
//     by the 2-term product modulo (t^2 - u)  (same algebra as marin.cl:378-392);
//   * m = R * S: one squaring is a strided forward pass (tiles of R rows x 4 pair columns), a
//     contiguous middle pass (one CTA per block of S = 4096 pairs: forward, pointwise, inverse) and a
//     strided inverse pass that also unweights, carries, hands carries from tile to tile, reweights and
//     (in a chain of squarings) already runs the next forward strided pass -- two or three launches
//     where the reference needs seven; transforms too long for one tile of rows get a second strided pass;
//   * each pass is a chain of radix-16 butterflies held in registers, one lane of the pairs per
//     thread, with shared-memory exchanges between stages;
//   * twiddles inside a radix-16 butterfly are shifts (2^12 is a 16th root of unity for our
//     generator), the ones between stages come from small tables.
// Everything here is __host__ __device__ and free of threadIdx/blockIdx so that tests/emu can run the
// very same phase functions on a CPU, thread by thread (test infrastructure; never a product path).
#pragma once
#include "gl64.cuh"

#if defined(__CUDACC__)
typedef ulonglong2 pair_t;
#else
struct alignas(16) pair_t { u64 x, y; };
#endif

// inter-CTA hand-off primitives (carry look-back in the strided inverse pass); the emulator runs the
// CTAs one after the other in order, so plain accesses are enough there.
#if defined(__CUDA_ARCH__)
#define MARIN_FENCE() __threadfence()
#define MARIN_LD_VOLATILE_U32(p) (*(volatile const unsigned int *)(p))
#define MARIN_ST_VOLATILE_U32(p, v) (*(volatile unsigned int *)(p) = (v))
#define MARIN_LD_CG_U64(p) __ldcg((const unsigned long long *)(p))
#define MARIN_NANOSLEEP(ns) __nanosleep(ns)
// 16-byte {carry, epoch} slots: one vector store publishes value and tag together (the single-word
// status+value of a decoupled look-back), one vector load reads both.  ASSUMPTION: an aligned 16-byte st.v2.u64 /
// ld.v2.u64 is performed as one access (not torn) on sm_100 -- the PTX memory model only promises that for scalar
// accesses, but the hardware moves aligned 16-byte vectors in one transaction and CUB's decoupled look-back relies on
// the same behaviour.  If that ever changes the parity tests catch it (a torn read would inject a stale carry), and the
// fix is the .b128 relaxed.gpu form or carry-then-epoch with st.release / ld.acquire.
static __device__ __forceinline__ void marin_st_slot(ulonglong2 *p, unsigned long long v, unsigned long long tag)
{ asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(v), "l"(tag) : "memory"); }
static __device__ __forceinline__ ulonglong2 marin_ld_slot(const ulonglong2 *p)
{ ulonglong2 r; asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory"); return r; }
#else
#define MARIN_FENCE() ((void)0)
#define MARIN_LD_VOLATILE_U32(p) (*(const unsigned int *)(p))
#define MARIN_ST_VOLATILE_U32(p, v) (*(unsigned int *)(p) = (v))
#define MARIN_LD_CG_U64(p) (*(const unsigned long long *)(p))
#define MARIN_NANOSLEEP(ns) ((void)0)
#endif

#if !defined(__CUDA_ARCH__)
static inline void marin_st_slot(pair_t *p, unsigned long long v, unsigned long long tag) { p->x = v; p->y = tag; }
static inline pair_t marin_ld_slot(const pair_t *p) { return *p; }
#endif

namespace marin {

static constexpr int MAX_STAGES = 8;
#ifndef MARIN_PAIR_BOUNDARY
#define MARIN_PAIR_BOUNDARY 1
#endif
#ifndef MARIN_RW_BATCH
#define MARIN_RW_BATCH 4 // weight loads in flight per thread in the reweight/store phase
#endif
#ifndef MARIN_TILE_C
#define MARIN_TILE_C 4
#endif
static constexpr int TILE_C = MARIN_TILE_C; // pair columns per strided tile: 4 * 16 B = two 32 B sectors per row; small
                                            // enough (R * 5 * 16 B) for two CTAs per SM, which overlaps load/compute/store phases
static constexpr int RUN = 2 * TILE_C;    // digits per carry run in the strided inverse pass
static constexpr int MAXRUN = 16;         // largest run any kernel carries in one thread

enum Mode : int { MODE_SQR = 0, MODE_MUL = 1, MODE_FWD = 2 };

struct Stage { u32 radix, L, s, toff, ls; }; // block length L = radix * s, s = 2^ls; toff: where this stage's twiddles start in the table

// Constant tables of one engine (device pointers).
struct Tables
{
	u32 p, n, m;                 // exponent, digits, pairs
	u32 wbase;                   // floor(p / n); digit j has wbase + bit_j(wbits) bits
	const u32 *wbits;            // n bits, natural order
	const u64 *w_nat, *winv_nat; // weight, 1/weight (natural digit order)
	const u64 *winvs_nat;        // 1/(weight m): unweighting fused with the 1/m of the inverse transform
	// the same two tables in the order the strided inverse pass walks them (multi-pass engines): pair (row, column)
	// of tile g at pair index (g R + row) TILE_C + column % TILE_C, so that a tile's weights are one contiguous
	// 16 R TILE_C byte stream instead of R segments of 16 TILE_C bytes one row stride apart
	const u64 *w_tile, *winvs_tile;
	u64 inv_m;                   // 1/m
	// block transforms
	const u64 *twS, *twSi; u32 S; // stage twiddles of the contiguous pass (see Tw), forward and inverse
	const u64 *bl_lo, *bl_hi, *bl_loi, *bl_hii; u32 bl_h; // m-th root: w^(+-e) = lo[e & (2^h - 1)] * hi[e >> h]
	const u64 *rowroot;          // w_m^(frequency index of the contiguous block) after the strided pass(es)
	const u64 *colroot;          // w_S^freq(sigma), position order inside a contiguous block
	u64 w5[5];                   // radix-5 constants (u, v, s1, s2 - s1, s1 + s2), see dft5
	u64 w5i[5];
};

// ------------------------------------------------------------------ programmatic dependent launch
// The transform kernels of an operation are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel
// touches nothing a predecessor wrote before griddepcontrol.wait returns (predecessor complete, its writes visible), and
// releases its successor with griddepcontrol.launch_dependents once its own main work is done, so that the successor's
// CTAs are resident and past their set-up when the last CTA of this grid exits.  The kernel boundary then costs neither
// the launch latency nor the ramp-up of the next grid: n = 4096: 20.5 -> 16.8 us per squaring, n = 65536: 26.7 -> 22.7,
// n = 2^20: 51.3 -> 48.2, n = 2^22: 168.5 -> 163.1 (profiles/r2/pdl_ab.log).  A successor launches only once every CTA of
// its predecessor is resident, so waiting CTAs never hold a slot the running grid needs.  Triggering at kernel entry
// instead (MARIN_PDL_MODE bits) was slower at the small sizes (n = 65536: 30 - 36 us) and no faster at the large ones.
// Rows of 4096 pairs (n >= 2^23) stay on plain stream order: with overlap the squaring took 309 instead of 296 us (the
// contiguous kernel wants 3 x 68 KB of shared memory per SM, the strided one 2 x 84 KB and the rest as L1: when the two
// grids share SMs the split cannot follow each kernel).
#ifndef MARIN_PDL_MODE
#define MARIN_PDL_MODE 0   // bit 0: contiguous-pass kernels trigger at entry (else after their main work), bit 1: strided-pass kernels
#endif
#if defined(__CUDA_ARCH__)
#define MARIN_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define MARIN_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#else
#define MARIN_PDL_TRIGGER() do { } while (0)
#define MARIN_PDL_WAIT() do { } while (0)
#endif

// ------------------------------------------------------------------ small helpers

HD u32 digit_width(const Tables &t, u32 j) { return t.wbase + ((t.wbits[j >> 5] >> (j & 31)) & 1u); }

HD u64 boundary_tw(const Tables &t, u64 e, bool inv)
{
	const u32 lo = (u32)e & ((1u << t.bl_h) - 1), hi = (u32)(e >> t.bl_h);
	return inv ? gl::mul(t.bl_loi[lo], t.bl_hii[hi]) : gl::mul(t.bl_lo[lo], t.bl_hi[hi]);
}

// ------------------------------------------------------------------ views

// Addressing of one FFT dimension for ONE lane of the pairs (lane 0 = even digit, 1 = odd digit):
// element i lives at p[2 * (i * mul + ((i >> 4) & pad))]; p already includes column and lane.
//   global / linear:           mul = pair stride, pad = 0
//   shared, contiguous pass:   mul = 1, pad = ~0   (one pair of padding per 16: conflict-free radix-16 stages)
//   shared, strided tile:      mul = TILE_PITCH (5 pairs per row of 4: conflict-free row-wise carry), pad = 0
// Adjacent threads take the two lanes of the same pair, so a warp touches 16 consecutive pairs.
static constexpr u32 TILE_PITCH = TILE_C + 1;
struct View
{
	u64 *p; u32 mul; u32 pad;
	HD size_t pidx(u32 i) const { return (size_t)i * mul + ((i >> 4) & pad); }
	HD u64 ld(u32 i) const { return p[2 * pidx(i)]; }
	HD void st(u32 i, u64 v) const { p[2 * pidx(i)] = v; }
	// elements base + s q, q < R, sit at constant distance when there is no padding, when s is a multiple
	// of 16, or when the whole butterfly lies inside one aligned group of 16 (s R <= 16, a power of two):
	// true for every hot stage, and it replaces a multiply-shift-mask-add per element by one add
	HD bool linear(u32 s, u32 R) const { const u32 L = s * R; return pad == 0 || (s & 15) == 0 || (L <= 16 && (L & (L - 1)) == 0); }
	HD size_t step(u32 s) const { return 2 * ((size_t)s * mul + ((s >> 4) & pad)); }
};
HD View view_global(const void *base, size_t pair_off, u32 stride, u32 lane) { return View{(u64 *)base + 2 * pair_off + lane, stride, 0}; }
HD View view_lin(void *sm, u32 lane) { return View{(u64 *)sm + lane, 1, 0xffffffffu}; }
HD View view_tile(void *sm, u32 col, u32 lane) { return View{(u64 *)sm + 2 * col + lane, TILE_PITCH, 0}; }

// whole pairs in the shared layouts (pointwise stage, carry, tile store)
HD size_t lin_pair(u32 i) { return i + (i >> 4); }
HD size_t tile_pair(u32 row, u32 col) { return (size_t)row * TILE_PITCH + col; }

HD u32 smem_pairs_lin(u32 S) { return S + (S >> 4) + 1; }

// ------------------------------------------------------------------ butterflies (scalar lanes)

HD void addsub(u64 &a, u64 &b) // (a, b) <- (a + b, a - b)
{
	const u64 nb = gl::pminus(b); // in (0, P]; P itself is a valid subtrahend for gl::sub
	const u64 d = gl::sub(a, b);
	a = gl::sub(a, nb);
	b = d;
}

// 4-point DFT, natural order in and out; the 4th root of unity is 2^48 (forward) / -2^48 (inverse).
// N1, N2, N3: input k holds the NEGATION of the true value (sign folded into the butterfly for free,
// which is how the negative shift twiddles -2^k of the 16-point transform cost nothing).
template <bool INV, bool N1 = false, bool N2 = false, bool N3 = false> HD void dft4(u64 &x0, u64 &x1, u64 &x2, u64 &x3)
{
	addsub(x0, x2);                  // (x0 + x2, x0 - x2)
	u64 t0 = N2 ? x2 : x0, t1 = N2 ? x0 : x2;
	addsub(x1, x3);                  // (S, D) = (x1 + x3, x1 - x3)
	u64 t2 = (N1 != N3) ? x3 : x1, t3 = (N1 != N3) ? x1 : x3; // |X1 + X3|, |X1 - X3|; both negated iff N1
	t3 = gl::mul_pow2<48>(t3);
	addsub(t0, t2);                  // t0 = y0 (or y2 if negated), t2 = y2 (or y0)
	addsub(t1, t3);                  // t1 + t3, t1 - t3
	x0 = N1 ? t2 : t0;
	x2 = N1 ? t0 : t2;
	const bool swap13 = (INV != N1);
	x1 = swap13 ? t3 : t1;           // forward: y1 = t1 + t3, y3 = t1 - t3; inverse: conjugate
	x3 = swap13 ? t1 : t3;
}

template <bool INV> HD void dft2(u64 &x0, u64 &x1) { addsub(x0, x1); }

// 16-point DFT.  Forward: natural slots in, output k in slot sw(k) = 4 (k & 3) + (k >> 2).
// Inverse: input y[k] in slot sw(k), natural slots out.  Unnormalised.  The 16th root of unity is 2^12,
// so the inner twiddles 2^(12 q k) are shifts; exponents >= 96 are -2^(e - 96) with the sign folded.
HD constexpr u32 sw16(u32 k) { return ((k & 3) << 2) | (k >> 2); }

template <bool INV> HD void dft16(u64 *x)
{
	if (!INV)
	{
		#pragma unroll
		for (int q = 0; q < 4; ++q) dft4<false>(x[q], x[q + 4], x[q + 8], x[q + 12]);
		// x[q + 4 k] *= 2^(12 q k)
		x[5] = gl::mul_pow2<12>(x[5]); x[9] = gl::mul_pow2<24>(x[9]); x[13] = gl::mul_pow2<36>(x[13]);
		x[6] = gl::mul_pow2<24>(x[6]); x[10] = gl::mul_pow2<48>(x[10]); x[14] = gl::mul_pow2<72>(x[14]);
		x[7] = gl::mul_pow2<36>(x[7]); x[11] = gl::mul_pow2<72>(x[11]); x[15] = gl::mul_pow2<12>(x[15]); // 2^108 = -2^12
		dft4<false>(x[0], x[1], x[2], x[3]);
		dft4<false>(x[4], x[5], x[6], x[7]);
		dft4<false>(x[8], x[9], x[10], x[11]);
		dft4<false, false, false, true>(x[12], x[13], x[14], x[15]);
	}
	else
	{
		#pragma unroll
		for (int k = 0; k < 4; ++k) dft4<true>(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
		// x[q + 4 k] *= 2^(-12 q k) = -2^(96 - 12 q k) (q k = 9: +2^84); signs folded below
		x[5] = gl::mul_pow2<84>(x[5]); x[9] = gl::mul_pow2<72>(x[9]); x[13] = gl::mul_pow2<60>(x[13]);
		x[6] = gl::mul_pow2<72>(x[6]); x[10] = gl::mul_pow2<48>(x[10]); x[14] = gl::mul_pow2<24>(x[14]);
		x[7] = gl::mul_pow2<60>(x[7]); x[11] = gl::mul_pow2<24>(x[11]); x[15] = gl::mul_pow2<84>(x[15]);
		dft4<true>(x[0], x[4], x[8], x[12]);
		dft4<true, true, true, true>(x[1], x[5], x[9], x[13]);
		dft4<true, true, true, true>(x[2], x[6], x[10], x[14]);
		dft4<true, true, true, false>(x[3], x[7], x[11], x[15]);
	}
}

// 5-point DFT (the only odd radix the Goldilocks IBDWT admits, ibdwt.h:17-22) with 5 multiplications (the reference's
// butterfly5 is a Winograd form too, marin.cl:349-376; this one is re-derived): with c1 = (w + w^4)/2, c2 = (w^2 + w^3)/2,
// s1 = (w - w^4)/2, s2 = (w^2 - w^3)/2 and a = x1 + x4, b = x1 - x4, e = x2 + x3, d = x2 - x3,
//   X1,4 = r1 +- i1, X2,3 = r2 +- i2,  r1,2 = x0 + u (a + e) +- v (a - e),  u = (c1 + c2)/2 = -1/4, v = (c1 - c2)/2,
//   i1 = s1 b + s2 d = P1 + P2, i2 = s2 b - s1 d = P3 - P1,  P1 = s1 (b + d), P2 = (s2 - s1) d, P3 = (s1 + s2) b.
// c = (u, v, s1, s2 - s1, s1 + s2); the inverse transform uses the same with the three s-constants negated.
HD void dft5(u64 *x, const u64 *c)
{
	const u64 a = gl::add(x[1], x[4]), b = gl::sub(x[1], x[4]);
	const u64 e = gl::add(x[2], x[3]), d = gl::sub(x[2], x[3]);
	const u64 t5 = gl::add(a, e);
	const u64 base = gl::add(x[0], gl::mul(t5, c[0])), vv = gl::mul(gl::sub(a, e), c[1]);
	const u64 r1 = gl::add(base, vv), r2 = gl::sub(base, vv);
	const u64 p1 = gl::mul(gl::add(b, d), c[2]), p2 = gl::mul(d, c[3]), p3 = gl::mul(b, c[4]);
	const u64 i1 = gl::add(p1, p2), i2 = gl::sub(p3, p1);
	x[0] = gl::add(x[0], t5);
	x[1] = gl::add(r1, i1); x[4] = gl::sub(r1, i1);
	x[2] = gl::add(r2, i2); x[3] = gl::sub(r2, i2);
}

// ------------------------------------------------------------------ one butterfly of one stage

// Twiddles of a stage with block length L = R s: tab[(k - 1) s + j] = w_L^(+-j k), 1 <= k < R, j < s -- one row per
// output index k, so that the threads of a warp (consecutive j) read consecutive words (coalesced; a flat table of
// powers indexed by j k made every load touch up to 30 sectors).  tab is the forward or the inverse table according
// to the direction; null when s = 1 (no twiddles).
struct Tw { const u64 *tab; u32 s; };

// Forward (DIF): y_k = sum_q x_q w_R^(qk), then y_k *= w_L^(jk); element q/k lives at base + s q/k.
// Inverse: the exact mirror (conjugate twiddle first, then conjugate DFT), unnormalised.
// PRE/POST: optional per-element factor applied to loaded inputs / stored outputs (boundary twiddles).
// Boundary: per-row factor w_m^(+-rowK[row] * colscale) applied to loaded inputs (inverse) or stored
// outputs (forward) of the stage that touches global memory in a strided pass.
// The stage that carries the boundary twiddle is the last forward stage (s = 1): its rows base + k have
// frequency indices K0 + k * Kstep, K0 = rowK[base], Kstep = R / radix.
struct Boundary { const Tables *t; u64 colscale; bool twiddle; const u64 *wpost; size_t wstride; u32 K0, Kstep; bool skip_wpost;
	const u64 *addp; size_t addstride; }; // addp: register added to the product before unweighting (mul_add), element (base + s q) at addp[(base + s q) addstride]

template <int R> HD u32 slot_of(u32 k) { return (R == 16) ? sw16(k) : k; }

template <int R, bool INV> HD void bfly_load(u64 *a, const View &in, u32 base, u32 s)
{
	if (in.linear(s, R))
	{
		const u64 *pi = in.p + 2 * in.pidx(base);
		const size_t di = in.step(s);
		#pragma unroll
		for (u32 q = 0; q < R; ++q) a[INV ? slot_of<R>(q) : q] = pi[q * di];
	}
	else
	{
		#pragma unroll
		for (u32 q = 0; q < R; ++q) a[INV ? slot_of<R>(q) : q] = in.ld(base + s * q);
	}
}

template <int R, bool INV> HD void bfly_core(u64 *a, const View &out, u32 base, u32 s, u32 j, const Tw &tw, const u64 *c5, const Boundary *bd);

template <int R, bool INV> HD void bfly(const View &in, const View &out, u32 base, u32 s, u32 j, const Tw &tw,
	const u64 *c5, const Boundary *bd)
{
	u64 a[R];
	bfly_load<R, INV>(a, in, base, s);
	bfly_core<R, INV>(a, out, base, s, j, tw, c5, bd);
}

template <int R, bool INV> HD void bfly_core(u64 *a, const View &out, u32 base, u32 s, u32 j, const Tw &tw, const u64 *c5, const Boundary *bd)
{
	if (INV && bd && bd->twiddle)
		#pragma unroll
		for (u32 q = 0; q < R; ++q)
		{
			const u64 f = boundary_tw(*bd->t, (u64)(bd->K0 + q * bd->Kstep) * bd->colscale, true);
			const u32 sl = slot_of<R>(q); a[sl] = gl::mul(a[sl], f);
		}

	if (INV && tw.tab)
		#pragma unroll
		for (u32 k = 1; k < R; ++k) { const u32 sl = slot_of<R>(k); a[sl] = gl::mul(a[sl], tw.tab[(k - 1) * tw.s + j]); }

	if (R == 16) dft16<INV>(a);
	else if (R == 4) dft4<INV>(a[0], a[1], a[2], a[3]);
	else if (R == 2) dft2<INV>(a[0], a[1]);
	else if (R == 5) dft5(a, c5);

	if (!INV && tw.tab)
		#pragma unroll
		for (u32 k = 1; k < R; ++k) { const u32 sl = slot_of<R>(k); a[sl] = gl::mul(a[sl], tw.tab[(k - 1) * tw.s + j]); }

	if (!INV && bd && bd->twiddle)
		#pragma unroll
		for (u32 k = 0; k < R; ++k)
		{
			const u64 f = boundary_tw(*bd->t, (u64)(bd->K0 + k * bd->Kstep) * bd->colscale, false);
			const u32 sl = slot_of<R>(k); a[sl] = gl::mul(a[sl], f);
		}

	// mul_add (carry_weight_muladd, marin.cl:2298-2365): the register to add joins the product while both are still
	// weighted -- times m, because the unweighting factor below carries the 1/m of the unnormalised transform
	if (INV && bd && bd->wpost && bd->addp)
		#pragma unroll
		for (u32 q = 0; q < R; ++q) a[q] = gl::add(a[q], gl::mul(bd->addp[(size_t)(base + s * q) * bd->addstride], (u64)bd->t->m));
	if (INV && bd && bd->wpost && !bd->skip_wpost) // unweight (and scale by 1/m) the outputs of the last inverse stage: natural order, coalesced
		#pragma unroll
		for (u32 q = 0; q < R; ++q) a[q] = gl::mul(a[q], bd->wpost[(size_t)(base + s * q) * bd->wstride]);

	if (out.linear(s, R))
	{
		u64 *po = out.p + 2 * out.pidx(base);
		const size_t dstep = out.step(s);
		#pragma unroll
		for (u32 k = 0; k < R; ++k) po[k * dstep] = a[INV ? k : slot_of<R>(k)];
	}
	else
	{
		#pragma unroll
		for (u32 k = 0; k < R; ++k) out.st(base + s * k, a[INV ? k : slot_of<R>(k)]);
	}
}

// ------------------------------------------------------------------ pointwise, modulo (t^2 - u)

HD pair_t sqr22(pair_t v, u64 u)
{
	pair_t r;
	const u64 ab = gl::mul(v.x, v.y);
	r.x = gl::add(gl::sqr(v.x), gl::mul(gl::sqr(v.y), u));
	r.y = gl::add(ab, ab);
	return r;
}

HD pair_t mul22(pair_t v, pair_t y, u64 u)
{
	pair_t r;
	r.x = gl::add(gl::mul(v.x, y.x), gl::mul(gl::mul(v.y, y.y), u));
	r.y = gl::add(gl::mul(v.x, y.y), gl::mul(v.y, y.x));
	return r;
}

// ------------------------------------------------------------------ carry

// Turn exact convolution values v[0..D) (integers < P) times `a` into digits of the given widths;
// returns the carry out of the run.  Restates adc / adc_mul (arith.h:75-91) on a run of digits.
HD u64 carry_run(u64 *v, u32 D, u32 wbase, u32 bits, u32 a, u64 cin)
{
	if (a == 1)
	{
		u64 c = cin;
		for (u32 i = 0; i < D; ++i)
		{
			const u32 w = wbase + ((bits >> i) & 1u);
			const u64 s = v[i] + c;
			const u64 co = (s < c) ? 1 : 0;
			c = (s >> w) | (co << (64 - w));
			v[i] = s & (((u64)1 << w) - 1);
		}
		return c;
	}
	unsigned __int128 c = cin;
	for (u32 i = 0; i < D; ++i)
	{
		const u32 w = wbase + ((bits >> i) & 1u);
		c += (unsigned __int128)v[i] * a;
		v[i] = (u64)c & (((u64)1 << w) - 1);
		c >>= w;
	}
	return (u64)c;
}

// Add an incoming carry to the first digits of a run; the last one absorbs what is left without
// renormalising (the reference's adc4 does the same, marin.cl:203-212).
HD void carry_absorb(u64 *d, u32 cnt, u32 wbase, u32 bits, u64 c)
{
	for (u32 i = 0; i + 1 < cnt; ++i)
	{
		const u32 w = wbase + ((bits >> i) & 1u);
		const u64 s = d[i] + c;
		d[i] = s & (((u64)1 << w) - 1);
		c = s >> w;
	}
	d[cnt - 1] += c;
}

HD u32 run_bits(const Tables &t, u32 j0, u32 D) { return (t.wbits[j0 >> 5] >> (j0 & 31)) & ((D >= 32) ? 0xffffffffu : ((1u << D) - 1)); }

// ------------------------------------------------------------------ kernel argument blocks

struct MidArgs     // contiguous pass: one CTA = one block of S pairs (or the whole register when small)
{
	Tables t;
	pair_t *x;          // register to transform (in place)
	const pair_t *y;    // multiplicand (MODE_MUL)
	pair_t *dst;        // output register (MODE_FWD), may equal x
	pair_t *dst2;       // single-CTA engines: also store the finished digits here (square_mul_copy / mul_copy), or null
	const pair_t *add_src; // single-CTA engines: register added to the product before the carry (mul_add, a = 1), or null
	u32 slab_pairs, slab_carry; // batched single-CTA engines (one ECM curve per CTA): CTA y works on the registers of slab y
	int mode;
	u32 S, nst;
	Stage st[MAX_STAGES];
	u32 use_rowroot;    // 0: colroot alone is the pointwise root (single-block engines)
	// fused carry for single-block engines (small n)
	u32 small; u32 a; u32 D;
	u64 *carry;
};

struct StridedArgs // strided pass: one CTA = R rows x TILE_C (4) pair-columns
{
	Tables t;
	const pair_t *src; pair_t *dst; // src == dst for in-place
	u32 R, nst;
	Stage st[MAX_STAGES];
	const u64 *tw, *twi; // stage twiddles of this pass (see Tw), forward and inverse
	const u32 *rowK;    // frequency index of row rho after this pass's stages
	u32 Lb;             // block length in pairs (= R * stride)
	u32 stride;         // pairs between rows
	u32 escale;         // m / Lb
	u32 tiles;          // stride / TILE_C tiles per block
	u32 a;              // multiplier for the carry
	pair_t *slots;      // slots[g * R + row] = {carry out of that run, epoch of the launch that wrote it}
	u32 with_carry;
	u32 *flags;         // flags[tiles]: error word (a hand-off that timed out)
	u32 epoch;
	pair_t *fwd_dst;    // inverse pass only: also run the next forward strided pass on each tile and write it here
	unsigned long long *ticket; // carry pass: tiles are handed out in START order (atomic counter, never reset), so the
	unsigned long long ticket_base; // tile a CTA waits for has always started before it, whatever the block scheduler does
	u32 skip_dst;       // carry pass of a squaring chain: leave the register alone, the parked transform in fwd_dst is the state
	pair_t *dst2;       // carry pass: second copy of the finished digits (square_mul_copy / mul_copy, marin.cl:1731-1769), or null
	const pair_t *add_src; // carry pass: register added to the product before the carry (mul_add, multiplier a = 1 only), or null
	u32 dbg;            // timing experiments only (builds with -DMARIN_DEBUG_SKIP): bit mask of steps to leave out
};
#if defined(MARIN_DEBUG_SKIP)
#define MARIN_DBG(A, bit) (((A).dbg >> (bit)) & 1u)
#else
#define MARIN_DBG(A, bit) 0
#endif

// ------------------------------------------------------------------ phase functions
// A kernel is a sequence of phases separated by block-wide barriers.  phase(ph, tid) does thread
// tid's share of phase ph.  GPU: `for ph: phase(ph, threadIdx.x); __syncthreads();`  Emulator:
// `for ph: for tid: phase(ph, tid)`.

template <int R, bool INV> HD void run_stage_lin(View in, View out, const Stage &sg, const Tw &tw, u32 S, u32 tid, u32 nthr, const Tables &t)
{
	const u32 nb = 2 * (S / R); // butterflies x lanes
	for (u32 b = tid; b < nb; b += nthr)
	{
		const u32 lane = b & 1, rb = b >> 1;
		const u32 blk = rb >> sg.ls, j = rb & (sg.s - 1); // s is a power of two in every stage
		View vi = in, vo = out;
		vi.p += lane; vo.p += lane;
		bfly<R, INV>(vi, vo, blk * sg.L + j, sg.s, j, tw, INV ? t.w5i : t.w5, nullptr);
	}
}

// NO16: the caller knows that the stage list holds no radix-16 stage (single-CTA engines: factor_stages gives them radix
// 4 / 2 / 5 lists, marin_plan.hpp) and does without that butterfly's code and registers
template <bool INV, bool NO16 = false> HD void dispatch_stage_lin(const View &in, const View &out, const Stage &sg, const Tw &tw, u32 S, u32 tid, u32 nthr, const Tables &t)
{
	switch (sg.radix)
	{
	case 16: if (!NO16) run_stage_lin<16, INV>(in, out, sg, tw, S, tid, nthr, t); break;
	case 4: run_stage_lin<4, INV>(in, out, sg, tw, S, tid, nthr, t); break;
	case 2: run_stage_lin<2, INV>(in, out, sg, tw, S, tid, nthr, t); break;
	case 5: run_stage_lin<5, INV>(in, out, sg, tw, S, tid, nthr, t); break;
	}
}

// ---- contiguous pass -------------------------------------------------------------------------
// phases: 0 .. nst-1 forward stages (the last has s = 1: no twiddles), nst pointwise on pairs,
// nst+1 .. 2 nst inverse stages, then for single-CTA engines: unweight, carry, absorb, re-weight + store.
// What differs from op to op on one engine, apart from the tables and stages of MidArgs: a recorded program (k_program)
// keeps ONE MidArgs in the kernel's parameter space (constant bank) and hands the registers of each op in here, instead
// of rewriting an argument block in shared memory per op and reading every field of it through LDS.
struct MidOp { pair_t *x; const pair_t *y; pair_t *dst; int mode; u32 a; };
HD MidOp mid_op_of(const MidArgs &A) { return MidOp{A.x, A.y, A.dst, A.mode, A.a}; }
HD u32 mid_num_phases(const MidArgs &A, int mode) { return (mode == MODE_FWD) ? A.nst : (2 * A.nst + 1 + (A.small ? 4 : 0)); }
HD u32 mid_num_phases(const MidArgs &A) { return mid_num_phases(A, A.mode); }

template <bool NO16 = false> HD void mid_phase(const MidArgs &A, const MidOp &o, u32 blk, u32 ph, u32 tid, u32 nthr, pair_t *sm, u64 *smc)
{
	const Tables &t = A.t;
	const u32 S = A.S, nst = A.nst;
	pair_t *xg = o.x + (size_t)blk * S;
	const View G = view_global(xg, 0, 1, 0), Sv = view_lin(sm, 0);
	if (ph < nst)
	{
		const Stage &sg = A.st[ph];
		const Tw tw{(sg.s > 1) ? t.twS + sg.toff : nullptr, sg.s};
		const View Gd = view_global(o.dst + (size_t)blk * S, 0, 1, 0);
		dispatch_stage_lin<false, NO16>((ph == 0) ? G : Sv, (o.mode == MODE_FWD && ph == nst - 1) ? Gd : Sv, sg, tw, S, tid, nthr, t);
		return;
	}
	if (ph == nst)
	{
		// pointwise square / multiply of the pairs modulo (t^2 - u), u = w_m^freq(position)
		const u64 rr = A.use_rowroot ? t.rowroot[blk] : 1;
		for (u32 i = tid; i < S; i += nthr)
		{
			const u64 cr = t.colroot[i];
			const u64 u = A.use_rowroot ? gl::mul(rr, cr) : cr;
			pair_t &v = sm[lin_pair(i)];
			v = (o.mode == MODE_SQR) ? sqr22(v, u) : mul22(v, o.y[(size_t)blk * S + i], u);
		}
		return;
	}
	if (ph <= 2 * nst)
	{
		const u32 si = 2 * nst - ph; // nst-1 .. 0
		const Stage &sg = A.st[si];
		const Tw tw{(sg.s > 1) ? t.twSi + sg.toff : nullptr, sg.s};
		dispatch_stage_lin<true, NO16>(Sv, (si == 0 && !A.small) ? G : Sv, sg, tw, S, tid, nthr, t);
		return;
	}
	// ---- small engines only: whole register is in shared memory (natural order).  Four phases: unweight (every thread,
	// a pair each), carry (one thread per run of D digits: additions only), absorb the carry of the run before (the same
	// thread, four digits), re-weight and store (every thread).  With the multiplications inside the per-run loops a
	// thread did 2 D of them one after the other while most of the CTA idled: at the ECM sizes that was a third of an op.
	const u32 D = A.D, nruns = t.n / D;
	if (ph == 2 * nst + 1)
	{
		for (u32 i = tid; i < S; i += nthr)
		{
			pair_t pv = sm[lin_pair(i)];
			pv.x = gl::mul(pv.x, t.winvs_nat[2 * i]); pv.y = gl::mul(pv.y, t.winvs_nat[2 * i + 1]);
			if (A.add_src)
			{
				const pair_t av = A.add_src[i];
				pv.x = gl::add(pv.x, gl::mul(av.x, t.winv_nat[2 * i])); pv.y = gl::add(pv.y, gl::mul(av.y, t.winv_nat[2 * i + 1]));
			}
			sm[lin_pair(i)] = pv;
		}
		return;
	}
	if (ph == 2 * nst + 2)
	{
		for (u32 r = tid; r < nruns; r += nthr)
		{
			u64 v[MAXRUN];
			const u32 j0 = r * D;
			for (u32 i = 0; i < D; i += 2) { const pair_t pv = sm[lin_pair((j0 + i) >> 1)]; v[i] = pv.x; v[i + 1] = pv.y; }
			smc[r] = carry_run(v, D, t.wbase, run_bits(t, j0, D), o.a, 0);
			for (u32 i = 0; i < D; i += 2) { pair_t pv; pv.x = v[i]; pv.y = v[i + 1]; sm[lin_pair((j0 + i) >> 1)] = pv; }
		}
		return;
	}
	if (ph == 2 * nst + 3)
	{
		for (u32 r = tid; r < nruns; r += nthr)
		{
			u64 v[4];
			const u32 j0 = r * D, na = (D < 4) ? D : 4;
			for (u32 i = 0; i < na; i += 2) { const pair_t pv = sm[lin_pair((j0 + i) >> 1)]; v[i] = pv.x; v[i + 1] = pv.y; }
			carry_absorb(v, na, t.wbase, run_bits(t, j0, D), smc[(r == 0) ? nruns - 1 : r - 1]);
			for (u32 i = 0; i < na; i += 2) { pair_t pv; pv.x = v[i]; pv.y = v[i + 1]; sm[lin_pair((j0 + i) >> 1)] = pv; }
		}
		return;
	}
	for (u32 i = tid; i < S; i += nthr)
	{
		pair_t pv = sm[lin_pair(i)];
		pv.x = gl::mul(pv.x, t.w_nat[2 * i]); pv.y = gl::mul(pv.y, t.w_nat[2 * i + 1]);
		xg[i] = pv;
		if (A.dst2) A.dst2[i] = pv;
	}
}

HD void mid_phase(const MidArgs &A, u32 blk, u32 ph, u32 tid, u32 nthr, pair_t *sm, u64 *smc) { mid_phase<false>(A, mid_op_of(A), blk, ph, tid, nthr, sm, smc); }

// ---- strided passes ----------------------------------------------------------------------------
// Tile = R rows x TILE_C (4) pair columns (64-byte row segments); butterfly id b -> lane b & 1, column (b >> 1) & 3, row-butterfly b >> 3.

template <int R, bool INV> HD void run_stage_tile(const StridedArgs &A, const Stage &sg, bool first, bool last, size_t blockbase /*pairs*/, u32 c0,
	pair_t *sm, u32 tid, u32 nthr, const pair_t *gsrc, pair_t *gdst, bool in_smem_always, bool out_smem_always)
{
	const Tables &t = A.t;
	// The radix-4 stage next to global memory (s = 1: boundary twiddles, no stage twiddles): one thread takes BOTH lanes
	// of its four pairs, so the boundary twiddle (a product of two table entries) is formed once per pair instead of
	// once per lane, and global memory is touched 16 bytes at a time (405 -> 386 us per squaring at n = 2^23).  The
	// same on a radix-16 stage needs 64 registers of data alone and spills: 579 -> 701 us at n = 5 * 2^21, not done.
	if (R == 4 && MARIN_PAIR_BOUNDARY && last && !first && sg.s == 1)
	{
		const u32 npb = (A.R / R) * TILE_C, Kstep = A.R / R;
		for (u32 b = tid; b < npb; b += nthr)
		{
			const u32 c = b & (TILE_C - 1), base = (b / TILE_C) * R;
			const size_t gcol = blockbase + c0 + c;
			const u64 colscale = (u64)(c0 + c) * A.escale;
			const u32 K0 = A.rowK[base];
			u64 ax[R], ay[R];
			if (INV)
			{
				pair_t v[R];
				#pragma unroll
				for (u32 q = 0; q < R; ++q) v[q] = gsrc[gcol + (size_t)(base + q) * A.stride];
				#pragma unroll
				for (u32 q = 0; q < R; ++q)
				{
					const u64 f = boundary_tw(t, (u64)(K0 + q * Kstep) * colscale, true);
					const u32 sl = slot_of<R>(q);
					ax[sl] = gl::mul(v[q].x, f); ay[sl] = gl::mul(v[q].y, f);
				}
				dft4<true>(ax[0], ax[1], ax[2], ax[3]); dft4<true>(ay[0], ay[1], ay[2], ay[3]);
				#pragma unroll
				for (u32 q = 0; q < R; ++q) { pair_t o; o.x = ax[q]; o.y = ay[q]; sm[tile_pair(base + q, c)] = o; }
			}
			else
			{
				#pragma unroll
				for (u32 q = 0; q < R; ++q) { const pair_t v = sm[tile_pair(base + q, c)]; ax[q] = v.x; ay[q] = v.y; }
				dft4<false>(ax[0], ax[1], ax[2], ax[3]); dft4<false>(ay[0], ay[1], ay[2], ay[3]);
				#pragma unroll
				for (u32 k = 0; k < R; ++k)
				{
					const u64 f = boundary_tw(t, (u64)(K0 + k * Kstep) * colscale, false);
					const u32 sl = slot_of<R>(k);
					pair_t o; o.x = gl::mul(ax[sl], f); o.y = gl::mul(ay[sl], f);
					gdst[gcol + (size_t)(base + k) * A.stride] = o;
				}
			}
		}
		return;
	}
	const u32 nb = (A.R / R) * TILE_C * 2;
	const Tw tw{(sg.s > 1) ? (INV ? A.twi : A.tw) + sg.toff : nullptr, sg.s};
	for (u32 b = tid; b < nb; b += nthr)
	{
		const u32 lane = b & 1, c = (b >> 1) & (TILE_C - 1), rb = b / (2 * TILE_C);
		const u32 blk = rb >> sg.ls, j = rb & (sg.s - 1); // s is a power of two in every stage
		const u32 base = blk * sg.L + j;
		const View Gi = view_global(gsrc, blockbase + c0 + c, A.stride, lane);
		const View Go = view_global(gdst, blockbase + c0 + c, A.stride, lane);
		const View Sv = view_tile(sm, c, lane);
		// the last inverse stage of a carry pass also unweights: digit index 2 (row stride + c0 + c) + lane
		const bool wp = INV && first && out_smem_always;
		const Boundary bd{&t, (u64)(c0 + c) * A.escale, last, wp ? t.winvs_tile + 2 * ((size_t)(c0 / TILE_C) * A.R * TILE_C + c) + lane : nullptr, 2 * (size_t)TILE_C,
			last ? A.rowK[base] : 0u, A.R / R, MARIN_DBG(A, 3) != 0,
			(wp && A.add_src) ? reinterpret_cast<const u64 *>(A.add_src) + 2 * (blockbase + c0 + c) + lane : nullptr, 2 * (size_t)A.stride};
		if (!INV)
			bfly<R, false>((first && !in_smem_always) ? Gi : Sv, last ? Go : Sv, base, sg.s, j, tw, t.w5, last ? &bd : nullptr);
		else // inverse order: the stage that is `last` in forward order comes first and reads global
			bfly<R, true>(last ? Gi : Sv, (first && !out_smem_always) ? Go : Sv, base, sg.s, j, tw, t.w5i, (last || wp) ? &bd : nullptr);
	}
}

template <bool INV> HD void dispatch_stage_tile(const StridedArgs &A, const Stage &sg, bool first, bool last, size_t blockbase, u32 c0,
	pair_t *sm, u32 tid, u32 nthr, const pair_t *gsrc, pair_t *gdst, bool in_smem_always, bool out_smem_always)
{
	switch (sg.radix)
	{
	case 16: run_stage_tile<16, INV>(A, sg, first, last, blockbase, c0, sm, tid, nthr, gsrc, gdst, in_smem_always, out_smem_always); break;
	case 4: run_stage_tile<4, INV>(A, sg, first, last, blockbase, c0, sm, tid, nthr, gsrc, gdst, in_smem_always, out_smem_always); break;
	case 2: run_stage_tile<2, INV>(A, sg, first, last, blockbase, c0, sm, tid, nthr, gsrc, gdst, in_smem_always, out_smem_always); break;
	case 5: run_stage_tile<5, INV>(A, sg, first, last, blockbase, c0, sm, tid, nthr, gsrc, gdst, in_smem_always, out_smem_always); break;
	}
}

// Phases of the inverse pass with carry (nst = number of stages):
//   0 .. nst-1   inverse stages; the last one unweights (x 1/(w m)) into shared memory
//   nst          carry each row's run of 2*TILE_C digits and publish the carry out as a {carry, epoch} slot (one
//                16-byte store: value and tag arrive together, no fence, no separate flag); thread 0 then waits for
//                the first slot it will need, so that a tile whose predecessor is far behind (tile 0 needs the LAST
//                tile: end-around, 2^p = 1) parks one thread instead of 256
//   nst + 1      absorb the carry of the run before into the first digits (each thread checks its own slot's tag)
//   nst + 2      reweight, store the digits to the register (coalesced); keep the tile if the forward pass follows
//   nst + 3 ..   optional: forward stages of the NEXT squaring on the tile still in shared memory -> A.fwd_dst
// One-hop look-back only: a tile publishes before it waits, and tile numbers are handed out in start order, so the
// tile it waits for is already running.  Tile 0 parks on the last tile, which costs one CTA slot out of 148 x 2.
HD u32 strided_num_phases(const StridedArgs &A, bool inv)
{
	if (!inv || !A.with_carry) return A.nst;
	return A.nst + 3 + (A.fwd_dst ? A.nst : 0);
}
HD u32 strided_publish_phase(const StridedArgs &A) { return A.nst; } // every slot of the tile is published when this phase ends

static constexpr u32 SPIN_LIMIT = 1u << 27; // bounded wait: a lost flag raises the error word instead of hanging the GPU

// carry published in slot idx by this launch; bounded wait (an error word and a poisoned carry instead of a hung GPU)
HD u64 wait_slot(const StridedArgs &A, size_t idx, bool backoff)
{
	u32 spins = 0;
	for (;;)
	{
		const pair_t v = marin_ld_slot(&A.slots[idx]);
		if (v.y == (u64)A.epoch) return v.x;
		if (++spins > SPIN_LIMIT) { MARIN_ST_VOLATILE_U32(&A.flags[A.tiles], 1u); return ~0ull; } // poisoned carry + sticky error word: every later read-back throws
		if (backoff) MARIN_NANOSLEEP(200);
	}
}

// the run before (row, g) is (row, g - 1); before (row, 0) comes (row - 1, last tile); before (0, 0) the very last run
HD size_t slot_before(const StridedArgs &A, u32 g, u32 row)
{
	return (g > 0) ? (size_t)(g - 1) * A.R + row : (size_t)(A.tiles - 1) * A.R + ((row == 0) ? A.R - 1 : row - 1);
}

template <bool INV> HD void strided_phase(const StridedArgs &A, u32 blkid, u32 ph, u32 tid, u32 nthr, pair_t *sm)
{
	const Tables &t = A.t;
	const u32 ob = blkid / A.tiles, g = blkid % A.tiles; // outer block, tile
	const size_t blockbase = (size_t)ob * A.Lb; const u32 c0 = g * TILE_C;
	if (ph < A.nst)
	{
		const u32 si = INV ? (A.nst - 1 - ph) : ph;
		dispatch_stage_tile<INV>(A, A.st[si], si == 0, si == A.nst - 1, blockbase, c0, sm, tid, nthr, A.src, A.dst, false, A.with_carry != 0);
		return;
	}
	if (!INV) return;
	const u32 cph = ph - A.nst;
	if (cph == 0)
	{
		// the tile holds exact convolution values: carry each row's run of consecutive digits
		if (MARIN_DBG(A, 6)) return;
		for (u32 row = tid; row < A.R; row += nthr)
		{
			u64 v[RUN];
			#pragma unroll
			for (u32 c = 0; c < TILE_C; ++c) { const pair_t pv = sm[tile_pair(row, c)]; v[2 * c] = pv.x; v[2 * c + 1] = pv.y; }
			const u32 j0 = 2 * (row * A.stride + c0);
			const u64 cout = MARIN_DBG(A, 4) ? 0 : carry_run(v, RUN, t.wbase, run_bits(t, j0, RUN), A.a, 0);
			if (!MARIN_DBG(A, 7)) marin_st_slot(&A.slots[(size_t)g * A.R + row], cout, A.epoch);
			#pragma unroll
			for (u32 c = 0; c < TILE_C; ++c) { pair_t pv; pv.x = v[2 * c]; pv.y = v[2 * c + 1]; sm[tile_pair(row, c)] = pv; }
		}
#if defined(__CUDA_ARCH__)
		if (tid == 0 && !MARIN_DBG(A, 1)) wait_slot(A, slot_before(A, g, 0), true); // a gate for performance only
#endif
		return;
	}
	if (cph == 1)
	{
		if (MARIN_DBG(A, 8)) return;
		for (u32 row = tid; row < A.R; row += nthr)
		{
			const u64 c = wait_slot(A, slot_before(A, g, row), false);
			if (c == 0) continue;
			const pair_t p0 = sm[tile_pair(row, 0)], p1 = sm[tile_pair(row, 1)];
			u64 d[4] = {p0.x, p0.y, p1.x, p1.y};
			carry_absorb(d, 4, t.wbase, run_bits(t, 2 * (row * A.stride + c0), 4), c);
			pair_t q0, q1; q0.x = d[0]; q0.y = d[1]; q1.x = d[2]; q1.y = d[3];
			sm[tile_pair(row, 0)] = q0; sm[tile_pair(row, 1)] = q1;
		}
		return;
	}
	if (cph == 2)
	{
		// reweight and store, coalesced (TILE_C threads cover one row segment); eight elements in flight per thread so
		// that the weight loads (a stream no cache holds) overlap instead of costing one memory latency each
		constexpr u32 B = MARIN_RW_BATCH;
		const u32 total = A.R * TILE_C;
		if (MARIN_DBG(A, 5)) return;
		for (u32 e0 = tid; e0 < total; e0 += B * nthr)
		{
			pair_t w[B];
			#pragma unroll
			for (u32 u = 0; u < B; ++u)
			{
				const u32 e = e0 + u * nthr;
				if (e < total) w[u] = reinterpret_cast<const pair_t *>(t.w_tile)[(size_t)g * total + e];
			}
			#pragma unroll
			for (u32 u = 0; u < B; ++u)
			{
				const u32 e = e0 + u * nthr;
				if (e < total)
				{
					const u32 row = e / TILE_C, c = e % TILE_C;
					const pair_t d = sm[tile_pair(row, c)];
					pair_t o; o.x = gl::mul(d.x, w[u].x); o.y = gl::mul(d.y, w[u].y);
					if (MARIN_DBG(A, 2)) o = d;
					if (!A.skip_dst) A.dst[blockbase + (size_t)row * A.stride + c0 + c] = o;
					if (A.dst2) A.dst2[blockbase + (size_t)row * A.stride + c0 + c] = o;
					if (A.fwd_dst) sm[tile_pair(row, c)] = o;
				}
			}
		}
		return;
	}
	// forward strided pass of the next squaring, straight from the tile in shared memory
	const u32 si = cph - 3;
	dispatch_stage_tile<false>(A, A.st[si], si == 0, si == A.nst - 1, blockbase, c0, sm, tid, nthr, nullptr, A.fwd_dst, true, false);
}

// ---- linear operations on digits (natural order, runs of D digits) -----------------------------
enum LinOp : int { LIN_ADD = 0, LIN_SUBREG = 1, LIN_SUBSMALL = 2, LIN_SETDIGITS = 3, LIN_SETSMALL = 4, LIN_ADDSUB = 5 };

struct LinArgs
{
	Tables t;
	u64 *dst; const u64 *src; // registers (weighted digits); src unused for SUBSMALL / SETSMALL
	const u64 *digits;        // LIN_SETDIGITS: encoded digits (low 32 bits used)
	u64 *carry;               // one per run
	int op; u32 a; u32 D;
	// LIN_ADDSUB (one pass for S = A + B and D = A - B, marin.cl:1891-2154): dst = S, dst2 = D, inputs srcA and src (= B);
	// cpy / cpy2: optional copies of S and D (addsub_copy); carry2: the D carries.  Any output may alias any input.
	const u64 *srcA; u64 *dst2, *cpy, *cpy2, *carry2;
	u32 slab_words, slab_carry; // batched engines: CTA y of the grid works on slab y (0 = not batched)
};

// digits of the small constant `a` that fall in digit j (a < 2^32)
HD u64 small_digit(const Tables &t, u32 j, u32 a)
{
	u32 pos = 0;
	for (u32 i = 0; i < j; ++i) { pos += digit_width(t, i); if (pos >= 32) return 0; }
	const u32 w = digit_width(t, j);
	return ((u64)a >> pos) & (((u64)1 << w) - 1);
}

HD void lin_p1_thread(const LinArgs &A, u32 r)
{
	const Tables &t = A.t;
	const u32 D = A.D, nruns = t.n / D;
	if (r >= nruns) return;
	const u32 j0 = r * D;
	const u32 bits = run_bits(t, j0, D);
	u64 v[MAXRUN];
	for (u32 i = 0; i < D; ++i)
	{
		const u32 j = j0 + i;
		const u32 w = t.wbase + ((bits >> i) & 1u);
		const u64 mask = ((u64)1 << w) - 1;
		switch (A.op)
		{
		case LIN_ADD: v[i] = gl::mul(gl::add(A.dst[j], A.src[j]), t.winv_nat[j]); break;
		// y - x + 2 (2^w - 1) per digit keeps every digit non-negative (marin.cl:244-257,1876)
		case LIN_SUBREG: v[i] = gl::mul(gl::sub(A.dst[j], A.src[j]), t.winv_nat[j]); v[i] = gl::add(v[i], 2 * mask); break;
		// x - a == x + (2^p - 1 - a): digit-wise mask - a_j, never negative
		case LIN_SUBSMALL: v[i] = gl::mul(A.dst[j], t.winv_nat[j]) + (mask - ((j < 64) ? small_digit(t, j, A.a) : 0)); break;
		case LIN_SETDIGITS: v[i] = (u32)A.digits[j]; break;
		default: v[i] = (j < 64) ? small_digit(t, j, A.a) : 0; break;
		}
	}
	A.carry[r] = carry_run(v, D, t.wbase, bits, 1, 0);
	for (u32 i = 0; i < D; ++i) A.dst[j0 + i] = gl::mul(v[i], t.w_nat[j0 + i]);
}

HD void lin_p2_thread(const LinArgs &A, u32 r)
{
	const Tables &t = A.t;
	const u32 D = A.D, nruns = t.n / D;
	if (r >= nruns) return;
	const u64 c = A.carry[(r == 0) ? nruns - 1 : r - 1];
	if (c == 0) return;
	const u32 j0 = r * D, cnt = (D < 4) ? D : 4;
	u64 d[4];
	for (u32 i = 0; i < cnt; ++i) d[i] = gl::mul(A.dst[j0 + i], t.winv_nat[j0 + i]);
	carry_absorb(d, cnt, t.wbase, run_bits(t, j0, D), c);
	for (u32 i = 0; i < cnt; ++i) A.dst[j0 + i] = gl::mul(d[i], t.w_nat[j0 + i]);
}

// ---- fast path for add / sub_reg / sub(small): one thread per PAIR, fully coalesced, single-hop carry.
// The sums are tiny (a few units above 2^w), so each digit takes the carry of the digit before it and
// stops there; what is left keeps digits within a few units of their width, which is the same
// semi-carried state the reference leaves (adc4 / addc4, marin.cl:203-212,268-279).
HD void lin_pair_compute(const LinArgs &A, u32 i, u64 &u0, u64 &u1, u32 &w0, u32 &w1)
{
	const Tables &t = A.t;
	const u32 j = 2 * i;
	w0 = digit_width(t, j); w1 = digit_width(t, j + 1);
	const pair_t x = reinterpret_cast<const pair_t *>(A.dst)[i], wi = reinterpret_cast<const pair_t *>(t.winv_nat)[i];
	const u64 m0 = ((u64)1 << w0) - 1, m1 = ((u64)1 << w1) - 1;
	if (A.op == LIN_ADD)
	{
		const pair_t y = reinterpret_cast<const pair_t *>(A.src)[i];
		u0 = gl::mul(gl::add(x.x, y.x), wi.x); u1 = gl::mul(gl::add(x.y, y.y), wi.y);
	}
	else if (A.op == LIN_SUBREG)
	{
		const pair_t y = reinterpret_cast<const pair_t *>(A.src)[i];
		u0 = gl::add(gl::mul(gl::sub(x.x, y.x), wi.x), 2 * m0); u1 = gl::add(gl::mul(gl::sub(x.y, y.y), wi.y), 2 * m1);
	}
	else // LIN_SUBSMALL
	{
		u0 = gl::mul(x.x, wi.x) + (m0 - ((j < 64) ? small_digit(t, j, A.a) : 0));
		u1 = gl::mul(x.y, wi.y) + (m1 - ((j + 1 < 64) ? small_digit(t, j + 1, A.a) : 0));
	}
}

HD void lin_pair_finish(const LinArgs &A, u32 i, u64 u0, u64 u1, u32 w0, u32 w1, u64 cin)
{
	const pair_t w = reinterpret_cast<const pair_t *>(A.t.w_nat)[i];
	const u64 v0 = (u0 & (((u64)1 << w0) - 1)) + cin;
	const u64 v1 = (u1 & (((u64)1 << w1) - 1)) + (u0 >> w0);
	pair_t o; o.x = gl::mul(v0, w.x); o.y = gl::mul(v1, w.y);
	reinterpret_cast<pair_t *>(A.dst)[i] = o;
}

// carry out of pair i (what the first digit of pair i + 1 receives)
HD u64 lin_pair_carry(const LinArgs &A, u32 i)
{
	u64 u0, u1; u32 w0, w1;
	lin_pair_compute(A, i, u0, u1, w0, w1);
	return u1 >> w1;
}

// ---- addsub on pairs: same single-hop carry, both results from one read of A and B
HD void addsub_pair_compute(const LinArgs &A, u32 i, u64 *us, u64 *ud, u32 &w0, u32 &w1)
{
	const Tables &t = A.t;
	const u32 j = 2 * i;
	w0 = digit_width(t, j); w1 = digit_width(t, j + 1);
	const pair_t x = reinterpret_cast<const pair_t *>(A.srcA)[i], y = reinterpret_cast<const pair_t *>(A.src)[i];
	const pair_t wi = reinterpret_cast<const pair_t *>(t.winv_nat)[i];
	const u64 m0 = ((u64)1 << w0) - 1, m1 = ((u64)1 << w1) - 1;
	us[0] = gl::mul(gl::add(x.x, y.x), wi.x); us[1] = gl::mul(gl::add(x.y, y.y), wi.y);
	ud[0] = gl::add(gl::mul(gl::sub(x.x, y.x), wi.x), 2 * m0); ud[1] = gl::add(gl::mul(gl::sub(x.y, y.y), wi.y), 2 * m1);
}
HD void addsub_pair_finish(const LinArgs &A, u32 i, const u64 *us, const u64 *ud, u32 w0, u32 w1, u64 cs, u64 cd)
{
	const pair_t w = reinterpret_cast<const pair_t *>(A.t.w_nat)[i];
	const u64 m0 = ((u64)1 << w0) - 1, m1 = ((u64)1 << w1) - 1;
	pair_t o;
	o.x = gl::mul((us[0] & m0) + cs, w.x); o.y = gl::mul((us[1] & m1) + (us[0] >> w0), w.y);
	reinterpret_cast<pair_t *>(A.dst)[i] = o;
	if (A.cpy) reinterpret_cast<pair_t *>(A.cpy)[i] = o;
	o.x = gl::mul((ud[0] & m0) + cd, w.x); o.y = gl::mul((ud[1] & m1) + (ud[0] >> w0), w.y);
	reinterpret_cast<pair_t *>(A.dst2)[i] = o;
	if (A.cpy2) reinterpret_cast<pair_t *>(A.cpy2)[i] = o;
}
HD void addsub_pair_carry(const LinArgs &A, u32 i, u64 &cs, u64 &cd)
{
	u64 us[2], ud[2]; u32 w0, w1;
	addsub_pair_compute(A, i, us, ud, w0, w1);
	cs = us[1] >> w1; cd = ud[1] >> w1;
}

static constexpr u32 LIN_BLOCK = 256; // pairs per CTA of the fast linear kernel

// x <- x * f, elementwise (the plain inverse strided pass is unnormalised: times 1/R it is the exact inverse)
HD void scale_thread(u64 *x, u64 f, u32 j, u32 n) { if (j < n) x[j] = gl::mul(x[j], f); }

// dst <- a + b element by element in the field: the sum of two MULTIPLICANDS (transform-domain values) is the multiplicand of
// the sum of their residues, because the transform is linear
HD void gl_add_thread(u64 *dst, const u64 *a, const u64 *b, u32 j, u32 n) { if (j < n) dst[j] = gl::add(a[j], b[j]); }

// unweighted digit values (not carried) for the host-side canonical export (engine_gpu.h:1534-1561)
HD void unweight_thread(const Tables &t, const u64 *x, u64 *out, u32 j) { if (j < t.n) out[j] = gl::mul(x[j], t.winv_nat[j]); }

// ---- canonical digits on the device (engine_gpu.h:1534-1561 does this on the host, serially): each thread unweights a run
// of CANON_D digits and carries it strongly from a zero carry-in (canon_p1); then every run adds the carry of the run
// before it (the last one's wraps to digit 0: 2^p = 1) and ripples it as far as it goes inside the run (canon_p2); a carry
// that leaves a run -- the run was all ones -- is counted in *pending and handled by another canon_p2 round (rare: the host
// looks at the counter).  Output: value | width << 32, the reference's encoding.
static constexpr u32 CANON_D = 16;
struct CanonArgs { Tables t; const u64 *x; u64 *enc; u64 *cin, *cout; u32 *pending; u32 D; };

HD void canon_p1_thread(const CanonArgs &A, u32 r)
{
	const Tables &t = A.t;
	const u32 D = A.D, nruns = t.n / D;
	if (r >= nruns) return;
	const u32 j0 = r * D, bits = run_bits(t, j0, D);
	u64 c = 0;
	for (u32 i = 0; i < D; ++i)
	{
		const u32 w = t.wbase + ((bits >> i) & 1u);
		const u64 u = gl::mul(A.x[j0 + i], t.winv_nat[j0 + i]);
		const u64 s = u + c, co = (s < c) ? 1 : 0;
		c = (s >> w) | (co << (64 - w));
		A.enc[j0 + i] = (s & (((u64)1 << w) - 1)) | ((u64)w << 32);
	}
	A.cout[r] = c;
}

HD void canon_p2_thread(const CanonArgs &A, u32 r)
{
	const Tables &t = A.t;
	const u32 D = A.D, nruns = t.n / D;
	if (r >= nruns) return;
	u64 c = A.cin[(r == 0) ? nruns - 1 : r - 1];
	const u32 j0 = r * D;
	for (u32 i = 0; i < D && c != 0; ++i)
	{
		const u64 e = A.enc[j0 + i];
		const u32 w = (u32)(e >> 32);
		const u64 s = (e & 0xffffffffull) + c; // value < 2^32, c < 2^64 - 2^32 after the first digit; the first add is checked
		const u64 co = (s < c) ? 1 : 0;
		c = (s >> w) | (co << (64 - w));
		A.enc[j0 + i] = (s & (((u64)1 << w) - 1)) | ((u64)w << 32);
	}
	A.cout[r] = c;
	if (c != 0)
	{
#if defined(__CUDA_ARCH__)
		atomicAdd(A.pending, 1u);
#else
		*A.pending += 1;
#endif
	}
}

// compare two canonical digit vectors; flags: bit 0 a digit differs, bit 1 l is not all ones, bit 2 l is not zero,
// bit 3 r is not all ones, bit 4 r is not zero (2^p - 1 and 0 are the same residue, engine.h:181-196)
HD u32 canon_cmp_digit(u64 el, u64 er)
{
	const u32 w = (u32)(el >> 32), vl = (u32)el, vr = (u32)er, ones = (u32)((((u64)1) << w) - 1);
	return (vl != vr ? 1u : 0u) | (vl != ones ? 2u : 0u) | (vl != 0 ? 4u : 0u) | (vr != ones ? 8u : 0u) | (vr != 0 ? 16u : 0u);
}

// ---- recorded op sequences for single-CTA engines ("programs"): one kernel launch runs the whole list ------------
enum ProgCode : u32 { PO_COPY = 0, PO_SQR, PO_SETMUL, PO_MUL, PO_ADD, PO_SUBREG, PO_SUBSMALL, PO_SETSMALL };
// A program is stored as a list schedule over up to PROG_MAX_GROUPS warp groups (marin_engine_core.hpp,
// schedule_program): the list is in start-time order, every op names its group (bits 8-9 of `code`) and, per group, how
// many ops of that group must have finished before it may start (`wait`: the ops it conflicts with -- a register one of
// them writes and the other reads or writes).  A group runs its own ops in list order; the groups meet nowhere else.
static constexpr u32 PO_CODE_MASK = 0xffu, PO_GROUP_SHIFT = 8;
static constexpr u32 PROG_MAX_GROUPS = 4;   // warp groups of k_program (named barriers 1 .. 4)
struct ProgOp { u32 code, dst, src, a; unsigned short wait[PROG_MAX_GROUPS]; };
static_assert(sizeof(ProgOp) == 24, "ProgOp is copied as three 64-bit words");
static constexpr u32 PROG_SMEM_OPS = 128;   // programs up to this length are staged in shared memory by k_program
struct ProgArgs
{
	MidArgs M;          // tables and stages of the engine; x / y / dst / mode / a are set per op
	u64 *reg;           // register slab
	const ProgOp *ops; u32 nops;
	u32 D;              // digits per run of the set-small path
	u64 *carry;         // one word per run (set-small), one per pair (linear ops)
	u32 slab_words, slab_carry; // batched engines: CTA x of the grid runs the program on slab x
	u32 groups, group_threads, group_smem; // warp groups per CTA, threads and bytes of dynamic shared memory of each
	u32 *err;                   // the engine's sticky device error word
	u64 *hid; u32 nregs, nhid;  // registers nregs .. nregs + nhid - 1 of a program are the engine's own (renamed lifetimes), nhid per slab
};
static constexpr u32 PROG_HIDDEN = 4;
HD u64 *prog_reg(const ProgArgs &A, u64 *reg, u64 *hid, u32 r) { return (r < A.nregs) ? reg + (size_t)r * A.M.t.n : hid + (size_t)(r - A.nregs) * A.M.t.n; }

// batched engines: the argument block of CTA `y`, every register pointer moved to slab y
HD MidArgs slab_of(const MidArgs &A, u32 y)
{
	MidArgs B = A;
	if (y == 0 || A.slab_pairs == 0) return B;
	const size_t o = (size_t)y * A.slab_pairs;
	if (B.x) B.x += o; if (B.y) B.y += o; if (B.dst) B.dst += o; if (B.dst2) B.dst2 += o; if (B.add_src) B.add_src += o;
	if (B.carry) B.carry += (size_t)y * A.slab_carry;
	return B;
}
HD LinArgs slab_of(const LinArgs &A, u32 y)
{
	LinArgs B = A;
	if (y == 0 || A.slab_words == 0) return B;
	const size_t o = (size_t)y * A.slab_words, c = (size_t)y * A.slab_carry;
	if (B.dst) B.dst += o; if (B.src) B.src += o; if (B.srcA) B.srcA += o; if (B.dst2) B.dst2 += o; if (B.cpy) B.cpy += o; if (B.cpy2) B.cpy2 += o;
	if (B.carry) B.carry += c; if (B.carry2) B.carry2 += c;
	return B;
}

// weight tables: digit j has weight r^((n - p j mod n) mod n), r the n-th root of two (ibdwt.h:111-147)
struct WeightArgs { u32 p, n; u64 nr2, nr2_inv, inv_m; u64 *w_nat, *winv_nat, *winvs_nat; u64 *w_tile, *winvs_tile; u32 R, stride; };

HD void weight_thread(const WeightArgs &A, u32 j)
{
	if (j >= A.n) return;
	const u64 r = ((u64)A.p * j) % A.n;
	const u64 e = r ? (u64)A.n - r : 0;
	const u64 w = gl::pow(A.nr2, e), wi = gl::pow(A.nr2_inv, e);
	const u64 wis = gl::mul(wi, A.inv_m);
	A.w_nat[j] = w; A.winv_nat[j] = wi;
	if (A.winvs_nat) A.winvs_nat[j] = wis;
	if (A.w_tile)
	{
		const u32 i = j >> 1, row = i / A.stride, col = i % A.stride;
		const size_t k = 2 * (((size_t)(col / TILE_C) * A.R + row) * TILE_C + col % TILE_C) + (j & 1);
		A.w_tile[k] = w; A.winvs_tile[k] = wis;
	}
}

} // namespace marin
