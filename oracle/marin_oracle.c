/*
 * marin_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the Marin IBDWT register
 * engine (PrMers, include/marin/*, kernels/marin.cl).  Nothing in the product path may link,
 * import or call this file: it is the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED.  The field ops, planner, digit widths and weights are checked
 * bit-for-bit against the reference's own headers compiled in place (oracle/_ref, see
 * oracle/ref_shim.cpp), and the engine results are checked against the reference's
 * known-answer vectors (unit_tests.sh:54-69,137-187,204-213) in tests/test_oracle_golden.py.
 *
 * What is restated (reference file:line given at each function):
 *   - Goldilocks field  P = 2^64 - 2^32 + 1             include/marin/arith.h:24-72
 *   - digit carry helpers adc / adc_mul / sbc            include/marin/arith.h:75-100
 *   - transform-size planner                             include/marin/ibdwt.h:17-43
 *   - digit widths and IBDWT weights                     include/marin/ibdwt.h:111-147
 *   - register ops (set/copy/square_mul/set_multiplicand/mul/sub/add/sub_reg)
 *                                                        include/marin/engine.h:45-74,
 *                                                        include/marin/engine_gpu.h:1432-2133
 *   - canonical residue export (unweight + end-around carry, value | width << 32)
 *                                                        include/marin/engine_gpu.h:1534-1561
 *
 * What is NOT restated: the raw (weighted, semi-carried, work-group-size dependent) register
 * image of the OpenCL engine.  The NTT is exact integer arithmetic, so every exact
 * implementation yields the same canonical digits; the oracle keeps its registers fully
 * carried (canonical digits times weight) and uses a plain full-length radix-2 (+ one radix-5
 * stage for 5*2^k) transform instead of the reference's paired radix-4 passes.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

#define MOD_P 0xffffffff00000001ull
#define MOD_MP64 0xffffffffull /* -P mod 2^64 */

/* ---------------------------------------------------------------- field (arith.h:24-72) */

uint64_t mo_mod_add(uint64_t a, uint64_t b) { return a + b + ((a >= MOD_P - b) ? MOD_MP64 : 0); }
uint64_t mo_mod_sub(uint64_t a, uint64_t b) { return a - b - ((a < b) ? MOD_MP64 : 0); }

/* t = hi * 2^64 + lo modulo P, t < P^2 (arith.h:31-36): 2^64 = 2^32 - 1, 2^96 = -1 (mod P) */
static inline uint64_t reduce128(uint64_t lo, uint64_t hi)
{
	const uint64_t r = (lo >= MOD_P) ? lo - MOD_P : lo;
	return mo_mod_sub(mo_mod_add(r, (hi << 32) - (uint32_t)hi), hi >> 32);
}

uint64_t mo_mod_mul(uint64_t a, uint64_t b)
{
	const u128 t = (u128)a * b;
	return reduce128((uint64_t)t, (uint64_t)(t >> 64));
}

uint64_t mo_mod_pow(uint64_t a, uint64_t e)
{
	uint64_t r = 1, y = a;
	for (; e != 0; e >>= 1) { if (e & 1) r = mo_mod_mul(r, y); y = mo_mod_mul(y, y); }
	return r;
}

uint64_t mo_mod_inv(uint64_t a) { return mo_mod_pow(a, MOD_P - 2); }
/* primitive n-th root of unity, generator 7 (arith.h:72) */
uint64_t mo_root_nth(uint64_t n) { return mo_mod_pow(7, (MOD_P - 1) / n); }

/* ---------------------------------------------------------------- digit helpers (arith.h:75-100) */

/* add a carry onto lhs, return the low `width` bits, leave the rest in *carry */
uint32_t mo_adc(uint64_t lhs, uint8_t width, uint64_t *carry)
{
	const uint64_t s = lhs + *carry;
	const uint64_t c = (s < lhs) ? 1 : 0;
	*carry = (s >> width) + (c << (64 - width));
	return (uint32_t)s & ((1u << width) - 1);
}

uint32_t mo_adc_mul(uint64_t lhs, uint32_t a, uint8_t width, uint64_t *carry)
{
	uint64_t c = 0;
	const uint32_t d = mo_adc(lhs, width, &c);
	const uint32_t r = mo_adc((uint64_t)d * a, width, carry);
	*carry += a * c;
	return r;
}

/* ---------------------------------------------------------------- planner (ibdwt.h:17-43) */

size_t mo_transform_size(uint32_t exponent)
{
	uint32_t w = 0, log2_n = 1, log2_n5 = 2;
	do { ++log2_n; w = exponent >> log2_n; } while ((w + 1) * 2 + log2_n >= 64);
	do { ++log2_n5; w = exponent / (5u << log2_n5); } while ((w + 1) * 2 + (log2_n5 + 2.4) >= 64);
	const size_t invalid = (size_t)-1;
	const size_t n2 = (log2_n <= 26) ? ((size_t)1 << log2_n) : invalid;
	const size_t n5 = (log2_n5 <= 26) ? ((size_t)5 << log2_n5) : invalid;
	return (n2 < n5) ? n2 : n5;
}

/* ---------------------------------------------------------------- widths, weights (ibdwt.h:111-147) */

/* width[j] = ceil(p (j+1) / n) - ceil(p j / n), natural digit order */
int mo_widths(size_t n, uint32_t p, uint8_t *width)
{
	const uint32_t q_n = p / (uint32_t)n;
	uint32_t prev = 0;
	for (size_t j = 1; j <= n; ++j)
	{
		const uint64_t qj = (uint64_t)p * j;
		const uint32_t c = (uint32_t)((qj - 1) / n + 1);
		const uint32_t d = c - prev;
		if (d != q_n && d != q_n + 1) return -1;
		width[j - 1] = (uint8_t)d;
		prev = c;
	}
	return 0;
}

/* weight[j] = 2^(ceil(p j / n) - p j / n) = r^((n - p j mod n) mod n), r = 554^((P-1)/192/n) the
   n-th root of two; natural digit order (the reference permutes the table for its 4-digit
   work-items, ibdwt.h:142; the permutation is not part of the contract). */
int mo_weights(size_t n, uint32_t p, uint64_t *w, uint64_t *winv)
{
	if (((MOD_P - 1) / 192) % n != 0) return -1;
	const uint64_t nr2 = mo_mod_pow(554, (MOD_P - 1) / 192 / n);
	#pragma omp parallel for schedule(static) if (n >= 4096)
	for (size_t j = 0; j < n; ++j)
	{
		const uint64_t r = ((uint64_t)p * j) % n;
		const uint64_t x = (r != 0) ? mo_mod_pow(nr2, n - r) : 1;
		w[j] = x;
		if (winv) winv[j] = mo_mod_inv(x);
	}
	return 0;
}

/* ---------------------------------------------------------------- transform */

typedef struct
{
	uint32_t p;
	size_t n, m, regs; /* n = m or 5 m, m a power of two */
	int five;
	uint8_t *width;
	uint64_t *w, *winv; /* winv already multiplied by 1/n */
	uint64_t *tw, *twi; /* m/2 powers of the primitive m-th root and of its inverse */
	uint64_t *tw5, *tw5i; /* n powers of the n-th root (radix-5 stage twiddles), 5*2^k only */
	uint64_t w5[5], w5i[5]; /* powers of the 5th root of unity */
	uint64_t *reg; /* regs * n */
	uint64_t *tmp; /* n */
} mo_engine;

/* The same radix-2 levels, cache-blocked for large m (bench.py's CPU baseline should not be a strawman: the plain
   loops below stream the whole array through memory once per level, 23 times at m = 2^23).  m = R * C, R = 2^t rows:
   the top t levels (stride >= C) are done on tiles of R rows x 8 columns (one cache line per row: the tile stays in
   L2 for all t levels), the remaining levels inside the R contiguous blocks of C elements (each stays in L1/L2).  Two
   passes over memory instead of log2 m.  Butterflies, twiddles and level order are exactly those of the plain loops,
   so the results are bit-identical. */
#define MO_TILE 8
static int blocked_split(size_t m, size_t *R, size_t *C)
{
	if (m < ((size_t)1 << 16) || getenv("MO_NO_BLOCK")) return 0;
	size_t lg = 0; while (((size_t)1 << lg) < m) ++lg;
	*R = (size_t)1 << (lg / 2); *C = m / *R;
	return 1;
}

static void fwd_pow2_blocked(uint64_t *x, size_t m, const uint64_t *tw, size_t R, size_t C)
{
	#pragma omp parallel
	{
		uint64_t *tile = (uint64_t *)malloc(R * MO_TILE * 8);
		#pragma omp for schedule(static)
		for (size_t c0 = 0; c0 < C; c0 += MO_TILE)
		{
			for (size_t r = 0; r < R; ++r) memcpy(tile + r * MO_TILE, x + r * C + c0, MO_TILE * 8);
			for (size_t hr = R / 2, step = 1; hr >= 1; hr /= 2, step *= 2)         /* half = hr * C elements */
				for (size_t blk = 0; blk < R / (2 * hr); ++blk)
					for (size_t rr = 0; rr < hr; ++rr)
					{
						uint64_t *p0 = tile + (blk * 2 * hr + rr) * MO_TILE, *p1 = p0 + hr * MO_TILE;
						const size_t jbase = rr * C + c0;
						for (size_t c = 0; c < MO_TILE; ++c)
						{
							const uint64_t a = p0[c], b = p1[c];
							p0[c] = mo_mod_add(a, b);
							p1[c] = mo_mod_mul(mo_mod_sub(a, b), tw[(jbase + c) * step]);
						}
					}
			for (size_t r = 0; r < R; ++r) memcpy(x + r * C + c0, tile + r * MO_TILE, MO_TILE * 8);
		}
		free(tile);
		#pragma omp for schedule(static)
		for (size_t b = 0; b < R; ++b)
		{
			uint64_t *y = x + b * C;
			for (size_t half = C / 2, step = R; half >= 1; half /= 2, step *= 2)
				for (size_t base = 0; base < C; base += 2 * half)
					for (size_t j = 0; j < half; ++j)
					{
						const size_t i0 = base + j, i1 = i0 + half;
						const uint64_t a = y[i0], bb = y[i1];
						y[i0] = mo_mod_add(a, bb);
						y[i1] = mo_mod_mul(mo_mod_sub(a, bb), tw[j * step]);
					}
		}
	}
}

static void inv_pow2_blocked(uint64_t *x, size_t m, const uint64_t *twi, size_t R, size_t C)
{
	#pragma omp parallel
	{
		#pragma omp for schedule(static)
		for (size_t b = 0; b < R; ++b)
		{
			uint64_t *y = x + b * C;
			for (size_t half = 1, step = m / 2; half <= C / 2; half *= 2, step /= 2)
				for (size_t base = 0; base < C; base += 2 * half)
					for (size_t j = 0; j < half; ++j)
					{
						const size_t i0 = base + j, i1 = i0 + half;
						const uint64_t a = y[i0], bb = mo_mod_mul(y[i1], twi[j * step]);
						y[i0] = mo_mod_add(a, bb);
						y[i1] = mo_mod_sub(a, bb);
					}
		}
		uint64_t *tile = (uint64_t *)malloc(R * MO_TILE * 8);
		#pragma omp for schedule(static)
		for (size_t c0 = 0; c0 < C; c0 += MO_TILE)
		{
			for (size_t r = 0; r < R; ++r) memcpy(tile + r * MO_TILE, x + r * C + c0, MO_TILE * 8);
			for (size_t hr = 1, step = R / 2; hr <= R / 2; hr *= 2, step /= 2)             /* half = hr * C, twiddle step = m / (2 half) */
				for (size_t blk = 0; blk < R / (2 * hr); ++blk)
					for (size_t rr = 0; rr < hr; ++rr)
					{
						uint64_t *p0 = tile + (blk * 2 * hr + rr) * MO_TILE, *p1 = p0 + hr * MO_TILE;
						const size_t jbase = rr * C + c0;
						for (size_t c = 0; c < MO_TILE; ++c)
						{
							const uint64_t a = p0[c], bb = mo_mod_mul(p1[c], twi[(jbase + c) * step]);
							p0[c] = mo_mod_add(a, bb);
							p1[c] = mo_mod_sub(a, bb);
						}
					}
			for (size_t r = 0; r < R; ++r) memcpy(x + r * C + c0, tile + r * MO_TILE, MO_TILE * 8);
		}
		free(tile);
	}
}

static void fwd_pow2(uint64_t *x, size_t m, const uint64_t *tw)
{
	size_t R, C;
	if (blocked_split(m, &R, &C)) { fwd_pow2_blocked(x, m, tw, R, C); return; }
	/* radix-2 decimation in frequency, natural in, bit-reversed out */
	for (size_t half = m / 2, step = 1; half >= 1; half /= 2, step *= 2)
	{
		#pragma omp parallel for schedule(static) if (m >= 16384)
		for (size_t t = 0; t < m / 2; ++t)
		{
			const size_t blk = t / half, j = t % half, i0 = blk * 2 * half + j, i1 = i0 + half;
			const uint64_t a = x[i0], b = x[i1];
			x[i0] = mo_mod_add(a, b);
			x[i1] = mo_mod_mul(mo_mod_sub(a, b), tw[j * step]);
		}
	}
}

static void inv_pow2(uint64_t *x, size_t m, const uint64_t *twi)
{
	size_t R, C;
	if (blocked_split(m, &R, &C)) { inv_pow2_blocked(x, m, twi, R, C); return; }
	/* radix-2 decimation in time, bit-reversed in, natural out, unnormalised */
	for (size_t half = 1, step = m / 2; half <= m / 2; half *= 2, step /= 2)
	{
		#pragma omp parallel for schedule(static) if (m >= 16384)
		for (size_t t = 0; t < m / 2; ++t)
		{
			const size_t blk = t / half, j = t % half, i0 = blk * 2 * half + j, i1 = i0 + half;
			const uint64_t a = x[i0], b = mo_mod_mul(x[i1], twi[j * step]);
			x[i0] = mo_mod_add(a, b);
			x[i1] = mo_mod_sub(a, b);
		}
	}
}

static void dft5(uint64_t v[5], const uint64_t w5[5])
{
	uint64_t o[5];
	for (int r = 0; r < 5; ++r)
	{
		uint64_t s = v[0];
		for (int q = 1; q < 5; ++q) s = mo_mod_add(s, mo_mod_mul(v[q], w5[(q * r) % 5]));
		o[r] = s;
	}
	for (int r = 0; r < 5; ++r) v[r] = o[r];
}

static void forward(const mo_engine *e, uint64_t *x)
{
	const size_t m = e->m;
	if (e->five)
	{
		/* radix-5 first (as the reference does, ibdwt.h:78): j = j1 + m q  ->  block r, then
		   five independent length-m transforms */
		#pragma omp parallel for schedule(static) if (m >= 4096)
		for (size_t j = 0; j < m; ++j)
		{
			uint64_t v[5];
			for (int q = 0; q < 5; ++q) v[q] = x[j + m * q];
			dft5(v, e->w5);
			for (int r = 0; r < 5; ++r) x[j + m * r] = mo_mod_mul(v[r], e->tw5[(j * r) % e->n]);
		}
		for (int r = 0; r < 5; ++r) fwd_pow2(x + m * r, m, e->tw);
	}
	else fwd_pow2(x, m, e->tw);
}

static void inverse(const mo_engine *e, uint64_t *x)
{
	const size_t m = e->m;
	if (e->five)
	{
		for (int r = 0; r < 5; ++r) inv_pow2(x + m * r, m, e->twi);
		#pragma omp parallel for schedule(static) if (m >= 4096)
		for (size_t j = 0; j < m; ++j)
		{
			uint64_t v[5];
			for (int r = 0; r < 5; ++r) v[r] = mo_mod_mul(x[j + m * r], e->tw5i[(j * r) % e->n]);
			dft5(v, e->w5i);
			for (int q = 0; q < 5; ++q) x[j + m * q] = v[q];
		}
	}
	else inv_pow2(x, m, e->twi);
}

/* ---------------------------------------------------------------- engine */

void mo_destroy(mo_engine *e)
{
	if (!e) return;
	free(e->width); free(e->w); free(e->winv); free(e->tw); free(e->twi);
	free(e->tw5); free(e->tw5i); free(e->reg); free(e->tmp); free(e);
}

mo_engine *mo_create_n(uint32_t p, size_t regs, size_t n)
{
	if (n < 4 || n == (size_t)-1 || regs == 0) return NULL;
	mo_engine *e = (mo_engine *)calloc(1, sizeof(mo_engine));
	e->p = p; e->n = n; e->regs = regs;
	e->five = (n % 5 == 0);
	e->m = e->five ? n / 5 : n;
	e->width = (uint8_t *)malloc(n);
	e->w = (uint64_t *)malloc(n * 8); e->winv = (uint64_t *)malloc(n * 8);
	e->tw = (uint64_t *)malloc((e->m / 2 + 1) * 8); e->twi = (uint64_t *)malloc((e->m / 2 + 1) * 8);
	e->reg = (uint64_t *)calloc(regs * n, 8); e->tmp = (uint64_t *)malloc(n * 8);
	if (mo_widths(n, p, e->width) != 0 || mo_weights(n, p, e->w, e->winv) != 0) { mo_destroy(e); return NULL; }
	const uint64_t inv_n = mo_mod_inv((uint64_t)n);
	for (size_t j = 0; j < n; ++j) e->winv[j] = mo_mod_mul(e->winv[j], inv_n);
	const uint64_t rm = mo_root_nth(e->m), rmi = mo_mod_inv(rm);
	uint64_t a = 1, b = 1;
	for (size_t j = 0; j < e->m / 2 + 1; ++j) { e->tw[j] = a; e->twi[j] = b; a = mo_mod_mul(a, rm); b = mo_mod_mul(b, rmi); }
	if (e->five)
	{
		e->tw5 = (uint64_t *)malloc(n * 8); e->tw5i = (uint64_t *)malloc(n * 8);
		const uint64_t rn = mo_root_nth(n), rni = mo_mod_inv(rn);
		a = 1; b = 1;
		for (size_t j = 0; j < n; ++j) { e->tw5[j] = a; e->tw5i[j] = b; a = mo_mod_mul(a, rn); b = mo_mod_mul(b, rni); }
		const uint64_t r5 = mo_root_nth(5), r5i = mo_mod_inv(r5);
		a = 1; b = 1;
		for (int j = 0; j < 5; ++j) { e->w5[j] = a; e->w5i[j] = b; a = mo_mod_mul(a, r5); b = mo_mod_mul(b, r5i); }
	}
	return e;
}

mo_engine *mo_create(uint32_t p, size_t regs) { return mo_create_n(p, regs, mo_transform_size(p)); }

size_t mo_size(const mo_engine *e) { return e->n; }
uint32_t mo_exponent(const mo_engine *e) { return e->p; }
const uint8_t *mo_width_table(const mo_engine *e) { return e->width; }

/* carry integer digit values v[j] * a into canonical digits of x (mod 2^p - 1, end-around),
   then weight.  Restates carry_weight_mul_p1 + carry_weight_p2 (marin.cl:1693-1728,2196-2216)
   and the strong carry of engine_gpu.h:1543-1557 with full propagation. */
static void carry_weight(const mo_engine *e, uint64_t *x, const uint64_t *v, uint32_t a)
{
	const size_t n = e->n;
	u128 c = 0;
	for (size_t j = 0; j < n; ++j)
	{
		c += (u128)v[j] * a;
		x[j] = (uint64_t)c & (((uint64_t)1 << e->width[j]) - 1);
		c >>= e->width[j];
	}
	while (c != 0)
	{
		for (size_t j = 0; j < n && c != 0; ++j)
		{
			c += x[j];
			x[j] = (uint64_t)c & (((uint64_t)1 << e->width[j]) - 1);
			c >>= e->width[j];
		}
	}
	#pragma omp parallel for schedule(static) if (n >= 16384)
	for (size_t j = 0; j < n; ++j) x[j] = mo_mod_mul(x[j], e->w[j]);
}

static void unweight(const mo_engine *e, uint64_t *out, const uint64_t *x, int scaled)
{
	const size_t n = e->n;
	const uint64_t nn = (uint64_t)n;
	#pragma omp parallel for schedule(static) if (n >= 16384)
	for (size_t j = 0; j < n; ++j)
	{
		uint64_t u = mo_mod_mul(x[j], e->winv[j]); /* includes 1/n */
		if (!scaled) u = mo_mod_mul(u, nn);
		out[j] = u;
	}
}

/* dst = a (engine_gpu.h:1432-1449); a may exceed the first digit, so carry it. */
void mo_set_u32(mo_engine *e, size_t dst, uint32_t a)
{
	uint64_t *x = e->reg + dst * e->n;
	memset(e->tmp, 0, e->n * 8);
	e->tmp[0] = a;
	carry_weight(e, x, e->tmp, 1);
}

void mo_copy(mo_engine *e, size_t dst, size_t src)
{
	if (dst != src) memcpy(e->reg + dst * e->n, e->reg + src * e->n, e->n * 8);
}

/* r = r^2 * a (engine.h:51, engine_gpu.h:1568-1630) */
void mo_square_mul(mo_engine *e, size_t r, uint32_t a)
{
	uint64_t *x = e->reg + r * e->n;
	const size_t n = e->n;
	forward(e, x);
	#pragma omp parallel for schedule(static) if (n >= 16384)
	for (size_t j = 0; j < n; ++j) x[j] = mo_mod_mul(x[j], x[j]);
	inverse(e, x);
	unweight(e, e->tmp, x, 1);
	carry_weight(e, x, e->tmp, a);
}

/* dst = transform(src) (engine.h:53, engine_gpu.h:1695-1756) */
void mo_set_multiplicand(mo_engine *e, size_t dst, size_t src)
{
	mo_copy(e, dst, src);
	forward(e, e->reg + dst * e->n);
}

/* dst = dst * src * a, src a multiplicand and left unchanged (engine.h:60, engine_gpu.h:1823-1884) */
void mo_mul(mo_engine *e, size_t dst, size_t src, uint32_t a)
{
	uint64_t *x = e->reg + dst * e->n;
	const uint64_t *y = e->reg + src * e->n;
	const size_t n = e->n;
	forward(e, x);
	#pragma omp parallel for schedule(static) if (n >= 16384)
	for (size_t j = 0; j < n; ++j) x[j] = mo_mod_mul(x[j], y[j]);
	inverse(e, x);
	unweight(e, e->tmp, x, 1);
	carry_weight(e, x, e->tmp, a);
}

/* dst = dst + src (engine.h:64, marin.cl:2157-2194) */
void mo_add(mo_engine *e, size_t dst, size_t src)
{
	uint64_t *x = e->reg + dst * e->n;
	uint64_t *t = (uint64_t *)malloc(e->n * 8);
	unweight(e, e->tmp, x, 0);
	unweight(e, t, e->reg + src * e->n, 0);
	for (size_t j = 0; j < e->n; ++j) e->tmp[j] += t[j];
	free(t);
	carry_weight(e, x, e->tmp, 1);
}

/* dst = dst - src, kept non-negative by adding 2 (2^p - 1) digit-wise
   (engine.h:71, marin.cl:244-257,1854-1888) */
void mo_sub_reg(mo_engine *e, size_t dst, size_t src)
{
	uint64_t *x = e->reg + dst * e->n;
	uint64_t *t = (uint64_t *)malloc(e->n * 8);
	unweight(e, e->tmp, x, 0);
	unweight(e, t, e->reg + src * e->n, 0);
	for (size_t j = 0; j < e->n; ++j) e->tmp[j] += 2 * (((uint64_t)1 << e->width[j]) - 1) - t[j];
	free(t);
	carry_weight(e, x, e->tmp, 1);
}

/* r = r - a (engine.h:62, marin.cl:2375-2393): serial borrow, end-around */
void mo_sub_u32(mo_engine *e, size_t r, uint32_t a)
{
	uint64_t *x = e->reg + r * e->n;
	unweight(e, e->tmp, x, 0);
	uint64_t c = a;
	while (c != 0)
	{
		for (size_t j = 0; j < e->n && c != 0; ++j)
		{
			const uint64_t base = (uint64_t)1 << e->width[j];
			const uint64_t take = c & (base - 1);
			c >>= e->width[j];
			if (e->tmp[j] < take) { e->tmp[j] += base - take; c += 1; }
			else e->tmp[j] -= take;
		}
	}
	carry_weight(e, x, e->tmp, 1);
}

/* canonical digits, value | width << 32 (engine_gpu.h:1534-1561) */
void mo_get_digits(mo_engine *e, size_t src, uint64_t *enc)
{
	const size_t n = e->n;
	unweight(e, e->tmp, e->reg + src * n, 0);
	uint64_t c = 0;
	for (size_t k = 0; k < n; ++k) enc[k] = mo_adc(e->tmp[k], e->width[k], &c);
	while (c != 0)
		for (size_t k = 0; k < n; ++k) { enc[k] = mo_adc(enc[k], e->width[k], &c); if (c == 0) break; }
	for (size_t k = 0; k < n; ++k) enc[k] = (uint32_t)enc[k] | ((uint64_t)e->width[k] << 32);
}

/* dst = digits (low 32 bits of each entry are used, engine_gpu.h:1452-1485) */
void mo_set_digits(mo_engine *e, size_t dst, const uint64_t *enc)
{
	for (size_t k = 0; k < e->n; ++k) e->tmp[k] = (uint32_t)enc[k];
	carry_weight(e, e->reg + dst * e->n, e->tmp, 1);
}

/* raw register image (weighted digits / transform), for tests of get_data/set_data only */
void mo_get_raw(const mo_engine *e, size_t src, uint64_t *out) { memcpy(out, e->reg + src * e->n, e->n * 8); }
void mo_set_raw(mo_engine *e, size_t dst, const uint64_t *in) { memcpy(e->reg + dst * e->n, in, e->n * 8); }
