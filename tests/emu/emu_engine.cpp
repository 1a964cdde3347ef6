// emu_engine.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the product's own kernel phase functions
// (prmers_b200/csrc/marin_kernels.cuh) and host logic (marin_engine_core.hpp) on the CPU, one
// emulated thread after the other with a barrier between phases, so that indexing, twiddles and
// carry logic can be checked against the oracle in the CPU-only test tier.  It is never built into
// or loaded by the product library; the product has no CPU path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../prmers_b200/csrc/marin_engine_core.hpp"

using namespace marin;

struct CpuBackend
{
	explicit CpuBackend(int) {}
	void *alloc(size_t bytes) { void *p = nullptr; if (posix_memalign(&p, 64, bytes ? bytes : 64)) throw std::bad_alloc(); return p; }
	void free(void *p) { ::free(p); }
	void memset0(void *p, size_t b) { std::memset(p, 0, b); }
	void h2d(void *d, const void *s, size_t b) { std::memcpy(d, s, b); }
	void d2h(void *d, const void *s, size_t b) { std::memcpy(d, s, b); }
	void d2d(void *d, const void *s, size_t b) { std::memcpy(d, s, b); }
	void sync() {}
	void sync_nothrow() noexcept {}
	void persist(void *, size_t) {}
	void persist_on() {}
	void persist_off() {}

	void mid(const MidArgs &A0, u32 grid, u32 thr, size_t smem, u32 batch = 1)
	{
		std::vector<pair_t> sm(smem / 16 + 1);
		u64 *smc = (u64 *)(sm.data() + smem_pairs_lin(A0.S));
		const u32 nph = mid_num_phases(A0);
		for (u32 y = 0; y < batch; ++y)
		{
			const MidArgs A = slab_of(A0, y);
			for (u32 b = 0; b < grid; ++b)
				for (u32 ph = 0; ph < nph; ++ph)
					for (u32 t = 0; t < thr; ++t) mid_phase(A, b, ph, t, thr, sm.data(), smc);
		}
	}
	void d2d_2d(void *d, const void *s, size_t bytes, size_t pitch, u32 rows) { for (u32 r = 0; r < rows; ++r) std::memcpy((char *)d + r * pitch, (const char *)s + r * pitch, bytes); }
	void strided_fwd(const StridedArgs &A, u32 grid, u32 thr, size_t smem)
	{
		std::vector<pair_t> sm(smem / 16 + 1);
		const u32 nph = strided_num_phases(A, false);
		for (u32 b = 0; b < grid; ++b)
			for (u32 ph = 0; ph < nph; ++ph)
				for (u32 t = 0; t < thr; ++t) strided_phase<false>(A, b, ph, t, thr, sm.data());
	}
	void strided_inv(const StridedArgs &A, u32 grid, u32 thr, size_t smem)
	{
		// Tile 0 waits for the last tile's flag (end-around carry), so on a sequential emulator its CTA runs up
		// to the publish phase, parks (keeping its shared memory), and resumes after every other CTA is done.
		std::vector<pair_t> sm(smem / 16 + 1), sm0(smem / 16 + 1);
		const u32 nph = strided_num_phases(A, true);
		const u32 park = A.with_carry ? strided_publish_phase(A) + 1 : nph;
		for (u32 ph = 0; ph < park; ++ph)
			for (u32 t = 0; t < thr; ++t) strided_phase<true>(A, 0, ph, t, thr, sm0.data());
		for (u32 b = 1; b < grid; ++b)
			for (u32 ph = 0; ph < nph; ++ph)
				for (u32 t = 0; t < thr; ++t) strided_phase<true>(A, b, ph, t, thr, sm.data());
		for (u32 ph = park; ph < nph; ++ph)
			for (u32 t = 0; t < thr; ++t) strided_phase<true>(A, 0, ph, t, thr, sm0.data());
	}
	void lin_pairs(const LinArgs &A0, u32 batch = 1) { for (u32 y = 0; y < batch; ++y) lin_pairs1(slab_of(A0, y)); }
	void addsub_pairs(const LinArgs &A0, u32 batch = 1) { for (u32 y = 0; y < batch; ++y) addsub_pairs1(slab_of(A0, y)); }
	void lin_pairs1(const LinArgs &A)
	{
		const u32 npairs = A.t.n / 2, nblocks = (npairs + LIN_BLOCK - 1) / LIN_BLOCK;
		for (u32 b = 0; b < nblocks; ++b) { const u32 first = b * LIN_BLOCK; A.carry[b] = lin_pair_carry(A, first == 0 ? npairs - 1 : first - 1); }
		std::vector<u64> cs(LIN_BLOCK);
		for (u32 b = 0; b < nblocks; ++b)
		{
			const u32 cnt = std::min<u32>(LIN_BLOCK, npairs - b * LIN_BLOCK);
			for (u32 t = 0; t < cnt; ++t) cs[t] = lin_pair_carry(A, b * LIN_BLOCK + t);
			for (u32 t = 0; t < cnt; ++t)
			{
				u64 u0, u1; u32 w0, w1;
				lin_pair_compute(A, b * LIN_BLOCK + t, u0, u1, w0, w1);
				lin_pair_finish(A, b * LIN_BLOCK + t, u0, u1, w0, w1, t ? cs[t - 1] : A.carry[b]);
			}
		}
	}
	void addsub_pairs1(const LinArgs &A)
	{
		const u32 npairs = A.t.n / 2, nblocks = (npairs + LIN_BLOCK - 1) / LIN_BLOCK;
		for (u32 b = 0; b < nblocks; ++b) { const u32 first = b * LIN_BLOCK; addsub_pair_carry(A, first == 0 ? npairs - 1 : first - 1, A.carry[b], A.carry2[b]); }
		std::vector<u64> cs(LIN_BLOCK), cd(LIN_BLOCK);
		for (u32 b = 0; b < nblocks; ++b)
		{
			const u32 cnt = std::min<u32>(LIN_BLOCK, npairs - b * LIN_BLOCK);
			for (u32 t = 0; t < cnt; ++t) addsub_pair_carry(A, b * LIN_BLOCK + t, cs[t], cd[t]);
			// all reads of a block come before its writes on the GPU (barrier): keep the results aside, then store
			std::vector<u64> us(2 * cnt), ud(2 * cnt); std::vector<u32> w(2 * cnt);
			for (u32 t = 0; t < cnt; ++t) addsub_pair_compute(A, b * LIN_BLOCK + t, &us[2 * t], &ud[2 * t], w[2 * t], w[2 * t + 1]);
			for (u32 t = 0; t < cnt; ++t)
				addsub_pair_finish(A, b * LIN_BLOCK + t, &us[2 * t], &ud[2 * t], w[2 * t], w[2 * t + 1], t ? cs[t - 1] : A.carry[b], t ? cd[t - 1] : A.carry2[b]);
		}
	}
	void canon_p1(const CanonArgs &A, u32 count) { for (u32 i = 0; i < count; ++i) canon_p1_thread(A, i); }
	void canon_p2(const CanonArgs &A, u32 count) { for (u32 i = 0; i < count; ++i) canon_p2_thread(A, i); }
	void canon_cmp(const u64 *l, const u64 *r, u32 n, u32 *flags) { u32 f = 0; for (u32 j = 0; j < n; ++j) f |= canon_cmp_digit(l[j], r[j]); *flags |= f; }
	void d2h_async(void *d, const void *s, size_t b) { std::memcpy(d, s, b); }
	void lin_p1(const LinArgs &A0, u32 count, u32 batch = 1) { for (u32 y = 0; y < batch; ++y) { const LinArgs A = slab_of(A0, y); for (u32 i = 0; i < count; ++i) lin_p1_thread(A, i); } }
	void lin_p2(const LinArgs &A0, u32 count, u32 batch = 1) { for (u32 y = 0; y < batch; ++y) { const LinArgs A = slab_of(A0, y); for (u32 i = 0; i < count; ++i) lin_p2_thread(A, i); } }
	void unweight(const Tables &t, const u64 *x, u64 *out, u32 n) { for (u32 j = 0; j < n; ++j) unweight_thread(t, x, out, j); }
	// recorded op sequences: the emulated k_program -- the same op order, each op through the same phase functions
	std::vector<std::vector<ProgOp>> progs;
	int prog_upload(const std::vector<ProgOp> &ops) { progs.push_back(ops); return (int)progs.size() - 1; }
	// The device runs the groups of a program side by side, ordered by nothing but the `wait` counts of the ops
	// (schedule_program).  Here the groups are played in an order as far from the recorded one as those counts allow: the
	// LAST group that can move runs as far as it can, then the one before it, ... -- a conflict the scheduler failed to
	// turn into a wait changes the registers.  A program that cannot make progress is an error.
	void prog_launch(int id, ProgArgs A0, u32 batch = 1)
	{
		if (id < 0 || (size_t)id >= progs.size()) throw std::runtime_error("marin_b200: unknown graph id");
		const u32 n = A0.M.t.n, npairs = n / 2;
		const u32 thr = A0.group_threads;
		const size_t smem = A0.group_smem;
		std::vector<ProgOp> order;
		{
			const std::vector<ProgOp> &pr = progs[(size_t)id];
			std::vector<size_t> next(PROG_MAX_GROUPS, 0);       // per group: index into pr of the next op to look at
			u32 done[PROG_MAX_GROUPS] = {0, 0, 0, 0};
			while (order.size() < pr.size())
			{
				bool moved = false;
				for (u32 g = PROG_MAX_GROUPS; g-- > 0;)
				{
					for (;;)
					{
						size_t q = next[g];
						while (q < pr.size() && ((pr[q].code >> PO_GROUP_SHIFT) & (PROG_MAX_GROUPS - 1)) != g) ++q;
						next[g] = q;
						if (q == pr.size()) break;
						bool ok = true;
						for (u32 c = 0; c < PROG_MAX_GROUPS; ++c) if (done[c] < pr[q].wait[c]) ok = false;
						if (!ok) break;
						ProgOp o = pr[q]; o.code &= PO_CODE_MASK; order.push_back(o);
						++done[g]; ++next[g]; moved = true;
					}
				}
				if (!moved) throw std::runtime_error("marin_b200 emulator: the program schedule deadlocks");
			}
		}
		for (u32 y = 0; y < batch; ++y)
		for (const ProgOp &op : order)
		{
			ProgArgs A = A0;
			A.reg += (size_t)y * A0.slab_words; A.carry += (size_t)y * A0.slab_carry; A.M.carry = A.carry; A.M.slab_pairs = 0;
			u64 *const hid = A0.hid + (size_t)y * A0.nhid * n;
			u64 *dst = prog_reg(A0, A.reg, hid, op.dst);
			const u64 *src = prog_reg(A0, A.reg, hid, op.src);
			LinArgs L{A.M.t, dst, src, nullptr, A.carry, 0, op.a, A.D};
			switch (op.code)
			{
			case PO_COPY: std::memcpy(dst, src, (size_t)n * 8); break;
			case PO_SQR: case PO_MUL: case PO_SETMUL:
			{
				MidArgs M = A.M;
				M.x = (pair_t *)((op.code == PO_SETMUL) ? const_cast<u64 *>(src) : dst);
				M.y = (op.code == PO_MUL) ? (const pair_t *)src : nullptr;
				M.dst = (pair_t *)dst;
				M.mode = (op.code == PO_SQR) ? MODE_SQR : ((op.code == PO_MUL) ? MODE_MUL : MODE_FWD);
				M.a = op.a;
				mid(M, 1, thr, smem);
				break;
			}
			case PO_ADD: case PO_SUBREG: case PO_SUBSMALL:
			{
				L.op = (op.code == PO_ADD) ? LIN_ADD : ((op.code == PO_SUBREG) ? LIN_SUBREG : LIN_SUBSMALL);
				for (u32 i = 0; i < npairs; ++i) A.carry[i] = lin_pair_carry(L, i);
				for (u32 i = 0; i < npairs; ++i)
				{
					u64 u0, u1; u32 w0, w1;
					lin_pair_compute(L, i, u0, u1, w0, w1);
					lin_pair_finish(L, i, u0, u1, w0, w1, A.carry[i ? i - 1 : npairs - 1]);
				}
				break;
			}
			default:
				L.src = dst; L.op = LIN_SETSMALL;
				lin_p1(L, n / A.D); lin_p2(L, n / A.D);
				break;
			}
		}
	}
	void scale(u64 *x, u64 f, u32 n) { for (u32 j = 0; j < n; ++j) scale_thread(x, f, j, n); }
	void gl_add(u64 *dst, const u64 *a, const u64 *b, u32 n) { for (u32 j = 0; j < n; ++j) gl_add_thread(dst, a, b, j, n); }
	void weights(const WeightArgs &A, u32 n)
	{
		#pragma omp parallel for schedule(static)
		for (u32 j = 0; j < n; ++j) weight_thread(A, j);
	}
};

typedef EngineCore<CpuBackend> Emu;
static thread_local std::string g_err;

#define GUARD(expr) try { expr; return 1; } catch (const std::exception &e) { g_err = e.what(); return 0; }

extern "C" {
const char *emu_last_error() { return g_err.c_str(); }
void *emu_create(uint32_t p, size_t regs, uint32_t force_S, uint32_t force_R2)
{
	try { return new Emu(p, regs, 0, force_S, force_R2); } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void *emu_create_batch(uint32_t p, size_t regs, uint32_t batch)
{
	try { return new Emu(p, regs, 0, 0, 0, batch); } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
int emu_get_digits_slab(void *h, uint32_t slab, size_t s, uint64_t *enc) { GUARD(((Emu *)h)->get_digits(s, (u64 *)enc, slab)) }
int emu_set_digits_slab(void *h, uint32_t slab, size_t d, const uint64_t *enc) { GUARD(((Emu *)h)->set_digits(d, (const u64 *)enc, slab)) }
void emu_destroy(void *h) { delete (Emu *)h; }
size_t emu_size(void *h) { return ((Emu *)h)->size(); }
int emu_is_small(void *h) { return ((Emu *)h)->pl.small ? 1 : 0; }
// the planner alone (no engine, no tables): out = {small, S, R, R2, nstS, nstR, radices of the contiguous pass [8], of the strided pass [8]}
int emu_plan(uint32_t p, uint32_t *out)
{
	marin::Plan pl; std::string err;
	if (!marin::make_plan(p, pl, err)) { g_err = err; return 0; }
	out[0] = pl.small; out[1] = pl.S; out[2] = pl.R; out[3] = pl.R2; out[4] = pl.nstS; out[5] = pl.nstR;
	for (uint32_t i = 0; i < 8; ++i) { out[6 + i] = i < pl.nstS ? pl.stS[i].radix : 0; out[14 + i] = i < pl.nstR ? pl.stR[i].radix : 0; }
	return 1;
}
uint32_t emu_plan_R(void *h) { return ((Emu *)h)->pl.R; }
uint32_t emu_plan_S(void *h) { return ((Emu *)h)->pl.S; }
uint32_t emu_plan_R2(void *h) { return ((Emu *)h)->pl.R2; }
uint32_t emu_a_max(void *h) { return ((Emu *)h)->pl.a_max; }
int emu_set_u32(void *h, size_t d, uint32_t a) { GUARD(((Emu *)h)->set_u32(d, a)) }
int emu_copy(void *h, size_t d, size_t s) { GUARD(((Emu *)h)->copy(d, s)) }
int emu_square_mul(void *h, size_t r, uint32_t a) { GUARD(((Emu *)h)->square_mul(r, a)) }
int emu_set_multiplicand(void *h, size_t d, size_t s) { GUARD(((Emu *)h)->set_multiplicand(d, s)) }
int emu_mul(void *h, size_t d, size_t s, uint32_t a) { GUARD(((Emu *)h)->mul(d, s, a)) }
int emu_add_multiplicand(void *h, size_t d, size_t a, size_t b) { GUARD(((Emu *)h)->add_multiplicand(d, a, b)) }
int emu_sub_u32(void *h, size_t r, uint32_t a) { GUARD(((Emu *)h)->sub_u32(r, a)) }
int emu_add(void *h, size_t d, size_t s) { GUARD(((Emu *)h)->add(d, s)) }
int emu_sub_reg(void *h, size_t d, size_t s) { GUARD(((Emu *)h)->sub_reg(d, s)) }
int emu_addsub(void *h, size_t s, size_t d, size_t a, size_t b) { GUARD(((Emu *)h)->addsub(s, d, a, b)) }
int emu_addsub_copy(void *h, size_t s, size_t d, size_t sc, size_t dc, size_t a, size_t b) { GUARD(((Emu *)h)->addsub(s, d, a, b, sc, dc)) }
int emu_square_mul_copy(void *h, size_t r, size_t c, uint32_t a) { GUARD(((Emu *)h)->square_mul_copy(r, c, a)) }
int emu_mul_add(void *h, size_t d, size_t ms, size_t as, uint32_t a) { GUARD(((Emu *)h)->mul_add(d, ms, as, a)) }
int emu_mul_copy(void *h, size_t d, size_t s, size_t c, uint32_t a) { GUARD(((Emu *)h)->mul_copy(d, s, c, a)) }
int emu_get_digits(void *h, size_t s, uint64_t *enc) { GUARD(((Emu *)h)->get_digits(s, (u64 *)enc)) }
int emu_set_digits(void *h, size_t d, const uint64_t *enc) { GUARD(((Emu *)h)->set_digits(d, (const u64 *)enc)) }
int emu_graph_begin(void *h) { GUARD(((Emu *)h)->graph_begin()) }
int emu_graph_end(void *h, int *id) { GUARD(*id = ((Emu *)h)->graph_end()) }
int emu_graph_launch(void *h, int id) { GUARD(((Emu *)h)->graph_launch(id)) }
// test aid: levels and ops of a stored program (schedule_program)
int emu_graph_shape(void *h, int id, int *length, int *ops)
{
	// length of the stored schedule in the units of the scheduler's cost model (0.1 us): the start times replayed
	try
	{
		const auto &pr = ((Emu *)h)->be.progs.at((size_t)id);
		*ops = (int)pr.size();
		std::vector<std::vector<unsigned>> fin(PROG_MAX_GROUPS);
		unsigned gfree[PROG_MAX_GROUPS] = {0, 0, 0, 0}, end = 0;
		for (const ProgOp &o : pr)
		{
			const unsigned g = (o.code >> PO_GROUP_SHIFT) & (PROG_MAX_GROUPS - 1);
			unsigned st = gfree[g];
			for (unsigned c = 0; c < PROG_MAX_GROUPS; ++c) if (o.wait[c]) st = std::max(st, fin[c].at(o.wait[c] - 1u));
			const unsigned f = st + prog_cost(o);
			fin[g].push_back(f); gfree[g] = f; end = std::max(end, f);
		}
		*length = (int)end;
		return 1;
	}
	catch (const std::exception &e) { g_err = e.what(); return 0; }
}
int emu_equal(void *h, size_t l, size_t r, int *out) { GUARD(*out = ((Emu *)h)->equal(l, r) ? 1 : 0) }
size_t emu_checkpoint_size(void *h) { Emu *e = (Emu *)h; return e->size() * 8 * e->regs() * e->batch(); }
int emu_get_checkpoint(void *h, void *buf) { GUARD(((Emu *)h)->get_all(buf)) }
int emu_set_checkpoint(void *h, const void *buf) { GUARD(((Emu *)h)->set_all(buf)) }
size_t emu_transform_size(uint32_t p) { return transform_size(p); }
}
