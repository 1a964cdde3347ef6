// marin_engine_core.hpp -- the register engine: which kernels run, in which order, for each call of
// the reference's engine API (include/marin/engine.h:37-170; per-op kernel sequences of
// engine_gpu_flat, include/marin/engine_gpu.h:1432-2164).  Templated on a Backend that owns memory
// and launches: the CUDA backend in marin_engine.cu is the product; tests/emu instantiates the same
// template with a CPU emulator of the kernels to check the host logic without a GPU.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include "marin_plan.hpp"

namespace marin {

// Schedule of a recorded op sequence (k_program), two steps.
// (1) Renaming.  The schedules of the callers reuse one scratch register over and over (the Edwards formulas of ECM:
// set_multiplicand(11, x); mul(., 11) five times in a point addition), which chains otherwise independent products one
// behind the other.  A LIFETIME of register r starts at an op that writes r without reading it (copy, set_multiplicand
// from another register, set) and ends at the last read before the next such op.  The first lifetime of r holds
// what the caller put there, the last one what the caller finds there afterwards; the ones in between are the engine's
// business, and each moves to one of `nhidden` registers behind the caller's (index nregs + h) that is free at that point
// of the recorded order, if there is one.
// (2) List scheduling over G warp groups.  Op i must wait for every earlier op it conflicts with -- it reads a register
// that op writes, writes one that op reads, or both write the same one; nothing else orders two ops.  With the measured
// cost of an op inside the kernel (a transform with its carry ~ 4.5 us at the ECM sizes, a forward transform alone ~ 2.5,
// a linear op ~ 1) the ops go, longest remaining chain first, to the group that can start them earliest.  An Edwards
// doubling (23 ops, 65 us one after the other) has a critical chain of 8 ops.
// Output: the ops in start-time order with their group and, per other group, the number of that group's ops that must be
// complete before the op starts (a group runs its own ops in list order).
struct ProgRW { u32 w, r0, r1; };                     // register written, registers read (~0u = none)
inline ProgRW prog_rw(const ProgOp &o)
{
	switch (o.code & PO_CODE_MASK)
	{
		case PO_COPY: case PO_SETMUL: return ProgRW{o.dst, o.src, ~0u};
		case PO_SQR: case PO_SUBSMALL: return ProgRW{o.dst, o.dst, ~0u};
		case PO_SETSMALL: return ProgRW{o.dst, ~0u, ~0u};
		default: return ProgRW{o.dst, o.dst, o.src};      // PO_MUL, PO_ADD, PO_SUBREG
	}
}
inline bool prog_pure_def(const ProgOp &o)
{
	const u32 c = o.code & PO_CODE_MASK;
	return c == PO_SETSMALL || ((c == PO_COPY || c == PO_SETMUL) && o.dst != o.src);
}
inline u32 prog_cost(const ProgOp &o)
{
	const u32 c = o.code & PO_CODE_MASK;
	return (c == PO_SQR || c == PO_MUL) ? 45u : ((c == PO_SETMUL) ? 17u : ((c == PO_COPY) ? 5u : 10u));   // 0.1 us, profiles/r2/prog_probe_*.log
}

inline std::vector<ProgOp> schedule_program(const std::vector<ProgOp> &recorded, u32 nregs = 0, u32 nhidden = 0, u32 G = PROG_MAX_GROUPS)
{
	std::vector<ProgOp> ops = recorded;
	const size_t k = ops.size();
	for (ProgOp &o : ops) { o.code &= PO_CODE_MASK; for (u32 g = 0; g < PROG_MAX_GROUPS; ++g) o.wait[g] = 0; }
	if (k > 4096 || G <= 1) return ops;                  // one group, recorded order (the conflict search below is quadratic)
	if (G > PROG_MAX_GROUPS) G = PROG_MAX_GROUPS;
	if (nhidden)
	{
		std::vector<size_t> busy_until(nhidden, 0);       // hidden register h is free for a lifetime that starts after this op index
		std::vector<bool> used(nhidden, false);
		for (size_t i = 0; i < k; ++i)
		{
			if (!prog_pure_def(ops[i]) || ops[i].dst >= nregs) continue;
			const u32 r = ops[i].dst;
			// the lifetime that op i starts: up to the next pure definition of r (none: it is the last one and stays in r)
			size_t next_def = k, last_use = i;
			for (size_t j = i + 1; j < k; ++j)
			{
				if (prog_pure_def(ops[j]) && ops[j].dst == r) { next_def = j; break; }
				const ProgRW b = prog_rw(ops[j]);
				if (b.w == r || b.r0 == r || b.r1 == r) last_use = j;
			}
			if (next_def == k) continue;
			// the free hidden register that has been idle longest: a fresh one does not tie this lifetime to the last one's readers
			u32 h = nhidden;
			for (u32 c = 0; c < nhidden; ++c)
				if (!used[c]) { h = c; break; }
				else if (busy_until[c] < i && (h == nhidden || busy_until[c] < busy_until[h])) h = c;
			if (h == nhidden) continue;
			used[h] = true; busy_until[h] = last_use;
			const u32 nr = nregs + h;
			// every reference to r from op i to last_use is a reference to this lifetime
			for (size_t j = i; j <= last_use; ++j)
			{
				if (ops[j].dst == r) ops[j].dst = nr;
				if (ops[j].src == r) ops[j].src = nr;
			}
		}
	}
	// conflicts, and the length of the longest chain of ops that hangs on each op
	std::vector<std::vector<u32>> pred(k);
	std::vector<u32> chain(k), npend(k, 0);
	for (size_t i = 0; i < k; ++i)
	{
		const ProgRW a = prog_rw(ops[i]);
		for (size_t j = 0; j < i; ++j)
		{
			const ProgRW b = prog_rw(ops[j]);
			const bool raw = (a.r0 == b.w || a.r1 == b.w), waw = (a.w == b.w), war = (a.w == b.r0 || a.w == b.r1);
			if (raw || waw || war) pred[i].push_back((u32)j);
		}
		chain[i] = prog_cost(ops[i]);
	}
	for (size_t i = k; i-- > 0;)
		for (u32 j : pred[i]) if (chain[j] < prog_cost(ops[j]) + chain[i]) chain[j] = prog_cost(ops[j]) + chain[i];
	std::vector<u32> start(k, 0), finish(k, 0), grp(k, 0), pos(k, 0), gfree(G, 0), gcount(G, 0);
	std::vector<bool> done(k, false);
	std::vector<u32> order;
	order.reserve(k);
	for (size_t step = 0; step < k; ++step)
	{
		// among the ops whose predecessors are all placed: the one that can start first, the longest chain on a tie
		size_t best = k; u32 best_start = 0, best_g = 0;
		for (size_t i = 0; i < k; ++i)
		{
			if (done[i]) continue;
			u32 ready = 0, home = 0; bool ok = true;
			for (u32 j : pred[i]) { if (!done[j]) { ok = false; break; } if (finish[j] >= ready) { ready = finish[j]; home = grp[j]; } }
			if (!ok) continue;
			// the group that can start it first; on a tie the group of the op it waits for longest (no hand-off between groups)
			u32 g = 0, st = ~0u;
			for (u32 c = 0; c < G; ++c) { const u32 s0 = (gfree[c] > ready) ? gfree[c] : ready; if (s0 < st || (s0 == st && c == home && !pred[i].empty())) { st = s0; g = c; } }
			if (best == k || st < best_start || (st == best_start && chain[i] > chain[best])) { best = i; best_start = st; best_g = g; }
		}
		const size_t i = best;
		done[i] = true; start[i] = best_start; finish[i] = best_start + prog_cost(ops[i]); grp[i] = best_g;
		gfree[best_g] = finish[i]; pos[i] = gcount[best_g]++;
		order.push_back((u32)i);
	}
	// start-time order (placement order is not: an op with a long chain may be placed before an earlier-starting one of
	// another group); a group's own ops keep their order, and every predecessor starts before its successor
	std::stable_sort(order.begin(), order.end(), [&](u32 x, u32 y) { return start[x] < start[y]; });
	// positions inside the groups follow the sorted list
	std::fill(gcount.begin(), gcount.end(), 0u);
	for (u32 i : order) pos[i] = gcount[grp[i]]++;
	std::vector<ProgOp> out;
	out.reserve(k);
	for (u32 i : order)
	{
		ProgOp o = ops[i];
		o.code |= grp[i] << PO_GROUP_SHIFT;
		for (u32 j : pred[i]) if (grp[j] != grp[i] && o.wait[grp[j]] < pos[j] + 1) o.wait[grp[j]] = (unsigned short)(pos[j] + 1);
		out.push_back(o);
	}
	return out;
}

template <class Backend> class EngineCore
{
public:
	Plan pl;
	Backend be;

	// batch > 1: `batch` independent register sets ("slabs") driven in lockstep -- every operation runs on all of them,
	// one CTA per slab in the same launch.  For the single-CTA sizes only (the ECM sizes: one curve per slab, so a B200
	// works on hundreds of curves in the time of one; the reference creates one engine per curve and runs them one
	// after the other, RunEcmTwistedEdwards.cpp:1736-1755).  Digits are read and written per slab.
	EngineCore(u32 p, size_t regs, int device, u32 force_S = 0, u32 force_R2 = 0, u32 batch = 1) : be(device), _regs(regs), _batch(batch ? batch : 1)
	{
		std::string err;
		if (regs == 0) throw std::runtime_error("marin_b200: register count must be >= 1");
		if (!make_plan(p, pl, err, force_S, force_R2)) throw std::runtime_error("marin_b200: " + err);
		if (_batch > 1 && !pl.small) throw std::runtime_error("marin_b200: batched engines are for single-CTA transform sizes (n <= 2048) only");
		if (_batch > 65535) throw std::runtime_error("marin_b200: at most 65535 slabs");
		HostTables h;
		build_tables(pl, h);
		_wbits = h.wbits;
		const size_t n = pl.n;
		_reg = (u64 *)dev_alloc(_batch * regs * n * 8);
		be.memset0(_reg, _batch * regs * n * 8);
		const size_t ncarry = std::max<size_t>(std::max<size_t>(n / pl.D, n / RUN + 1), pl.small ? n / 2 : 0); // recorded linear ops keep one word per pair
		_ncarry = ncarry;
		_carry = (u64 *)dev_alloc(_batch * ncarry * 8);
		_carry2 = (u64 *)dev_alloc(_batch * ncarry * 8);
		_scratch = (u64 *)dev_alloc(n * 8);
		if (pl.small) { _hid = (u64 *)dev_alloc(_batch * PROG_HIDDEN * n * 8); be.memset0(_hid, _batch * PROG_HIDDEN * n * 8); } // recorded programs: renamed lifetimes
		_canon_D = (n % 16 == 0) ? 16 : ((n % 8 == 0) ? 8 : 4);
		_cc[0] = (u64 *)dev_alloc((n / _canon_D) * 8); _cc[1] = (u64 *)dev_alloc((n / _canon_D) * 8);
		_words = (u32 *)dev_alloc(16); be.memset0(_words, 16);   // [0], [1] carries still pending after a canon round, [2] comparison flags
		_nslots = n / RUN + 1;
		_slots = (pair_t *)dev_alloc(_nslots * 16);   // {carry, epoch} hand-off slots of the inverse strided pass
		be.memset0(_slots, _nslots * 16);
		if (!pl.small) { _fwd = (u64 *)dev_alloc(n * 8); be.persist(_fwd, n * 8); _inv_R = gl::inv(pl.R); }
		Tables &t = _t;
		std::memset(&t, 0, sizeof(t));
		t.p = pl.p; t.n = pl.n; t.m = pl.m; t.wbase = pl.wbase; t.inv_m = h.inv_m;
		t.wbits = (const u32 *)upload(h.wbits.data(), h.wbits.size() * 4);
		t.twS = (const u64 *)upload(h.twS.data(), h.twS.size() * 8); t.S = pl.S;
		t.twSi = (const u64 *)upload(h.twSi.data(), h.twSi.size() * 8);
		_twR = (const u64 *)upload(h.twR.data(), h.twR.size() * 8);
		_twRi = (const u64 *)upload(h.twRi.data(), h.twRi.size() * 8);
		_twR2 = (const u64 *)upload(h.twR2.data(), h.twR2.size() * 8);
		_twR2i = (const u64 *)upload(h.twR2i.data(), h.twR2i.size() * 8);
		_rowK2 = (const u32 *)upload(h.rowK2.data(), h.rowK2.size() * 4);
		t.bl_lo = (const u64 *)upload(h.bl_lo.data(), h.bl_lo.size() * 8);
		t.bl_hi = (const u64 *)upload(h.bl_hi.data(), h.bl_hi.size() * 8); t.bl_h = h.bl_h;
		t.bl_loi = (const u64 *)upload(h.bl_loi.data(), h.bl_loi.size() * 8);
		t.bl_hii = (const u64 *)upload(h.bl_hii.data(), h.bl_hii.size() * 8);
		_rowK = (const u32 *)upload(h.rowK.data(), h.rowK.size() * 4);
		t.rowroot = (const u64 *)upload(h.rowroot.data(), h.rowroot.size() * 8);
		t.colroot = (const u64 *)upload(h.colroot.data(), h.colroot.size() * 8);
		for (int i = 0; i < 5; ++i) { t.w5[i] = h.w5[i]; t.w5i[i] = h.w5i[i]; }
		// natural-order tables for the linear and single-CTA kernels; multi-pass engines unweight / reweight from
		// the tile-ordered copies instead
		u64 *w_nat = (u64 *)dev_alloc(n * 8), *winv_nat = (u64 *)dev_alloc(n * 8);
		u64 *winvs_nat = pl.small ? (u64 *)dev_alloc(n * 8) : nullptr;
		u64 *w_tile = pl.small ? nullptr : (u64 *)dev_alloc(n * 8), *winvs_tile = pl.small ? nullptr : (u64 *)dev_alloc(n * 8);
		WeightArgs W{pl.p, pl.n, h.nr2, h.nr2_inv, h.inv_m, w_nat, winv_nat, winvs_nat, w_tile, winvs_tile, pl.R, pl.small ? 1u : pl.R2 * pl.S};
		be.weights(W, pl.n);
		t.w_nat = w_nat; t.winv_nat = winv_nat; t.winvs_nat = winvs_nat; t.w_tile = w_tile; t.winvs_tile = winvs_tile;
		const size_t nflags = (size_t)(pl.R2 * pl.S / TILE_C) + 4; // published[tiles], error word, then the 8-byte aligned 64-bit tile ticket
		_flags = (u32 *)dev_alloc(nflags * 4);
		be.memset0(_flags, nflags * 4);
		_nflags = nflags;
		be.sync();
	}

	~EngineCore()
	{
		be.sync_nothrow();   // a destructor must not throw: a sticky CUDA error here is reported by whoever created the next engine
	}

	size_t size() const { return pl.n; }
	size_t regs() const { return _regs; }
	u32 batch() const { return _batch; }
	void sync()
	{
		blocking("sync");
		check_device_error();
	}

	// The error word of the inverse pass (a carry hand-off that timed out poisons the carry it returns and raises this
	// word): read, after a stream sync, by EVERY call that hands register contents or a comparison to the host, so a
	// wrong residue cannot leave the engine silently.  Sticky: once set the engine refuses to export anything.
	void check_device_error()
	{
		u32 err = 0;
		be.d2h(&err, _flags + (size_t)(pl.R2 * pl.S / TILE_C), 4); // synchronises the stream
		if (err) throw std::runtime_error("marin_b200: carry hand-off timed out inside the inverse pass (device error flag set)");
	}

	// Record the engine calls that follow instead of running them (graph_end returns the id of the recorded sequence,
	// graph_launch replays it as ONE kernel launch, k_program).  Only engines whose every op fits one CTA can be
	// recorded: the multi-pass transforms need grid-wide passes.
	void graph_begin()
	{
		if (!pl.small) throw std::runtime_error("marin_b200: recording is for single-CTA engines (n <= 2048) only");
		blocking("graph_begin");
		_rec.clear();
		_recording = true;
	}
	int graph_end()
	{
		if (!_recording) throw std::runtime_error("marin_b200: graph_end without graph_begin");
		_recording = false;
		return be.prog_upload(schedule_program(_rec, (u32)_regs, PROG_HIDDEN, prog_groups()));
	}
	// warp groups of a program launch: 4, 4, 4, 2, 1 for groups of 32, 64, 128, 256, 512 threads (a power of two; the CTA has 512 threads at most)
	u32 prog_group_threads() const { return pl.thrMid < 32 ? 32 : pl.thrMid; }   // one warp per group up to n = 128: its barrier is a __syncwarp
	u32 prog_groups() const { const u32 thr = prog_group_threads(); u32 g = 1; while (2 * g <= PROG_MAX_GROUPS && 2 * g * thr <= 512) g *= 2; return g; }
	void graph_launch(int id)
	{
		blocking("graph_launch");
		ProgArgs A;
		std::memset(&A, 0, sizeof(A));
		A.M = mid_args(nullptr, nullptr, nullptr, MODE_SQR, 1);
		A.reg = _reg; A.D = pl.D; A.carry = _carry;
		A.slab_words = (u32)slab_words(); A.slab_carry = (u32)_ncarry;
		A.hid = _hid; A.nregs = (u32)_regs; A.nhid = PROG_HIDDEN;
		// warp groups: the independent ops of a level side by side, each group with its own scratch in shared memory
		// (the transform buffer doubles as the carry words of the linear ops: one per pair <= the buffer's 2 S words)
		A.group_threads = prog_group_threads();
		A.groups = prog_groups();
		A.err = _flags + (size_t)(pl.R2 * pl.S / TILE_C);
		A.group_smem = (u32)((mid_smem() + 15) / 16 * 16);
		be.prog_launch(id, A, _batch);
	}

	void set_u32(size_t dst, u32 a)
	{
		check(dst);
		if (pl.p < 32) a = (u32)((u64)a % (((u64)1 << pl.p) - 1));
		if (rec(PO_SETSMALL, dst, dst, a)) return;
		lin(LIN_SETSMALL, dst, dst, a, nullptr);
	}

	void copy(size_t dst, size_t src)
	{
		check(dst); check(src);
		if (dst == src || rec(PO_COPY, dst, src, 0)) return;
		touch(dst);
		if (_batch > 1) be.d2d_2d(R(dst), R(src), (size_t)pl.n * 8, slab_words() * 8, _batch);
		else be.d2d(R(dst), R(src), pl.n * 8);
	}

	// r = r^2 * a.  In a chain of squarings of the same register (PRP, LL, P-1 stage 1) the inverse
	// pass of one squaring also runs the forward strided pass of the next on the tile it still holds in
	// shared memory and parks it in _fwd: two launches per squaring instead of three, and one 16n-byte
	// round trip through HBM less.  The register itself always holds the finished digits, so any other
	// operation simply ignores (and invalidates) the parked transform.
	void square_mul(size_t r, u32 a)
	{
		check_range(r); check_a(a);
		if (rec(PO_SQR, r, r, a)) return;
		if (pl.small) { MidArgs M = mid_args(R(r), nullptr, R(r), MODE_SQR, a); launch_mid(M); return; }
		if (_lazy != r) materialize();                   // another register's state lives in _fwd: write it out first
		const bool chained = (_last_sq == r);            // second squaring in a row: start speculating
		if (_fwd_reg != r) strided(false, 0, R(r), _fwd, 1, nullptr);
		_fwd_reg = (size_t)-1;
		if (pl.R2 > 1) strided(false, 1, _fwd, _fwd, 1, nullptr);
		MidArgs M = mid_args(_fwd, nullptr, _fwd, MODE_SQR, a); launch_mid(M);
		if (pl.R2 > 1) strided(true, 1, _fwd, _fwd, 1, nullptr);
		// In a chain the finished digits are not even written back: the parked transform of them IS the register's
		// state (8n bytes of HBM writes less per squaring); materialize() recovers the digits, exactly, the first
		// time anything else looks at the register.
		strided(true, 0, _fwd, R(r), a, chained ? _fwd : nullptr, chained);
		if (chained) { _fwd_reg = r; _lazy = r; } else _lazy = (size_t)-1;
		_last_sq = r;
		if (chained) { if (++_chain_len == CHAIN_PERSIST) be.persist_on(); } else calm();
	}

	// Register _lazy's memory is stale; _fwd holds the forward strided pass of its digits.  The plain inverse
	// strided pass times 1/R is the exact inverse (field arithmetic), so this writes the very digits the carry pass
	// would have stored.  _fwd is left intact: the chain may go on.
	void materialize()
	{
		if (_lazy == (size_t)-1) return;
		const size_t r = _lazy;
		_lazy = (size_t)-1;
		strided(true, 0, _fwd, R(r), 1, nullptr, false, true);
		be.scale(R(r), _inv_R, pl.n);
	}

	void set_multiplicand(size_t dst, size_t src)
	{
		check(dst); check(src);
		if (rec(PO_SETMUL, dst, src, 1)) return;
		if (pl.small) { MidArgs M = mid_args(R(src), nullptr, R(dst), MODE_FWD, 1); launch_mid(M); return; }
		touch(dst);
		strided(false, 0, R(src), R(dst), 1, nullptr);
		if (pl.R2 > 1) strided(false, 1, R(dst), R(dst), 1, nullptr);
		MidArgs M = mid_args(R(dst), nullptr, R(dst), MODE_FWD, 1); launch_mid(M);
	}

	// dst <- multiplicand of (x + y), from the multiplicands a of x and b of y: one streaming pass instead of
	// copy + add + set_multiplicand (no reference counterpart; the V-trace term V_kD - V_j of P-1 stage 2 is its use).
	// The sum has digits of up to twice the size, which a later mul must absorb: allowed only where the carry chain has
	// that bit of headroom (max multiplier >= 2), and the sum itself must not be summed again.
	void add_multiplicand(size_t dst, size_t a, size_t b)
	{
		check(dst); check(a); check(b);
		if (_recording) throw std::runtime_error("marin_b200: add_multiplicand cannot be recorded");
		if (_batch > 1) throw std::runtime_error("marin_b200: add_multiplicand is not available on a batched engine");
		if (pl.a_max < 2) throw std::runtime_error("marin_b200: add_multiplicand needs one bit of headroom in the carry chain (max multiplier >= 2)");
		touch(dst);
		be.gl_add(R(dst), R(a), R(b), pl.n);
	}

	void mul(size_t dst, size_t src, u32 a)
	{
		check(dst); check(src); check_a(a);
		if (dst == src) throw std::runtime_error("marin_b200: mul(dst, src) needs src != dst (src is a multiplicand)");
		if (rec(PO_MUL, dst, src, a)) return;
		if (pl.small) { MidArgs M = mid_args(R(dst), R(src), R(dst), MODE_MUL, a); launch_mid(M); return; }
		materialize();
		if (_fwd_reg != dst) strided(false, 0, R(dst), _fwd, 1, nullptr);
		touch(dst);
		_fwd_reg = (size_t)-1;                           // _fwd is the scratch buffer of this product: whatever was parked there is gone
		if (pl.R2 > 1) strided(false, 1, _fwd, _fwd, 1, nullptr);
		MidArgs M = mid_args(_fwd, R(src), _fwd, MODE_MUL, a); launch_mid(M);
		if (pl.R2 > 1) strided(true, 1, _fwd, _fwd, 1, nullptr);
		strided(true, 0, _fwd, R(dst), a, nullptr);
	}

	// r = r^2 * a (or dst = dst * src * a) and a second copy of the result, stored by the same carry pass
	// (carry_weight_mul_p1_copy / p2_copy, marin.cl:1731-1769,1824-1851; engine_gpu.h:1131)
	void square_mul_copy(size_t r, size_t c, u32 a)
	{
		check_range(r); check_range(c); check_a(a);
		if (_recording || c == r) { square_mul(r, a); copy(c, r); return; }
		check(r); check(c);
		touch(c); touch(r);
		if (pl.small) { MidArgs M = mid_args(R(r), nullptr, R(r), MODE_SQR, a); M.dst2 = (pair_t *)R(c); launch_mid(M); return; }
		materialize();
		_fwd_reg = (size_t)-1;
		strided(false, 0, R(r), _fwd, 1, nullptr);
		if (pl.R2 > 1) strided(false, 1, _fwd, _fwd, 1, nullptr);
		MidArgs M = mid_args(_fwd, nullptr, _fwd, MODE_SQR, a); launch_mid(M);
		if (pl.R2 > 1) strided(true, 1, _fwd, _fwd, 1, nullptr);
		strided(true, 0, _fwd, R(r), a, nullptr, false, false, R(c));
	}
	void mul_copy(size_t dst, size_t src, size_t c, u32 a)
	{
		check_range(dst); check_range(src); check_range(c); check_a(a);
		if (_recording || c == dst || c == src) { mul(dst, src, a); copy(c, dst); return; }
		check(dst); check(src); check(c);
		if (dst == src) throw std::runtime_error("marin_b200: mul(dst, src) needs src != dst (src is a multiplicand)");
		touch(c);
		if (pl.small) { MidArgs M = mid_args(R(dst), R(src), R(dst), MODE_MUL, a); M.dst2 = (pair_t *)R(c); launch_mid(M); return; }
		materialize();
		if (_fwd_reg != dst) strided(false, 0, R(dst), _fwd, 1, nullptr);
		touch(dst);
		_fwd_reg = (size_t)-1;
		if (pl.R2 > 1) strided(false, 1, _fwd, _fwd, 1, nullptr);
		MidArgs M = mid_args(_fwd, R(src), _fwd, MODE_MUL, a); launch_mid(M);
		if (pl.R2 > 1) strided(true, 1, _fwd, _fwd, 1, nullptr);
		strided(true, 0, _fwd, R(dst), a, nullptr, false, false, R(c));
	}

	// dst = dst * mul_src + add_src with the addition inside the carry pass (carry_weight_muladd_p1/p2,
	// marin.cl:2298-2365); a != 1 multiplies the whole sum in the carry chain, so that case keeps the composition
	void mul_add(size_t dst, size_t mul_src, size_t add_src, u32 a)
	{
		check_range(dst); check_range(mul_src); check_range(add_src); check_a(a);
		if (_recording || a != 1 || add_src == dst) { mul(dst, mul_src, a); add(dst, add_src); return; }
		check(dst); check(mul_src); check(add_src);
		if (dst == mul_src) throw std::runtime_error("marin_b200: mul(dst, src) needs src != dst (src is a multiplicand)");
		if (pl.small) { MidArgs M = mid_args(R(dst), R(mul_src), R(dst), MODE_MUL, 1); M.add_src = (const pair_t *)R(add_src); launch_mid(M); return; }
		materialize();
		if (_fwd_reg != dst) strided(false, 0, R(dst), _fwd, 1, nullptr);
		touch(dst);
		_fwd_reg = (size_t)-1;
		if (pl.R2 > 1) strided(false, 1, _fwd, _fwd, 1, nullptr);
		MidArgs M = mid_args(_fwd, R(mul_src), _fwd, MODE_MUL, 1); launch_mid(M);
		if (pl.R2 > 1) strided(true, 1, _fwd, _fwd, 1, nullptr);
		strided(true, 0, _fwd, R(dst), 1, nullptr, false, false, nullptr, R(add_src));
	}

	// S = A + B and D = A - B in one pass over the registers, optionally with a second copy of each
	// (carry_weight_addsub_p1/p2 and their _copy forms, marin.cl:1891-2154; engine_gpu.h:1064-1072).  Outputs may alias
	// inputs (the Edwards schedules call it in place); S and D must differ.
	void addsub(size_t s, size_t d, size_t a, size_t b, size_t sc = (size_t)-1, size_t dc = (size_t)-1)
	{
		const bool copies = sc != (size_t)-1;
		check(s); check(d); check(a); check(b);
		if (copies) { check(sc); check(dc); }
		if (_recording || s == d || (copies && (sc == dc || sc == d || dc == s || sc == a || sc == b || dc == a || dc == b)))
		{
			// the reference's default composition (engine.h:72-79,125-131), also what a recording captures
			copy(s, a); copy(d, a); add(s, b); sub_reg(d, b);
			if (copies) { copy(sc, s); copy(dc, d); }
			return;
		}
		touch(s); touch(d);
		if (copies) { touch(sc); touch(dc); }
		LinArgs L{_t, R(s), R(b), nullptr, _carry, LIN_ADDSUB, 0, pl.D, R(a), R(d), copies ? R(sc) : nullptr, copies ? R(dc) : nullptr, _carry2};
		if (_batch > 1) { L.slab_words = (u32)slab_words(); L.slab_carry = (u32)_ncarry; }
		be.addsub_pairs(L, _batch);
	}

	void sub_u32(size_t r, u32 a)
	{
		check(r);
		if (pl.p < 32) a = (u32)((u64)a % (((u64)1 << pl.p) - 1));
		if (rec(PO_SUBSMALL, r, r, a)) return;
		lin(LIN_SUBSMALL, r, r, a, nullptr);
	}
	void add(size_t dst, size_t src) { check(dst); check(src); if (!rec(PO_ADD, dst, src, 0)) lin(LIN_ADD, dst, src, 0, nullptr); }
	void sub_reg(size_t dst, size_t src) { check(dst); check(src); if (!rec(PO_SUBREG, dst, src, 0)) lin(LIN_SUBREG, dst, src, 0, nullptr); }

	// Canonical digits, value | width << 32 (engine_gpu.h:1534-1561 carries serially on the host: 8.4 M digits, ~10 ms).
	// Here: unweight + strong carry per run of 16 digits, then one round in which every run takes in the carry of the
	// run before it (end-around); a carry that crosses a whole run (all ones) needs another round, which the host sees
	// in a 4-byte counter.  After CANON_ROUNDS rounds (a pathological all-ones residue plus a stray unit) the rest is
	// finished on the host exactly as the reference does.
	static constexpr u32 CANON_ROUNDS = 8;
	// first round only, nothing read back: the caller looks at word `slot` of _words later (several of these can be in
	// flight before one 4-byte read) and calls canon_finish if carries are still pending
	void canon_start(size_t src, u64 *enc_dev, u64 *cc0, u64 *cc1, u32 slot, u32 slab = 0)
	{
		const u32 runs = pl.n / _canon_D;
		CanonArgs C{_t, R(src) + (size_t)slab * slab_words(), enc_dev, cc0, cc0, _words + slot, _canon_D};
		be.canon_p1(C, runs);
		C.cin = cc0; C.cout = cc1;
		be.canon_p2(C, runs);
	}
	// carries cc1[r] (leaving run r) are still pending: more rounds, then the host
	void canon_finish(u64 *enc_dev, u64 *cc0, u64 *cc1, u32 slot)
	{
		const u32 runs = pl.n / _canon_D;
		CanonArgs C{_t, nullptr, enc_dev, cc0, cc0, _words + slot, _canon_D};
		u64 *cc[2] = {cc1, cc0};
		u32 cur = 0;
		for (u32 round = 1; round < CANON_ROUNDS; ++round)
		{
			be.memset0(_words + slot, 4);
			C.cin = cc[cur]; C.cout = cc[cur ^ 1];
			be.canon_p2(C, runs);
			cur ^= 1;
			u32 pending = 0;
			be.d2h(&pending, _words + slot, 4);
			if (pending == 0) return;
		}
		std::vector<u64> enc(pl.n), cv(runs);
		be.d2h(enc.data(), enc_dev, (size_t)pl.n * 8);
		be.d2h(cv.data(), cc[cur], (size_t)runs * 8);
		bool any = true;
		while (any)
		{
			any = false;
			for (u32 r = 0; r < runs; ++r)
			{
				u64 c = cv[(r == 0) ? runs - 1 : r - 1];
				if (c == 0) continue;
				cv[(r == 0) ? runs - 1 : r - 1] = 0;
				for (u32 k = r * _canon_D; c != 0; k = (k + 1 == pl.n) ? 0 : k + 1) { const u32 w = width(k); enc[k] = (u64)adc((u32)enc[k], w, c) | ((u64)w << 32); }
				any = true;
			}
		}
		be.h2d(enc_dev, enc.data(), (size_t)pl.n * 8);
	}
	// one blocking read of the four status words: [0], [1] pending carries of two canon_start calls, [2] comparison
	// flags, and the error word of the inverse pass
	void read_words(u32 *w4)
	{
		be.d2h_async(w4, _words, 12);
		be.d2h(w4 + 3, _flags + (size_t)(pl.R2 * pl.S / TILE_C), 4); // synchronises the stream
		if (w4[3]) throw std::runtime_error("marin_b200: carry hand-off timed out inside the inverse pass (device error flag set)");
	}

	void get_digits(size_t src, u64 *enc, u32 slab = 0)
	{
		check(src); check_slab(slab); blocking("get");
		be.memset0(_words, 12);
		canon_start(src, _scratch, _cc[0], _cc[1], 0, slab);
		u32 w[4];
		read_words(w);
		if (w[0]) canon_finish(_scratch, _cc[0], _cc[1], 0);
		be.d2h(enc, _scratch, (size_t)pl.n * 8);
	}

	void set_digits(size_t dst, const u64 *enc, u32 slab = 0)
	{
		check(dst); check_slab(slab); blocking("set");
		be.h2d(_scratch, enc, (size_t)pl.n * 8);
		lin(LIN_SETDIGITS, dst, dst, 0, _scratch, (int)slab);
	}

	// is_equal (engine.h:148-157 converts both registers to GMP integers on the host): canonical digits of both on the
	// device, one compare-and-reduce kernel, 4 bytes back.  2^p - 1 (all ones) and 0 are the same residue.
	bool equal(size_t l, size_t r)
	{
		check(l); check(r); blocking("is_equal");
		if (_batch > 1) throw std::runtime_error("marin_b200: is_equal compares one pair of registers; on a batched engine read the digits per slab");
		if (!_scratch2)
		{
			_scratch2 = (u64 *)dev_alloc((size_t)pl.n * 8);
			_cc2[0] = (u64 *)dev_alloc((pl.n / _canon_D) * 8); _cc2[1] = (u64 *)dev_alloc((pl.n / _canon_D) * 8);
		}
		u32 w[4];
		be.memset0(_words, 12);
		canon_start(l, _scratch, _cc[0], _cc[1], 0);
		canon_start(r, _scratch2, _cc2[0], _cc2[1], 1);
		be.canon_cmp(_scratch, _scratch2, pl.n, _words + 2);     // speculative: valid when no carry is pending (the usual case)
		read_words(w);
		if (w[0] || w[1])
		{
			if (w[0]) canon_finish(_scratch, _cc[0], _cc[1], 0);
			if (w[1]) canon_finish(_scratch2, _cc2[0], _cc2[1], 1);
			be.memset0(_words, 12);
			be.canon_cmp(_scratch, _scratch2, pl.n, _words + 2);
			read_words(w);
		}
		const u32 f = w[2];
		if (!(f & 1u)) return true;
		const bool l_zero = !(f & 2u) || !(f & 4u), r_zero = !(f & 8u) || !(f & 16u);
		return l_zero && r_zero;
	}

	void get_raw(size_t src, void *buf, u32 slab = 0) { check(src); check_slab(slab); blocking("get_data"); be.d2h(buf, R(src) + (size_t)slab * slab_words(), (size_t)pl.n * 8); check_device_error(); }
	void set_raw(size_t dst, const void *buf, u32 slab = 0) { check(dst); check_slab(slab); blocking("set_data"); touch(dst); be.h2d(R(dst) + (size_t)slab * slab_words(), buf, (size_t)pl.n * 8); }
	void get_all(void *buf) { blocking("get_checkpoint"); calm(); materialize(); be.d2h(buf, _reg, _batch * _regs * (size_t)pl.n * 8); check_device_error(); }
	void set_all(const void *buf) { blocking("set_checkpoint"); _fwd_reg = _last_sq = _lazy = (size_t)-1; be.h2d(_reg, buf, _batch * _regs * (size_t)pl.n * 8); }

	// single passes of square_mul, for per-kernel timing (3-pass engines only): 0 strided forward,
	// 1 contiguous middle, 2 strided inverse + carry, 3 carry fix-up
	bool run_pass(size_t r, int which)
	{
		check(r);
		if (pl.small) { if (which != 1) return false; MidArgs M = mid_args(R(r), nullptr, R(r), MODE_SQR, 1); launch_mid(M); return true; }
		materialize();
		touch(r);
		StridedArgs A;
		std::memset(&A, 0, sizeof(A));
		fill_strided_args(pl, A, 0);
		set_pass_tables(A, 0);
		A.src = (const pair_t *)R(r); A.dst = (pair_t *)R(r); A.a = 1; A.slots = _slots;
#if defined(MARIN_DEBUG_SKIP)
		if (const char *s = getenv("MARIN_B200_DBG")) A.dbg = (u32)atoi(s);
#endif
		const size_t smem = (size_t)pl.R * TILE_PITCH * 16;
		if (which == 0) { be.strided_fwd(A, A.tiles, pl.thrStr, smem); return true; }
		if (which == 1) { MidArgs M = mid_args(R(r), nullptr, R(r), MODE_SQR, 1); launch_mid(M); return true; }
		if (which == 2 || which == 3)
		{
			// 2: inverse pass + carry; 3: the same with the next forward strided pass fused in (chained squarings)
			A.with_carry = 1; A.flags = _flags; A.epoch = next_epoch(); A.fwd_dst = (which == 3) ? (pair_t *)_fwd : nullptr; A.skip_dst = (which == 3); set_ticket(A);
			_fwd_reg = (size_t)-1;
			be.strided_inv(A, A.tiles, pl.thrStr, smem); return true;
		}
		if (which == 4) { A.with_carry = 0; be.strided_inv(A, A.tiles, pl.thrStr, smem); return true; } // plain inverse pass (timing aid)
		return false;
	}

	u32 width(u32 k) const { return pl.wbase + ((_wbits[k >> 5] >> (k & 31)) & 1u); }
	u64 *R(size_t r) { return _reg + r * (size_t)pl.n; }                  // register r of slab 0
	size_t slab_words() const { return _regs * (size_t)pl.n; }
	const Tables &tables() const { return _t; }

private:
	size_t _regs;
	u32 _batch = 1; size_t _ncarry = 0;
	void check_slab(u32 slab) const { if (slab >= _batch) throw std::runtime_error("marin_b200: slab index out of range"); }
	u64 *_hid = nullptr;
	u64 *_reg = nullptr, *_carry = nullptr, *_carry2 = nullptr, *_scratch = nullptr, *_scratch2 = nullptr, *_cc[2] = {nullptr, nullptr}, *_cc2[2] = {nullptr, nullptr};
	u32 *_words = nullptr, _canon_D = 16;
	pair_t *_slots = nullptr; size_t _nslots = 0;
	u32 *_flags = nullptr, _epoch = 0;
	// speculation for squaring chains: _fwd holds the forward strided pass of register _fwd_reg's current value
	u64 *_fwd = nullptr;
	size_t _fwd_reg = (size_t)-1, _last_sq = (size_t)-1;
	static constexpr u32 CHAIN_PERSIST = 64; // chained squarings after which the parked transform is pinned in L2
	u32 _chain_len = 0;
	size_t _lazy = (size_t)-1;   // register whose digits exist only as the parked transform in _fwd (see square_mul)
	u64 _inv_R = 0;
	size_t _nflags = 0;
	unsigned long long _ticket_next = 0;
	Tables _t;
	const u64 *_twR = nullptr, *_twRi = nullptr, *_twR2 = nullptr, *_twR2i = nullptr;
	const u32 *_rowK = nullptr, *_rowK2 = nullptr;
	std::vector<u32> _wbits;

	// every device allocation of the engine: owned from the moment it exists, so a constructor that throws half-way
	// (out of memory on a later table) leaves nothing behind -- Allocs frees in its own destructor, which runs for a
	// partially constructed engine too
	struct Allocs
	{
		Backend *be = nullptr; std::vector<void *> v;
		~Allocs() { if (be) for (void *q : v) be->free(q); }
	} _allocs;
	void *dev_alloc(size_t bytes) { _allocs.be = &be; void *d = be.alloc(bytes ? bytes : 8); _allocs.v.push_back(d); return d; }

	void *upload(const void *src, size_t bytes)
	{
		void *d = dev_alloc(bytes ? bytes : 8);
		if (bytes) be.h2d(d, src, bytes);
		return d;
	}

	bool _recording = false;
	std::vector<ProgOp> _rec;
	bool rec(u32 code, size_t dst, size_t src, u32 a) { if (!_recording) return false; _rec.push_back(ProgOp{code, (u32)dst, (u32)src, a, {0, 0, 0, 0}}); return true; }
	size_t mid_smem() const { return (size_t)smem_pairs_lin(pl.S) * 16 + (pl.small ? ((size_t)pl.n / pl.D) * 8 : 0); }
	void blocking(const char *what) const { if (_recording) throw std::runtime_error(std::string("marin_b200: ") + what + " is not allowed while a graph is being recorded"); }
	void check_range(size_t r) const { if (r >= _regs) throw std::runtime_error("marin_b200: register index out of range"); }
	void check(size_t r) { check_range(r); calm(); if (r == _lazy) materialize(); }
	// any operation other than the next squaring of a chain: the chain is broken, give the L2 set-aside back to everybody else
	void calm() { if (_chain_len) { _chain_len = 0; be.persist_off(); } }
	void check_a(u32 a) const
	{
		if (a == 0 || a > pl.a_max)
			throw std::runtime_error("marin_b200: multiplier a = " + std::to_string(a) + " outside [1, " + std::to_string(pl.a_max) + "] for this transform");
	}

	// arith.h:75-81
	static u32 adc(u64 lhs, u32 w, u64 &carry)
	{
		const u64 s = lhs + carry, c = (s < lhs) ? 1 : 0;
		carry = (s >> w) + (c << (64 - w));
		return (u32)s & ((1u << w) - 1);
	}

	MidArgs mid_args(u64 *x, const u64 *y, u64 *dst, int mode, u32 a)
	{
		MidArgs M;
		std::memset(&M, 0, sizeof(M));
		fill_mid_args(pl, M);
		M.t = _t; M.x = (pair_t *)x; M.y = (const pair_t *)y; M.dst = (pair_t *)dst; M.mode = mode; M.a = a; M.carry = _carry;
		if (_batch > 1) { M.slab_pairs = (u32)(slab_words() / 2); M.slab_carry = (u32)_ncarry; }
		return M;
	}

	void launch_mid(const MidArgs &M)
	{
		be.mid(M, pl.m / pl.S, pl.thrMid, mid_smem(), _batch);
	}

	// any write to a register other than by square_mul drops the parked forward transform of it
	void touch(size_t r) { if (_fwd_reg == r) _fwd_reg = (size_t)-1; if (_last_sq == r) _last_sq = (size_t)-1; }

	void set_pass_tables(StridedArgs &A, int level)
	{
		A.t = _t;
		A.tw = level ? _twR2 : _twR; A.twi = level ? _twR2i : _twRi; A.rowK = level ? _rowK2 : _rowK;
	}

	// level 0 = strided pass over the whole register (the inverse one carries), level 1 = second strided
	// pass inside each of the R blocks (only for transforms with more rows than one tile can hold)
	void strided(bool inv, int level, const u64 *src, u64 *dst, u32 a, u64 *fwd_dst, bool skip_dst = false, bool exact_inverse = false, u64 *dst2 = nullptr, const u64 *add_src = nullptr)
	{
		StridedArgs A;
		std::memset(&A, 0, sizeof(A));
		fill_strided_args(pl, A, level);
		set_pass_tables(A, level);
		A.src = (const pair_t *)src; A.dst = (pair_t *)dst; A.a = a; A.slots = _slots;
#if defined(MARIN_DEBUG_SKIP)
		if (const char *s = getenv("MARIN_B200_DBG")) A.dbg = (u32)atoi(s);
#endif
		const size_t smem = (size_t)A.R * TILE_PITCH * 16;
		const u32 grid = A.tiles * (level ? pl.R : 1), thr = level ? pl.thrStr2 : pl.thrStr;
		if (!inv) { be.strided_fwd(A, grid, thr, smem); return; }
		if (exact_inverse) { be.strided_inv(A, grid, thr, smem); return; } // plain, unnormalised: the caller scales by 1/R
		A.skip_dst = skip_dst ? 1u : 0u;
		A.dst2 = (pair_t *)dst2; A.add_src = (const pair_t *)add_src;
		if (level == 0) { A.with_carry = 1; A.flags = _flags; A.epoch = next_epoch(); A.fwd_dst = (pair_t *)fwd_dst; set_ticket(A); }
		be.strided_inv(A, grid, thr, smem);
	}

	// tile numbers of a carry pass come from a 64-bit counter that only grows: launch k hands out
	// [k * tiles, (k + 1) * tiles)
	void set_ticket(StridedArgs &A)
	{
		const size_t off = ((size_t)(pl.R2 * pl.S / TILE_C) + 2) & ~(size_t)1; // 8-byte aligned slot after the error word
		A.ticket = (unsigned long long *)(_flags + off);
		A.ticket_base = _ticket_next;
		_ticket_next += A.tiles;
	}

	u32 next_epoch()
	{
		if (++_epoch == 0) { be.memset0(_slots, _nslots * 16); _epoch = 1; } // wrapped: start over from clean slots
		return _epoch;
	}

	// only_slab >= 0: the operation touches that slab alone (set_digits); otherwise every slab, in lockstep
	void lin(int op, size_t dst, size_t src, u32 a, const u64 *digits, int only_slab = -1)
	{
		touch(dst);
		LinArgs L{_t, R(dst), R(src), digits, _carry, op, a, pl.D};
		u32 b = _batch;
		if (only_slab >= 0)
		{
			const size_t o = (size_t)only_slab * slab_words();
			L.dst += o; L.src += o; L.carry += (size_t)only_slab * _ncarry; b = 1;
		}
		else if (_batch > 1) { L.slab_words = (u32)slab_words(); L.slab_carry = (u32)_ncarry; }
		if (op == LIN_ADD || op == LIN_SUBREG || op == LIN_SUBSMALL) { be.lin_pairs(L, b); return; }
		be.lin_p1(L, pl.n / pl.D, b);
		be.lin_p2(L, pl.n / pl.D, b);
	}
};

} // namespace marin
