// marin_engine.cu -- CUDA backend and C ABI of the B200 Marin engine (sm_100a only).
//
// Replaces src/marin/gpu.cpp + include/marin/ocl.h + the launch plumbing of class gpu
// (include/marin/engine_gpu.h:36-1249): context/queue -> one CUDA stream per engine, JIT'd OpenCL
// program -> the kernels below (compiled ahead of time for sm_100a), cl_mem slabs -> one cudaMalloc
// slab of reg_count * n words.  There is no CPU fallback: without a CUDA device creation fails.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "marin_engine_core.hpp"
#include "marin_fast.cuh"
#include "../../include/marin_b200.h"

using namespace marin;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw std::runtime_error(std::string("cuda error: ") + cudaGetErrorString(e_) + " at " #call); } while (0)

// ------------------------------------------------------------------ kernels

#ifndef MARIN_MID_MAXT
#define MARIN_MID_MAXT 256
#define MARIN_MID_MINB 3
#endif
#ifndef MARIN_STR_MAXT
#define MARIN_STR_MAXT 256
#endif
#ifndef MARIN_STR_MINB
#define MARIN_STR_MINB 2
#endif

__global__ void __launch_bounds__(MARIN_MID_MAXT, MARIN_MID_MINB) k_mid(const __grid_constant__ MidArgs A0)
{
	if (MARIN_PDL_MODE & 1) MARIN_PDL_TRIGGER();
	MARIN_PDL_WAIT();
	const MidArgs A = slab_of(A0, blockIdx.y);
	extern __shared__ __align__(16) unsigned char smraw[];
	pair_t *sm = reinterpret_cast<pair_t *>(smraw);
	u64 *smc = reinterpret_cast<u64 *>(sm + smem_pairs_lin(A.S));
	const u32 nph = mid_num_phases(A);
	for (u32 ph = 0; ph < nph; ++ph)
	{
		mid_phase(A, blockIdx.x, ph, threadIdx.x, blockDim.x, sm, smc);
		__syncthreads();
	}
	if (!(MARIN_PDL_MODE & 1)) MARIN_PDL_TRIGGER();
}

template <bool INV> __global__ void __launch_bounds__(MARIN_STR_MAXT, MARIN_STR_MINB) k_strided(const __grid_constant__ StridedArgs A)
{
	if (MARIN_PDL_MODE & 2) MARIN_PDL_TRIGGER();
	MARIN_PDL_WAIT();
	extern __shared__ __align__(16) unsigned char smraw[];
	pair_t *sm = reinterpret_cast<pair_t *>(smraw);
	__shared__ u32 s_tile;
	u32 blk = blockIdx.x;
	if (INV && A.ticket)
	{
		// decoupled look-back needs "the tile I wait for has started": take tile numbers from an atomic ticket
		if (threadIdx.x == 0) s_tile = (u32)(atomicAdd(A.ticket, 1ull) - A.ticket_base);
		__syncthreads();
		blk = s_tile;
	}
	const u32 nph = strided_num_phases(A, INV);
	for (u32 ph = 0; ph < nph; ++ph)
	{
		strided_phase<INV>(A, blk, ph, threadIdx.x, blockDim.x, sm);
		__syncthreads();
	}
	if (!(MARIN_PDL_MODE & 2)) MARIN_PDL_TRIGGER();
}

__global__ void k_lin_p1(const __grid_constant__ LinArgs A0, u32 count)
{
	const LinArgs A = slab_of(A0, blockIdx.y);
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) lin_p1_thread(A, i);
}
__global__ void k_lin_p2(const __grid_constant__ LinArgs A0, u32 count)
{
	const LinArgs A = slab_of(A0, blockIdx.y);
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) lin_p2_thread(A, i);
}
// carry into the first pair of every block, taken from the ORIGINAL register content (runs before k_lin_pairs)
__global__ void k_lin_boundary(const __grid_constant__ LinArgs A0, u32 nblocks)
{
	const LinArgs A = slab_of(A0, blockIdx.y);
	const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblocks) return;
	const u32 npairs = A.t.n / 2;
	const u32 first = b * LIN_BLOCK;
	A.carry[b] = lin_pair_carry(A, (first == 0) ? npairs - 1 : first - 1);
}

__global__ void __launch_bounds__(LIN_BLOCK) k_lin_pairs(const __grid_constant__ LinArgs A0)
{
	const LinArgs A = slab_of(A0, blockIdx.y);
	__shared__ u64 cs[LIN_BLOCK];
	const u32 npairs = A.t.n / 2;
	const u32 i = blockIdx.x * LIN_BLOCK + threadIdx.x;
	u64 u0 = 0, u1 = 0; u32 w0 = 1, w1 = 1;
	if (i < npairs) { lin_pair_compute(A, i, u0, u1, w0, w1); cs[threadIdx.x] = u1 >> w1; }
	__syncthreads();
	// a register that fits one block wraps its own end-around carry (no boundary kernel needed)
	if (i < npairs) lin_pair_finish(A, i, u0, u1, w0, w1, threadIdx.x ? cs[threadIdx.x - 1] : ((gridDim.x == 1) ? cs[npairs - 1] : A.carry[blockIdx.x]));
}

// S = A + B and D = A - B (+ copies) in one pass over the registers (marin.cl:1891-2154)
__global__ void k_addsub_boundary(const __grid_constant__ LinArgs A0, u32 nblocks)
{
	const LinArgs A = slab_of(A0, blockIdx.y);
	const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblocks) return;
	const u32 npairs = A.t.n / 2;
	const u32 first = b * LIN_BLOCK;
	addsub_pair_carry(A, (first == 0) ? npairs - 1 : first - 1, A.carry[b], A.carry2[b]);
}
__global__ void __launch_bounds__(LIN_BLOCK) k_addsub_pairs(const __grid_constant__ LinArgs A0)
{
	const LinArgs A = slab_of(A0, blockIdx.y);
	__shared__ u64 cs[LIN_BLOCK], cd[LIN_BLOCK];
	const u32 npairs = A.t.n / 2;
	const u32 i = blockIdx.x * LIN_BLOCK + threadIdx.x;
	u64 us[2] = {0, 0}, ud[2] = {0, 0}; u32 w0 = 1, w1 = 1;
	if (i < npairs) { addsub_pair_compute(A, i, us, ud, w0, w1); cs[threadIdx.x] = us[1] >> w1; cd[threadIdx.x] = ud[1] >> w1; }
	__syncthreads();
	if (i < npairs)
	{
		const bool wrap = (gridDim.x == 1);
		const u64 c_s = threadIdx.x ? cs[threadIdx.x - 1] : (wrap ? cs[npairs - 1] : A.carry[blockIdx.x]);
		const u64 c_d = threadIdx.x ? cd[threadIdx.x - 1] : (wrap ? cd[npairs - 1] : A.carry2[blockIdx.x]);
		addsub_pair_finish(A, i, us, ud, w0, w1, c_s, c_d);
	}
}

// canonical digits and comparison on the device (see canon_p1_thread / canon_p2_thread)
__global__ void k_canon_p1(const __grid_constant__ CanonArgs A, u32 count)
{
	const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < count) canon_p1_thread(A, r);
}
__global__ void k_canon_p2(const __grid_constant__ CanonArgs A, u32 count)
{
	const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < count) canon_p2_thread(A, r);
}
__global__ void k_canon_cmp(const u64 *l, const u64 *r, u32 n, u32 *flags)
{
	u32 f = 0;
	for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) f |= canon_cmp_digit(l[j], r[j]);
	f = __reduce_or_sync(0xffffffffu, f);
	if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

// A recorded sequence of engine ops on a single-CTA engine, run by ONE launch: every op is the same phase functions the
// stand-alone kernels run (registers live in global memory).  At these sizes an op is a few microseconds of launch
// latency, which is what an ECM point operation (23-40 ops) otherwise consists of, and inside the kernel it is the serial
// instruction latency of a tiny transform in a handful of threads (staging registers and tables in shared memory was
// tried and is no faster).  So the CTA is A.groups warp groups and the list is a schedule over them (schedule_program):
// a group runs the ops that carry its number, in list order, each with the group's own argument blocks, scratch in
// shared memory and named barrier between the phases; before an op starts, the group's first thread waits until the
// other groups have completed as many ops as the op's conflicts ask for (counters in shared memory; release / acquire
// by __threadfence_block around them, the group barrier passes the order on to the other threads).
// (immediate barrier numbers: with the number in a register ptxas reserves all 16 barriers for every CTA, and barriers
// are a per-SM resource that limits how many curves' CTAs are resident)
__device__ __forceinline__ void group_sync(u32 g, u32 gthr)
{
	if (gthr == 32) { __syncwarp(); return; }                  // a group of one warp
	switch (g)
	{
		case 0: asm volatile("bar.sync 1, %0;" :: "r"(gthr) : "memory"); break;
		case 1: asm volatile("bar.sync 2, %0;" :: "r"(gthr) : "memory"); break;
		case 2: asm volatile("bar.sync 3, %0;" :: "r"(gthr) : "memory"); break;
		default: asm volatile("bar.sync 4, %0;" :: "r"(gthr) : "memory"); break;
	}
}
static_assert(PROG_MAX_GROUPS == 4, "group_sync names four barriers");

__global__ void __launch_bounds__(512, 1) k_program(const __grid_constant__ ProgArgs A)
{
	extern __shared__ __align__(16) unsigned char smraw[];
	__shared__ LinArgs sLs[PROG_MAX_GROUPS];
	__shared__ volatile u32 s_done[PROG_MAX_GROUPS];          // ops completed, per group of the schedule
	const u32 gthr = A.group_threads, grp = threadIdx.x / gthr, tid = threadIdx.x - grp * gthr, ngr = A.groups;
	const u32 n = A.M.t.n, npairs = n / 2;
	pair_t *sm = reinterpret_cast<pair_t *>(smraw + (size_t)grp * A.group_smem);
	u64 *smc = reinterpret_cast<u64 *>(sm + smem_pairs_lin(A.M.S));
	u64 *const carry = reinterpret_cast<u64 *>(sm);           // linear ops: one word per pair / per run, in the idle transform buffer
	LinArgs &sL = sLs[grp];
	// batched engines (one ECM curve per CTA): CTA x runs the same op list on its own slab of registers
	u64 *const reg = A.reg + (size_t)blockIdx.x * A.slab_words, *const hid = A.hid + (size_t)blockIdx.x * A.nhid * A.M.t.n;
	if (tid == 0) { sL.t = A.M.t; sL.carry = carry; sL.D = A.D; sL.digits = nullptr; sL.slab_words = 0; sL.srcA = nullptr; sL.dst2 = sL.cpy = sL.cpy2 = sL.carry2 = nullptr; }
	// programmatic dependent launch: the next replay's CTA may become resident and stage its op list while this one runs;
	// nothing a predecessor wrote (the registers) is touched before the wait below
	MARIN_PDL_TRIGGER();
	if (threadIdx.x < PROG_MAX_GROUPS) s_done[threadIdx.x] = 0;
	// the op list itself: every group walks all of it (to find its own ops), so a short one is staged in shared memory
	__shared__ ProgOp s_ops[PROG_SMEM_OPS];
	const bool staged = A.nops <= PROG_SMEM_OPS;
	if (staged)
		for (u32 w = threadIdx.x; w < A.nops * (u32)(sizeof(ProgOp) / 8); w += blockDim.x) reinterpret_cast<u64 *>(s_ops)[w] = reinterpret_cast<const u64 *>(A.ops)[w];
	MARIN_PDL_WAIT();
	__syncthreads();
	const ProgOp *const ops = staged ? s_ops : A.ops;
	for (u32 q = 0; q < A.nops; ++q)
	{
		const u32 sgq = (ops[q].code >> PO_GROUP_SHIFT) & (PROG_MAX_GROUPS - 1);
		if ((sgq & (ngr - 1)) != grp) continue;                 // fewer groups in the launch than in the schedule: they share (ngr is a power of two)
		const ProgOp op = ops[q];
		const u32 code = op.code & PO_CODE_MASK, sg = (op.code >> PO_GROUP_SHIFT) & (PROG_MAX_GROUPS - 1);
		u64 *dst = prog_reg(A, reg, hid, op.dst);
		const u64 *src = prog_reg(A, reg, hid, op.src);
		if (tid == 0)
		{
			#pragma unroll
			for (u32 g = 0; g < PROG_MAX_GROUPS; ++g)
				if (op.wait[g])
				{
					// (bounded: a schedule that could not make progress raises the engine's error word instead of hanging the GPU)
					u32 spins = 0;
					while (s_done[g] < op.wait[g]) { if (++spins > 64) MARIN_NANOSLEEP(20); if (spins > (1u << 22)) { MARIN_ST_VOLATILE_U32(A.err, 1u); break; } }
				}
			__threadfence_block();
			switch (code)
			{
			case PO_ADD: case PO_SUBREG: case PO_SUBSMALL:
				sL.dst = dst; sL.src = src; sL.a = op.a; sL.op = (code == PO_ADD) ? LIN_ADD : ((code == PO_SUBREG) ? LIN_SUBREG : LIN_SUBSMALL);
				break;
			case PO_SETSMALL:
				sL.dst = dst; sL.src = dst; sL.a = op.a; sL.op = LIN_SETSMALL;
				break;
			default: break;
			}
		}
		group_sync(grp, gthr);
		switch (code)
		{
		case PO_COPY:
			for (u32 i = tid; i < npairs; i += gthr) reinterpret_cast<pair_t *>(dst)[i] = reinterpret_cast<const pair_t *>(src)[i];
			break;
		case PO_SQR: case PO_MUL: case PO_SETMUL:
		{
			// the engine's tables and stages straight from the parameter space, the op's registers in registers
			const MidOp o{(pair_t *)((code == PO_SETMUL) ? const_cast<u64 *>(src) : dst), (code == PO_MUL) ? (const pair_t *)src : nullptr, (pair_t *)dst,
				(code == PO_SQR) ? MODE_SQR : ((code == PO_MUL) ? MODE_MUL : MODE_FWD), op.a};
			const u32 nph = mid_num_phases(A.M, o.mode);
			for (u32 ph = 0; ph < nph; ++ph) { mid_phase<true>(A.M, o, 0, ph, tid, gthr, sm, smc); if (ph + 1 < nph) group_sync(grp, gthr); }   // no radix-16 stage in a single-CTA engine
			break;
		}
		case PO_ADD: case PO_SUBREG: case PO_SUBSMALL:
		{
			for (u32 i = tid; i < npairs; i += gthr) carry[i] = lin_pair_carry(sL, i);      // from the ORIGINAL content
			group_sync(grp, gthr);
			for (u32 i = tid; i < npairs; i += gthr)
			{
				u64 u0, u1; u32 w0, w1;
				lin_pair_compute(sL, i, u0, u1, w0, w1);
				lin_pair_finish(sL, i, u0, u1, w0, w1, carry[i ? i - 1 : npairs - 1]);
			}
			break;
		}
		default: // PO_SETSMALL
		{
			const u32 runs = n / A.D;
			for (u32 r = tid; r < runs; r += gthr) lin_p1_thread(sL, r);
			group_sync(grp, gthr);
			for (u32 r = tid; r < runs; r += gthr) lin_p2_thread(sL, r);
			break;
		}
		}
		group_sync(grp, gthr);   // every thread of the group is through: the op is complete (and the group's scratch is free)
		if (tid == 0) { __threadfence_block(); s_done[sg] = s_done[sg] + 1; }
	}
}

__global__ void k_unweight(const __grid_constant__ Tables t, const u64 *x, u64 *out, u32 n)
{
	const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < n) unweight_thread(t, x, out, j);
}
__global__ void k_scale(u64 *x, u64 f, u32 n)
{
	scale_thread(x, f, blockIdx.x * blockDim.x + threadIdx.x, n);
}
__global__ void k_gl_add(u64 *dst, const u64 *a, const u64 *b, u32 n)
{
	gl_add_thread(dst, a, b, blockIdx.x * blockDim.x + threadIdx.x, n);
}
__global__ void k_weights(const __grid_constant__ WeightArgs A, u32 n)
{
	const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < n) weight_thread(A, j);
}

// Load-time self-test of the field arithmetic: the device paths of gl64.cuh are carry-chain PTX whose correctness rests
// on how ptxas lowers add.cc / sub.cc / madc sequences (a borrow after sub.cc is read by madc as "carry = no borrow"); the
// host paths are plain C.  Every engine's first creation in a process runs both on edge values and refuses to go on if
// they differ -- a toolkit that lowers the chains differently fails here, loudly, instead of producing wrong residues.
// (Its first run caught one: an unused helper whose subc-after-add.cc mask came out inverted on the device.  Removed.)
__global__ void k_selftest(const u64 *a, const u64 *b, u64 *out, u32 cnt)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cnt) return;
	const u64 x = a[i], y = b[i];
	u64 *o = out + (size_t)i * 12;
	o[0] = gl::add(x, y); o[1] = gl::sub(x, y); o[2] = gl::mul(x, y); o[3] = gl::pminus(x); o[4] = gl::neg(x);
	o[5] = gl::mul_pow2<12>(x); o[6] = gl::mul_pow2<48>(x); o[7] = gl::mul_pow2<60>(x); o[8] = gl::mul_pow2<84>(x); o[9] = gl::mul_pow2<180>(x);
	o[10] = gl::reduce128(x, y); o[11] = fast::mul_small((u32)y, x);
}
static void selftest_host(u64 x, u64 y, u64 *o)
{
	o[0] = gl::add(x, y); o[1] = gl::sub(x, y); o[2] = gl::mul(x, y); o[3] = gl::pminus(x); o[4] = gl::neg(x);
	o[5] = gl::mul_pow2<12>(x); o[6] = gl::mul_pow2<48>(x); o[7] = gl::mul_pow2<60>(x); o[8] = gl::mul_pow2<84>(x); o[9] = gl::mul_pow2<180>(x);
	o[10] = gl::reduce128(x, y); o[11] = gl::mul((u64)(u32)y, x);
}
static void field_selftest(cudaStream_t st)
{
	const u64 P = gl::P;
	const u64 edge[] = {0, 1, 2, 0xffffffffull, 0x100000000ull, 0x100000001ull, 0xfffffffeffffffffull, 0xffffffff00000000ull /* P - 1 */,
		P - 2, P - 0xffffffffull, 0x8000000000000000ull, 0x7fffffffffffffffull, 0xfffffffe00000001ull, 0x00000001ffffffffull, 0xdeadbeefcafef00dull % P, 0x0123456789abcdefull};
	const u32 ne = sizeof(edge) / sizeof(edge[0]), cnt = ne * ne;
	std::vector<u64> a(cnt), b(cnt), got((size_t)cnt * 12), want((size_t)cnt * 12);
	for (u32 i = 0; i < ne; ++i) for (u32 j = 0; j < ne; ++j) { a[i * ne + j] = edge[i]; b[i * ne + j] = edge[j]; }
	u64 *da = nullptr, *db = nullptr, *dout = nullptr;
	CK(cudaMalloc(&da, cnt * 8)); CK(cudaMalloc(&db, cnt * 8)); CK(cudaMalloc(&dout, (size_t)cnt * 12 * 8));
	CK(cudaMemcpyAsync(da, a.data(), cnt * 8, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(db, b.data(), cnt * 8, cudaMemcpyHostToDevice, st));
	k_selftest<<<(cnt + 127) / 128, 128, 0, st>>>(da, db, dout, cnt);
	CK(cudaMemcpyAsync(got.data(), dout, (size_t)cnt * 12 * 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	cudaFree(da); cudaFree(db); cudaFree(dout);
	for (u32 k = 0; k < cnt; ++k)
	{
		selftest_host(a[k], b[k], &want[(size_t)k * 12]);
		for (u32 f = 0; f < 12; ++f)
		{
			// add / sub / mul / shifts are defined on canonical inputs only (< P); reduce128 and mul take anything
			const bool canon = a[k] < P && b[k] < P;
			if (!canon && (f == 0 || f == 1 || (f >= 5 && f <= 9) || f == 4 || f == 3 || f == 11)) continue;
			if (got[(size_t)k * 12 + f] != want[(size_t)k * 12 + f])
				throw std::runtime_error("marin_b200: field arithmetic self-test failed (device and host disagree on op " + std::to_string(f) +
					"): this build of the kernels is not usable with this toolkit / driver");
		}
	}
}

// ------------------------------------------------------------------ backend

struct CudaBackend
{
	int dev;
	cudaStream_t st = nullptr;
	uint64_t launches = 0;

	explicit CudaBackend(int device) : dev(device)
	{
		int cnt = 0;
		if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) throw std::runtime_error("marin_b200: no CUDA device (this engine has no CPU path)");
		if (device < 0 || device >= cnt) throw std::runtime_error("marin_b200: device index out of range");
		CK(cudaSetDevice(dev));
		cudaDeviceProp pr;
		CK(cudaGetDeviceProperties(&pr, dev));
		if (pr.major < 10) throw std::runtime_error("marin_b200: built for sm_100a (Blackwell B200) only");
		CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
		const int maxsm = (int)pr.sharedMemPerBlockOptin - 1024; // room for the kernels' few static shared words
		CK(cudaFuncSetAttribute(k_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
		CK(cudaFuncSetAttribute(k_program, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm - 8192)); // it also keeps the argument blocks of its warp groups in static shared memory
		CK(cudaFuncSetAttribute(k_strided<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
		CK(cudaFuncSetAttribute(k_strided<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
		mid_fast_attr<0>(); mid_fast_attr<1>(); mid_fast_attr<2>(); mid_fast_attr<3>(); mid_fast_attr<4>(); mid_fast_attr<5>(); mid_fast_attr<6>(); mid_fast_attr<7>();
		CK(cudaFuncSetAttribute(fast::k_inv_fast_T4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast::INV_SMEM_T4));
		CK(cudaFuncSetAttribute(fast::k_inv_fast_T4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast::INV_SMEM_T4));
		CK(cudaFuncSetAttribute(fast::k_fwd_fast_T4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast::SC<4>::SMEM));
		use_fast = getenv("MARIN_B200_NO_FAST") == nullptr;
		use_pdl = getenv("MARIN_B200_NO_PDL") == nullptr;
		static bool tested[64] = {false};
		if (dev < 64 && !tested[dev]) { field_selftest(st); tested[dev] = true; }
	}
	bool use_pdl = true;  // MARIN_B200_NO_PDL=1: plain stream order (A/B timing)
	bool use_fast = true; // the specialised kernels of marin_fast.cuh (same results; MARIN_B200_NO_FAST=1 runs the generic ones: tests, A/B timing)

	template <int SHAPE> static void mid_fast_attr()
	{
		const int smem = (int)fast::mid_shape_smem(SHAPE);
		CK(cudaFuncSetAttribute(fast::k_mid_fast<MODE_SQR, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		CK(cudaFuncSetAttribute(fast::k_mid_fast<MODE_MUL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		CK(cudaFuncSetAttribute(fast::k_mid_fast<MODE_FWD, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	}
	template <int SHAPE> void mid_fast_launch(const MidArgs &A, u32 grid)
	{
		const u32 thr = fast::mid_shape_threads(SHAPE); const size_t smem = fast::mid_shape_smem(SHAPE);
		if (A.mode == MODE_SQR) launch_windowed(fast::k_mid_fast<MODE_SQR, SHAPE>, A, grid, thr, smem);
		else if (A.mode == MODE_MUL) launch_windowed(fast::k_mid_fast<MODE_MUL, SHAPE>, A, grid, thr, smem);
		else launch_windowed(fast::k_mid_fast<MODE_FWD, SHAPE>, A, grid, thr, smem);
	}
	// which specialised contiguous-pass kernel serves this row shape, or -1 (the stage list must be the one the kernel hard-codes)
	static int fast_mid_shape(const MidArgs &A)
	{
		if (A.small) return -1;
		auto is = [&](std::initializer_list<u32> r) { if (A.nst != r.size()) return false; u32 i = 0; for (u32 x : r) if (A.st[i++].radix != x) return false; return true; };
		if (A.S == 4096 && is({16, 16, 16})) return (int)fast::MID_SHAPE_4096;
		if (A.S == 5120 && is({5, 4, 16, 16})) return (int)fast::MID_SHAPE_5120;
		if (A.S == 2048 && is({2, 4, 16, 16})) return (int)fast::MID_SHAPE_2048;
		if (A.S == 1024 && is({4, 16, 16})) return (int)fast::MID_SHAPE_1024;
		if (A.S == 2560 && is({5, 2, 16, 16})) return (int)fast::MID_SHAPE_2560;
		if (A.S == 1280 && is({5, 16, 16})) return (int)fast::MID_SHAPE_1280;
		if (A.S == 8192 && is({2, 16, 16, 16})) return (int)fast::MID_SHAPE_8192;
		if (A.S == 10240 && is({5, 2, 4, 16, 16})) return (int)fast::MID_SHAPE_10240;
		return -1;
	}
	static bool fast_str_shape(const StridedArgs &A)
	{
		return TILE_C == 4 && A.R == 1024 && A.nst == 3 && A.st[0].radix == 16 && A.st[1].radix == 16 && A.st[2].radix == 4
			&& A.escale == 1 && (size_t)A.Lb == (size_t)A.R * A.stride && A.tiles == A.stride / TILE_C;
	}
	~CudaBackend()
	{
		cudaSetDevice(dev);
		persist_off();
		for (ProgOp *d : progs) cudaFree(d);
		if (st) cudaStreamDestroy(st);
	}

	// ---- recorded op sequences ("programs") of single-CTA engines: uploaded once, run by one k_program launch each
	// (the reference pays one enqueue per engine op, include/marin/engine_gpu.h:886-1248)
	std::vector<ProgOp *> progs;
	std::vector<u32> prog_len;
	int prog_upload(const std::vector<ProgOp> &ops)
	{
		use();
		ProgOp *d = nullptr;
		CK(cudaMalloc(&d, (ops.size() ? ops.size() : 1) * sizeof(ProgOp)));
		if (!ops.empty()) CK(cudaMemcpy(d, ops.data(), ops.size() * sizeof(ProgOp), cudaMemcpyHostToDevice));
		progs.push_back(d); prog_len.push_back((u32)ops.size());
		return (int)progs.size() - 1;
	}
	void prog_launch(int id, ProgArgs A, u32 batch = 1)
	{
		if (id < 0 || (size_t)id >= progs.size()) throw std::runtime_error("marin_b200: unknown graph id");
		use();
		A.ops = progs[(size_t)id]; A.nops = prog_len[(size_t)id];
		if (A.nops == 0) return;
		// warp groups shorten one slab's chain of ops; a batch that no longer fits the GPU in ONE wave of CTAs with them
		// (hundreds of ECM curves) is after throughput instead: fewer groups per CTA, more CTAs resident
		u32 g = A.groups;
		while (g > 1 && (uint64_t)prog_resident_ctas(g, A.group_threads, A.group_smem) < batch) g >>= 1;
		A.groups = g;
		launch_windowed(k_program, A, batch, g * A.group_threads, (size_t)g * A.group_smem); post();
	}
	int sm_count = 0;
	u32 occ_key[PROG_MAX_GROUPS + 1] = {0, 0, 0, 0, 0}, occ_val[PROG_MAX_GROUPS + 1] = {0, 0, 0, 0, 0};
	u32 prog_resident_ctas(u32 g, u32 thr, u32 smem)
	{
		const u32 key = thr * 65536u + smem / 16u + 1u;
		if (occ_key[g] != key)
		{
			int nb = 0;
			if (!sm_count) CK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
			CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_program, (int)(g * thr), (size_t)g * smem));
			occ_key[g] = key; occ_val[g] = (u32)nb * (u32)sm_count;
		}
		return occ_val[g];
	}

	void use() { CK(cudaSetDevice(dev)); }
	void *alloc(size_t b) { use(); void *p = nullptr; CK(cudaMalloc(&p, b ? b : 8)); return p; }
	void free(void *p) { cudaSetDevice(dev); cudaFree(p); }
	void memset0(void *p, size_t b) { use(); CK(cudaMemsetAsync(p, 0, b, st)); }
	void h2d(void *d, const void *s, size_t b) { use(); CK(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st)); }
	void d2h(void *d, const void *s, size_t b) { use(); CK(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
	void d2d(void *d, const void *s, size_t b) { use(); CK(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, st)); ++launches; }
	void sync() { use(); CK(cudaStreamSynchronize(st)); }
	void sync_nothrow() noexcept { cudaSetDevice(dev); cudaStreamSynchronize(st); }

	// Optional (MARIN_B200_L2_PERSIST=1): pin the parked transform of a squaring chain in L2 with an access-policy
	// window.  It bought the round-1 kernels 1.8 %; with the specialised kernels (which already mark the weight
	// streams and the dead tile input evict-first) it costs 15 % (374 vs 326 us per squaring at n = 2^23: the
	// set-aside starves everything else), and the limit is device-global, so engines sharing a GPU would fight
	// over it.  Off by default.
	void persist(void *p, size_t bytes)
	{
		use();
		if (!getenv("MARIN_B200_L2_PERSIST")) return;
		int maxPersist = 0, maxWindow = 0;
		cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, dev);
		cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, dev);
		if (maxPersist <= 0 || maxWindow <= 0) return;
		persist_bytes = (size_t)maxPersist;
		if (const char *e = getenv("MARIN_B200_L2_PERSIST_MB")) { const size_t mb = (size_t)atoi(e) << 20; if (mb < persist_bytes) persist_bytes = mb; }
		if (persist_bytes == 0) return;
		const size_t winb = bytes < (size_t)maxWindow ? bytes : (size_t)maxWindow;
		memset(&win, 0, sizeof(win));
		win.base_ptr = p;
		win.num_bytes = winb;
		win.hitRatio = (winb <= persist_bytes) ? 1.0f : (float)((double)persist_bytes / (double)winb);
		win.hitProp = cudaAccessPropertyPersisting;
		win.missProp = cudaAccessPropertyStreaming;
		if (getenv("MARIN_B200_VERBOSE")) fprintf(stderr, "marin_b200: L2 persistence window %zu bytes (set-aside %zu, max window %d)\n", winb, persist_bytes, maxWindow);
	}
	size_t persist_bytes = 0;
	void persist_on()
	{
		if (has_win || persist_bytes == 0) return;
		use();
		if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist_bytes) != cudaSuccess) { cudaGetLastError(); persist_bytes = 0; return; }
		has_win = true;
	}
	void persist_off()
	{
		if (!has_win) return;
		use();
		has_win = false;
		cudaCtxResetPersistingL2Cache();
		cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
		cudaGetLastError();
	}
	// The window is a LAUNCH attribute of the transform kernels only: as a stream attribute it also made every other
	// kernel's accesses (copies, linear ops) "streaming" misses and slowed them down by up to 2x.
	cudaAccessPolicyWindow win;
	bool has_win = false;
	// rows of 4096 pairs: see MARIN_PDL_MODE in marin_kernels.cuh
	static bool pdl_shape(const MidArgs &A) { return A.small || A.t.S != 4096; }
	static bool pdl_shape(const StridedArgs &A) { return A.t.S != 4096; }
	template <class X> static bool pdl_shape(const X &) { return true; }
	template <class K, class A> void launch_windowed(K kernel, const A &args, u32 grid, u32 thr, size_t smem, u32 gridy = 1)
	{
		cudaLaunchConfig_t cfg;
		memset(&cfg, 0, sizeof(cfg));
		cfg.gridDim = dim3(grid, gridy); cfg.blockDim = dim3(thr); cfg.dynamicSmemBytes = smem; cfg.stream = st;
		cudaLaunchAttribute at[2];
		u32 na = 0;
		if (has_win) { at[na].id = cudaLaunchAttributeAccessPolicyWindow; at[na].val.accessPolicyWindow = win; ++na; }
		// programmatic dependent launch (MARIN_PDL_* in marin_kernels.cuh): every kernel launched from here waits for its
		// predecessor with griddepcontrol.wait before it reads or writes anything
		if (use_pdl && pdl_shape(args)) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
		cfg.attrs = at; cfg.numAttrs = na;
		CK(cudaLaunchKernelEx(&cfg, kernel, args));
	}
	void post() { ++launches; CK(cudaGetLastError()); }

	static u32 cap(u32 thr, u32 maxt) { return thr > maxt ? maxt : thr; }
	static u32 env_u32(const char *name) { const char *s = getenv(name); return s ? (u32)atoi(s) : 0u; }
	void mid(const MidArgs &A, u32 grid, u32 thr, size_t smem, u32 batch = 1)
	{
		// timing probes only (results are wrong with a truncated grid): MARIN_B200_MID_GRID / MARIN_B200_MID_THREADS
		static const u32 fg = env_u32("MARIN_B200_MID_GRID"), ft = env_u32("MARIN_B200_MID_THREADS");
		if (fg && fg < grid) grid = fg;
		if (ft) thr = ft;
		use();
		const int shape = (use_fast && !ft && batch == 1) ? fast_mid_shape(A) : -1;
		if (shape >= 0)
		{
			switch (shape)
			{
				case 0: mid_fast_launch<0>(A, grid); break;
				case 1: mid_fast_launch<1>(A, grid); break;
				case 2: mid_fast_launch<2>(A, grid); break;
				case 3: mid_fast_launch<3>(A, grid); break;
				case 4: mid_fast_launch<4>(A, grid); break;
				case 5: mid_fast_launch<5>(A, grid); break;
				case 6: mid_fast_launch<6>(A, grid); break;
				default: mid_fast_launch<7>(A, grid); break;
			}
			post(); return;
		}
		launch_windowed(k_mid, A, grid, cap(thr, MARIN_MID_MAXT), smem, batch); post();
	}
	// second-level pass of the n >= 2^25 decompositions m = 1024 x 4 x S: one radix-4 stage, a streaming kernel
	static u32 fast_level1_shape(const StridedArgs &A, u32 grid)   // rows of the streaming second-level kernel that serves A, or 0
	{
		if (A.with_carry || A.fwd_dst || A.escale <= 1 || (size_t)A.Lb != (size_t)A.R * A.stride || A.tiles != A.stride / TILE_C
			|| grid % A.tiles != 0 || grid / A.tiles > 65535) return 0;
		if (A.R == 4 && A.nst == 1 && A.st[0].radix == 4) return 4;
		if (A.R == 8 && A.nst == 2 && A.st[0].radix == 4 && A.st[1].radix == 2) return 8;
		if (A.R == 16 && A.nst == 1 && A.st[0].radix == 16) return 16;
		return 0;
	}
	template <bool INV> void level1_launch(const StridedArgs &A, u32 grid, u32 rows)
	{
		const u32 gx = (A.stride + fast::L1_THREADS - 1) / fast::L1_THREADS, gy = grid / A.tiles;
		if (rows == 4) launch_windowed(fast::k_level1_r4<INV>, A, gx, fast::L1_THREADS, 0, gy);
		else if (rows == 8) launch_windowed(fast::k_level1_r8<INV>, A, gx, fast::L1_THREADS, 0, gy);
		else launch_windowed(fast::k_level1_r16<INV>, A, 2 * gx, fast::L1_THREADS, 0, gy);   // one lane per thread
		post();
	}
	void strided_fwd(const StridedArgs &A, u32 grid, u32 thr, size_t smem)
	{
		use();
		if (const u32 rows = use_fast ? fast_level1_shape(A, grid) : 0u) { level1_launch<false>(A, grid, rows); return; }
		if (use_fast && fast_str_shape(A) && grid == A.tiles) { launch_windowed(fast::k_fwd_fast_T4, A, grid, fast::STR_THREADS, fast::SC<4>::SMEM); post(); return; }
		launch_windowed(k_strided<false>, A, grid, cap(thr, MARIN_STR_MAXT), smem); post();
	}
	void strided_inv(const StridedArgs &A, u32 grid, u32 thr, size_t smem)
	{
		use();
		if (const u32 rows = use_fast ? fast_level1_shape(A, grid) : 0u) { level1_launch<true>(A, grid, rows); return; }
		if (use_fast && A.with_carry && A.ticket && fast_str_shape(A) && grid == A.tiles)
		{
			if (A.fwd_dst) launch_windowed(fast::k_inv_fast_T4<true>, A, grid, fast::STR_THREADS, fast::INV_SMEM_T4);
			else launch_windowed(fast::k_inv_fast_T4<false>, A, grid, fast::STR_THREADS, fast::INV_SMEM_T4);
			post(); return;
		}
		launch_windowed(k_strided<true>, A, grid, cap(thr, MARIN_STR_MAXT), smem); post();
	}
	void lin_pairs(const LinArgs &A, u32 batch = 1)
	{
		use();
		const u32 npairs = A.t.n / 2, nblocks = (npairs + LIN_BLOCK - 1) / LIN_BLOCK;
		if (nblocks > 1) { k_lin_boundary<<<dim3((nblocks + 127) / 128, batch), 128, 0, st>>>(A, nblocks); post(); }
		k_lin_pairs<<<dim3(nblocks, batch), LIN_BLOCK, 0, st>>>(A); post();
	}
	void addsub_pairs(const LinArgs &A, u32 batch = 1)
	{
		use();
		const u32 npairs = A.t.n / 2, nblocks = (npairs + LIN_BLOCK - 1) / LIN_BLOCK;
		if (nblocks > 1) { k_addsub_boundary<<<dim3((nblocks + 127) / 128, batch), 128, 0, st>>>(A, nblocks); post(); }
		k_addsub_pairs<<<dim3(nblocks, batch), LIN_BLOCK, 0, st>>>(A); post();
	}
	// registers of every slab: `rows` copies of `bytes`, `pitch` bytes apart
	void d2d_2d(void *d, const void *s, size_t bytes, size_t pitch, u32 rows)
	{ use(); CK(cudaMemcpy2DAsync(d, pitch, s, pitch, bytes, rows, cudaMemcpyDeviceToDevice, st)); ++launches; }
	void canon_p1(const CanonArgs &A, u32 count) { use(); k_canon_p1<<<(count + 127) / 128, 128, 0, st>>>(A, count); post(); }
	void canon_p2(const CanonArgs &A, u32 count) { use(); k_canon_p2<<<(count + 127) / 128, 128, 0, st>>>(A, count); post(); }
	void canon_cmp(const u64 *l, const u64 *r, u32 n, u32 *flags) { use(); k_canon_cmp<<<592, 256, 0, st>>>(l, r, n, flags); post(); }
	// device -> host without waiting (the caller syncs once for several copies)
	void d2h_async(void *d, const void *s, size_t b) { use(); CK(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, st)); }
	void lin_p1(const LinArgs &A, u32 count, u32 batch = 1) { use(); k_lin_p1<<<dim3((count + 127) / 128, batch), 128, 0, st>>>(A, count); post(); }
	void lin_p2(const LinArgs &A, u32 count, u32 batch = 1) { use(); k_lin_p2<<<dim3((count + 127) / 128, batch), 128, 0, st>>>(A, count); post(); }
	void unweight(const Tables &t, const u64 *x, u64 *out, u32 n) { use(); k_unweight<<<(n + 255) / 256, 256, 0, st>>>(t, x, out, n); post(); }
	void scale(u64 *x, u64 f, u32 n) { use(); k_scale<<<(n + 255) / 256, 256, 0, st>>>(x, f, n); post(); }
	void gl_add(u64 *dst, const u64 *a, const u64 *b, u32 n) { use(); k_gl_add<<<(n + 255) / 256, 256, 0, st>>>(dst, a, b, n); post(); }
	void weights(const WeightArgs &A, u32 n) { use(); k_weights<<<(n + 255) / 256, 256, 0, st>>>(A, n); post(); }
};

typedef EngineCore<CudaBackend> Eng;

// ------------------------------------------------------------------ C ABI

static thread_local std::string g_err;
#define E(h) (static_cast<Eng *>(h))
#define GUARD(body) do { if (!eng) { g_err = "null engine"; return 0; } try { g_err.clear(); body; return 1; } \
	catch (const std::exception &ex) { g_err = ex.what(); return 0; } catch (...) { g_err = "unknown error"; return 0; } } while (0)

extern "C" {

const char *mb200_version(void) { return "marin_b200 0.1 (sm_100a)"; }
const char *mb200_last_error(void) { return g_err.c_str(); }
size_t mb200_transform_size_for(uint32_t p) { return transform_size(p); }

void *mb200_create(uint32_t p, size_t reg_count, uint32_t device, int verbose) { return mb200_create_split(p, reg_count, device, verbose, 0, 0); }

// the same with the decomposition of the transform chosen by the caller (tests and tuning; results do not depend on it)
void *mb200_create_split(uint32_t p, size_t reg_count, uint32_t device, int verbose, uint32_t row_length, uint32_t second_level_rows)
{
	try
	{
		g_err.clear();
		Eng *e = new Eng(p, reg_count, (int)device, row_length, second_level_rows);
		if (verbose)
			fprintf(stderr, "marin_b200: p = %u, n = %u, %s, R = %u x %u, S = %u, %zu registers, device %u\n", p, e->pl.n,
				e->pl.small ? "single-CTA path" : (e->pl.R2 > 1 ? "5-pass path" : "3-pass path"), e->pl.R, e->pl.R2, e->pl.S, reg_count, device);
		return e;
	}
	catch (const std::exception &ex) { g_err = ex.what(); return nullptr; }
	catch (...) { g_err = "unknown error"; return nullptr; }
}

// `batch` independent register sets driven in lockstep, one CTA per set in every launch (single-CTA sizes only)
void *mb200_create_batch(uint32_t p, size_t reg_count, uint32_t device, int verbose, uint32_t batch)
{
	try
	{
		g_err.clear();
		Eng *e = new Eng(p, reg_count, (int)device, 0, 0, batch);
		if (verbose) fprintf(stderr, "marin_b200: p = %u, n = %u, single-CTA path x %u slabs, %zu registers each, device %u\n", p, e->pl.n, e->batch(), reg_count, device);
		return e;
	}
	catch (const std::exception &ex) { g_err = ex.what(); return nullptr; }
	catch (...) { g_err = "unknown error"; return nullptr; }
}
uint32_t mb200_batch(void *eng) { return eng ? E(eng)->batch() : 0; }
int mb200_get_digits_slab(void *eng, uint32_t slab, size_t src, uint64_t *enc, size_t n)
{ GUARD(if (n != E(eng)->size()) throw std::runtime_error("marin_b200: digit count must equal the transform size"); E(eng)->get_digits(src, (u64 *)enc, slab)); }
int mb200_set_digits_slab(void *eng, uint32_t slab, size_t dst, const uint64_t *enc, size_t n)
{ GUARD(if (n != E(eng)->size()) throw std::runtime_error("marin_b200: digit count must equal the transform size"); E(eng)->set_digits(dst, (const u64 *)enc, slab)); }

void mb200_destroy(void *eng) { try { delete E(eng); } catch (...) {} }
size_t mb200_transform_size(void *eng) { return eng ? E(eng)->size() : 0; }
size_t mb200_register_count(void *eng) { return eng ? E(eng)->regs() : 0; }
uint32_t mb200_max_multiplier(void *eng) { return eng ? E(eng)->pl.a_max : 0; }
int mb200_sync(void *eng) { GUARD(E(eng)->sync()); }

int mb200_set_u32(void *eng, size_t dst, uint32_t a) { GUARD(E(eng)->set_u32(dst, a)); }
int mb200_copy(void *eng, size_t dst, size_t src) { GUARD(E(eng)->copy(dst, src)); }
int mb200_square_mul(void *eng, size_t r, uint32_t a) { GUARD(E(eng)->square_mul(r, a)); }
int mb200_set_multiplicand(void *eng, size_t dst, size_t src) { GUARD(E(eng)->set_multiplicand(dst, src)); }
int mb200_add_multiplicand(void *eng, size_t dst, size_t a, size_t b) { GUARD(E(eng)->add_multiplicand(dst, a, b)); }
int mb200_mul(void *eng, size_t dst, size_t src, uint32_t a) { GUARD(E(eng)->mul(dst, src, a)); }
int mb200_sub_u32(void *eng, size_t r, uint32_t a) { GUARD(E(eng)->sub_u32(r, a)); }
int mb200_add(void *eng, size_t dst, size_t src) { GUARD(E(eng)->add(dst, src)); }
int mb200_sub_reg(void *eng, size_t dst, size_t src) { GUARD(E(eng)->sub_reg(dst, src)); }

int mb200_mul_add(void *eng, size_t dst, size_t mul_src, size_t add_src, uint32_t a) { GUARD(E(eng)->mul_add(dst, mul_src, add_src, a)); }
int mb200_addsub(void *eng, size_t s, size_t d, size_t a, size_t b) { GUARD(E(eng)->addsub(s, d, a, b)); }
int mb200_addsub_copy(void *eng, size_t s, size_t d, size_t sc, size_t dc, size_t a, size_t b) { GUARD(E(eng)->addsub(s, d, a, b, sc, dc)); }
int mb200_square_mul_copy(void *eng, size_t r, size_t c, uint32_t a) { GUARD(E(eng)->square_mul_copy(r, c, a)); }
int mb200_mul_copy(void *eng, size_t dst, size_t src, size_t c, uint32_t a) { GUARD(E(eng)->mul_copy(dst, src, c, a)); }

int mb200_equal(void *eng, size_t l, size_t r, int *out) { GUARD(*out = E(eng)->equal(l, r) ? 1 : 0); }

int mb200_get_digits(void *eng, size_t src, uint64_t *enc, size_t n)
{ GUARD(if (n != E(eng)->size()) throw std::runtime_error("marin_b200: digit count must equal the transform size"); E(eng)->get_digits(src, (u64 *)enc)); }
int mb200_set_digits(void *eng, size_t dst, const uint64_t *enc, size_t n)
{ GUARD(if (n != E(eng)->size()) throw std::runtime_error("marin_b200: digit count must equal the transform size"); E(eng)->set_digits(dst, (const u64 *)enc)); }

size_t mb200_register_data_size(void *eng) { return eng ? E(eng)->size() * 8 : 0; }
int mb200_get_raw(void *eng, size_t src, void *buf, size_t bytes)
{ if (!eng || bytes != E(eng)->size() * 8) { g_err = "size mismatch"; return 0; } GUARD(E(eng)->get_raw(src, buf)); }
int mb200_set_raw(void *eng, size_t dst, const void *buf, size_t bytes)
{ if (!eng || bytes != E(eng)->size() * 8) { g_err = "size mismatch"; return 0; } GUARD(E(eng)->set_raw(dst, buf)); }
size_t mb200_checkpoint_size(void *eng) { return eng ? E(eng)->size() * 8 * E(eng)->regs() * E(eng)->batch() : 0; }
int mb200_get_checkpoint(void *eng, void *buf, size_t bytes)
{ if (!eng || bytes != mb200_checkpoint_size(eng)) { g_err = "size mismatch"; return 0; } GUARD(E(eng)->get_all(buf)); }
int mb200_set_checkpoint(void *eng, const void *buf, size_t bytes)
{ if (!eng || bytes != mb200_checkpoint_size(eng)) { g_err = "size mismatch"; return 0; } GUARD(E(eng)->set_all(buf)); }

int mb200_square_mul_repeat(void *eng, size_t r, uint32_t a, uint32_t count)
{ GUARD(for (uint32_t i = 0; i < count; ++i) E(eng)->square_mul(r, a)); }

static void bench_square_mul(Eng *e, size_t r, uint32_t a, uint32_t count, float *ms)
{
	e->be.use();
	cudaEvent_t t0;
	cudaEvent_t t1;
	CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
	CK(cudaEventRecord(t0, e->be.st));
	for (uint32_t i = 0; i < count; ++i) e->square_mul(r, a);
	CK(cudaEventRecord(t1, e->be.st));
	CK(cudaEventSynchronize(t1));
	CK(cudaEventElapsedTime(ms, t0, t1));
	cudaEventDestroy(t0); cudaEventDestroy(t1);
}

int mb200_bench_square_mul(void *eng, size_t r, uint32_t a, uint32_t count, float *ms) { GUARD(bench_square_mul(E(eng), r, a, count, ms)); }

static void bench_pass(Eng *e, size_t r, int which, uint32_t count, float *ms)
{
	e->be.use();
	cudaEvent_t t0;
	cudaEvent_t t1;
	CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
	if (!e->run_pass(r, which)) throw std::runtime_error("marin_b200: this engine has no such pass");
	CK(cudaEventRecord(t0, e->be.st));
	for (uint32_t i = 0; i < count; ++i) e->run_pass(r, which);
	CK(cudaEventRecord(t1, e->be.st));
	CK(cudaEventSynchronize(t1));
	CK(cudaEventElapsedTime(ms, t0, t1));
	cudaEventDestroy(t0); cudaEventDestroy(t1);
}

int mb200_bench_pass(void *eng, size_t r, int which, uint32_t count, float *ms) { GUARD(bench_pass(E(eng), r, which, count, ms)); }
int mb200_is_single_cta(void *eng) { return (eng && E(eng)->pl.small) ? 1 : 0; }

int mb200_graph_begin(void *eng) { GUARD(E(eng)->graph_begin()); }
int mb200_graph_end(void *eng, int *id)
{
	if (!eng) { g_err = "null engine"; return 0; }
	try { g_err.clear(); *id = E(eng)->graph_end(); return 1; }
	catch (const std::exception &ex) { g_err = ex.what(); return 0; }
	catch (...) { g_err = "unknown error"; return 0; }
}
int mb200_graph_launch(void *eng, int id) { GUARD(E(eng)->graph_launch(id)); }

uint64_t mb200_launch_count(void *eng) { return eng ? E(eng)->be.launches : 0; }

int mb200_device_mem_info(uint32_t device, size_t *free_bytes, size_t *total_bytes)
{
	g_err.clear();
	if (cudaSetDevice((int)device) != cudaSuccess || cudaMemGetInfo(free_bytes, total_bytes) != cudaSuccess)
	{ g_err = "cudaMemGetInfo failed"; cudaGetLastError(); return 0; }
	return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// The reference's plugin ABI (third_party/aevum/src/EngineApi.h:28-62), exported by this library as well, so that the
// reference's own run-time loader (src/aevum/EngineAevum.cpp:89-243: dlopen of $AEVUM_ENGINE_LIB, the 19 symbols below)
// and its `engine_aevum` adapter drive the B200 engine UNCHANGED.  Differences from mb200_*: create takes an FFT spec
// and a tune directory (ignored here: the transform size is the planner's), set_multiplicand is called `prepare`, and
// residues cross the boundary as packed little-endian 32-bit words (ceil(p / 32) of them), not digits.
static void words_from_digits(Eng *e, const std::vector<u64> &enc, uint32_t *words, size_t count)
{
	memset(words, 0, count * 4);
	bool all_ones = true;
	for (u32 k = 0; k < e->pl.n; ++k) if ((u32)enc[k] != (u32)((((u64)1) << e->width(k)) - 1)) { all_ones = false; break; }
	if (all_ones) return;                                   // 2^p - 1 is 0
	size_t bit = 0;
	for (u32 k = 0; k < e->pl.n; ++k)
	{
		const u32 w = e->width(k);
		const u64 v = (u32)enc[k];
		const size_t wi = bit / 32; const u32 sh = (u32)(bit % 32);
		if (wi < count) words[wi] |= (uint32_t)(v << sh);
		if (sh + w > 32 && wi + 1 < count) words[wi + 1] |= (uint32_t)(v >> (32 - sh));
		bit += w;
	}
}
static void digits_from_words(Eng *e, const uint32_t *words, size_t count, std::vector<u64> &enc)
{
	size_t bit = 0;
	for (u32 k = 0; k < e->pl.n; ++k)
	{
		const u32 w = e->width(k);
		const size_t wi = bit / 32; const u32 sh = (u32)(bit % 32);
		u64 chunk = (wi < count) ? words[wi] : 0;
		if (wi + 1 < count) chunk |= (u64)words[wi + 1] << 32;
		enc[k] = ((chunk >> sh) & ((((u64)1) << w) - 1)) | ((u64)w << 32);
		bit += w;
	}
}

const char *aevum_engine_version(void) { return "marin_b200 0.2 (sm_100a) behind the Aevum plugin ABI"; }
const char *aevum_engine_last_error(void) { return g_err.c_str(); }
int aevum_engine_resolve_fft(uint32_t exponent, const char *, char *output, size_t output_size)
{
	const size_t n = transform_size(exponent);
	if (n == (size_t)-1) { g_err = "exponent outside the Goldilocks IBDWT range"; return 0; }
	snprintf(output, output_size, "marin-b200:%zu", n);
	return 1;
}
void *aevum_engine_create(uint32_t exponent, size_t register_count, uint32_t device, int verbose, const char *, const char *)
{ return mb200_create(exponent, register_count, device, verbose); }
void aevum_engine_destroy(void *h) { mb200_destroy(h); }
size_t aevum_engine_transform_size(void *h) { return mb200_transform_size(h); }
size_t aevum_engine_word_count(void *h) { return h ? ((size_t)E(h)->pl.p + 31) / 32 : 0; }
int aevum_engine_sync(void *h) { return mb200_sync(h); }
int aevum_engine_set_u32(void *h, size_t dst, uint32_t v) { return mb200_set_u32(h, dst, v); }
int aevum_engine_copy(void *h, size_t dst, size_t src) { return mb200_copy(h, dst, src); }
int aevum_engine_prepare(void *h, size_t dst, size_t src) { return mb200_set_multiplicand(h, dst, src); }
int aevum_engine_square_mul(void *h, size_t reg, uint32_t f) { return mb200_square_mul(h, reg, f); }
int aevum_engine_mul(void *h, size_t dst, size_t src, uint32_t f) { return mb200_mul(h, dst, src, f); }
int aevum_engine_add(void *h, size_t dst, size_t src) { return mb200_add(h, dst, src); }
int aevum_engine_sub_reg(void *h, size_t dst, size_t src) { return mb200_sub_reg(h, dst, src); }
int aevum_engine_sub_u32(void *h, size_t dst, uint32_t v) { return mb200_sub_u32(h, dst, v); }
int aevum_engine_equal(void *h, size_t l, size_t r, int *out) { return mb200_equal(h, l, r, out); }
int aevum_engine_set_words(void *eng, size_t dst, const uint32_t *words, size_t count)
{ GUARD(std::vector<u64> enc(E(eng)->pl.n); digits_from_words(E(eng), words, count, enc); E(eng)->set_digits(dst, enc.data())); }
int aevum_engine_get_words(void *eng, size_t src, uint32_t *words, size_t count)
{ GUARD(std::vector<u64> enc(E(eng)->pl.n); E(eng)->get_digits(src, enc.data()); words_from_digits(E(eng), enc, words, count)); }

void *mb200_host_alloc(size_t bytes) { void *p = nullptr; if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr; return p; }
void mb200_host_free(void *p) { if (p) cudaFreeHost(p); }

}

#ifdef MARIN_PHASE_TIMING
extern "C" void mb200_debug_phase_clocks(long long *out) { cudaMemcpyFromSymbol(out, marin::fast::g_phase_clk, sizeof(marin::fast::g_phase_clk)); }
#endif
