// Microbenchmark (development aid): throughput of the register-resident radix-16 butterfly core of the engine
// (dft16 forward + inverse, optionally with 15 table-free twiddle multiplies), with a check of the device arithmetic
// against the host formulation.  It is how the Goldilocks add / sub / reduce formulations in gl64.cuh were chosen
// (profiles/bench_r1j/micro2.log: all-ALU carry chains 35.0 G butterflies/s, carry consumers as IMAD.X 51.1).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -diag-suppress 550 -o bfly_bench bfly_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../prmers_b200/csrc/marin_kernels.cuh"

using namespace marin;

template <int MODE> __global__ void __launch_bounds__(256) k_bench(u64 *out, const u64 *in, int iters, u64 tw)
{
	u64 x[16];
	const u32 gid = blockIdx.x * blockDim.x + threadIdx.x;
	#pragma unroll
	for (int i = 0; i < 16; ++i) x[i] = in[(gid * 16 + i) & 4095];
	for (int it = 0; it < iters; ++it)
	{
		dft16<false>(x);
		if (MODE == 1) { u64 t = tw;
			#pragma unroll
			for (int k = 1; k < 16; ++k) { x[k] = gl::mul(x[k], t); t += 0x9e3779b97f4a7c15ull; } }
		dft16<true>(x);
	}
	u64 acc = 0;
	#pragma unroll
	for (int i = 0; i < 16; ++i) acc ^= x[i];
	out[gid] = acc;
}

static u64 host_ref(const u64 *in, u32 gid, int iters)
{
	u64 x[16];
	for (int i = 0; i < 16; ++i) x[i] = in[(gid * 16 + i) & 4095];
	for (int it = 0; it < iters; ++it) { dft16<false>(x); dft16<true>(x); }
	u64 acc = 0;
	for (int i = 0; i < 16; ++i) acc ^= x[i];
	return acc;
}

int main(int argc, char **argv)
{
	const int iters = argc > 1 ? atoi(argv[1]) : 2000;
	u64 *in, *out, hin[4096];
	u64 s = 88172645463325252ull;
	for (int i = 0; i < 4096; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; hin[i] = s % gl::P; }
	hin[0] = 0; hin[1] = gl::P - 1; hin[2] = 1; hin[3] = 0xffffffffull; hin[4] = 0xffffffff00000000ull; hin[5] = gl::P - 0xffffffffull;
	cudaMalloc(&in, sizeof(hin)); cudaMalloc(&out, 148 * 8 * 256 * 8);
	cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
	// correctness of this variant against the host formulation (2 iterations)
	k_bench<0><<<4, 256>>>(out, in, 2, 0);
	u64 hout[1024];
	cudaMemcpy(hout, out, sizeof(hout), cudaMemcpyDeviceToHost);
	int bad = 0;
	for (u32 g = 0; g < 1024; ++g) if (hout[g] != host_ref(hin, g, 2)) ++bad;
	printf("device arithmetic %s\n", bad ? "MISMATCHES the host formulation" : "matches the host formulation");
	cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
	for (int mode = 0; mode < 2; ++mode)
		for (int cps = 1; cps <= 4; ++cps) // CTAs (of 8 warps) per SM
		{
			const int grid = 148 * cps;
			float best = 1e30f;
			for (int rep = 0; rep < 3; ++rep)
			{
				cudaEventRecord(t0);
				if (mode == 0) k_bench<0><<<grid, 256>>>(out, in, iters, 3); else k_bench<1><<<grid, 256>>>(out, in, iters, 3);
				cudaEventRecord(t1); cudaEventSynchronize(t1);
				float ms; cudaEventElapsedTime(&ms, t0, t1); if (ms < best) best = ms;
			}
			const double bf = (double)grid * 256 * iters * 2; // butterflies
			printf("  mode %d (%s) %d CTA/SM: %.3f ms  %.2f Gbfly/s  %.1f ns/bfly/SM-warp-slot\n", mode, mode ? "dft16+15mul+idft16" : "dft16+idft16", cps, best, bf / best / 1e6,
				best * 1e6 / (iters * 2.0));
		}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) printf("cuda error %s\n", cudaGetErrorString(e));
	return bad;
}
