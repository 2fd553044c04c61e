// probes.cu -- measurement and self-test probes of libadmm_b200.so that need no solver context.
//
//  admmb_probe_fp64        the FP64 roofline denominators of the local step, MEASURED on the device the bench runs on:
//                          issue rate of independent DFMA / DADD / DMUL chains (the local-step translation unit is compiled
//                          with -fmad=false to stay bit-exact with the reference's x86 arithmetic, so the kernel can only
//                          ever reach the DADD / DMUL rate, half the DFMA flop rate) and the dependent-issue latency of
//                          a single DFMA chain (what a warp with no instruction-level parallelism pays per operation).
//  admmb_debug_fastmath_selftest
//                          div_by / rcp_x / sqrt_x of elastic_math.h (the exact fast paths with their shared fallback)
//                          against the plain operators `/` and sqrt() on pseudo-random operands of every class (random bit
//                          patterns incl. NaN / inf / subnormals, moderate magnitudes, values near 1, equal operands,
//                          zeros): any result that differs in a bit is counted.
#include <cuda_runtime.h>

#include <cstdio>

#include "elastic_math.h"
#include "../../include/admm_b200.h"

namespace admmb {

template <int KIND, int CHAINS>
__global__ void __launch_bounds__(256) k_fp64_rate(int iters, double seed, double *sink) {
	double acc[CHAINS];
	const double a = 1.0 + seed * 1e-9, b = seed * 1e-12;
#pragma unroll
	for (int c = 0; c < CHAINS; ++c) acc[c] = seed + c + threadIdx.x * 1e-3;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int c = 0; c < CHAINS; ++c) {
			if (KIND == 0) acc[c] = __fma_rn(acc[c], a, b);
			else if (KIND == 1) acc[c] = __dadd_rn(acc[c], b);
			else acc[c] = __dmul_rn(acc[c], a);
		}
	}
	double s = 0.0;
#pragma unroll
	for (int c = 0; c < CHAINS; ++c) s += acc[c];
	if (s == 12345.678) sink[0] = s; // never true: keeps the chains alive
}

__global__ void k_fp64_latency(int iters, double seed, double *sink, long long *cycles) {
	double acc = seed;
	const double a = 1.0 + seed * 1e-9, b = seed * 1e-12;
	const long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int c = 0; c < 16; ++c) acc = __fma_rn(acc, a, b);
	}
	const long long t1 = clock64();
	if (threadIdx.x == 0) cycles[0] = t1 - t0;
	if (acc == 12345.678) sink[0] = acc;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long &s) {
	unsigned long long z = (s += 0x9e3779b97f4a7c15ULL);
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return z ^ (z >> 31);
}
// operand classes: 0 any bit pattern, 1 moderate magnitude 2^-40 .. 2^40, 2 within 2^-20 of +-1, 3 tiny / huge exponents,
// 4 exact zeros and powers of two
__device__ double operand(unsigned long long &s, int cls) {
	const unsigned long long r = splitmix64(s);
	const unsigned long long mant = r & 0x000fffffffffffffULL, sign = r & 0x8000000000000000ULL;
	switch (cls) {
	case 0: return __longlong_as_double((long long)r);
	case 1: return __longlong_as_double((long long)(sign | ((unsigned long long)(1023 - 40 + (r >> 52) % 81) << 52) | mant));
	case 2: return __longlong_as_double((long long)(sign | (1023ULL << 52) | (mant >> 20))) - ((r >> 60 & 1) ? 0.0 : __longlong_as_double((long long)((1023ULL - 21) << 52 | (mant >> 3))));
	case 3: { const unsigned long long e = (r >> 52 & 1) ? (r >> 53) % 80 : 2046 - (r >> 53) % 80; return __longlong_as_double((long long)(sign | (e << 52) | mant)); }
	default: return (r >> 52 & 3) == 0 ? __longlong_as_double((long long)sign) : __longlong_as_double((long long)(sign | ((unsigned long long)(1023 - 30 + (r >> 54) % 61) << 52)));
	}
}
__device__ __forceinline__ bool same_bits(double a, double b) {
	return __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b);
}

// counts[0..3]: mismatches of a/y, (a,b,c)/y with one shared reciprocal, 1/x, sqrt(x); counts[4..7]: how often each took
// its fallback (information only)
__global__ void __launch_bounds__(256) k_fastmath_selftest(unsigned long long seed, int per_thread, unsigned long long *counts) {
	unsigned long long s = seed + 0x632be59bd9b4e019ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
	unsigned long long bad_div = 0, bad_div3 = 0, bad_rcp = 0, bad_sqrt = 0, fb_div = 0, fb_div3 = 0, fb_rcp = 0, fb_sqrt = 0;
	for (int i = 0; i < per_thread; ++i) {
		const int ca = (int)(splitmix64(s) % 5), cy = (int)(splitmix64(s) % 5);
		double a = operand(s, ca), y = operand(s, cy);
		const double b = operand(s, ca), c = operand(s, (ca + 1) % 5);
		if ((i & 15) == 0) a = y;                       // equal operands
		if ((i & 15) == 1) a = y * 3.0;
		if ((i & 15) == 2) a = 0.0;                     // exact zeros of both signs over every class of denominator
		if ((i & 15) == 3) a = -0.0;
		{
			bool bad = false;
			double q = div_by(a, recip_of(y), bad);
			if (bad) { q = ref_div(a, y); ++fb_div; }
			if (!same_bits(q, a / y)) ++bad_div;
		}
		{
			bool bad = false;
			const Recip R = recip_of(y);
			double q0 = div_by(a, R, bad), q1 = div_by(b, R, bad), q2 = div_by(c, R, bad);
			if (bad) { q0 = ref_div(a, y); q1 = ref_div(b, y); q2 = ref_div(c, y); ++fb_div3; }
			if (!same_bits(q0, a / y) || !same_bits(q1, b / y) || !same_bits(q2, c / y)) ++bad_div3;
		}
		{
			bool bad = false;
			double r = rcp_x(y, bad);
			if (bad) { r = ref_div(1.0, y); ++fb_rcp; }
			if (!same_bits(r, 1.0 / y)) ++bad_rcp;
		}
		{
			bool bad = false;
			const double x = (i & 1) ? fabs(a) : a;
			double r = sqrt_x(x, bad);
			if (bad) { r = ref_sqrt(x); ++fb_sqrt; }
			if (!same_bits(r, sqrt(x))) ++bad_sqrt;
		}
	}
	atomicAdd(counts + 0, bad_div); atomicAdd(counts + 1, bad_div3); atomicAdd(counts + 2, bad_rcp); atomicAdd(counts + 3, bad_sqrt);
	atomicAdd(counts + 4, fb_div); atomicAdd(counts + 5, fb_div3); atomicAdd(counts + 6, fb_rcp); atomicAdd(counts + 7, fb_sqrt);
}

} // namespace admmb

using namespace admmb;

#define PROBE_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "admmb probe: %s: %s\n", #x, cudaGetErrorString(e_)); return ADMMB_E_CUDA; } } while (0)

extern "C" int admmb_probe_fp64(int device, double *out6) {
	if (!out6) return ADMMB_E_ARG;
	PROBE_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	PROBE_CUDA(cudaGetDeviceProperties(&prop, device));
	double *d_sink = nullptr;
	long long *d_cyc = nullptr;
	PROBE_CUDA(cudaMalloc(&d_sink, 8));
	PROBE_CUDA(cudaMalloc(&d_cyc, 8));
	cudaEvent_t e0, e1;
	PROBE_CUDA(cudaEventCreate(&e0));
	PROBE_CUDA(cudaEventCreate(&e1));
	const int CH = 8, iters = 1 << 14, threads = 256, blocks = prop.multiProcessorCount * 8;
	for (int kind = 0; kind < 3; ++kind) {
		float best = 1e30f;
		for (int rep = 0; rep < 5; ++rep) {
			PROBE_CUDA(cudaEventRecord(e0));
			if (kind == 0) k_fp64_rate<0, CH><<<blocks, threads>>>(iters, 1.0 + rep, d_sink);
			else if (kind == 1) k_fp64_rate<1, CH><<<blocks, threads>>>(iters, 1.0 + rep, d_sink);
			else k_fp64_rate<2, CH><<<blocks, threads>>>(iters, 1.0 + rep, d_sink);
			PROBE_CUDA(cudaEventRecord(e1));
			PROBE_CUDA(cudaEventSynchronize(e1));
			float ms = 0;
			PROBE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
			if (rep > 0 && ms < best) best = ms;
		}
		// instructions (per thread) per second, in units of 1e12
		out6[kind] = (double)blocks * threads * iters * CH / (best * 1e-3) / 1e12;
	}
	k_fp64_latency<<<1, 32>>>(1 << 12, 1.0, d_sink, d_cyc);
	PROBE_CUDA(cudaDeviceSynchronize());
	long long cyc = 0;
	PROBE_CUDA(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
	out6[3] = (double)cyc / ((1 << 12) * 16.0);  // SM cycles per dependent DFMA
	out6[4] = prop.multiProcessorCount;
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
	out6[5] = khz * 1e-3;                         // max SM clock, MHz
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(d_sink); cudaFree(d_cyc);
	return ADMMB_OK;
}

extern "C" int admmb_debug_fastmath_selftest(int device, unsigned long long seed, long samples, unsigned long long *counts8) {
	if (!counts8 || samples <= 0) return ADMMB_E_ARG;
	PROBE_CUDA(cudaSetDevice(device));
	unsigned long long *d = nullptr;
	PROBE_CUDA(cudaMalloc(&d, 64));
	PROBE_CUDA(cudaMemset(d, 0, 64));
	const int threads = 256, blocks = 148 * 8, per_thread = (int)((samples + (long)threads * blocks - 1) / ((long)threads * blocks));
	k_fastmath_selftest<<<blocks, threads>>>(seed, per_thread, d);
	PROBE_CUDA(cudaGetLastError());
	PROBE_CUDA(cudaDeviceSynchronize());
	PROBE_CUDA(cudaMemcpy(counts8, d, 64, cudaMemcpyDeviceToHost));
	cudaFree(d);
	return ADMMB_OK;
}
