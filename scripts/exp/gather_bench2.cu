// Micro-benchmark, second pass: does the FLAVOUR of the load change the random-gather ceiling (scripts/exp/gather_bench.cu measured
// 280 G/s with ld.global.nc)?  Same access pattern (17.8 MB L2-resident table, 8 independent 4-byte loads in flight per thread), loads:
//   0 ld.global.nc (__ldg)   1 ld.global.ca   2 ld.global.cg (L2 only)   3 ld.global.nc.L1::no_allocate   4 ld.global.L1::evict_first
//   5 8-byte entries (two levels' worth per lookup would need this)  6 ld.global.cv
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE>
__device__ __forceinline__ uint32_t load(const uint32_t* p)
{
	uint32_t v;
	if (MODE == 0) v = __ldg(p);
	else if (MODE == 1) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
	else if (MODE == 2) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
	else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	else if (MODE == 4) asm volatile("ld.global.L1::evict_first.u32 %0, [%1];" : "=r"(v) : "l"(p));
	else if (MODE == 6) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
	else { uint2 w; asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(reinterpret_cast<const uint2*>(p))); v = w.x + w.y; }
	return v;
}
template <int MODE>
__global__ void k(const uint32_t* __restrict__ table, uint32_t n_entries, int rounds, uint32_t* out)
{
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t acc = 0;
	for (int j = 0; j < rounds; j++) {
		uint32_t v[8];
#pragma unroll
		for (int d = 0; d < 8; d++) {
			uint32_t idx = hash32(tid * 977u + j * 8 + d) % n_entries;
			if (MODE == 5) idx &= ~1u;
			v[d] = load<MODE>(table + idx);
		}
#pragma unroll
		for (int d = 0; d < 8; d++) acc += v[d];
	}
	if (acc == 0x12345678u) out[0] = acc;
}
template <int MODE>
void run(const char* name, uint32_t* t, uint32_t n_entries, uint32_t* o)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	const int rounds = 16, threads = 128, blocks = 3072 * 2;
	float best = 1e9f;
	for (int rep = 0; rep < 4; rep++) {
		cudaEventRecord(a);
		k<MODE><<<blocks, threads>>>(t, n_entries, rounds, o);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b);
		if (rep) best = ms < best ? ms : best;
	}
	printf("%-40s %8.1f us  %6.1f G gathers/s\n", name, best * 1e3, double(blocks) * threads * rounds * 8 / (best * 1e-3) / 1e9);
}
int main()
{
	const uint32_t n_entries = 17u << 18;
	uint32_t *t, *o; cudaMalloc(&t, size_t(n_entries) * 4); cudaMemset(t, 1, size_t(n_entries) * 4); cudaMalloc(&o, 4);
	run<0>("ld.global.nc (__ldg)", t, n_entries, o);
	run<1>("ld.global.ca", t, n_entries, o);
	run<2>("ld.global.cg", t, n_entries, o);
	run<3>("ld.global.nc.L1::no_allocate", t, n_entries, o);
	run<4>("ld.global.L1::evict_first", t, n_entries, o);
	run<6>("ld.global.cv", t, n_entries, o);
	run<5>("ld.global.nc.v2 (8-byte entries)", t, n_entries, o);
	printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
