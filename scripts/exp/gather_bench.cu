// Micro-benchmark: random 4-byte gather throughput on B200 from an L2-resident 17 MB table (the fp16 hash-table shadow's size),
// 8 independent loads in flight per thread like hash_fwd_kernel.  Gives the hardware ceiling hash_fwd is measured against.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void k(const uint32_t* __restrict__ table, uint32_t n_entries, int rounds, uint32_t* out)
{
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t acc = 0;
	for (int j = 0; j < rounds; j++) {
		uint32_t v[8];
#pragma unroll
		for (int d = 0; d < 8; d++) v[d] = __ldg(table + hash32(tid * 977u + j * 8 + d) % n_entries);
#pragma unroll
		for (int d = 0; d < 8; d++) acc += v[d];
	}
	if (acc == 0x12345678u) out[0] = acc;
}
int main()
{
	const uint32_t n_entries = 17u << 18;   // 4.46 M entries of 4 B = 17.8 MB
	uint32_t *t, *o; cudaMalloc(&t, size_t(n_entries) * 4); cudaMemset(t, 1, size_t(n_entries) * 4); cudaMalloc(&o, 4);
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	const int rounds = 16;                   // 128 gathers per thread = one point's 16 levels x 8 corners
	for (int threads : {128, 256}) for (int blocks : {3072 * 256 / threads, 1024 * 256 / threads}) {
		for (int rep = 0; rep < 3; rep++) {
			cudaEventRecord(a);
			k<<<blocks, threads>>>(t, n_entries, rounds, o);
			cudaEventRecord(b); cudaEventSynchronize(b);
			float ms; cudaEventElapsedTime(&ms, a, b);
			if (rep == 2) printf("%d x %d threads, 128 random 4B gathers each: %8.1f us  %6.1f G gathers/s\n", blocks, threads, ms * 1e3, double(blocks) * threads * rounds * 8 / (ms * 1e-3) / 1e9);
		}
	}
	printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
