// Micro-benchmark: L2 RED throughput on B200 for random addresses in a 34 MB table: v2.f32 (8 B) vs f32 (4 B) vs f16x2 (4 B) vs bf16x2 (4 B)
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE>
__global__ void k(float* table, uint32_t n_entries, int per_thread)
{
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	for (int j = 0; j < per_thread; j++) {
		const uint32_t idx = hash32(tid * 131u + j) % n_entries;   // entry = 8 bytes
		float* p = table + size_t(idx) * 2;
		if (MODE == 0) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.f), "f"(2.f) : "memory");
		if (MODE == 1) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.f) : "memory"); asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 1), "f"(2.f) : "memory"); }
		if (MODE == 2) { __half2 v = __floats2half2_rn(1.f, 2.f); asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(p), "r"(*reinterpret_cast<uint32_t*>(&v)) : "memory"); }
		if (MODE == 3) { __nv_bfloat162 v = __floats2bfloat162_rn(1.f, 2.f); asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(p), "r"(*reinterpret_cast<uint32_t*>(&v)) : "memory"); }
		if (MODE == 4) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.f) : "memory");
	}
}
int main()
{
	const uint32_t n_entries = 4456448;   // 17 * 2^18 entries of 8 B = 35.6 MB
	float* t; cudaMalloc(&t, size_t(n_entries) * 8); cudaMemset(t, 0, size_t(n_entries) * 8);
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	const int blocks = 3072, threads = 256, per = 64;   // 50.3 M ops
	const char* names[] = {"v2.f32 (8B)", "2 x f32 (4B+4B)", "f16x2 (4B)", "bf16x2 (4B)", "1 x f32 (4B)"};
	for (int mode = 0; mode < 5; mode++) {
		for (int rep = 0; rep < 3; rep++) {
			cudaEventRecord(a);
			if (mode == 0) k<0><<<blocks, threads>>>(t, n_entries, per);
			if (mode == 1) k<1><<<blocks, threads>>>(t, n_entries, per);
			if (mode == 2) k<2><<<blocks, threads>>>(t, n_entries, per);
			if (mode == 3) k<3><<<blocks, threads>>>(t, n_entries, per);
			if (mode == 4) k<4><<<blocks, threads>>>(t, n_entries, per);
			cudaEventRecord(b); cudaEventSynchronize(b);
			float ms; cudaEventElapsedTime(&ms, a, b);
			if (rep == 2) printf("%-18s %8.1f us  %6.1f G ops/s\n", names[mode], ms * 1e3, double(blocks) * threads * per / (ms * 1e-3) / 1e9);
		}
	}
	printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
