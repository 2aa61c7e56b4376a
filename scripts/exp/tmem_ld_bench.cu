// Microbenchmark: tensor-memory read rate of one SM (tcgen05.ld), B200.
//
// Round 1 explained the 25 % tensor-pipe figure of the fused NeRFSmall forward by "accumulators leave TMEM at 64 B/clk/SM" (one line of
// B300_MICROARCH.md).  This measures it: one CTA per SM allocates all 512 columns; W warps (4: one per lane quarter; 8 / 16: two / four per quarter)
// read all 512 columns of their lanes ITER times with tcgen05.ld.32x32b.x{8,16,32}, K loads issued before each tcgen05.wait::ld, nothing else
// (the values are XOR-ed into one register so that the loads are not dead).  One JSON line per configuration: bytes per clock per SM
// (a warp instruction of shape .xC moves 32 lanes x C columns x 4 B).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/exp/bin/tmem_ld_bench scripts/exp/tmem_ld_bench.cu && scripts/exp/bin/tmem_ld_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int C>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ld<8>(uint32_t taddr, uint32_t* r)
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* r)
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
		  "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* r)
{
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		"{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
		  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
		  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
		  "=r"(r[31]) : "r"(taddr));
}

template <int C, int K>   // C columns per load, K loads in flight per wait
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(int iters, unsigned long long* cycles, uint32_t* sink)
{
	__shared__ uint32_t tmem_base;
	const int warp = threadIdx.x >> 5;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) << 5) << 16);
	uint32_t x = 0;
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int c = 0; c < 512; c += C * K) {
			uint32_t r[C * K];
#pragma unroll
			for (int k = 0; k < K; k++) ld<C>(t_lane + c + C * k, r + C * k);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
			for (int i = 0; i < C * K; i++) x ^= r[i];
		}
	}
	__syncthreads();
	const long long t1 = clock64();
	if (threadIdx.x == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
	if (x == 0x12345678u) sink[0] = x;
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int C, int K>
static void run(int warps)
{
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	unsigned long long* cyc;
	uint32_t* sink;
	cudaMalloc(&cyc, sms * sizeof(unsigned long long));
	cudaMalloc(&sink, 4);
	const int iters = 1000;
	tmem_ld_kernel<C, K><<<sms, warps * 32>>>(10, cyc, sink);
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	tmem_ld_kernel<C, K><<<sms, warps * 32>>>(iters, cyc, sink);
	cudaEventRecord(e1);
	cudaError_t err = cudaDeviceSynchronize();
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	unsigned long long h[256];
	cudaMemcpy(h, cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
	double mean = 0;
	for (int i = 0; i < sms; i++) mean += double(h[i]);
	mean /= sms;
	const double bytes = double(iters) * warps * 32.0 * 512.0 * 4.0;      // per SM: every warp reads all 512 columns of its 32 lanes per iteration
	const double per_wait = mean / (double(iters) * (512 / (C * K)));   // cycles between waits of one warp
	printf("{\"shape\": \"32x32b.x%d\", \"loads_per_wait\": %d, \"warps_per_sm\": %d, \"bytes_per_clk_per_sm\": %.1f, \"cycles_per_wait_group\": %.1f, \"tb_per_s_chip\": %.1f, \"error\": \"%s\"}\n",
		C, K, warps, bytes / mean, per_wait, bytes * sms / (ms * 1e-3) / 1e12, cudaGetErrorString(err));
	cudaFree(cyc); cudaFree(sink);
}

int main()
{
	for (int w : {4, 8, 16}) {
		run<8, 1>(w); run<8, 4>(w);
		run<16, 1>(w); run<16, 2>(w); run<16, 4>(w);
		run<32, 1>(w); run<32, 2>(w); run<32, 4>(w);
	}
	return 0;
}
