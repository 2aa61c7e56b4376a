// Packed-weight blob of the fused NeRFSmall kernels (nrf_mlp_small_pack): layout constants shared by the mma.sync kernels
// (mlp_small.cu) and the tcgen05 kernel (mlp_small_tc.cu).
//
//   words [0, kFwdWords)            forward B fragments for mma.sync m16n8k16, fp16, [layer][ks][nt][lane][2]
//   words [kFwdWords, kBlobWords)   backward (transposed) B fragments, bf16
//   words [kUmmaBase, kTotalWords)  forward weights as UMMA shared-memory operands (tcgen05.mma B, K-major, no swizzle), fp16:
//                                   layer l at kU<l>; element (n, k) at byte (k/8)*(N*16) + n*16 + (k%8)*2, i.e. 8x16-byte
//                                   core matrices with SBO = 128 B between 8-row groups and LBO = N*16 B between K chunks
#pragma once
#include "common.cuh"

namespace nrf {

constexpr int kW0 = 0, kW1 = 2048, kW2 = 3072, kW3 = 5056, kW4 = 9152, kParamCount = 9344;  // flat fp32 offsets
// fragment blob, offsets in 32-bit words.  fwd: [ks][nt][lane][2]
constexpr int kF0 = 0, kF1 = 1024, kF2 = 1536, kF3 = 2560, kF4 = 4608, kFwdWords = 4864;
constexpr int kB4 = 4864, kB3 = 5376, kB2 = 7424, kB1 = 8448, kB0 = 8960, kBlobWords = 9984;

// padded logical weight matrices Wp_l(n, k)
__device__ __forceinline__ float wp(const float* __restrict__ p, int layer, int n, int k)
{
	switch (layer) {
		case 0: return p[kW0 + n * 32 + k];
		case 1: return p[kW1 + n * 64 + k];
		case 2: return k < 16 ? p[kW2 + n * 31 + k] : (k == 16 ? 0.f : p[kW2 + n * 31 + k - 1]);
		case 3: return p[kW3 + n * 64 + k];
		default: return n < 3 ? p[kW4 + n * 64 + k] : 0.f;
	}
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
	__nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
	return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi)
{
	__half2 v = __floats2half2_rn(lo, hi);
	return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi)
{
	uint32_t r;
	asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
	return r;
}


__device__ __forceinline__ uint32_t pack_f16_relu(float lo, float hi)
{
	uint32_t r;
	asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
	return r;
}

// UMMA operand region (fp16), offsets in 32-bit words; padded N: layer 1 -> 16, layer 4 -> 16 (rows 3..15 zero)
constexpr int kUmmaBase = kBlobWords;
constexpr int kU0 = kUmmaBase, kU1 = kU0 + 64 * 32 / 2, kU2 = kU1 + 16 * 64 / 2, kU3 = kU2 + 64 * 32 / 2, kU4 = kU3 + 64 * 64 / 2;
constexpr int kTotalWords = kU4 + 16 * 64 / 2;
constexpr int kUmmaWords = kTotalWords - kUmmaBase;   // 5120 words = 20 KiB

}  // namespace nrf
