// Direction / position encoders for sm_100a.
//   nrf_sh_encode_fwd  replaces CuSHKernel               (reference src/CuSHEncoder.cu:4-118)
//   nrf_posenc_fwd     replaces EmbedderImpl::forward    (reference src/NeRF.cpp:22-39)
//
// Both are pure streaming kernels (12 B in, deg^2*4 or 63*4 B out per row).  Rows are staged through shared
// memory so that global stores are full 128-byte lines instead of one row-strided 4-byte store per thread
// (the reference writes 64 B/thread strided).  In the fused training path the SH table is
// evaluated once per RAY (not per sample, the reference expands directions 192x, src/NeRFRenderer.h:179-181).
#include "common.cuh"

namespace nrf {

// Real spherical-harmonics basis, same polynomial forms and constants as src/CuSHEncoder.cu:26-103.
template <int DEG>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float* o)
{
	const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
	const float x4 = x2 * x2, y4 = y2 * y2, z4 = z2 * z2;
	const float x6 = x4 * x2, y6 = y4 * y2, z6 = z4 * z2;
	o[0] = 0.28209479177387814f;
	if (DEG <= 1) return;
	o[1] = -0.48860251190291987f * y;
	o[2] = 0.48860251190291987f * z;
	o[3] = -0.48860251190291987f * x;
	if (DEG <= 2) return;
	o[4] = 1.0925484305920792f * xy;
	o[5] = -1.0925484305920792f * yz;
	o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
	o[7] = -1.0925484305920792f * xz;
	o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
	if (DEG <= 3) return;
	o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
	o[10] = 2.8906114426405538f * xy * z;
	o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
	o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
	o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
	o[14] = 1.4453057213202769f * z * (x2 - y2);
	o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
	if (DEG <= 4) return;
	o[16] = 2.5033429417967046f * xy * (x2 - y2);
	o[17] = 1.7701307697799304f * yz * (-3.0f * x2 + y2);
	o[18] = 0.94617469575756008f * xy * (7.0f * z2 - 1.0f);
	o[19] = 0.66904654355728921f * yz * (3.0f - 7.0f * z2);
	o[20] = -3.1735664074561294f * z2 + 3.7024941420321507f * z4 + 0.31735664074561293f;
	o[21] = 0.66904654355728921f * xz * (3.0f - 7.0f * z2);
	o[22] = 0.47308734787878004f * (x2 - y2) * (7.0f * z2 - 1.0f);
	o[23] = 1.7701307697799304f * xz * (-x2 + 3.0f * y2);
	o[24] = -3.7550144126950569f * x2 * y2 + 0.62583573544917614f * x4 + 0.62583573544917614f * y4;
	if (DEG <= 5) return;
	o[25] = 0.65638205684017015f * y * (10.0f * x2 * y2 - 5.0f * x4 - y4);
	o[26] = 8.3026492595241645f * xy * z * (x2 - y2);
	o[27] = -0.48923829943525038f * y * (3.0f * x2 - y2) * (9.0f * z2 - 1.0f);
	o[28] = 4.7935367849733241f * xy * z * (3.0f * z2 - 1.0f);
	o[29] = 0.45294665119569694f * y * (14.0f * z2 - 21.0f * z4 - 1.0f);
	o[30] = 0.1169503224534236f * z * (-70.0f * z2 + 63.0f * z4 + 15.0f);
	o[31] = 0.45294665119569694f * x * (14.0f * z2 - 21.0f * z4 - 1.0f);
	o[32] = 2.3967683924866621f * z * (x2 - y2) * (3.0f * z2 - 1.0f);
	o[33] = -0.48923829943525038f * x * (x2 - 3.0f * y2) * (9.0f * z2 - 1.0f);
	o[34] = 2.0756623148810411f * z * (-6.0f * x2 * y2 + x4 + y4);
	o[35] = 0.65638205684017015f * x * (10.0f * x2 * y2 - x4 - 5.0f * y4);
	if (DEG <= 6) return;
	o[36] = 1.3663682103838286f * xy * (-10.0f * x2 * y2 + 3.0f * x4 + 3.0f * y4);
	o[37] = 2.3666191622317521f * yz * (10.0f * x2 * y2 - 5.0f * x4 - y4);
	o[38] = 2.0182596029148963f * xy * (x2 - y2) * (11.0f * z2 - 1.0f);
	o[39] = -0.92120525951492349f * yz * (3.0f * x2 - y2) * (11.0f * z2 - 3.0f);
	o[40] = 0.92120525951492349f * xy * (-18.0f * z2 + 33.0f * z4 + 1.0f);
	o[41] = 0.58262136251873131f * yz * (30.0f * z2 - 33.0f * z4 - 5.0f);
	o[42] = 6.6747662381009842f * z2 - 20.024298714302954f * z4 + 14.684485723822165f * z6 - 0.31784601133814211f;
	o[43] = 0.58262136251873131f * xz * (30.0f * z2 - 33.0f * z4 - 5.0f);
	o[44] = 0.46060262975746175f * (x2 - y2) * (11.0f * z2 * (3.0f * z2 - 1.0f) - 7.0f * z2 + 1.0f);
	o[45] = -0.92120525951492349f * xz * (x2 - 3.0f * y2) * (11.0f * z2 - 3.0f);
	o[46] = 0.50456490072872406f * (11.0f * z2 - 1.0f) * (-6.0f * x2 * y2 + x4 + y4);
	o[47] = 2.3666191622317521f * xz * (10.0f * x2 * y2 - x4 - 5.0f * y4);
	o[48] = 10.247761577878714f * x2 * y4 - 10.247761577878714f * x4 * y2 + 0.6831841051919143f * x6 - 0.6831841051919143f * y6;
	if (DEG <= 7) return;
	o[49] = 0.70716273252459627f * y * (-21.0f * x2 * y4 + 35.0f * x4 * y2 - 7.0f * x6 + y6);
	o[50] = 5.2919213236038001f * xy * z * (-10.0f * x2 * y2 + 3.0f * x4 + 3.0f * y4);
	o[51] = -0.51891557872026028f * y * (13.0f * z2 - 1.0f) * (-10.0f * x2 * y2 + 5.0f * x4 + y4);
	o[52] = 4.1513246297620823f * xy * z * (x2 - y2) * (13.0f * z2 - 3.0f);
	o[53] = -0.15645893386229404f * y * (3.0f * x2 - y2) * (13.0f * z2 * (11.0f * z2 - 3.0f) - 27.0f * z2 + 3.0f);
	o[54] = 0.44253269244498261f * xy * z * (-110.0f * z2 + 143.0f * z4 + 15.0f);
	o[55] = 0.090331607582517306f * y * (-135.0f * z2 + 495.0f * z4 - 429.0f * z6 + 5.0f);
	o[56] = 0.068284276912004949f * z * (315.0f * z2 - 693.0f * z4 + 429.0f * z6 - 35.0f);
	o[57] = 0.090331607582517306f * x * (-135.0f * z2 + 495.0f * z4 - 429.0f * z6 + 5.0f);
	o[58] = 0.07375544874083044f * z * (x2 - y2) * (143.0f * z2 * (3.0f * z2 - 1.0f) - 187.0f * z2 + 45.0f);
	o[59] = -0.15645893386229404f * x * (x2 - 3.0f * y2) * (13.0f * z2 * (11.0f * z2 - 3.0f) - 27.0f * z2 + 3.0f);
	o[60] = 1.0378311574405206f * z * (13.0f * z2 - 3.0f) * (-6.0f * x2 * y2 + x4 + y4);
	o[61] = -0.51891557872026028f * x * (13.0f * z2 - 1.0f) * (-10.0f * x2 * y2 + x4 + 5.0f * y4);
	o[62] = 2.6459606618019f * z * (15.0f * x2 * y4 - 15.0f * x4 * y2 + x6 - y6);
	o[63] = 0.70716273252459627f * x * (-35.0f * x2 * y4 + 21.0f * x4 * y2 - x6 + 7.0f * y6);
}

constexpr int kEncThreads = 128;

// One thread evaluates one direction into registers; the CTA's [128, DEG^2] block is transposed through shared
// memory (row pitch DEG^2+1 words: conflict-free) and leaves as contiguous, fully coalesced stores.
template <int DEG>
__global__ void __launch_bounds__(kEncThreads) sh_kernel(const float* __restrict__ dirs, int stride, int64_t n, float* __restrict__ out)
{
	constexpr int D2 = DEG * DEG;
	__shared__ float tile[kEncThreads * (D2 + 1)];
	const int64_t base = static_cast<int64_t>(blockIdx.x) * kEncThreads;
	const int64_t i = base + threadIdx.x;
	if (i < n) {
		float o[D2];
		sh_basis<DEG>(dirs[i * stride], dirs[i * stride + 1], dirs[i * stride + 2], o);
#pragma unroll
		for (int k = 0; k < D2; k++) tile[threadIdx.x * (D2 + 1) + k] = o[k];
	}
	__syncthreads();
	const int64_t rows = min(static_cast<int64_t>(kEncThreads), n - base);
	const int64_t total = rows * D2;
	float* dst = out + base * D2;
	for (int64_t e = threadIdx.x; e < total; e += kEncThreads) {
		const int r = static_cast<int>(e / D2), k = static_cast<int>(e % D2);
		dst[e] = tile[r * (D2 + 1) + k];
	}
}

struct FreqBands {
	float f[32];
};

// [x, sin(f0 x), cos(f0 x), ...] (src/NeRF.cpp:33-37).  One CTA handles 128 rows; element e of the CTA's output
// block maps to (row, channel) so that stores are contiguous.  sinf/cosf are the accurate libdevice versions
// (ATen's CUDA sin/cos kernels call the same functions), arguments are the fp32 product x*freq as in the reference.
__global__ void __launch_bounds__(kEncThreads) posenc_kernel(const float* __restrict__ x, int64_t n, int in_dims, int num_freqs,
	FreqBands fb, int include_input, float* __restrict__ out)
{
	extern __shared__ float xs[];  // [128 * in_dims] then [num_freqs] frequency bands
	const int64_t base = static_cast<int64_t>(blockIdx.x) * kEncThreads;
	const int rows = static_cast<int>(min(static_cast<int64_t>(kEncThreads), n - base));
	float* fs = xs + kEncThreads * in_dims;
	for (int e = threadIdx.x; e < rows * in_dims; e += kEncThreads) xs[e] = x[base * in_dims + e];
	for (int e = threadIdx.x; e < num_freqs; e += kEncThreads) fs[e] = fb.f[e];
	__syncthreads();
	// 32-bit index arithmetic throughout (a CTA's block has at most 128 x out_dims elements): the 64-bit divisions of the first
	// version cost more than the sinf/cosf they addressed (50 -> see profiles: 196 608 x 63 outputs)
	const int out_dims = in_dims * (include_input ? 1 : 0) + 2 * num_freqs * in_dims;
	const int total = rows * out_dims;
	const int first = include_input ? in_dims : 0, two_in = 2 * in_dims;
	float* dst = out + base * out_dims;
	for (int e = threadIdx.x; e < total; e += kEncThreads) {
		const int r = e / out_dims;
		const int c = e - r * out_dims;
		float v;
		if (c < first) {
			v = xs[r * in_dims + c];
		} else {
			const int cc = c - first;
			const int band = cc / two_in;
			const int rem = cc - band * two_in;
			const bool is_sin = rem < in_dims;
			const float arg = __fmul_rn(xs[r * in_dims + (is_sin ? rem : rem - in_dims)], fs[band]);
			v = is_sin ? sinf(arg) : cosf(arg);
		}
		dst[e] = v;
	}
}

}  // namespace nrf

using namespace nrf;

extern "C" {

int nrf_sh_encode_fwd(const float* dirs, int32_t dir_stride, int64_t n, int32_t degree, float* out, nrf_stream stream)
{
	NRF_REQUIRE(degree >= 1 && degree <= 8, "degree must be 1..8");
	NRF_REQUIRE(dir_stride >= 3, "dir_stride must be >= 3");
	NRF_REQUIRE(n >= 0, "negative n");
	if (n == 0) return NRF_OK;
	NRF_REQUIRE(dirs && out, "null pointer");
	const unsigned blocks = static_cast<unsigned>((n + kEncThreads - 1) / kEncThreads);
	cudaStream_t s = as_stream(stream);
	switch (degree) {
		case 1: sh_kernel<1><<<blocks, kEncThreads, 0, s>>>(dirs, dir_stride, n, out); break;
		case 2: sh_kernel<2><<<blocks, kEncThreads, 0, s>>>(dirs, dir_stride, n, out); break;
		case 3: sh_kernel<3><<<blocks, kEncThreads, 0, s>>>(dirs, dir_stride, n, out); break;
		case 4: sh_kernel<4><<<blocks, kEncThreads, 0, s>>>(dirs, dir_stride, n, out); break;
		case 5: sh_kernel<5><<<blocks, kEncThreads, 0, s>>>(dirs, dir_stride, n, out); break;
		case 6: sh_kernel<6><<<blocks, kEncThreads, 0, s>>>(dirs, dir_stride, n, out); break;
		case 7: sh_kernel<7><<<blocks, kEncThreads, 0, s>>>(dirs, dir_stride, n, out); break;
		default: sh_kernel<8><<<blocks, kEncThreads, 0, s>>>(dirs, dir_stride, n, out); break;
	}
	NRF_CHECK_LAUNCH("sh_kernel");
	return NRF_OK;
}

int nrf_posenc_fwd(const float* x, int64_t n, int32_t input_dims, int32_t num_freqs, const float* freq_bands_host,
	int32_t include_input, float* out, nrf_stream stream)
{
	NRF_REQUIRE(input_dims >= 1 && input_dims <= 16, "input_dims out of range");
	NRF_REQUIRE(num_freqs >= 0 && num_freqs <= 32, "num_freqs out of range");
	NRF_REQUIRE(n >= 0, "negative n");
	if (n == 0) return NRF_OK;
	NRF_REQUIRE(x && out && (freq_bands_host || num_freqs == 0), "null pointer");
	FreqBands fb;
	for (int i = 0; i < num_freqs; i++) fb.f[i] = freq_bands_host[i];
	const unsigned blocks = static_cast<unsigned>((n + kEncThreads - 1) / kEncThreads);
	posenc_kernel<<<blocks, kEncThreads, (kEncThreads * input_dims + num_freqs) * sizeof(float), as_stream(stream)>>>(x, n, input_dims, num_freqs, fb, include_input, out);
	NRF_CHECK_LAUNCH("posenc_kernel");
	return NRF_OK;
}

}
