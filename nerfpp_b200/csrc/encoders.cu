// Direction / position encoders for sm_100a.
//   nrf_sh_encode_fwd  replaces CuSHKernel               (reference src/CuSHEncoder.cu:4-118)
//   nrf_posenc_fwd     replaces EmbedderImpl::forward    (reference src/NeRF.cpp:22-39)
//
// Both are pure streaming kernels (12 B in, deg^2*4 or 63*4 B out per row).  Rows are staged through shared
// memory so that global stores are full 128-byte lines instead of one row-strided 4-byte store per thread
// (the reference writes 64 B/thread strided).  In the fused training path the SH table is
// evaluated once per RAY (not per sample, the reference expands directions 192x, src/NeRFRenderer.h:179-181).
#include "common.cuh"
#include "sh_basis.cuh"

namespace nrf {

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
