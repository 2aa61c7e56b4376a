// Glue between torch::Tensor and the C ABI of include/nerfpp_b200.h.  The host layer keeps torch::Tensor at the API
// boundary (like the reference, src/*.h) and calls ONLY extern "C" entry points of libnerfpp_b200.so for arithmetic on
// the hot path; there is no CPU path: a CPU tensor is an error, not a fallback.
#pragma once
#include <torch/torch.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>

#include "nerfpp_b200.h"

namespace nrfhost {

inline nrf_stream Stream() { return reinterpret_cast<nrf_stream>(c10::cuda::getCurrentCUDAStream().stream()); }

inline void Check(int status, const char* op)
{
	TORCH_CHECK(status == NRF_OK, "nerfpp_b200: ", op, " failed (", status, "): ", nrf_last_error());
}

// contiguous CUDA tensor of the given dtype (copies only when the caller handed over a strided / other-dtype view)
inline torch::Tensor Dense(const torch::Tensor& t, torch::ScalarType dtype, const char* what)
{
	TORCH_CHECK(t.defined(), "nerfpp_b200: ", what, " is undefined");
	TORCH_CHECK(t.is_cuda(), "nerfpp_b200: ", what, " must be a CUDA tensor (the sm_100a path has no CPU fallback)");
	torch::Tensor r = t;
	if (r.scalar_type() != dtype) r = r.to(dtype);
	return r.contiguous();
}

template <class T>
inline const T* CPtr(const torch::Tensor& t) { return t.defined() && t.numel() ? reinterpret_cast<const T*>(t.data_ptr()) : nullptr; }
template <class T>
inline T* Ptr(torch::Tensor& t) { return t.defined() && t.numel() ? reinterpret_cast<T*>(t.data_ptr()) : nullptr; }

inline torch::TensorOptions F32Like(const torch::Tensor& t) { return torch::TensorOptions().dtype(torch::kFloat32).device(t.device()); }

// six floats of a bounding-box tensor on the host (one small D2H copy when it lives on the GPU)
inline std::array<float, 6> HostBox(const torch::Tensor& bbox)
{
	TORCH_CHECK(bbox.defined() && bbox.numel() == 6, "nerfpp_b200: bounding box must hold 6 values");
	torch::Tensor b = bbox.detach().to(torch::kCPU, torch::kFloat32).contiguous().reshape({6});
	std::array<float, 6> r;
	for (int i = 0; i < 6; i++) r[i] = b.data_ptr<float>()[i];
	return r;
}

}  // namespace nrfhost
