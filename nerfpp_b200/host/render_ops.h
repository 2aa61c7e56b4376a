// Drop-in rendering operators: TruncExp, RawToOutputs (as one fused autograd op), SamplePDF, GetRays / IntersectWithAABB /
// NDCRays, TangentScatter, ReflectBoundary — same names and argument meaning as the reference
// (src/CustomOps.h, src/Sampler.h:6, src/RayUtils.h:5-126, src/NeRFRenderer.h:285-362), bodies behind the C ABI.
#pragma once
#include "nrf_torch.h"

namespace torch::autograd {
/// exp(x) forward, g * exp(clamp(x, -100, 5)) backward (src/CustomOps.cpp:5-16).  Kept for source compatibility; the
/// compositing kernels have it folded in.
class TruncExp : public Function<TruncExp> {
public:
	static variable_list forward(AutogradContext* ctx, torch::Tensor input);
	static variable_list backward(AutogradContext* ctx, variable_list grad_output);
};
}  // namespace torch::autograd

/// Hierarchical sampling (src/Sampler.h:6-43).  bins [R,B], weights [R,B-1] -> samples [R,nsamples]
torch::Tensor SamplePDF(torch::Tensor bins, torch::Tensor weights, const int nsamples, const bool det);

/// pinhole directions / rays (src/RayUtils.h:5-46); returns (rays_o [h,w,3], rays_d [h,w,3], cone_angle scalar tensor)
torch::Tensor GetDirections(const int h, const int w, torch::Tensor k);
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> GetRays(const int h, const int w, torch::Tensor k, torch::Tensor c2w);
/// forward-facing NDC warp (src/RayUtils.h:49-83); ATen, not on the HashNeRF hot path
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> NDCRays(const int h, const int w, const float focal, const float near,
	torch::Tensor rays_o, torch::Tensor rays_d, torch::Tensor cone_angle);
/// slab test (src/RayUtils.h:87-126); returns (near [N], far [N])
std::pair<torch::Tensor, torch::Tensor> IntersectWithAABB(const torch::Tensor& rays_o, const torch::Tensor& rays_d,
	const torch::Tensor& bounding_box, float near_plane = 0.f);
/// reflect-at-bounds used by stochastic preconditioning (src/NeRFRenderer.h:285-304)
torch::Tensor ReflectBoundary(torch::Tensor pts, torch::Tensor min_bound, torch::Tensor max_bound);
/// in-cone jitter of the sample points (src/NeRFRenderer.h:307-362); no-op when cone_angle is undefined / empty
torch::Tensor TangentScatter(torch::Tensor pts, torch::Tensor z_vals, torch::Tensor cone_angle, torch::Tensor rays_d, torch::Device device,
	torch::Tensor bounding_box);

namespace nrfhost {

struct CompositeResult {
	torch::Tensor rgb, depth, disp, acc, weights;
};
/// RawToOutputs (src/NeRFRenderer.h:199-282) as one differentiable op: raw [R,S,C>=4], z [R,S], rays_d [R,3]
CompositeResult Composite(const torch::Tensor& raw, const torch::Tensor& z_vals, const torch::Tensor& rays_d, float raw_noise_std,
	bool white_bkgr);
/// Render prologue (src/NeRFRenderer.h:549-583): [o, d, near, far (, viewdirs)] rows
torch::Tensor RaysPrepare(const torch::Tensor& rays_o, const torch::Tensor& rays_d, const torch::Tensor& bounding_box, float near_plane,
	bool use_viewdirs);
/// z = near(1-t)+far t (src/NeRFRenderer.h:393-402) and pts = o + d z (:419)
torch::Tensor ZSample(const torch::Tensor& ray_batch, int n_samples, bool lin_disp);
torch::Tensor SamplePoints(const torch::Tensor& ray_batch, const torch::Tensor& z_vals);
/// z_mid, SamplePDF(z_mid, w[:,1:-1]) and sort(cat(z, z_samples)) in one kernel (src/NeRFRenderer.h:427-431); det only
torch::Tensor SamplePdfMerge(const torch::Tensor& z_vals, const torch::Tensor& weights, int n_importance);
/// The same merge, and the coarse pass's raw rows [R,S,4] travel to their merged positions (nrf_sample_pdf_merge_rows): returns
/// {z_merged [R,S+N], perm [R,S+N] int16 (entry j < N: merged position of importance sample j), raw_merged [R,S+N,4] with the coarse rows filled}
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> SamplePdfMergeRows(const torch::Tensor& z_vals, const torch::Tensor& weights, int n_importance,
	const torch::Tensor& raw_coarse);
/// linspace(0,1,n) on the device, built once per (n, device)
torch::Tensor UnitLinspace(int n, const torch::Device& device);

}  // namespace nrfhost
