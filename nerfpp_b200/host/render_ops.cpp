// Host side of the drop-in rendering operators (see render_ops.h).
#include "render_ops.h"

#include <map>
#include <mutex>

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

// ------------------------------------------------------------------------------------------------ TruncExp
variable_list torch::autograd::TruncExp::forward(AutogradContext* ctx, Tensor input)
{
	ctx->save_for_backward({input});
	return {torch::exp(input)};                                              // NOT truncated (src/CustomOps.cpp:8)
}

variable_list torch::autograd::TruncExp::backward(AutogradContext* ctx, variable_list grad_output)
{
	Tensor x = ctx->get_saved_variables()[0];
	return {grad_output[0] * torch::exp(torch::clamp(x, -100.f, 5.f))};     // :14
}

// ------------------------------------------------------------------------------------------------ helpers
Tensor nrfhost::UnitLinspace(int n, const torch::Device& device)
{
	static std::mutex mu;
	static std::map<std::pair<int, int>, Tensor> cache;
	std::lock_guard<std::mutex> lock(mu);
	auto key = std::make_pair(n, int(device.index()));
	auto it = cache.find(key);
	if (it == cache.end()) {
		// built on the CPU like the reference (src/Sampler.h:20, src/NeRFRenderer.h:393) so the fp32 values are the same, uploaded once
		it = cache.emplace(key, torch::linspace(0.f, 1.f, n, torch::kFloat).to(device)).first;
	}
	return it->second;
}

// ------------------------------------------------------------------------------------------------ sampler
Tensor SamplePDF(Tensor bins, Tensor weights, const int nsamples, const bool det)
{
	Tensor b = nrfhost::Dense(bins.detach(), torch::kFloat32, "bins"), w = nrfhost::Dense(weights.detach(), torch::kFloat32, "weights");
	TORCH_CHECK(b.dim() == 2 && w.dim() == 2 && w.size(0) == b.size(0) && w.size(1) + 1 == b.size(1), "SamplePDF: bins [R,B], weights [R,B-1]");
	// det: one shared linspace; else per-ray uniforms drawn like the reference (torch::rand on the CPU, src/Sampler.h:23)
	Tensor u = det ? nrfhost::UnitLinspace(nsamples, b.device()) : torch::rand({b.size(0), nsamples}).to(b.device()).contiguous();
	Tensor out = torch::empty({b.size(0), nsamples}, nrfhost::F32Like(b));
	nrfhost::Check(nrf_sample_pdf(nrfhost::CPtr<float>(b), nrfhost::CPtr<float>(w), int32_t(b.size(1)), nrfhost::CPtr<float>(u), det ? 0 : 1,
		b.size(0), nsamples, nrfhost::Ptr<float>(out), nrfhost::Stream()), "nrf_sample_pdf");
	return out;
}

Tensor nrfhost::SamplePdfMerge(const Tensor& z_vals, const Tensor& weights, int n_importance)
{
	Tensor z = Dense(z_vals.detach(), torch::kFloat32, "z_vals"), w = Dense(weights.detach(), torch::kFloat32, "weights");
	Tensor u = UnitLinspace(n_importance, z.device());
	Tensor merged = torch::empty({z.size(0), z.size(1) + n_importance}, F32Like(z));
	Check(nrf_sample_pdf_merge(CPtr<float>(z), CPtr<float>(w), CPtr<float>(u), 0, z.size(0), int32_t(z.size(1)), n_importance, nullptr,
		Ptr<float>(merged), Stream()), "nrf_sample_pdf_merge");
	return merged;
}

std::tuple<Tensor, Tensor, Tensor> nrfhost::SamplePdfMergeRows(const Tensor& z_vals, const Tensor& weights, int n_importance, const Tensor& raw_coarse)
{
	Tensor z = Dense(z_vals.detach(), torch::kFloat32, "z_vals"), w = Dense(weights.detach(), torch::kFloat32, "weights");
	Tensor rc = Dense(raw_coarse.detach(), torch::kFloat32, "raw");
	TORCH_CHECK(rc.dim() == 3 && rc.size(0) == z.size(0) && rc.size(1) == z.size(1) && rc.size(2) == 4, "SamplePdfMergeRows: raw must be [R,S,4]");
	Tensor u = UnitLinspace(n_importance, z.device());
	const int64_t R = z.size(0), T = z.size(1) + n_importance;
	Tensor merged = torch::empty({R, T}, F32Like(z)), raw = torch::empty({R, T, 4}, F32Like(z));
	Tensor perm = torch::empty({R, T}, torch::TensorOptions().dtype(torch::kInt16).device(z.device()));
	Check(nrf_sample_pdf_merge_rows(CPtr<float>(z), CPtr<float>(w), CPtr<float>(u), 0, R, int32_t(z.size(1)), n_importance, nullptr, Ptr<float>(merged),
		perm.data_ptr<int16_t>(), CPtr<float>(rc), Ptr<float>(raw), Stream()), "nrf_sample_pdf_merge_rows");
	return {merged, perm, raw};
}

// ------------------------------------------------------------------------------------------------ rays
Tensor GetDirections(const int h, const int w, Tensor k)
{
	// only the intrinsics matter: build the rays of an identity camera and keep the directions
	Tensor eye = torch::eye(4, torch::kFloat32).narrow(0, 0, 3);
	return std::get<1>(GetRays(h, w, k, eye.to(k.device())));
}

std::tuple<Tensor, Tensor, Tensor> GetRays(const int h, const int w, Tensor k, Tensor c2w)
{
	TORCH_CHECK(c2w.is_cuda(), "GetRays: c2w must be a CUDA tensor (the sm_100a path has no CPU fallback)");
	Tensor kh = k.detach().to(torch::kCPU, torch::kFloat32).contiguous();
	Tensor ch = c2w.detach().to(torch::kCPU, torch::kFloat32).narrow(0, 0, 3).narrow(1, 0, 4).contiguous();
	Tensor rays_o = torch::empty({h, w, 3}, nrfhost::F32Like(c2w)), rays_d = torch::empty({h, w, 3}, nrfhost::F32Like(c2w));
	nrfhost::Check(nrf_get_rays(h, w, kh.data_ptr<float>(), ch.data_ptr<float>(), 0, h, nrfhost::Ptr<float>(rays_o), nrfhost::Ptr<float>(rays_d),
		nrfhost::Stream()), "nrf_get_rays");
	// cone_angle = 1.1 * (1/fx + 1/fy) / 2 (src/RayUtils.h:36-43), a 0-dim tensor on k's device like the reference's
	const float fx = kh[0][0].item<float>(), fy = kh[1][1].item<float>();
	Tensor cone_angle = torch::tensor((1.0f / fx + 1.0f / fy) / 2.0f * 1.1f, torch::TensorOptions().dtype(torch::kFloat32).device(k.device()));
	return {rays_o, rays_d, cone_angle};
}

std::tuple<Tensor, Tensor, Tensor> NDCRays(const int h, const int w, const float focal, const float near, Tensor rays_o, Tensor rays_d,
	Tensor cone_angle)
{
	using torch::indexing::None;
	auto comp = [](const Tensor& t, int i) { return t.select(-1, i); };
	// shift origins onto the near plane, then project (src/RayUtils.h:58-72)
	Tensor t = -(near + comp(rays_o, 2)) / comp(rays_d, 2);
	rays_o = rays_o + t.unsqueeze(-1) * rays_d;
	const double sx = -1. / (w / (2. * focal)), sy = -1. / (h / (2. * focal));
	Tensor oz = comp(rays_o, 2);
	Tensor o0 = sx * comp(rays_o, 0) / oz, o1 = sy * comp(rays_o, 1) / oz, o2 = 1. + 2. * near / oz;
	Tensor d0 = sx * (comp(rays_d, 0) / comp(rays_d, 2) - comp(rays_o, 0) / oz);
	Tensor d1 = sy * (comp(rays_d, 1) / comp(rays_d, 2) - comp(rays_o, 1) / oz);
	Tensor d2 = -2. * near / oz;
	Tensor ndc_o = torch::stack({o0, o1, o2}, -1), ndc_d = torch::stack({d0, d1, d2}, -1);
	if (cone_angle.defined() && cone_angle.numel()) {
		// NB the reference divides by the norm of the ALREADY re-assigned rays_d (:79-80), i.e. the ratio is 1: reproduced
		Tensor scale = torch::norm(ndc_d, 2, -1) / torch::norm(ndc_d, 2, -1);
		cone_angle = cone_angle * scale.unsqueeze(-1);
	}
	return {ndc_o, ndc_d, cone_angle};
}

Tensor nrfhost::RaysPrepare(const Tensor& rays_o, const Tensor& rays_d, const Tensor& bounding_box, float near_plane, bool use_viewdirs)
{
	Tensor o = Dense(rays_o.detach(), torch::kFloat32, "rays_o").reshape({-1, 3}), d = Dense(rays_d.detach(), torch::kFloat32, "rays_d").reshape({-1, 3});
	TORCH_CHECK(o.size(0) == d.size(0), "rays_o / rays_d size mismatch");
	const std::array<float, 6> box = HostBox(bounding_box);
	Tensor rb = torch::empty({o.size(0), use_viewdirs ? 11 : 8}, F32Like(o));
	Check(nrf_rays_prepare(CPtr<float>(o), CPtr<float>(d), o.size(0), box.data(), near_plane, use_viewdirs ? 1 : 0, Ptr<float>(rb), Stream()),
		"nrf_rays_prepare");
	return rb;
}

std::pair<Tensor, Tensor> IntersectWithAABB(const Tensor& rays_o, const Tensor& rays_d, const Tensor& bounding_box, float near_plane)
{
	Tensor rb = nrfhost::RaysPrepare(rays_o, rays_d, bounding_box, near_plane, false);
	return {rb.select(1, 6).contiguous(), rb.select(1, 7).contiguous()};
}

Tensor nrfhost::ZSample(const Tensor& ray_batch, int n_samples, bool lin_disp)
{
	Tensor rb = Dense(ray_batch.detach(), torch::kFloat32, "ray_batch");
	Tensor t = UnitLinspace(n_samples, rb.device());
	Tensor z = torch::empty({rb.size(0), n_samples}, F32Like(rb));
	Check(nrf_z_sample(CPtr<float>(rb), int32_t(rb.size(1)), CPtr<float>(t), rb.size(0), n_samples, lin_disp ? 1 : 0, Ptr<float>(z), Stream()),
		"nrf_z_sample");
	return z;
}

Tensor nrfhost::SamplePoints(const Tensor& ray_batch, const Tensor& z_vals)
{
	Tensor rb = Dense(ray_batch.detach(), torch::kFloat32, "ray_batch"), z = Dense(z_vals.detach(), torch::kFloat32, "z_vals");
	Tensor pts = torch::empty({z.size(0), z.size(1), 3}, F32Like(z));
	Check(nrf_sample_points(CPtr<float>(rb), int32_t(rb.size(1)), CPtr<float>(z), z.size(0), int32_t(z.size(1)), Ptr<float>(pts), Stream()),
		"nrf_sample_points");
	return pts;
}

Tensor ReflectBoundary(Tensor pts, Tensor min_bound, Tensor max_bound)
{
	// normalise to [0,1]^3, fold with period 2 (x -> 2 - x above 1), map back (src/NeRFRenderer.h:285-304)
	Tensor extent = max_bound - min_bound;
	Tensor q = torch::fmod((pts - min_bound) / extent, 2.0f);
	q = torch::where(q > 1.0f, 2.0f - q, q);
	return q * extent + min_bound;
}

Tensor TangentScatter(Tensor pts, Tensor z_vals, Tensor cone_angle, Tensor rays_d, torch::Device device, Tensor bounding_box)
{
	if (!(cone_angle.defined() && cone_angle.numel())) return pts;
	Tensor p = nrfhost::Dense(pts.detach(), torch::kFloat32, "pts").clone();   // the kernel updates in place
	Tensor z = nrfhost::Dense(z_vals.detach(), torch::kFloat32, "z_vals"), d = nrfhost::Dense(rays_d.detach(), torch::kFloat32, "rays_d");
	const int64_t n_rays = z.size(0), n_samples = z.size(1);
	// the same two draws, in the same order, as src/NeRFRenderer.h:342-343: the Philox stream stays in lock-step
	Tensor u_r = torch::rand({n_rays, n_samples, 1}, torch::TensorOptions().dtype(torch::kFloat32).device(device));
	Tensor u_t = torch::rand({n_rays, n_samples, 1}, torch::TensorOptions().dtype(torch::kFloat32).device(device));
	Tensor ca = nrfhost::Dense(cone_angle.detach(), torch::kFloat32, "cone_angle").reshape({-1});
	TORCH_CHECK(ca.numel() == 1 || ca.numel() == n_rays, "TangentScatter: cone_angle must be a scalar or one value per ray");
	std::array<float, 6> box{};
	const bool have_box = bounding_box.defined() && bounding_box.numel() == 6;
	if (have_box) box = nrfhost::HostBox(bounding_box);
	nrfhost::Check(nrf_tangent_scatter(nrfhost::Ptr<float>(p), nrfhost::CPtr<float>(z), nrfhost::CPtr<float>(ca), ca.numel() == 1 ? 0 : 1,
		nrfhost::CPtr<float>(d), int32_t(d.size(1)), nrfhost::CPtr<float>(u_r), nrfhost::CPtr<float>(u_t), have_box ? box.data() : nullptr,
		n_rays, int32_t(n_samples), nrfhost::Stream()), "nrf_tangent_scatter");
	return p;
}

// ------------------------------------------------------------------------------------------------ compositing
namespace {

struct CompositeFn : public torch::autograd::Function<CompositeFn> {
	static variable_list forward(AutogradContext* ctx, Tensor raw, Tensor z, Tensor rays_d, Tensor noise, double noise_std, bool white)
	{
		Tensor r = nrfhost::Dense(raw, torch::kFloat32, "raw"), zz = nrfhost::Dense(z, torch::kFloat32, "z_vals");
		Tensor d = nrfhost::Dense(rays_d, torch::kFloat32, "rays_d");
		TORCH_CHECK(r.dim() == 3 && r.size(2) >= 4 && zz.dim() == 2 && zz.size(0) == r.size(0) && zz.size(1) == r.size(1), "RawToOutputs: raw [R,S,>=4], z [R,S]");
		const int64_t R = r.size(0), S = r.size(1);
		const auto opt = nrfhost::F32Like(r);
		Tensor rgb = torch::empty({R, 3}, opt), depth = torch::empty({R}, opt), disp = torch::empty({R}, opt), acc = torch::empty({R}, opt);
		Tensor weights = torch::empty({R, S}, opt);
		nrfhost::Check(nrf_composite_fwd(nrfhost::CPtr<float>(r), int32_t(r.size(2)), nrfhost::CPtr<float>(zz), nrfhost::CPtr<float>(d),
			nrfhost::CPtr<float>(noise), float(noise_std), white ? 1 : 0, R, int32_t(S), nrfhost::Ptr<float>(rgb), nrfhost::Ptr<float>(depth),
			nrfhost::Ptr<float>(disp), nrfhost::Ptr<float>(acc), nrfhost::Ptr<float>(weights), nrfhost::Stream()), "nrf_composite_fwd");
		ctx->save_for_backward({r, zz, d, noise});
		ctx->saved_data["noise_std"] = noise_std;
		ctx->saved_data["white"] = white;
		return {rgb, depth, disp, acc, weights};
	}

	static variable_list backward(AutogradContext* ctx, variable_list g)
	{
		auto saved = ctx->get_saved_variables();
		Tensor r = saved[0], z = saved[1], d = saved[2], noise = saved[3];
		auto dense = [](const Tensor& t) { return t.defined() ? nrfhost::Dense(t, torch::kFloat32, "upstream gradient") : t; };
		Tensor g_rgb = dense(g[0]), g_depth = dense(g[1]), g_disp = dense(g[2]), g_acc = dense(g[3]), g_w = dense(g[4]);
		Tensor d_raw4 = torch::empty({r.size(0), r.size(1), 4}, nrfhost::F32Like(r));
		nrfhost::Check(nrf_composite_bwd(nrfhost::CPtr<float>(r), int32_t(r.size(2)), nrfhost::CPtr<float>(z), nrfhost::CPtr<float>(d),
			nrfhost::CPtr<float>(noise), float(ctx->saved_data["noise_std"].toDouble()), ctx->saved_data["white"].toBool() ? 1 : 0, r.size(0),
			int32_t(r.size(1)), nrfhost::CPtr<float>(g_rgb), nrfhost::CPtr<float>(g_depth), nrfhost::CPtr<float>(g_disp), nrfhost::CPtr<float>(g_acc),
			nrfhost::CPtr<float>(g_w), nrfhost::Ptr<float>(d_raw4), nrfhost::Stream()), "nrf_composite_bwd");
		Tensor d_raw = d_raw4;
		if (r.size(2) > 4) {   // extra channels (normals) take no part in the compositing
			d_raw = torch::zeros_like(r);
			d_raw.narrow(2, 0, 4).copy_(d_raw4);
		}
		return {d_raw, Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};   // z is detached upstream (src/NeRFRenderer.h:429)
	}
};

}  // namespace

nrfhost::CompositeResult nrfhost::Composite(const Tensor& raw, const Tensor& z_vals, const Tensor& rays_d, float raw_noise_std, bool white_bkgr)
{
	// the noise is drawn with the reference's call (randn_like on the density channel, src/NeRFRenderer.h:253-254)
	// (an EMPTY tensor, not an undefined one, stands for "no noise": autograd::Function inputs must have a device)
	Tensor noise = raw_noise_std > 0.f ? torch::randn_like(raw.detach().select(-1, 3)).contiguous() : torch::empty({0}, F32Like(raw));
	variable_list out = CompositeFn::apply(raw, z_vals.detach(), rays_d.detach(), noise, double(raw_noise_std), white_bkgr);
	return {out[0], out[1], out[2], out[3], out[4]};
}
