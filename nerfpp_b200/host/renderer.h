// Drop-in NeRFRenderer<TEmbedder, TEmbedDirs, TNeRF>: the reference's public surface (src/NeRFRenderer.h:12-159) — result /
// parameter structs, Render / BatchifyRays / RenderRays public virtual, RunNetwork / RawToOutputs protected virtual — with the
// per-chunk pipeline running on the sm_100a kernels:
//   * ray prologue, z sampling and point generation are one kernel each (not ~40 ATen launches);
//   * RawToOutputs is ONE differentiable op forward and ONE backward (not ~25 launches + autograd bookkeeping);
//   * SamplePDF + sort(cat) is ONE kernel;
//   * for <CuHashEmbedder, CuSHEncoder, NeRFSmall> RunNetwork is hash-encode -> fused MLP with the SH basis evaluated once
//     per ray and the keep mask applied in the MLP epilogue; the coarse pass, which never receives a gradient
//     (SURVEY §9-Q3), runs without an autograd graph;
//   * for the same trio, a chunk rendered without an autograd graph in the parity configuration is ONE C-ABI call (nrf_render_raybatch_fwd):
//     the fine pass evaluates the importance samples only and the gathers walk neighbouring rays together;
//   * for <Embedder, Embedder, NeRF> at the BASELINE shape RunNetwork is ONE kernel: the positional embeddings of points and
//     directions are evaluated in the fused MLP's input stage (directions per ray), in inference and in training.
// Everything else (generic embedders / models, NDC, perturb > 0) follows the reference's ATen formulation.
#pragma once
#include <type_traits>

#include "embedders.h"
#include "models.h"
#include "render_ops.h"

struct NeRFRendererOutputs {
	torch::Tensor RGBMap,   ///< [num_rays, 3]
		DispMap,            ///< [num_rays]
		AccMap,             ///< [num_rays]
		Weights,            ///< [num_rays, num_samples]
		DepthMap;           ///< [num_rays]
};

struct NeRFRenderResult {
	NeRFRendererOutputs Outputs;
	torch::Tensor Raw;      ///< [num_rays, num_samples, 4]
	float Near, Far;
};

struct NeRFRenderParams {
	int NSamples{64};
	int NImportance{192};
	int Chunk{1024 * 32};
	bool ReturnRaw{false};
	bool LinDisp{false};
	float Perturb{0.f};
	bool WhiteBkgr{false};
	float RawNoiseStd{0.};
	bool Ndc{true};
	bool UseViewdirs{false};
	bool ReturnWeights{false};
	bool ThinRay{false};
	float RenderFactor{0};
	torch::Tensor BoundingBox{torch::Tensor()};
	float StochasticPreconditioningAlpha{0};
};

template <class TEmbedder, class TEmbedDirs, class TNeRF>
class NeRFRenderer {
protected:
	TEmbedder EmbedFn;
	TEmbedDirs EmbeddirsFn;
	TNeRF NeRF;

	static constexpr bool kFusedHashPath =
		std::is_same_v<TEmbedder, CuHashEmbedder> && std::is_same_v<TEmbedDirs, CuSHEncoder> && std::is_same_v<TNeRF, NeRFSmall>;

	static constexpr bool kFusedClassicPath =
		std::is_same_v<TEmbedder, ::Embedder> && std::is_same_v<TEmbedDirs, ::Embedder> && std::is_same_v<TNeRF, ::NeRF>;

	/// embed -> (dir embed, cat) -> model -> sigma := 0 outside the box (src/NeRFRenderer.h:164-194)
	virtual torch::Tensor RunNetwork(torch::Tensor inputs, torch::Tensor view_dirs, TNeRF fn, TEmbedder embed_fn, TEmbedDirs embeddirs_fn)
	{
		std::vector<int64_t> shape = inputs.sizes().vec();
		torch::Tensor flat = inputs.detach().reshape({-1, shape.back()});     // no gradient to sample positions (:173)
		const bool dirs = view_dirs.defined() && view_dirs.numel() != 0;
		torch::Tensor out;
		if constexpr (kFusedHashPath) {
			if (dirs && fn->Fused() && embed_fn->GetOutputDims() == fn->GetInputCh() && embeddirs_fn->GetOutputDims() == fn->GetInputChViews()
				&& inputs.dim() == 3) {
				// SH once per RAY (the reference expands the directions per sample, :179-181)
				torch::Tensor ray_sh = embeddirs_fn->forward(view_dirs.detach()).first;
				out = nrfhost::HashNeRFNetwork(*embed_fn, *fn, flat, ray_sh, int(shape[1]));
				shape.back() = out.size(-1);
				return out.view(shape);
			}
		}
		if constexpr (kFusedClassicPath) {
			// positional embeddings evaluated inside the MLP kernel's input stage, directions per RAY: no [N,63] / [N,27] / [N,90] arrays.
			// When autograd records and the fused training kernels are switched off, or the shape is not the built one, fall through.
			if (dirs && inputs.dim() == 3 && flat.is_cuda() && fn->FusedEmbeddingShape(*embed_fn, *embeddirs_fn) &&
				(!torch::GradMode::is_enabled() || fn->FusedTraining)) {
				out = fn->ForwardPoints(flat, view_dirs, int(shape[1]), embed_fn->GetFreqBands(), embeddirs_fn->GetFreqBands());
				shape.back() = out.size(-1);
				return out.view(shape);
			}
		}
		auto [embedded, keep] = embed_fn->forward(flat);
		if (dirs) {
			torch::Tensor d = view_dirs.unsqueeze(-2).expand(inputs.sizes()).reshape({-1, view_dirs.size(-1)});
			embedded = torch::cat({embedded, embeddirs_fn->forward(d.contiguous()).first}, -1);
		}
		out = fn->forward(embedded);
		if (keep.defined() && keep.numel() != 0) {
			// same effect as out.index_put_({~keep, -1}, 0) (:188) without the nonzero() host sync
			torch::Tensor sigma = out.select(-1, out.size(-1) - 1) * keep.to(out.dtype());
			out = torch::cat({out.narrow(-1, 0, out.size(-1) - 1), sigma.unsqueeze(-1)}, -1);
		}
		shape.back() = out.size(-1);
		return out.view(shape);
	}

	/// alpha compositing (src/NeRFRenderer.h:199-282); cone_angle is accepted and unused, as in the reference (:246-248)
	virtual NeRFRendererOutputs RawToOutputs(torch::Tensor raw, torch::Tensor cone_angle, torch::Tensor z_vals, torch::Tensor rays_d,
		const float raw_noise_std = 0.f, const bool white_bkgr = false)
	{
		nrfhost::CompositeResult c = nrfhost::Composite(raw, z_vals, rays_d, raw_noise_std, white_bkgr);
		NeRFRendererOutputs o;
		o.RGBMap = c.rgb; o.DispMap = c.disp; o.AccMap = c.acc; o.Weights = c.weights; o.DepthMap = c.depth;
		return o;
	}

	torch::Tensor RenderWs;   ///< scratch of the one-call inference path, grown on demand and reused by every chunk

	/// true when RenderRays can run as ONE C-ABI call (nrf_render_raybatch_fwd): <CuHashEmbedder, CuSHEncoder, NeRFSmall> at a shape the fused
	/// kernels cover, no autograd graph, parity configuration (thin rays, no jitter / noise / preconditioning), rows [o d near far viewdirs]
	bool FusedInference(const torch::Tensor& ray_batch, const torch::Tensor& cone_angle, int n_samples, bool return_raw, float perturb, int n_importance,
		float raw_noise_std, float sp_alpha)
	{
		if constexpr (!kFusedHashPath) return false;
		else {
			if (!UseFusedInference || torch::GradMode::is_enabled()) return false;
			if (!ray_batch.is_cuda() || ray_batch.dim() != 2 || ray_batch.size(1) != 11) return false;
			if ((cone_angle.defined() && cone_angle.numel() != 0) || perturb != 0.f || raw_noise_std != 0.f || sp_alpha != 0.f || return_raw) return false;
			if (n_importance < 1 || n_samples < 3 || n_samples + n_importance > 1024) return false;
			const int deg = EmbeddirsFn->GetDegree();
			return NeRF->FusedPerRay() && EmbedFn->GetOutputDims() == NeRF->GetInputCh() && NeRF->GetInputChViews() == deg * deg && deg >= 1 && deg <= 8;
		}
	}

	/// RenderRays as one call: ray prologue from the prepared batch, both passes (the fine one evaluates the importance samples only: one network
	/// for both, src/NeRFRenderer.h:422,447), SamplePDF + merge, RawToOutputs.  Same kernels as the staged path below, hence the same bits.
	NeRFRenderResult RenderRaysFused(const torch::Tensor& ray_batch, int n_samples, bool lin_disp, int n_importance, bool white_bkgr, bool return_weights)
	{
		NeRFRenderResult result;
		if constexpr (kFusedHashPath) {
			torch::NoGradGuard no_grad;
			torch::Tensor rb = nrfhost::Dense(ray_batch.detach(), torch::kFloat32, "ray_batch");
			const int64_t R = rb.size(0);
			nrf_render_config cfg{};
			cfg.n_samples = n_samples; cfg.n_importance = n_importance; cfg.white_bkgr = white_bkgr ? 1 : 0; cfg.lin_disp = lin_disp ? 1 : 0;
			cfg.sh_degree = EmbeddirsFn->GetDegree(); cfg.near_plane = 0.f;                       // bbox / near_plane are not read for a prepared batch
			const nrf_hash_grid grid = EmbedFn->Grid();
			const nrf_mlp_small_shape shape = NeRF->Shape();
			const int64_t need = nrf_render_rays_workspace_bytes(&cfg, &grid, R);
			TORCH_CHECK(need >= 0, "nrf_render_rays_workspace_bytes: ", nrf_last_error());
			if (!RenderWs.defined() || RenderWs.numel() < need || RenderWs.device() != rb.device())
				RenderWs = torch::empty({std::max<int64_t>(need, 256)}, torch::TensorOptions().dtype(torch::kUInt8).device(rb.device()));
			torch::Tensor t_vals = nrfhost::UnitLinspace(n_samples, rb.device()), u = nrfhost::UnitLinspace(n_importance, rb.device());
			torch::Tensor table = EmbedFn->ShadowF16(), packed = NeRF->Packed();
			NeRFRendererOutputs& o = result.Outputs;
			o.RGBMap = torch::empty({R, 3}, nrfhost::F32Like(rb));
			o.DepthMap = torch::empty({R}, nrfhost::F32Like(rb));
			o.DispMap = torch::empty({R}, nrfhost::F32Like(rb));
			o.AccMap = torch::empty({R}, nrfhost::F32Like(rb));
			if (return_weights) o.Weights = torch::empty({R, n_samples + n_importance}, nrfhost::F32Like(rb));
			nrfhost::Check(nrf_render_raybatch_fwd(&cfg, &grid, table.data_ptr(), &shape, packed.data_ptr(), rb.data_ptr<float>(), int32_t(rb.size(1)), R,
				t_vals.data_ptr<float>(), u.data_ptr<float>(), RenderWs.data_ptr(), RenderWs.numel(), o.RGBMap.data_ptr<float>(), o.DepthMap.data_ptr<float>(),
				o.DispMap.data_ptr<float>(), o.AccMap.data_ptr<float>(), return_weights ? o.Weights.data_ptr<float>() : nullptr, nullptr, nrfhost::Stream()),
				"nrf_render_raybatch_fwd");
		}
		return result;
	}

public:
	/// false: inference goes stage by stage like training (A/B and tests); true (default): one C-ABI call per chunk when FusedInference() holds
	bool UseFusedInference = true;
	/// staged inference: the fine pass evaluates the importance samples only and takes the coarse samples' raw rows from the coarse pass
	bool ReuseCoarseRows = true;

	NeRFRenderer(TEmbedder embed_fn, TEmbedDirs embeddirs_fn, TNeRF nerf) : EmbedFn(embed_fn), EmbeddirsFn(embeddirs_fn), NeRF(nerf) {}
	virtual ~NeRFRenderer() {}

	/// one chunk (src/NeRFRenderer.h:366-459); ray_batch rows are [o(3) d(3) near far (viewdirs 3)]
	virtual NeRFRenderResult RenderRays(torch::Tensor ray_batch, torch::Tensor cone_angle, const int n_samples, const bool return_raw = false,
		const bool lin_disp = false, const float perturb = 0.f, const int n_importance = 0, const bool white_bkgr = false,
		const float raw_noise_std = 0.f, const float stochastic_preconditioning_alpha = 0.f, torch::Tensor bounding_box = torch::Tensor(),
		const bool return_weights = true)
	{
		if (FusedInference(ray_batch, cone_angle, n_samples, return_raw, perturb, n_importance, raw_noise_std, stochastic_preconditioning_alpha))
			return RenderRaysFused(ray_batch, n_samples, lin_disp, n_importance, white_bkgr, return_weights);
		NeRFRenderResult result;
		const torch::Device device = ray_batch.device();
		torch::Tensor rb = nrfhost::Dense(ray_batch.detach(), torch::kFloat32, "ray_batch");
		torch::Tensor rays_d = rb.narrow(1, 3, 3).contiguous();
		torch::Tensor viewdirs = rb.size(1) > 8 ? rb.narrow(1, rb.size(1) - 3, 3).contiguous() : torch::Tensor();

		torch::Tensor z_vals = nrfhost::ZSample(rb, n_samples, lin_disp);                                  // :393-402
		if (perturb > 0.f) {                                                                                // stratified jitter (:404-417)
			torch::Tensor mids = 0.5 * (z_vals.slice(1, 1) + z_vals.slice(1, 0, -1));
			torch::Tensor upper = torch::cat({mids, z_vals.slice(1, -1)}, -1), lower = torch::cat({z_vals.slice(1, 0, 1), mids}, -1);
			torch::Tensor width = upper - lower;
			torch::Tensor t_rand = torch::rand(z_vals.sizes(), torch::TensorOptions().dtype(torch::kFloat32).device(device));
			z_vals = (lower + torch::where(width > 1e-8f, width * t_rand, torch::zeros_like(width))).contiguous();
		}
		torch::Tensor pts = nrfhost::SamplePoints(rb, z_vals);                                              // :419
		pts = TangentScatter(pts, z_vals, cone_angle, rays_d, device, bounding_box);                        // :420

		NeRFRendererOutputs coarse;
		torch::Tensor raw;
		{
			// with importance sampling the coarse pass only feeds SamplePDF, whose output is detached (:429): no graph needed
			torch::AutoGradMode grad(n_importance > 0 ? false : torch::GradMode::is_enabled());
			raw = RunNetwork(pts, viewdirs, NeRF, EmbedFn, EmbeddirsFn);                                    // :422
			coarse = RawToOutputs(raw, cone_angle, z_vals, rays_d, raw_noise_std, white_bkgr);              // :423
		}
		// Inference in the parity configuration: the merged list holds the coarse samples bit for bit and ONE network serves both passes (:422,447),
		// so their raw rows are the coarse pass's own — the fine pass evaluates the n_importance NEW samples only (any embedder / model with 4-channel
		// rows; a model evaluated through cuBLAS may round a row differently in another batch shape, the fused models do not).
		if (n_importance > 0 && ReuseCoarseRows && !torch::GradMode::is_enabled() && perturb == 0.f && stochastic_preconditioning_alpha == 0.f &&
			!(cone_angle.defined() && cone_angle.numel() != 0) && raw.dim() == 3 && raw.size(-1) == 4) {
			auto [z_merged, perm, raw_merged] = nrfhost::SamplePdfMergeRows(z_vals, coarse.Weights, n_importance, raw);
			torch::Tensor at = perm.narrow(1, 0, n_importance).to(torch::kLong);                           // merged positions of the importance samples
			torch::Tensor z_new = torch::gather(z_merged, 1, at).contiguous();
			torch::Tensor raw_new = RunNetwork(nrfhost::SamplePoints(rb, z_new), viewdirs, NeRF, EmbedFn, EmbeddirsFn);   // [R, N, 4]
			raw_merged.scatter_(1, at.unsqueeze(-1).expand({-1, -1, 4}), raw_new.to(torch::kFloat32));
			z_vals = z_merged;
			raw = raw_merged;
			result.Outputs = RawToOutputs(raw, cone_angle, z_vals, rays_d, raw_noise_std, white_bkgr);      // :448
		} else if (n_importance > 0) {
			if (perturb == 0.f) {
				z_vals = nrfhost::SamplePdfMerge(z_vals, coarse.Weights, n_importance);                     // :427-431 in one kernel
			} else {
				torch::Tensor mids = 0.5 * (z_vals.slice(1, 1) + z_vals.slice(1, 0, -1));
				torch::Tensor z_samples = SamplePDF(mids, coarse.Weights.slice(1, 1, -1), n_importance, false).detach();
				z_vals = std::get<0>(torch::sort(torch::cat({z_vals, z_samples}, -1), -1)).contiguous();
			}
			pts = nrfhost::SamplePoints(rb, z_vals);                                                        // :432
			if (stochastic_preconditioning_alpha > 0.f) {                                                   // :435-443
				std::vector<torch::Tensor> bounds = torch::split(bounding_box.to(device), {3, 3}, -1);
				pts = ReflectBoundary(pts + torch::randn_like(pts) * stochastic_preconditioning_alpha, bounds[0], bounds[1]).contiguous();
			}
			pts = TangentScatter(pts, z_vals, cone_angle, rays_d, device, bounding_box);                    // :445
			raw = RunNetwork(pts, viewdirs, NeRF, EmbedFn, EmbeddirsFn);                                    // :447
			result.Outputs = RawToOutputs(raw, cone_angle, z_vals, rays_d, raw_noise_std, white_bkgr);      // :448
		} else {
			result.Outputs = coarse;   // the reference leaves Outputs empty here; the coarse result is a superset
		}
		if (return_raw) result.Raw = raw;
		if (!return_weights) result.Outputs.Weights = torch::Tensor();
		return result;
	}

	/// chunk loop + concatenation (src/NeRFRenderer.h:465-525)
	virtual NeRFRenderResult BatchifyRays(torch::Tensor rays_flat, torch::Tensor cone_angle, const int n_samples, const int chunk = 1024 * 32,
		const bool return_raw = false, const bool lin_disp = false, const float perturb = 0.f, const int n_importance = 0,
		const bool white_bkgr = false, const float raw_noise_std = 0., const float stochastic_preconditioning_alpha = 0.f,
		torch::Tensor bounding_box = torch::Tensor(), const bool return_weights = true)
	{
		const int64_t n = rays_flat.size(0);
		std::vector<torch::Tensor> rgb, disp, acc, weights, depth, raw;
		for (int64_t i = 0; i < n; i += chunk) {
			NeRFRenderResult part = RenderRays(rays_flat.slice(0, i, std::min<int64_t>(i + chunk, n)), cone_angle, n_samples, return_raw, lin_disp,
				perturb, n_importance, white_bkgr, raw_noise_std, stochastic_preconditioning_alpha, bounding_box, return_weights);
			auto keep = [](std::vector<torch::Tensor>& v, const torch::Tensor& t) { if (t.defined()) v.push_back(t); };
			keep(rgb, part.Outputs.RGBMap); keep(disp, part.Outputs.DispMap); keep(acc, part.Outputs.AccMap);
			keep(weights, part.Outputs.Weights); keep(depth, part.Outputs.DepthMap); keep(raw, part.Raw);
		}
		auto join = [](const std::vector<torch::Tensor>& v) { return v.empty() ? torch::Tensor() : (v.size() == 1 ? v[0] : torch::cat(v, 0)); };
		NeRFRenderResult result;
		result.Outputs.RGBMap = join(rgb); result.Outputs.DispMap = join(disp); result.Outputs.AccMap = join(acc);
		result.Outputs.Weights = join(weights); result.Outputs.DepthMap = join(depth); result.Raw = join(raw);
		return result;
	}

	/// whole image (c2w given) or a ray batch (src/NeRFRenderer.h:530-604)
	virtual NeRFRenderResult Render(const int h, const int w, torch::Tensor k, const NeRFRenderParams& render_params,
		std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> rays = {torch::Tensor(), torch::Tensor(), torch::Tensor()},
		torch::Tensor c2w = torch::Tensor(), torch::Tensor c2w_staticcam = torch::Tensor())
	{
		torch::Tensor rays_o, rays_d, cone_angle;
		if (c2w.defined() && c2w.numel() != 0) std::tie(rays_o, rays_d, cone_angle) = GetRays(h, w, k, c2w);
		else std::tie(rays_o, rays_d, cone_angle) = rays;
		torch::Tensor view_src = rays_d;                                       // viewing directions come from c2w (:553)
		if (render_params.UseViewdirs && c2w_staticcam.defined() && c2w_staticcam.numel() != 0)
			std::tie(rays_o, rays_d, cone_angle) = GetRays(h, w, k, c2w_staticcam);
		const std::vector<int64_t> image_shape = rays_d.sizes().vec();
		if (render_params.Ndc)
			std::tie(rays_o, rays_d, cone_angle) = NDCRays(h, w, k[0][0].template item<float>(), 1.f, rays_o, rays_d,
				render_params.ThinRay ? torch::Tensor() : cone_angle);

		// viewdirs, AABB near/far and the [o d near far viewdirs] rows in one kernel (:549-583)
		torch::Tensor rays_;
		const bool same_dirs = view_src.is_same(rays_d) || !render_params.UseViewdirs;
		rays_ = nrfhost::RaysPrepare(rays_o, rays_d, render_params.BoundingBox, 0.f, render_params.UseViewdirs && same_dirs);
		if (render_params.UseViewdirs && !same_dirs) {
			torch::Tensor v = view_src.reshape({-1, 3});
			rays_ = torch::cat({rays_, v / torch::norm(v, 2, -1, true)}, -1);
		}
		NeRFRenderResult all_ret = BatchifyRays(rays_, render_params.ThinRay ? torch::Tensor() : cone_angle, render_params.NSamples,
			render_params.Chunk, render_params.ReturnRaw, render_params.LinDisp, render_params.Perturb, render_params.NImportance,
			render_params.WhiteBkgr, render_params.RawNoiseStd, render_params.StochasticPreconditioningAlpha, render_params.BoundingBox,
			render_params.ReturnWeights);

		if (all_ret.Outputs.RGBMap.defined() && all_ret.Outputs.RGBMap.numel() != 0) all_ret.Outputs.RGBMap = all_ret.Outputs.RGBMap.reshape(image_shape);
		if (image_shape.size() > 2) {
			if (all_ret.Outputs.DispMap.defined()) all_ret.Outputs.DispMap = all_ret.Outputs.DispMap.reshape({image_shape[0], image_shape[1]});
			if (all_ret.Outputs.DepthMap.defined()) all_ret.Outputs.DepthMap = all_ret.Outputs.DepthMap.reshape({image_shape[0], image_shape[1]});
		}
		// one host read for both scalars (the reference issues two .item() syncs, :602-603)
		torch::Tensor nf = torch::stack({rays_.select(1, 6).min(), rays_.select(1, 7).max()}).cpu();
		all_ret.Near = nf[0].template item<float>();
		all_ret.Far = nf[1].template item<float>();
		return all_ret;
	}
};
