// LeRF head and LeRFRenderer of the drop-in layer (see lerf.h).  All hot-path arithmetic is behind the C ABI.
#include "lerf.h"

using torch::Tensor;
namespace F = torch::nn::functional;

// ------------------------------------------------------------------------------------------------ LeRF (src/LeRF.cpp)
LeRFImpl::LeRFImpl(const int geo_feat_dim_le, const int num_layers_le, const int hidden_dim_le, const int lang_embed_dim, const int input_ch_le,
	const std::string module_name)
	: BaseNeRFImpl(module_name), GeoFeatDimLE(geo_feat_dim_le), NumLayersLE(num_layers_le), HiddenDimLE(hidden_dim_le), LangEmbedDim(lang_embed_dim),
	  InputChLE(input_ch_le)
{
	// layer widths and registration order as src/LeRF.cpp:11-25 (the order fixes parameters() and therefore the optimizer state layout)
	for (int l = 0; l < NumLayersLE; l++)
		SigmaLENet->push_back(torch::nn::Linear(torch::nn::LinearOptions(l == 0 ? InputChLE : HiddenDimLE,
			l == NumLayersLE - 1 ? 1 + GeoFeatDimLE : HiddenDimLE).bias(false)));
	for (int l = 0; l < NumLayersLE; l++)
		LENet->push_back(torch::nn::Linear(torch::nn::LinearOptions(l == 0 ? GeoFeatDimLE + InputChLE : HiddenDimLE,
			l == NumLayersLE - 1 ? LangEmbedDim : HiddenDimLE).bias(false)));
	for (size_t i = 0; i < SigmaLENet->size(); i++) register_module(module_name + "_sigma_le_net_" + std::to_string(i), SigmaLENet[i]);
	for (size_t i = 0; i < LENet->size(); i++) register_module(module_name + "_le_net_" + std::to_string(i), LENet[i]);
}

bool LeRFImpl::Fused() const
{
	const nrf_lerf_shape s = Shape();
	return nrf_lerf_packed_bytes(&s) > 0;
}

nrf_lerf_shape LeRFImpl::Shape() const
{
	nrf_lerf_shape s;
	s.geo_feat_dim = GeoFeatDimLE; s.num_layers = NumLayersLE; s.hidden_dim = HiddenDimLE; s.lang_embed_dim = LangEmbedDim; s.input_ch = InputChLE;
	return s;
}

std::vector<Tensor> LeRFImpl::Weights()
{
	std::vector<Tensor> w;
	for (size_t i = 0; i < SigmaLENet->size(); i++) w.push_back(SigmaLENet[i]->as<torch::nn::Linear>()->weight);
	for (size_t i = 0; i < LENet->size(); i++) w.push_back(LENet[i]->as<torch::nn::Linear>()->weight);
	return w;
}

Tensor LeRFImpl::Packed()
{
	std::vector<Tensor> w = Weights();
	std::vector<std::pair<const void*, uint32_t>> key;
	for (const Tensor& t : w) key.emplace_back(t.data_ptr(), t._version());
	if (!PackedBlob.defined() || key != PackedKey) {
		torch::NoGradGuard no_grad;
		std::vector<Tensor> dense;
		for (const Tensor& t : w) dense.push_back(nrfhost::Dense(t.detach(), torch::kFloat32, "LeRF weight"));
		const nrf_lerf_shape s = Shape();
		nrf_lerf_weights p{dense[0].data_ptr<float>(), dense[1].data_ptr<float>(), dense[2].data_ptr<float>(), dense[3].data_ptr<float>()};
		PackedBlob = torch::empty({nrf_lerf_packed_bytes(&s)}, torch::TensorOptions().dtype(torch::kUInt8).device(dense[0].device()));
		nrfhost::Check(nrf_lerf_pack(&s, &p, PackedBlob.data_ptr(), nrfhost::Stream()), "nrf_lerf_pack");
		PackedKey = key;
	}
	return PackedBlob;
}

Tensor LeRFImpl::ForwardAten(const Tensor& x)
{
	// src/LeRF.cpp:78-110
	Tensor h = x;
	for (size_t i = 0; i < SigmaLENet->size(); i++) {
		h = SigmaLENet[i]->as<torch::nn::Linear>()->forward(h);
		if (i != SigmaLENet->size() - 1) h = torch::relu(h);
	}
	Tensor sigma_le = h.select(-1, 0);
	Tensor geo = h.slice(-1, 1);
	h = torch::cat({geo, x}, -1);
	for (size_t i = 0; i < LENet->size(); i++) {
		h = LENet[i]->as<torch::nn::Linear>()->forward(h);
		if (i != LENet->size() - 1) h = torch::relu(h);
	}
	Tensor le = F::normalize(h, F::NormalizeFuncOptions().dim(-1).eps(1e-8));
	return torch::cat({le, sigma_le.unsqueeze(-1)}, -1);
}

static bool AnyRequiresGrad(const std::vector<Tensor>& ts)
{
	for (const Tensor& t : ts) if (t.requires_grad()) return true;
	return false;
}

Tensor LeRFImpl::forward(Tensor x)
{
	if (NumLayersLE <= 0 || x.numel() == 0) return ForwardAten(x);
	if (!Fused()) return ForwardAten(x);
	TORCH_CHECK(x.is_cuda(), "nerfpp_b200: LeRF::forward needs a CUDA tensor at the fused shape (the sm_100a path has no CPU fallback)");
	const bool recording = torch::GradMode::is_enabled() && (x.requires_grad() || AnyRequiresGrad(Weights()));
	// a bare LeRF::forward under autograd returns the [N, 513] tensor the reference returns, through torch::linear; the TRAINING path of the
	// drop-in is LeRFRenderer::RenderRays, whose fine pass is one fused autograd node (LerfFineFn) that never forms this tensor
	if (recording) return ForwardAten(x);
	std::vector<int64_t> shape = x.sizes().vec();
	Tensor enc = nrfhost::Dense(x.detach().reshape({-1, shape.back()}), torch::kFloat16, "LeRF input");
	Tensor packed = Packed();
	const nrf_lerf_shape s = Shape();
	Tensor out = torch::empty({enc.size(0), LangEmbedDim + 1}, torch::TensorOptions().dtype(torch::kFloat32).device(enc.device()));
	nrfhost::Check(nrf_lerf_fwd(&s, packed.data_ptr(), enc.data_ptr(), nullptr, enc.size(0), nrfhost::Ptr<float>(out), nrfhost::Stream()), "nrf_lerf_fwd");
	shape.back() = LangEmbedDim + 1;
	return out.view(shape);
}

// ------------------------------------------------------------------------------------------------ fused training of the fine pass
namespace {

using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

// RunLENetwork + RawToLEOutputs of the FINE pass under autograd as ONE node (src/LeRFRenderer.cpp:150,165-166 with src/LeRF.cpp:78-110 and
// src/LeRFRenderer.h:45-54 inside): hash encode (F = 8) -> bf16 tcgen05 head (nrf_lerf_fwd_train) -> compositing -> per-ray projection.  The
// backward is the fused chain of csrc/lerf_bwd_tc.cu; the [N, 512] embedding is formed in neither direction.  Inputs 0..4 are the leaves that
// receive gradients (lang_embedder Embeddings, the four Linear weights); points / z / rays_d are detached upstream (:147).
struct LerfFineFn : public torch::autograd::Function<LerfFineFn> {
	static variable_list forward(AutogradContext* ctx, Tensor embeddings, Tensor w_s0, Tensor w_s1, Tensor w_e0, Tensor w_e1, Tensor points, Tensor z,
		Tensor rays_d, int64_t lerf_module, int64_t embed_module)
	{
		auto* lerf = reinterpret_cast<LeRFImpl*>(lerf_module);
		auto* embed = reinterpret_cast<CuHashEmbedderImpl*>(embed_module);
		const int64_t r = z.size(0), ns = z.size(1), n = r * ns;
		const nrf_lerf_shape s = lerf->Shape();
		auto [enc, keep] = embed->EncodeF16(points);
		Tensor packed = lerf->Packed();
		const auto f32 = nrfhost::F32Like(z);
		const auto u8 = torch::TensorOptions().dtype(torch::kUInt8).device(z.device());
		Tensor raw4 = torch::empty({r, ns, 4}, f32), q = torch::empty({n}, f32);
		Tensor saved = torch::empty({nrf_lerf_train_saved_bytes(&s, n)}, u8);
		nrfhost::Check(nrf_lerf_fwd_train(&s, packed.data_ptr(), enc.data_ptr(), reinterpret_cast<const uint8_t*>(keep.data_ptr()), n, nrfhost::Ptr<float>(raw4),
			saved.data_ptr(), nrfhost::Ptr<float>(q), nrfhost::Stream()), "nrf_lerf_fwd_train");
		Tensor depth = torch::empty({r}, f32), disp = torch::empty({r}, f32), acc = torch::empty({r}, f32), weights = torch::empty({r, ns}, f32);
		Tensor rgb = torch::empty({r, 3}, f32);
		nrfhost::Check(nrf_composite_fwd(nrfhost::CPtr<float>(raw4), 4, nrfhost::CPtr<float>(z), nrfhost::CPtr<float>(rays_d), nullptr, 0.f, 0, r, int32_t(ns),
			nrfhost::Ptr<float>(rgb), nrfhost::Ptr<float>(depth), nrfhost::Ptr<float>(disp), nrfhost::Ptr<float>(acc), nrfhost::Ptr<float>(weights), nrfhost::Stream()),
			"nrf_composite_fwd");
		Tensor hsum = torch::empty({r, s.hidden_dim}, f32), rendered = torch::empty({r, s.lang_embed_dim}, f32), enorm = torch::empty({r}, f32);
		nrfhost::Check(nrf_lerf_render_embedding_train(&s, packed.data_ptr(), nrfhost::CPtr<float>(weights), saved.data_ptr(), nrfhost::CPtr<float>(q), r, int32_t(ns),
			nrfhost::Ptr<float>(hsum), nrfhost::Ptr<float>(rendered), nrfhost::Ptr<float>(enorm), nrfhost::Stream()), "nrf_lerf_render_embedding_train");
		ctx->save_for_backward({w_s0, w_s1, w_e0, w_e1, points, z, rays_d, keep, raw4, q, saved, weights, hsum, rendered, enorm, packed});
		ctx->saved_data["lerf"] = lerf_module;
		ctx->saved_data["embed"] = embed_module;
		ctx->saved_data["embed_numel"] = embeddings.numel();
		ctx->mark_non_differentiable({weights, depth, disp, acc});
		return {rendered, weights, depth, disp, acc};
	}

	static variable_list backward(AutogradContext* ctx, variable_list g)
	{
		auto sv = ctx->get_saved_variables();
		Tensor w_s0 = sv[0], w_s1 = sv[1], w_e0 = sv[2], w_e1 = sv[3], points = sv[4], z = sv[5], rays_d = sv[6], keep = sv[7], raw4 = sv[8], q = sv[9],
		       saved = sv[10], weights = sv[11], hsum = sv[12], rendered = sv[13], enorm = sv[14], packed = sv[15];
		auto* lerf = reinterpret_cast<LeRFImpl*>(ctx->saved_data["lerf"].toInt());
		auto* embed = reinterpret_cast<CuHashEmbedderImpl*>(ctx->saved_data["embed"].toInt());
		const int64_t r = z.size(0), ns = z.size(1), n = r * ns;
		const nrf_lerf_shape s = lerf->Shape();
		const auto f32 = nrfhost::F32Like(z);
		Tensor grad_table = torch::zeros({ctx->saved_data["embed_numel"].toInt()}, f32).view(embed->Embeddings.sizes());
		Tensor g0 = torch::zeros_like(w_s0), g1 = torch::zeros_like(w_s1), g2 = torch::zeros_like(w_e0), g3 = torch::zeros_like(w_e1);
		if (!g[0].defined()) return {grad_table, g0, g1, g2, g3, Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
		Tensor g_r = nrfhost::Dense(g[0], torch::kFloat32, "d loss / d RenderedLangEmbedding");
		// dense fp32 views of the weights as they were at forward time (the blob `packed` was built from them)
		Tensor d0 = nrfhost::Dense(w_s0.detach(), torch::kFloat32, "w"), d1 = nrfhost::Dense(w_s1.detach(), torch::kFloat32, "w"),
		       d2 = nrfhost::Dense(w_e0.detach(), torch::kFloat32, "w"), d3 = nrfhost::Dense(w_e1.detach(), torch::kFloat32, "w");
		const nrf_lerf_weights wp{d0.data_ptr<float>(), d1.data_ptr<float>(), d2.data_ptr<float>(), d3.data_ptr<float>()};
		const nrf_lerf_weights gp{g0.data_ptr<float>(), g1.data_ptr<float>(), g2.data_ptr<float>(), g3.data_ptr<float>()};
		Tensor ws = torch::empty({nrf_lerf_bwd_workspace_bytes(&s, n, r)}, torch::TensorOptions().dtype(torch::kUInt8).device(z.device()));
		Tensor dw = torch::empty({r, ns}, f32), d_raw4 = torch::empty({r, ns, 4}, f32);
		nrfhost::Check(nrf_lerf_bwd_rays(&s, &wp, saved.data_ptr(), nrfhost::CPtr<float>(q), nrfhost::CPtr<float>(weights), nrfhost::CPtr<float>(hsum),
			nrfhost::CPtr<float>(rendered), nrfhost::CPtr<float>(enorm), nullptr, nrfhost::CPtr<float>(g_r), r, int32_t(ns), 1.f, nullptr, g3.data_ptr<float>(),
			ws.data_ptr(), nrfhost::Ptr<float>(dw), nrfhost::Stream()), "nrf_lerf_bwd_rays");
		nrfhost::Check(nrf_composite_bwd(nrfhost::CPtr<float>(raw4), 4, nrfhost::CPtr<float>(z), nrfhost::CPtr<float>(rays_d), nullptr, 0.f, 0, r, int32_t(ns),
			nullptr, nullptr, nullptr, nullptr, nrfhost::CPtr<float>(dw), nrfhost::Ptr<float>(d_raw4), nrfhost::Stream()), "nrf_composite_bwd");
		Tensor d_enc = torch::empty({n, s.input_ch}, torch::TensorOptions().dtype(torch::kBFloat16).device(z.device()));
		nrfhost::Check(nrf_lerf_bwd_rows(&s, packed.data_ptr(), &wp, saved.data_ptr(), reinterpret_cast<const uint8_t*>(keep.data_ptr()), nrfhost::CPtr<float>(d_raw4),
			n, int32_t(ns), ws.data_ptr(), &gp, d_enc.data_ptr(), nrfhost::Stream()), "nrf_lerf_bwd_rows");
		embed->Backward(points, d_enc, grad_table);
		return {grad_table, g0, g1, g2, g3, Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
	}
};

}  // namespace

bool LeRFRenderer::FusedTraining(const Tensor& ray_batch, const Tensor& cone_angle, float perturb, int n_importance, float raw_noise_std,
	float stochastic_preconditioning_alpha)
{
	const bool recording = torch::GradMode::is_enabled() && (AnyRequiresGrad(Lerf->Weights()) || LangEmbedFn->Embeddings.requires_grad());
	const bool thin = !(cone_angle.defined() && cone_angle.numel() != 0);
	return UseFusedTraining && recording && ray_batch.is_cuda() && Lerf->Fused() && LangEmbedFn->GetOutputDims() == Lerf->Shape().input_ch && thin &&
		perturb == 0.f && n_importance > 0 && raw_noise_std == 0.f && stochastic_preconditioning_alpha == 0.f;
}

// ------------------------------------------------------------------------------------------------ LeRFRenderer (src/LeRFRenderer.cpp)
Tensor LeRFRenderer::RunLENetwork(Tensor inputs, LeRF lerf, CuHashEmbedder lang_embed_fn)
{
	std::vector<int64_t> shape = inputs.sizes().vec();
	Tensor flat = inputs.detach().reshape({-1, shape.back()});                                  // :11-12
	const bool recording = torch::GradMode::is_enabled() && (AnyRequiresGrad(lerf->Weights()) || lang_embed_fn->Embeddings.requires_grad());
	Tensor out;
	if (!recording && lerf->Fused() && lang_embed_fn->GetOutputDims() == lerf->Shape().input_ch && flat.is_cuda()) {
		auto [enc, keep] = lang_embed_fn->EncodeF16(flat);
		Tensor packed = lerf->Packed();
		const nrf_lerf_shape s = lerf->Shape();
		out = torch::empty({flat.size(0), lerf->GetLangEmbedDim() + 1}, nrfhost::F32Like(flat));
		nrfhost::Check(nrf_lerf_fwd(&s, packed.data_ptr(), enc.data_ptr(), reinterpret_cast<const uint8_t*>(keep.data_ptr()), flat.size(0),
			nrfhost::Ptr<float>(out), nrfhost::Stream()), "nrf_lerf_fwd");
	} else {
		auto [embedded, keep] = lang_embed_fn->forward(flat);                                   // :13
		out = lerf->forward(embedded);                                                          // :15
		if (keep.defined() && keep.numel() != 0) {
			// same effect as outputs_flat.index_put_({~keep_mask, -1}, 0) (:18-19) without the nonzero() host sync
			Tensor sigma = out.select(-1, out.size(-1) - 1) * keep.to(out.dtype());
			out = torch::cat({out.narrow(-1, 0, out.size(-1) - 1), sigma.unsqueeze(-1)}, -1);
		}
	}
	shape.back() = out.size(-1);
	return out.view(shape);
}

LeRFRendererOutputs LeRFRenderer::RawToLEOutputs(Tensor raw_le, Tensor z_vals_le, Tensor rays_d, const int lang_embed_dim, const float raw_noise_std)
{
	LeRFRendererOutputs result;
	result.LangEmbedding = raw_le.narrow(-1, 0, lang_embed_dim);                                 // :46
	// the density column in the [.., 4] layout of the compositing op (rgb unused): weights, depth, disp, acc as :53-72
	Tensor sigma = raw_le.select(-1, lang_embed_dim).unsqueeze(-1);
	Tensor raw4 = torch::cat({torch::zeros_like(sigma).expand({sigma.size(0), sigma.size(1), 3}), sigma}, -1);
	nrfhost::CompositeResult c = nrfhost::Composite(raw4, z_vals_le, rays_d, raw_noise_std, false);
	result.WeightsLE = c.weights; result.DepthMapLE = c.depth; result.DispMapLE = c.disp; result.AccMapLE = c.acc;
	result.RenderedLangEmbedding = RenderCLIPEmbedding(result.LangEmbedding, result.WeightsLE.unsqueeze(-1));   // :75
	if (RelevancyFn && LerfPositives.defined() && LerfNegatives.defined())
		result.Relevancy = RelevancyFn(result.RenderedLangEmbedding, LerfPositives.to(raw_le.device()), LerfNegatives.to(raw_le.device()));   // :79
	return result;
}

bool LeRFRenderer::FusedInference(const Tensor& ray_batch, const Tensor& cone_angle, float perturb, int n_importance, float raw_noise_std,
	float stochastic_preconditioning_alpha)
{
	const bool recording = torch::GradMode::is_enabled() && (AnyRequiresGrad(Lerf->Weights()) || LangEmbedFn->Embeddings.requires_grad());
	const bool thin = !(cone_angle.defined() && cone_angle.numel() != 0);
	return !recording && ray_batch.is_cuda() && Lerf->Fused() && LangEmbedFn->GetOutputDims() == Lerf->Shape().input_ch && thin && perturb == 0.f &&
		n_importance > 0 && raw_noise_std == 0.f && stochastic_preconditioning_alpha == 0.f;
}

LeRFRenderResult LeRFRenderer::RenderRays(Tensor ray_batch, Tensor cone_angle, const int n_samples, const bool return_raw, const bool lin_disp,
	const float perturb, const int n_importance, const bool white_bkgr, const float raw_noise_std, const float stochastic_preconditioning_alpha,
	Tensor bounding_box, const bool return_weights)
{
	LeRFRenderResult result;
	const torch::Device device = ray_batch.device();
	Tensor rb = nrfhost::Dense(ray_batch.detach(), torch::kFloat32, "ray_batch");
	Tensor rays_d = rb.narrow(1, 3, 3).contiguous();
	const int dim = Lerf->GetLangEmbedDim();

	if (FusedInference(rb, cone_angle, perturb, n_importance, raw_noise_std, stochastic_preconditioning_alpha)) {
		torch::NoGradGuard no_grad;
		const int64_t r = rb.size(0);
		const nrf_lerf_shape s = Lerf->Shape();
		Tensor packed = Lerf->Packed();
		Tensor z_vals = nrfhost::ZSample(rb, n_samples, lin_disp);                                          // :112-118
		// coarse pass: only the density of the language field is needed (its embedding is discarded, :140-147)
		auto [enc_c, keep_c] = LangEmbedFn->EncodeF16(nrfhost::SamplePoints(rb, z_vals).reshape({-1, 3}));
		Tensor raw4 = torch::empty({r, n_samples, 4}, nrfhost::F32Like(rb));
		nrfhost::Check(nrf_lerf_sigma_fwd(&s, packed.data_ptr(), enc_c.data_ptr(), reinterpret_cast<const uint8_t*>(keep_c.data_ptr()), r * n_samples,
			nrfhost::Ptr<float>(raw4), nrfhost::Stream()), "nrf_lerf_sigma_fwd");
		nrfhost::CompositeResult coarse = nrfhost::Composite(raw4, z_vals, rays_d, 0.f, false);
		z_vals = nrfhost::SamplePdfMerge(z_vals, coarse.weights, n_importance);                             // :145-149
		const int64_t ns = z_vals.size(1);
		auto [enc, keep] = LangEmbedFn->EncodeF16(nrfhost::SamplePoints(rb, z_vals).reshape({-1, 3}));      // :150
		raw4 = torch::empty({r, ns, 4}, nrfhost::F32Like(rb));
		Tensor hidden = torch::empty({nrf_lerf_hidden_bytes(&s, r * ns)}, torch::TensorOptions().dtype(torch::kUInt8).device(device));
		Tensor q = torch::empty({r * ns}, nrfhost::F32Like(rb));
		nrfhost::Check(nrf_lerf_hidden_fwd(&s, packed.data_ptr(), enc.data_ptr(), reinterpret_cast<const uint8_t*>(keep.data_ptr()), r * ns,
			nrfhost::Ptr<float>(raw4), hidden.data_ptr(), nrfhost::Ptr<float>(q), nrfhost::Stream()), "nrf_lerf_hidden_fwd");
		nrfhost::CompositeResult fine = nrfhost::Composite(raw4, z_vals, rays_d, 0.f, false);
		Tensor hsum = torch::empty({r, s.hidden_dim}, nrfhost::F32Like(rb));
		Tensor rendered = torch::empty({r, dim}, nrfhost::F32Like(rb));
		nrfhost::Check(nrf_lerf_render_embedding(&s, packed.data_ptr(), nrfhost::CPtr<float>(fine.weights), hidden.data_ptr(), nrfhost::CPtr<float>(q), r,
			int32_t(ns), nrfhost::Ptr<float>(hsum), nrfhost::Ptr<float>(rendered), nrfhost::Stream()), "nrf_lerf_render_embedding");
		result.Outputs.WeightsLE = fine.weights; result.Outputs.DepthMapLE = fine.depth; result.Outputs.DispMapLE = fine.disp;
		result.Outputs.AccMapLE = fine.acc; result.Outputs.RenderedLangEmbedding = rendered;
		if (return_raw || (MaterializeLangEmbedding && return_weights)) {
			Tensor raw = torch::empty({r, ns, dim + 1}, nrfhost::F32Like(rb));
			nrfhost::Check(nrf_lerf_fwd(&s, packed.data_ptr(), enc.data_ptr(), reinterpret_cast<const uint8_t*>(keep.data_ptr()), r * ns,
				nrfhost::Ptr<float>(raw), nrfhost::Stream()), "nrf_lerf_fwd");
			if (return_raw) result.Raw = raw;
			if (MaterializeLangEmbedding) result.Outputs.LangEmbedding = raw.narrow(-1, 0, dim);
		}
		if (RelevancyFn && LerfPositives.defined() && LerfNegatives.defined())
			result.Outputs.Relevancy = RelevancyFn(rendered, LerfPositives.to(device), LerfNegatives.to(device));
	} else if (FusedTraining(rb, cone_angle, perturb, n_importance, raw_noise_std, stochastic_preconditioning_alpha)) {
		// training at the built shape: the coarse pass only feeds SamplePDF, whose output is detached (:147) — density-only head, no graph;
		// the fine pass is ONE autograd node on the fused forward / backward kernels (LerfFineFn)
		const int64_t r = rb.size(0);
		const nrf_lerf_shape s = Lerf->Shape();
		Tensor z_vals;
		{
			torch::NoGradGuard no_grad;
			Tensor packed = Lerf->Packed();
			z_vals = nrfhost::ZSample(rb, n_samples, lin_disp);
			auto [enc_c, keep_c] = LangEmbedFn->EncodeF16(nrfhost::SamplePoints(rb, z_vals).reshape({-1, 3}));
			Tensor raw4 = torch::empty({r, n_samples, 4}, nrfhost::F32Like(rb));
			nrfhost::Check(nrf_lerf_sigma_fwd(&s, packed.data_ptr(), enc_c.data_ptr(), reinterpret_cast<const uint8_t*>(keep_c.data_ptr()), r * n_samples,
				nrfhost::Ptr<float>(raw4), nrfhost::Stream()), "nrf_lerf_sigma_fwd");
			nrfhost::CompositeResult coarse = nrfhost::Composite(raw4, z_vals, rays_d, 0.f, false);
			z_vals = nrfhost::SamplePdfMerge(z_vals, coarse.weights, n_importance);
		}
		Tensor pts = nrfhost::SamplePoints(rb, z_vals).reshape({-1, 3});
		std::vector<Tensor> w = Lerf->Weights();
		variable_list out = LerfFineFn::apply(LangEmbedFn->Embeddings, w[0], w[1], w[2], w[3], pts, z_vals, rays_d, reinterpret_cast<int64_t>(Lerf.get()),
			reinterpret_cast<int64_t>(LangEmbedFn.get()));
		result.Outputs.RenderedLangEmbedding = out[0]; result.Outputs.WeightsLE = out[1]; result.Outputs.DepthMapLE = out[2];
		result.Outputs.DispMapLE = out[3]; result.Outputs.AccMapLE = out[4];
		if (return_raw || MaterializeLangEmbedding) {
			torch::NoGradGuard no_grad;       // the reference's Raw / LangEmbedding tensors, values only (nothing downstream of them is differentiated)
			auto [enc, keep] = LangEmbedFn->EncodeF16(pts);
			Tensor packed = Lerf->Packed();
			Tensor raw = torch::empty({r, z_vals.size(1), dim + 1}, nrfhost::F32Like(rb));
			nrfhost::Check(nrf_lerf_fwd(&s, packed.data_ptr(), enc.data_ptr(), reinterpret_cast<const uint8_t*>(keep.data_ptr()), r * z_vals.size(1),
				nrfhost::Ptr<float>(raw), nrfhost::Stream()), "nrf_lerf_fwd");
			if (return_raw) result.Raw = raw;
			if (MaterializeLangEmbedding) result.Outputs.LangEmbedding = raw.narrow(-1, 0, dim);
		}
		if (RelevancyFn && LerfPositives.defined() && LerfNegatives.defined())
			result.Outputs.Relevancy = RelevancyFn(result.Outputs.RenderedLangEmbedding, LerfPositives.to(device), LerfNegatives.to(device));
	} else {
		// the reference's sequence (:112-172) on the drop-in pieces
		Tensor z_vals = nrfhost::ZSample(rb, n_samples, lin_disp);
		if (perturb > 0.f) {                                                                                // :120-135
			Tensor mids = 0.5 * (z_vals.slice(1, 1) + z_vals.slice(1, 0, -1));
			Tensor upper = torch::cat({mids, z_vals.slice(1, -1)}, -1), lower = torch::cat({z_vals.slice(1, 0, 1), mids}, -1);
			Tensor width = upper - lower;
			Tensor t_rand = torch::rand(z_vals.sizes(), torch::TensorOptions().dtype(torch::kFloat32).device(device));
			z_vals = (lower + torch::where(width > 1e-8f, width * t_rand, torch::zeros_like(width))).contiguous();
		}
		Tensor pts = TangentScatter(nrfhost::SamplePoints(rb, z_vals), z_vals, cone_angle, rays_d, device, bounding_box);   // :137-138
		LeRFRendererOutputs coarse;
		Tensor raw;
		{
			// with importance sampling the coarse pass only feeds SamplePDF, whose output is detached (:147): no graph needed
			torch::AutoGradMode grad(n_importance > 0 ? false : torch::GradMode::is_enabled());
			raw = RunLENetwork(pts, Lerf, LangEmbedFn);                                                     // :140
			coarse = RawToLEOutputs(raw, z_vals, rays_d, dim, raw_noise_std);                               // :141
		}
		if (n_importance > 0) {
			if (perturb == 0.f) {
				z_vals = nrfhost::SamplePdfMerge(z_vals, coarse.WeightsLE, n_importance);                   // :145-149 in one kernel
			} else {
				Tensor mids = 0.5 * (z_vals.slice(1, 1) + z_vals.slice(1, 0, -1));
				Tensor z_samples = SamplePDF(mids, coarse.WeightsLE.slice(1, 1, -1), n_importance, false).detach();
				z_vals = std::get<0>(torch::sort(torch::cat({z_vals, z_samples}, -1), -1)).contiguous();
			}
			pts = nrfhost::SamplePoints(rb, z_vals);                                                        // :150
			if (stochastic_preconditioning_alpha > 0.f) {                                                   // :153-161
				std::vector<Tensor> bounds = torch::split(bounding_box.to(device), {3, 3}, -1);
				pts = ReflectBoundary(pts + torch::randn_like(pts) * stochastic_preconditioning_alpha, bounds[0], bounds[1]).contiguous();
			}
			pts = TangentScatter(pts, z_vals, cone_angle, rays_d, device, bounding_box);                    // :163
			raw = RunLENetwork(pts, Lerf, LangEmbedFn);                                                     // :165
			result.Outputs = RawToLEOutputs(raw, z_vals, rays_d, dim, raw_noise_std);                       // :166
		}
		// n_importance == 0: the reference leaves Outputs empty (:143,168); kept
		if (return_raw) result.Raw = raw;
	}
	if (!return_weights) {                                                                                 // :174-179
		result.Outputs.WeightsLE = Tensor();
		result.Outputs.LangEmbedding = Tensor();
		result.Outputs.RenderedLangEmbedding = Tensor();
	}
	return result;
}

LeRFRenderResult LeRFRenderer::BatchifyRays(Tensor rays_flat, Tensor cone_angle, const int n_samples, const int chunk, const bool return_raw,
	const bool lin_disp, const float perturb, const int n_importance, const bool white_bkgr, const float raw_noise_std,
	const float stochastic_preconditioning_alpha, Tensor bounding_box, const bool return_weights)
{
	// src/LeRFRenderer.cpp:185-263
	const int64_t n = rays_flat.size(0);
	std::vector<Tensor> disp, acc, weights, depth, emb, rendered, relevancy, raw;
	for (int64_t i = 0; i < n; i += chunk) {
		LeRFRenderResult part = RenderRays(rays_flat.slice(0, i, std::min<int64_t>(i + chunk, n)), cone_angle, n_samples, return_raw, lin_disp, perturb,
			n_importance, white_bkgr, raw_noise_std, stochastic_preconditioning_alpha, bounding_box, return_weights);
		auto keep = [](std::vector<Tensor>& v, const Tensor& t) { if (t.defined()) v.push_back(t); };
		keep(disp, part.Outputs.DispMapLE); keep(acc, part.Outputs.AccMapLE); keep(weights, part.Outputs.WeightsLE); keep(depth, part.Outputs.DepthMapLE);
		keep(emb, part.Outputs.LangEmbedding); keep(rendered, part.Outputs.RenderedLangEmbedding); keep(relevancy, part.Outputs.Relevancy);
		keep(raw, part.Raw);
	}
	auto join = [](const std::vector<Tensor>& v) { return v.empty() ? Tensor() : (v.size() == 1 ? v[0] : torch::cat(v, 0)); };
	LeRFRenderResult result;
	result.Outputs.DispMapLE = join(disp); result.Outputs.AccMapLE = join(acc); result.Outputs.WeightsLE = join(weights);
	result.Outputs.DepthMapLE = join(depth); result.Outputs.LangEmbedding = join(emb); result.Outputs.RenderedLangEmbedding = join(rendered);
	result.Outputs.Relevancy = join(relevancy); result.Raw = join(raw);
	return result;
}

LeRFRenderResult LeRFRenderer::Render(const int h, const int w, Tensor k, const NeRFRenderParams& render_params,
	std::tuple<Tensor, Tensor, Tensor> rays, Tensor c2w, Tensor c2w_staticcam)
{
	// src/LeRFRenderer.cpp:266-331
	Tensor rays_o, rays_d, cone_angle;
	if (c2w.defined() && c2w.numel() != 0) std::tie(rays_o, rays_d, cone_angle) = GetRays(h, w, k, c2w);
	else std::tie(rays_o, rays_d, cone_angle) = rays;
	const std::vector<int64_t> sh = rays_d.sizes().vec();
	if (render_params.Ndc)
		std::tie(rays_o, rays_d, cone_angle) = NDCRays(h, w, k[0][0].item<float>(), 1.f, rays_o, rays_d, render_params.ThinRay ? Tensor() : cone_angle);
	// AABB near / far and the [o d near far] rows in one kernel (:293-304)
	Tensor rays_ = nrfhost::RaysPrepare(rays_o, rays_d, render_params.BoundingBox, 0.f, false);
	LeRFRenderResult all_ret = BatchifyRays(rays_, render_params.ThinRay ? Tensor() : cone_angle, render_params.NSamples, render_params.Chunk,
		render_params.ReturnRaw, render_params.LinDisp, render_params.Perturb, render_params.NImportance, render_params.WhiteBkgr,
		render_params.RawNoiseStd, render_params.StochasticPreconditioningAlpha, render_params.BoundingBox, render_params.ReturnWeights);
	auto has = [](const Tensor& t) { return t.defined() && t.numel() != 0; };
	const int dim = Lerf->GetLangEmbedDim();
	if (sh.size() > 2) {
		if (has(all_ret.Outputs.DispMapLE)) all_ret.Outputs.DispMapLE = all_ret.Outputs.DispMapLE.reshape({sh[0], sh[1]});
		if (has(all_ret.Outputs.DepthMapLE)) all_ret.Outputs.DepthMapLE = all_ret.Outputs.DepthMapLE.reshape({sh[0], sh[1]});
		if (has(all_ret.Outputs.RenderedLangEmbedding)) all_ret.Outputs.RenderedLangEmbedding = all_ret.Outputs.RenderedLangEmbedding.reshape({sh[0], sh[1], dim});
		if (has(all_ret.Outputs.Relevancy)) all_ret.Outputs.Relevancy = all_ret.Outputs.Relevancy.reshape({sh[0], sh[1], 2});
	} else {
		if (has(all_ret.Outputs.RenderedLangEmbedding)) all_ret.Outputs.RenderedLangEmbedding = all_ret.Outputs.RenderedLangEmbedding.reshape({sh[0], dim});
		if (has(all_ret.Outputs.Relevancy)) all_ret.Outputs.Relevancy = all_ret.Outputs.Relevancy.reshape({sh[0], 2});
	}
	// one host read for both scalars (the reference issues two .item() syncs, :329-330)
	Tensor nf = torch::stack({rays_.select(1, 6).min(), rays_.select(1, 7).max()}).cpu();
	all_ret.Near = nf[0].item<float>();
	all_ret.Far = nf[1].item<float>();
	return all_ret;
}
