// Drop-in LeRF language head and LeRFRenderer (BASELINE C5).
//
// Same class names, constructor signatures, registered sub-module names ("<name>_sigma_le_net_<i>", "<name>_le_net_<i>",
// src/LeRF.cpp:17-25), result structs and public / protected virtuals as the reference (src/LeRF.h:6-32, src/LeRFRenderer.h:9-133), so
// NeRFExecutor<..., CuHashEmbedder, LeRF, LeRFRenderer> instantiates unchanged and lang_model / lang_embedder checkpoints are
// interchangeable.
//
// What runs where:
//   * inference (no autograd graph) at the built shape LeRF(32, 2, 256, 512, 128): LeRF::forward is ONE tcgen05 kernel (nrf_lerf_fwd);
//     LeRFRenderer::RenderRays runs hash encode -> density-only head (coarse) -> compositing -> SamplePDF+merge -> hash encode ->
//     fused head without the [N,512] embedding -> compositing -> per-ray projection (nrf_lerf_sigma_fwd / _hidden_fwd /
//     _render_embedding), every step a C-ABI call;
//   * training (autograd recording) at the built shape, parity configuration (thin rays, no jitter / noise / preconditioning):
//     LeRFRenderer::RenderRays runs the coarse pass as inference and the fine pass as ONE autograd node on the fused kernels
//     (nrf_lerf_fwd_train -> nrf_composite_fwd -> nrf_lerf_render_embedding_train; backward nrf_lerf_bwd_rays -> nrf_composite_bwd ->
//     nrf_lerf_bwd_rows -> nrf_hash_encode_bwd at F = 8): gradients arrive at lang_embedder's Embeddings and the four Linear weights;
//   * other shapes / the jittered configurations: the reference's formulation on the drop-in pieces — hash encode forward / backward on the
//     sm_100a kernels (CuHashEmbedder::forward), compositing as one differentiable op, the Linear layers through torch::linear;
//   * Relevancy (src/LeRFRenderer.cpp:79) belongs to RuCLIP, which the reference does not vendor: LeRFRenderer::RelevancyFn is the
//     hook a build that has RuCLIP assigns; unset, LeRFRendererOutputs::Relevancy stays undefined.
#pragma once
#include <functional>

#include "embedders.h"
#include "models.h"
#include "render_ops.h"
#include "renderer.h"

/// Language Embedded Radiance Field MLP (src/LeRF.h:6-32, src/LeRF.cpp:3-111)
class LeRFImpl : public BaseNeRFImpl {
protected:
	int GeoFeatDimLE, NumLayersLE, HiddenDimLE, LangEmbedDim, InputChLE;
	torch::nn::ModuleList SigmaLENet, LENet;
public:
	LeRFImpl(const int geo_feat_dim_le = 32, const int num_layers_le = 3, const int hidden_dim_le = 64, const int lang_embed_dim = 768,
		const int input_ch_le = 0, const std::string module_name = "lerf");
	~LeRFImpl() override = default;
	/// x [.., input_ch_le] -> [.., lang_embed_dim + 1] = [normalize(le), sigma_le]
	torch::Tensor forward(torch::Tensor x) override;
	virtual int GetLangEmbedDim() const { return LangEmbedDim; }

	// ---- B200 additions
	bool Fused() const;                        ///< true when the fused kernels cover this shape
	nrf_lerf_shape Shape() const;
	std::vector<torch::Tensor> Weights();      ///< sigma_le_net_0, sigma_le_net_1, le_net_0, le_net_1 weights, each [out,in]
	torch::Tensor Packed();                    ///< operand blob of the current weights (re-packed only when a weight changed)
	torch::Tensor ForwardAten(const torch::Tensor& x);   ///< the reference's formulation (torch::linear), differentiable
	/// after a write to a weight that bypasses Tensor::_version() (param.data().copy_(), raw kernel, NCCL broadcast): re-pack on next use
	void InvalidateCaches() { PackedKey.clear(); PackedBlob = torch::Tensor(); }

private:
	torch::Tensor PackedBlob;
	std::vector<std::pair<const void*, uint32_t>> PackedKey;
};
TORCH_MODULE(LeRF);

struct LeRFRendererOutputs {
	torch::Tensor LangEmbedding,    ///< [num_rays, num_samples, lang_embed_dim]
		RenderedLangEmbedding,      ///< [num_rays, lang_embed_dim]
		DispMapLE,                  ///< [num_rays]
		AccMapLE,                   ///< [num_rays]
		WeightsLE,                  ///< [num_rays, num_samples]
		DepthMapLE,                 ///< [num_rays]
		Relevancy;                  ///< [num_rays, 2]
};

struct LeRFRenderResult {
	LeRFRendererOutputs Outputs;
	torch::Tensor Raw;              ///< [num_rays, num_samples, lang_embed_dim + 1]
	float Near, Far;
};

/// normalize(sum_s w_s e_s) (src/LeRFRenderer.h:45-54); embeds [bs, S, D], weights [bs, S, 1]
inline torch::Tensor RenderCLIPEmbedding(const torch::Tensor embeds, const torch::Tensor weights, const bool normalize = true)
{
	auto output = torch::sum(weights * embeds, -2);
	return torch::nn::functional::normalize(output, torch::nn::functional::NormalizeFuncOptions().dim(-1).eps(1e-8));
}

class LeRFRenderer {
protected:
	CuHashEmbedder LangEmbedFn = nullptr;
	LeRF Lerf = nullptr;
	torch::Tensor LerfPositives, LerfNegatives;

	virtual torch::Tensor RunLENetwork(torch::Tensor inputs, LeRF lerf, CuHashEmbedder lang_embed_fn);
	virtual LeRFRendererOutputs RawToLEOutputs(torch::Tensor raw_le, torch::Tensor z_vals_le, torch::Tensor rays_d, const int lang_embed_dim = 768,
		const float raw_noise_std = 0.f);

public:
	LeRFRenderer(CuHashEmbedder lang_embed_fn, LeRF lerf, torch::Tensor lerf_positives = torch::Tensor(), torch::Tensor lerf_negatives = torch::Tensor())
		: LangEmbedFn(lang_embed_fn), Lerf(lerf), LerfPositives(lerf_positives), LerfNegatives(lerf_negatives) {}
	virtual ~LeRFRenderer() {}

	std::tuple<torch::Tensor, torch::Tensor> GetLeRFPrompts() { return std::make_tuple(LerfPositives, LerfNegatives); }
	void SetLeRFPrompts(const torch::Tensor lerf_positives, const torch::Tensor lerf_negatives) { LerfPositives = lerf_positives; LerfNegatives = lerf_negatives; }

	virtual LeRFRenderResult RenderRays(torch::Tensor ray_batch, torch::Tensor cone_angle, const int n_samples, const bool return_raw = false,
		const bool lin_disp = false, const float perturb = 0.f, const int n_importance = 0, const bool white_bkgr = false,
		const float raw_noise_std = 0.f, const float stochastic_preconditioning_alpha = 0.f, torch::Tensor bounding_box = torch::Tensor(),
		const bool return_weights = true);
	virtual LeRFRenderResult BatchifyRays(torch::Tensor rays_flat, torch::Tensor cone_angle, const int n_samples, const int chunk = 1024 * 32,
		const bool return_raw = false, const bool lin_disp = false, const float perturb = 0.f, const int n_importance = 0,
		const bool white_bkgr = false, const float raw_noise_std = 0., const float stochastic_preconditioning_alpha = 0.f,
		torch::Tensor bounding_box = torch::Tensor(), const bool return_weights = true);
	virtual LeRFRenderResult Render(const int h, const int w, torch::Tensor k, const NeRFRenderParams& render_params,
		std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> rays = {torch::Tensor(), torch::Tensor(), torch::Tensor()},
		torch::Tensor c2w = torch::Tensor(), torch::Tensor c2w_staticcam = torch::Tensor());

	// ---- B200 additions
	/// Relevancy(rendered, positives, negatives) of RuCLIP, when the build has it (src/LeRFRenderer.cpp:79)
	std::function<torch::Tensor(torch::Tensor, torch::Tensor, torch::Tensor)> RelevancyFn;
	/// The fused inference path never forms LangEmbedding [R,S,D] (2 KB per sample that no caller of the reference reads:
	/// src/NeRFExecutor.h:642-650,706-720,957-983 use RenderedLangEmbedding / Relevancy only).  true: evaluate it as well (nrf_lerf_fwd).
	bool MaterializeLangEmbedding = false;
	/// false: training goes through the reference's formulation on torch::linear (A/B and tests); true (default): the fused backward
	bool UseFusedTraining = true;
	/// true when RenderRays can take the fused training path for these arguments
	bool FusedTraining(const torch::Tensor& ray_batch, const torch::Tensor& cone_angle, float perturb, int n_importance, float raw_noise_std,
		float stochastic_preconditioning_alpha);
	/// true when RenderRays can take the fused inference path for these arguments
	bool FusedInference(const torch::Tensor& ray_batch, const torch::Tensor& cone_angle, float perturb, int n_importance, float raw_noise_std,
		float stochastic_preconditioning_alpha);
};
