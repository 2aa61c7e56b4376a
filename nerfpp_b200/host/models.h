// Drop-in models: Trainable, BaseNeRF, NeRFSmall (fused sm_100a kernel), NeRF (classic 8x256).
//
// Constructor signatures, registered sub-module names ("<name>_sigma_net_<i>", "<name>_color_net_<i>",
// "<name>_pts_linears_<i>", ... — src/NeRF.cpp:76-89,350-360) and forward() contracts follow the reference
// (src/NeRF.h:33-77,213-252; src/LibTorchTraining/Trainable.h:6-54), so checkpoints are interchangeable and
// NeRFExecutor<> / NeRFRenderer<> instantiate unchanged.
#pragma once
#include <set>

#include "nrf_torch.h"

class CuHashEmbedderImpl;   // embedders.h

struct Trainable : public torch::nn::Module {
	explicit Trainable(const std::string module_name) : torch::nn::Module(module_name) {}
	~Trainable() override = default;
	virtual torch::Tensor forward(torch::Tensor x) = 0;

	/// number of scalar parameters
	template <typename T>
	static int ParamsCount(T& module)
	{
		int64_t n = 0;
		for (const auto& p : module->parameters()) n += p.numel();
		return int(n);
	}

	/// Xavier-normal(gain 0.1) on "*.weight", 1 on norm weights, 0 on "*.bias" (src/LibTorchTraining/Trainable.h:32-53)
	template <typename T>
	static void Initialize(T& module)
	{
		torch::NoGradGuard no_grad;
		for (auto& item : module->named_parameters()) {
			const std::string& key = item.key();
			const bool weight = key.find(".weight") != std::string::npos;
			if (weight && key.find("norm") != std::string::npos) torch::nn::init::constant_(item.value(), 1.);
			else if (weight) torch::nn::init::xavier_normal_(item.value(), 0.1);
			if (key.find(".bias") != std::string::npos) torch::nn::init::constant_(item.value(), 0.);
		}
	}
};

class BaseNeRFImpl : public Trainable {
public:
	explicit BaseNeRFImpl(const std::string& module_name) : Trainable(module_name) {}
	~BaseNeRFImpl() override = default;
	torch::Tensor forward(torch::Tensor x) override { return torch::Tensor(); }
};
TORCH_MODULE(BaseNeRF);

/// Classic NeRF MLP (src/NeRF.h:44-77, src/NeRF.cpp:41-126).  Inference (no autograd graph, CUDA input) at the BASELINE shape
/// D=8, W=256, 63+27 inputs, skips={4}, view branch runs as ONE fused tcgen05 kernel (nrf_mlp_nerf_fwd: activations stay in
/// tensor memory, weights stream through a TMA ring).  Training at that shape (autograd recording, input without a gradient of
/// its own — the positional embedder has no parameters) runs nrf_mlp_nerf_fwd_train / nrf_mlp_nerf_bwd behind one
/// torch::autograd::Function: bf16 tcgen05 forward that stores every layer's input, gradient chain and weight gradients on the
/// tensor cores.  Other shapes run the same maths through torch::linear (cuBLAS) with LibTorch autograd.
struct NeRFImpl : public BaseNeRFImpl {
	int D, W, InputCh, InputChViews, OutputCh;
	std::set<int> Skips;
	bool UseViewDirs;
	torch::nn::ModuleList PtsLinears, ViewsLinears;
	torch::nn::Linear FeatureLinear = nullptr, AlphaLinear = nullptr, RGBLinear = nullptr, OutputLinear = nullptr;

	NeRFImpl(const int d = 8, const int w = 256, const int input_ch = 3, const int input_ch_views = 3, const int output_ch = 4,
		const std::set<int>& skips = std::set<int>{4}, const bool use_viewdirs = false, const std::string module_name = "nerf");
	~NeRFImpl() override = default;
	torch::Tensor forward(torch::Tensor x) override;

	// ---- B200 additions
	bool FusedShape() const;                   ///< true when nrf_mlp_nerf_fwd covers this configuration
	torch::Tensor ForwardFused(const torch::Tensor& x);   ///< [.., 90] -> [.., 4] on the fused kernel (no autograd)
	torch::Tensor ForwardFusedTrain(const torch::Tensor& x);   ///< the same with the fused backward attached (parameter gradients)
	std::vector<torch::Tensor> FusedParams();  ///< pts_linears w,b x8, feature w,b, alpha w,b, views w,b, rgb w,b
	bool FusedTraining = true;                 ///< false: training goes through torch::linear + LibTorch autograd (fp32 cuBLAS)
	bool FusedEmbedding = true;                ///< false: RunNetwork embeds and concatenates with separate kernels (the A/B baseline)
	/// RunNetwork with the positional embeddings evaluated inside the MLP kernel (nrf_mlp_nerf_fwd[_train]_points): points [N,3],
	/// view_dirs [R,3], N = R * samples_per_ray -> [N,4].  Differentiable w.r.t. the parameters when autograd is recording.
	torch::Tensor ForwardPoints(const torch::Tensor& points, const torch::Tensor& view_dirs, int samples_per_ray,
		const std::vector<float>& freqs_pts, const std::vector<float>& freqs_views);
	/// true when ForwardPoints covers this model with these two embedders
	bool FusedEmbeddingShape(class EmbedderImpl& e_pts, class EmbedderImpl& e_dirs) const;
	/// The operand blobs are re-packed when a parameter's (data_ptr, Tensor::_version()) changes; after a write that bypasses the version
	/// counter (param.data().copy_(), a raw kernel, an NCCL broadcast into the storage) call this.
	void InvalidateCaches() { PackedKey.clear(); PackedBlob = torch::Tensor(); }
private:
	torch::Tensor PackedBlob;
	std::vector<std::pair<const void*, uint32_t>> PackedKey;
};
TORCH_MODULE(NeRF);

/// HashNeRF MLP (src/NeRF.h:213-252, src/NeRF.cpp:322-412): bias-free sigma net in->H..->1+geo, colour net
/// [views|geo]->H..->3, optional normals net.  At the BASELINE shape (32+16 -> 64 -> 16 | 31 -> 64 -> 64 -> 3, no normals)
/// forward/backward are ONE fused tensor-core kernel each (nrf_mlp_small_fwd/_bwd); other shapes run the same maths
/// through torch::linear.
class NeRFSmallImpl : public BaseNeRFImpl {
protected:
	int InputCh, InputChViews, NumLayers, HiddenDim, GeoFeatDim, NumLayersColor, HiddenDimColor, NumLayersNormals, HiddenDimNormals;
	bool UsePredNormal;
	torch::nn::ModuleList SigmaNet, ColorNet, NormalsNet;
public:
	NeRFSmallImpl(const int num_layers = 3, const int hidden_dim = 64, const int geo_feat_dim = 15, const int num_layers_color = 4,
		const int hidden_dim_color = 64, const bool use_pred_normal = true, const int num_layers_normals = 3,
		const int hidden_dim_normals = 64, const int input_ch = 3, const int input_ch_views = 3,
		const std::string module_name = "hashnerf");
	~NeRFSmallImpl() override = default;
	torch::Tensor forward(torch::Tensor x) override;

	// ---- B200 additions
	bool Fused() const;                        ///< true when the fused kernels cover this shape through the per-point interface (16 view channels)
	bool FusedPerRay() const;                  ///< true when they cover it with the view channels per RAY (any SH degree 1..8: nrf_render_raybatch_fwd)
	nrf_mlp_small_shape Shape() const;
	std::vector<torch::Tensor> Weights();      ///< sigma_net_0.., color_net_0.. weights, each [out,in]
	/// tensor-core operand blob of the current weights (re-packed only when a weight changed)
	torch::Tensor Packed();
	int GetInputCh() const { return InputCh; }
	int GetInputChViews() const { return InputChViews; }
	/// see NeRFImpl::InvalidateCaches
	void InvalidateCaches() { PackedKey.clear(); PackedBlob = torch::Tensor(); }

private:
	torch::Tensor PackedBlob, FlatParams;
	std::vector<std::pair<const void*, uint32_t>> PackedKey;
};
TORCH_MODULE(NeRFSmall);

namespace nrfhost {
/// Fused NeRFSmall on already-encoded inputs — the drop-in NeRFSmall::forward path: x [N, in+views] fp32 -> [N,4].
torch::Tensor MlpSmallF32(NeRFSmallImpl& model, const torch::Tensor& x);
/// Fused RunNetwork tail: enc fp16 [N,32], per-ray SH [R,16] (row n uses ray n / samples_per_ray), keep [N] -> raw [N,4].
/// Differentiable w.r.t. the weights and (through `hash`) the hash table; used by NeRFRenderer<CuHashEmbedder,CuSHEncoder,NeRFSmall>.
torch::Tensor HashNeRFNetwork(::CuHashEmbedderImpl& hash, NeRFSmallImpl& model, const torch::Tensor& points,
	const torch::Tensor& ray_sh, int samples_per_ray);
}  // namespace nrfhost
