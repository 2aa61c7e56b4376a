// Drop-in embedders: BaseEmbedder, CuHashEmbedder, CuSHEncoder, Embedder.
//
// Same class names, constructor signatures, public members, registered parameter / buffer names and forward() contract
// as the reference (src/BaseEmbedder.h:6-15, src/CuHashEmbedder.h:8-58, src/CuSHEncoder.h:6-29, src/NeRF.h:12-30), so
// NeRFRenderer<> / NeRFExecutor<> instantiate on them unchanged and torch::save/load archives are interchangeable.
// All arithmetic happens behind the C ABI (include/nerfpp_b200.h).
#pragma once
#include "nrf_torch.h"

/// embedding + keep-mask (may be undefined) — src/BaseEmbedder.h:6-15
class BaseEmbedderImpl : public torch::nn::Module {
public:
	explicit BaseEmbedderImpl(const std::string& module_name) : torch::nn::Module(module_name) {}
	~BaseEmbedderImpl() override = default;
	virtual int GetOutputDims() { return 0; }
	virtual std::pair<torch::Tensor, torch::Tensor> forward(torch::Tensor x) { return {torch::Tensor(), torch::Tensor()}; }
};
TORCH_MODULE(BaseEmbedder);

/// Multiresolution hash grid — replaces src/CuHashEmbedder.{h,cpp,cu}
class CuHashEmbedderImpl : public BaseEmbedderImpl {
public:
	// public surface of the reference (src/CuHashEmbedder.h:12-27)
	torch::Tensor BoundingBox;
	bool RandBias{false};
	int NLevels, NFeaturesPerLevel, Log2HashmapSize, BaseResolution, FinestResolution, OutputDims, NVolumes{1};
	torch::Tensor Embeddings, Primes, Biases, FeatLocalSize, FeatLocalIdx, QueryPoints, QueryVolumeIdx;

	int GetNLevels() const { return NLevels; }
	int GetNFeaturesPerLevel() const { return NFeaturesPerLevel; }
	int GetLog2HashmapSize() const { return Log2HashmapSize; }
	int GetBaseResolution() const { return BaseResolution; }
	int GetFinestResolution() const { return FinestResolution; }
	torch::Tensor GetBoundingBox() const { return BoundingBox; }

	CuHashEmbedderImpl(const std::string& module_name, torch::Tensor bounding_box, const int n_levels = 16,
		const int n_features_per_level = 2, const int log2_hashmap_size = 19, const int base_resolution = 16,
		const int finest_resolution = 512);
	~CuHashEmbedderImpl() override = default;

	void Initialize() {}
	int GetOutputDims() override { return OutputDims; }
	std::pair<torch::Tensor, torch::Tensor> forward(torch::Tensor x) override;

	// ---- B200 additions (not part of the reference surface)
	/// fp16 copy of Embeddings the gather kernels read; re-derived only when Embeddings changed (tensor version counter),
	/// instead of the reference's full-table cast on every forward (src/CuHashEmbedder.cu:257).
	torch::Tensor ShadowF16();
	/// descriptor for the C ABI; buffers are taken as they are (they pin the hash function)
	nrf_hash_grid Grid();
	/// fp16 [N, L*F] encodings + keep mask, no autograd: the fused RunNetwork path (renderer.h)
	std::pair<torch::Tensor, torch::Tensor> EncodeF16(const torch::Tensor& x);
	/// accumulates dL/dEmbeddings for bf16 / fp32 [N, L*F] encoding gradients into grad_table (fp32, same shape as Embeddings)
	void Backward(const torch::Tensor& points, const torch::Tensor& grad_enc, torch::Tensor& grad_table);
	/// The fp16 shadow is rebuilt when Embeddings' (data_ptr, Tensor::_version()) changes.  A write that bypasses the version counter —
	/// Embeddings.data().copy_(), variable_data(), a raw kernel or an NCCL broadcast into the storage — must be followed by this call.
	void InvalidateCaches() { ShadowSource = nullptr; }

private:
	torch::Tensor LevelScale, Shadow;
	std::array<float, 6> Box{};
	bool BoxCached{false};
	uint32_t ShadowVersion{0};
	const void* ShadowSource{nullptr};
};
TORCH_MODULE(CuHashEmbedder);

/// the reference's helper (src/CuHashEmbedder.h:75-95); safe here because points are saved per call
torch::Tensor TotalVariationLoss(CuHashEmbedder embedder);

/// Spherical harmonics, degree 1..8 — replaces src/CuSHEncoder.{h,cpp,cu}
class CuSHEncoderImpl : public BaseEmbedderImpl {
protected:
	int InputDim, Degree, OutputDims;
public:
	CuSHEncoderImpl(const std::string& module_name, const int input_dim = 3, const int degree = 4)
		: BaseEmbedderImpl(module_name), InputDim(input_dim), Degree(degree), OutputDims(degree * degree) {}
	~CuSHEncoderImpl() override = default;
	int GetOutputDims() override { return OutputDims; }
	int GetDegree() const { return Degree; }
	std::pair<torch::Tensor, torch::Tensor> forward(torch::Tensor input) override;
};
TORCH_MODULE(CuSHEncoder);

/// Positional encoding [x, sin(f0 x), cos(f0 x), ...] — replaces EmbedderImpl (src/NeRF.h:12-30, src/NeRF.cpp:4-39)
class EmbedderImpl : public BaseEmbedderImpl {
protected:
	int NumFreqs;
	float MaxFreq;
	bool IncludeInput;
	int InputDims, OutputDims = 0;
	bool LogSampling;
	std::vector<float> FreqBands;
public:
	EmbedderImpl(const std::string& module_name, int multires) : EmbedderImpl(module_name, multires, float(multires - 1)) {}
	EmbedderImpl(const std::string& module_name, int num_freqs, float max_freq_log2, bool include_input = true, int input_dims = 3,
		bool log_sampling = true);
	~EmbedderImpl() override = default;
	int GetOutputDims() override { return OutputDims; }
	std::pair<torch::Tensor, torch::Tensor> forward(torch::Tensor x) override;
	// ---- B200 additions: what a kernel that evaluates the embedding itself needs to know
	const std::vector<float>& GetFreqBands() const { return FreqBands; }
	bool GetIncludeInput() const { return IncludeInput; }
	int GetInputDims() const { return InputDims; }
};
TORCH_MODULE(Embedder);
