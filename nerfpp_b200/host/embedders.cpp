// Host side of the drop-in embedders (see embedders.h).  Every kernel is reached through the C ABI.
#include "embedders.h"

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

// ------------------------------------------------------------------------------------------------ CuHashEmbedder
CuHashEmbedderImpl::CuHashEmbedderImpl(const std::string& module_name, Tensor bounding_box, const int n_levels,
	const int n_features_per_level, const int log2_hashmap_size, const int base_resolution, const int finest_resolution)
	: BaseEmbedderImpl(module_name), BoundingBox(bounding_box), NLevels(n_levels), NFeaturesPerLevel(n_features_per_level),
	  Log2HashmapSize(log2_hashmap_size), BaseResolution(base_resolution), FinestResolution(finest_resolution),
	  OutputDims(n_levels * n_features_per_level)
{
	TORCH_CHECK(n_levels >= 1 && n_levels <= NRF_MAX_LEVELS, "CuHashEmbedder: n_levels out of range");
	TORCH_CHECK(n_features_per_level == 2 || n_features_per_level == 4 || n_features_per_level == 8,
		"CuHashEmbedder: the sm_100a kernels are built for 2, 4 or 8 features per level");
	const auto cuda_f32 = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA);
	const auto cuda_i32 = torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA);
	const int64_t entries_per_level = int64_t(1) << log2_hashmap_size;

	// The RNG calls below are made in the reference's order and with its arguments (src/CuHashEmbedder.cpp:24,45), so a
	// process seeded like src/main.cpp:174 draws the same table and the same primes from either implementation.
	Embeddings = register_parameter(module_name + "_embeddings",
		torch::rand({entries_per_level * NLevels, NFeaturesPerLevel}, cuda_f32) * 1e-4f, /*requires_grad=*/true);

	std::vector<int> primes;
	primes.reserve(size_t(3) * NLevels * NVolumes);
	auto is_prime = [](int v) {
		for (int d = 2; int64_t(d) * d <= v; d++)
			if (v % d == 0) return false;
		return true;
	};
	while (primes.size() < size_t(3) * NLevels * NVolumes) {
		const int v = torch::randint(1 << 28, 1 << 30, {1}, torch::TensorOptions().dtype(torch::kInt32).device(torch::kCPU)).item<int>();
		if (is_prime(v)) primes.push_back(v);
	}
	Primes = torch::tensor(primes, torch::kInt32).reshape({NLevels, NVolumes, 3}).to(torch::kCUDA).contiguous();
	Biases = RandBias ? (torch::rand({int64_t(NLevels) * NVolumes, 3}, cuda_f32) * 1000.f + 100.f).contiguous()
	                  : torch::zeros({int64_t(NLevels) * NVolumes, 3}, cuda_f32);
	// entries per level rounded down to a multiple of 16; offsets are the exclusive running sum (src/CuHashEmbedder.cpp:62-68)
	const int local_size = int((entries_per_level >> 4) << 4);
	FeatLocalSize = torch::full({NLevels}, local_size, cuda_i32);
	FeatLocalIdx = (torch::arange(NLevels, cuda_i32) * local_size).to(torch::kInt32).contiguous();

	Primes = register_buffer(module_name + "_primes", Primes);
	Biases = register_buffer(module_name + "_biases", Biases);
	FeatLocalSize = register_buffer(module_name + "_feat_local_size", FeatLocalSize);
	FeatLocalIdx = register_buffer(module_name + "_feat_local_idx", FeatLocalIdx);

	LevelScale = torch::empty({NLevels}, cuda_f32);
	nrfhost::Check(nrf_hash_level_scales(BaseResolution, FinestResolution, NLevels, LevelScale.data_ptr<float>(), nrfhost::Stream()),
		"nrf_hash_level_scales");
}

nrf_hash_grid CuHashEmbedderImpl::Grid()
{
	if (!BoxCached) {   // BoundingBox is set once at construction in the reference (src/NeRFExecutor.h:432)
		Box = nrfhost::HostBox(BoundingBox);
		BoxCached = true;
	}
	nrf_hash_grid g{};
	g.n_levels = NLevels;
	g.n_features = NFeaturesPerLevel;
	g.n_volumes = NVolumes;
	g.base_resolution = BaseResolution;
	g.finest_resolution = FinestResolution;
	for (int k = 0; k < 3; k++) { g.box_min[k] = Box[k]; g.box_max[k] = Box[3 + k]; }
	TORCH_CHECK(Primes.is_cuda() && Primes.is_contiguous() && Primes.scalar_type() == torch::kInt32, "CuHashEmbedder: Primes must be contiguous int32 CUDA");
	TORCH_CHECK(FeatLocalIdx.is_cuda() && FeatLocalSize.is_cuda() && Biases.is_cuda(), "CuHashEmbedder: buffers must live on the GPU");
	g.primes = Primes.data_ptr<int32_t>();
	g.biases = Biases.data_ptr<float>();
	g.feat_local_idx = FeatLocalIdx.data_ptr<int32_t>();
	g.feat_local_size = FeatLocalSize.data_ptr<int32_t>();
	g.level_scale = LevelScale.data_ptr<float>();
	g.table_scalars = Embeddings.numel();
	return g;
}

Tensor CuHashEmbedderImpl::ShadowF16()
{
	TORCH_CHECK(Embeddings.is_cuda() && Embeddings.is_contiguous() && Embeddings.scalar_type() == torch::kFloat32,
		"CuHashEmbedder: Embeddings must be a contiguous fp32 CUDA tensor");
	const void* src = Embeddings.data_ptr();
	const uint32_t version = Embeddings._version();   // bumped by every in-place update (optimizer step, copy_, load)
	if (!Shadow.defined() || Shadow.numel() != Embeddings.numel() || src != ShadowSource || version != ShadowVersion) {
		if (!Shadow.defined() || Shadow.numel() != Embeddings.numel())
			Shadow = torch::empty(Embeddings.sizes(), torch::TensorOptions().dtype(torch::kFloat16).device(Embeddings.device()));
		nrfhost::Check(nrf_table_to_half(Embeddings.data_ptr<float>(), Shadow.data_ptr(), Embeddings.numel(), nrfhost::Stream()), "nrf_table_to_half");
		ShadowSource = src;
		ShadowVersion = version;
	}
	return Shadow;
}

std::pair<Tensor, Tensor> CuHashEmbedderImpl::EncodeF16(const Tensor& x)
{
	Tensor pts = nrfhost::Dense(x, torch::kFloat32, "points");
	TORCH_CHECK(pts.dim() == 2 && pts.size(1) == 3, "CuHashEmbedder: points must be [N,3]");
	const int64_t n = pts.size(0);
	Tensor shadow = ShadowF16();
	Tensor enc = torch::empty({n, OutputDims}, torch::TensorOptions().dtype(torch::kFloat16).device(pts.device()));
	Tensor keep = torch::empty({n}, torch::TensorOptions().dtype(torch::kBool).device(pts.device()));
	nrf_hash_grid g = Grid();
	nrfhost::Check(nrf_hash_encode_fwd(&g, shadow.data_ptr(), nrfhost::CPtr<float>(pts), n, /*clamp=*/1,
		reinterpret_cast<uint8_t*>(keep.data_ptr()), enc.data_ptr(), NRF_ENC_F16, nrfhost::Stream()), "nrf_hash_encode_fwd");
	return {enc, keep};
}

void CuHashEmbedderImpl::Backward(const Tensor& points, const Tensor& grad_enc, Tensor& grad_table)
{
	TORCH_CHECK(grad_table.is_cuda() && grad_table.is_contiguous() && grad_table.scalar_type() == torch::kFloat32 &&
		grad_table.numel() == Embeddings.numel(), "CuHashEmbedder: grad_table must match Embeddings");
	const bool bf16 = grad_enc.scalar_type() == torch::kBFloat16;
	Tensor g = bf16 ? grad_enc.contiguous() : nrfhost::Dense(grad_enc, torch::kFloat32, "grad_enc");
	nrf_hash_grid grid = Grid();
	nrfhost::Check(nrf_hash_encode_bwd(&grid, nrfhost::CPtr<float>(points), points.size(0), /*clamp=*/1, g.data_ptr(),
		bf16 ? NRF_GRAD_BF16 : NRF_GRAD_F32, grad_table.data_ptr<float>(), nrfhost::Stream()), "nrf_hash_encode_bwd");
}

namespace {

// Replaces torch::autograd::CuHashEmbedderFunction (src/CuHashEmbedder.cu:221-325).  The query points are saved in the
// autograd context of THIS call (the reference keeps them in the module, so a second forward before backward corrupts
// the gradient).
struct HashEncodeFn : public torch::autograd::Function<HashEncodeFn> {
	static variable_list forward(AutogradContext* ctx, Tensor embeddings, Tensor points, int64_t module)
	{
		auto* self = reinterpret_cast<CuHashEmbedderImpl*>(module);
		const int64_t n = points.size(0);
		Tensor shadow = self->ShadowF16();
		Tensor enc = torch::empty({n, self->OutputDims}, nrfhost::F32Like(points));
		Tensor keep = torch::empty({n}, torch::TensorOptions().dtype(torch::kBool).device(points.device()));
		nrf_hash_grid g = self->Grid();
		nrfhost::Check(nrf_hash_encode_fwd(&g, shadow.data_ptr(), nrfhost::CPtr<float>(points), n, /*clamp=*/1,
			reinterpret_cast<uint8_t*>(keep.data_ptr()), enc.data_ptr(), NRF_ENC_F32, nrfhost::Stream()), "nrf_hash_encode_fwd");
		ctx->save_for_backward({points});
		ctx->saved_data["module"] = module;
		ctx->mark_non_differentiable({keep});
		return {enc, keep};
	}

	static variable_list backward(AutogradContext* ctx, variable_list grad_out)
	{
		auto* self = reinterpret_cast<CuHashEmbedderImpl*>(ctx->saved_data["module"].toInt());
		Tensor points = ctx->get_saved_variables()[0];
		Tensor grad_table = torch::zeros_like(self->Embeddings);
		if (grad_out[0].defined()) self->Backward(points, grad_out[0], grad_table);
		return {grad_table, Tensor(), Tensor()};   // no gradient to the sample positions (src/CuHashEmbedder.cu:323)
	}
};

}  // namespace

std::pair<Tensor, Tensor> CuHashEmbedderImpl::forward(Tensor x)
{
	Tensor pts = nrfhost::Dense(x.detach(), torch::kFloat32, "points");
	TORCH_CHECK(pts.dim() == 2 && pts.size(1) == 3, "CuHashEmbedder: input must be [N,3]");
	QueryPoints = pts;   // kept for source compatibility with code that inspects it; the backward does not read it
	variable_list out = HashEncodeFn::apply(Embeddings, pts, reinterpret_cast<int64_t>(this));
	return {out[0], out[1]};
}

Tensor TotalVariationLoss(CuHashEmbedder embedder)
{
	// src/CuHashEmbedder.h:75-95: (FinestResolution/100)^3 random points, squared forward differences one finest cell apart
	const std::array<float, 6> box = nrfhost::HostBox(embedder->BoundingBox);
	const int per_axis = embedder->GetFinestResolution() / 100;
	const int64_t n = std::max<int64_t>(1, int64_t(per_axis) * per_axis * per_axis);
	const auto opt = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA);
	Tensor lo = torch::tensor({box[0], box[1], box[2]}, opt), hi = torch::tensor({box[3], box[4], box[5]}, opt);
	Tensor samples = torch::rand({n, 3}, opt) * (hi - lo) + lo;
	Tensor base = embedder->forward(samples).first;
	Tensor total = torch::zeros({}, opt);
	for (int axis = 0; axis < 3; axis++) {
		Tensor step = torch::zeros({3}, opt);
		step[axis] = (box[3 + axis] - box[axis]) / float(embedder->GetFinestResolution());
		total = total + torch::pow(base - embedder->forward(samples + step).first, 2).sum();
	}
	return total;
}

// ------------------------------------------------------------------------------------------------ CuSHEncoder
std::pair<Tensor, Tensor> CuSHEncoderImpl::forward(Tensor input)
{
	TORCH_CHECK(Degree >= 1 && Degree <= 8, "CuSHEncoder: degree must be 1..8");
	Tensor dirs = nrfhost::Dense(input, torch::kFloat32, "directions");
	TORCH_CHECK(dirs.dim() == 2 && dirs.size(1) == 3, "CuSHEncoder: input must be [N,3]");
	Tensor out = torch::empty({dirs.size(0), OutputDims}, nrfhost::F32Like(dirs));
	nrfhost::Check(nrf_sh_encode_fwd(nrfhost::CPtr<float>(dirs), 3, dirs.size(0), Degree, nrfhost::Ptr<float>(out), nrfhost::Stream()),
		"nrf_sh_encode_fwd");
	return {out, Tensor()};
}

// ------------------------------------------------------------------------------------------------ Embedder
EmbedderImpl::EmbedderImpl(const std::string& module_name, int num_freqs, float max_freq_log2, bool include_input, int input_dims,
	bool log_sampling)
	: BaseEmbedderImpl(module_name), NumFreqs(num_freqs), MaxFreq(max_freq_log2), IncludeInput(include_input), InputDims(input_dims),
	  LogSampling(log_sampling)
{
	// band k of N: 2^(k*max/(N-1)) when log-sampled, else evenly spaced between 2^0 and 2^max — same scalar fp32
	// expressions as src/NeRF.cpp:11-19 so the bands are the reference's floats
	FreqBands.resize(NumFreqs);
	for (int k = 0; k < NumFreqs; k++) {
		if (LogSampling) FreqBands[k] = powf(2.f, MaxFreq / (NumFreqs - 1) * k);
		else FreqBands[k] = 1.f + (pow(2.f, MaxFreq) - 1.f) / (NumFreqs - 1) * k;
	}
	OutputDims = InputDims * ((IncludeInput ? 1 : 0) + 2 * NumFreqs);
}

std::pair<Tensor, Tensor> EmbedderImpl::forward(Tensor x)
{
	Tensor in = nrfhost::Dense(x, torch::kFloat32, "Embedder input");
	TORCH_CHECK(in.dim() == 2 && in.size(1) == InputDims, "Embedder: input must be [N,", InputDims, "]");
	Tensor out = torch::empty({in.size(0), OutputDims}, nrfhost::F32Like(in));
	nrfhost::Check(nrf_posenc_fwd(nrfhost::CPtr<float>(in), in.size(0), InputDims, NumFreqs, FreqBands.data(), IncludeInput ? 1 : 0,
		nrfhost::Ptr<float>(out), nrfhost::Stream()), "nrf_posenc_fwd");
	return {out, Tensor()};
}
