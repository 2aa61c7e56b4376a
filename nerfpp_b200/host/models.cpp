// Host side of the drop-in models (see models.h).
#include "models.h"
#include "embedders.h"

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;
namespace nn = torch::nn;

// ------------------------------------------------------------------------------------------------ NeRF (classic)
NeRFImpl::NeRFImpl(const int d, const int w, const int input_ch, const int input_ch_views, const int output_ch,
	const std::set<int>& skips, const bool use_viewdirs, const std::string module_name)
	: BaseNeRFImpl(module_name), D(d), W(w), InputCh(input_ch), InputChViews(input_ch_views), OutputCh(output_ch), Skips(skips),
	  UseViewDirs(use_viewdirs)
{
	// layer i+1 takes [x | h] when layer i is a skip layer (src/NeRF.cpp:52-57)
	for (int i = 0; i < d; i++) {
		const int in = i == 0 ? input_ch : (skips.count(i - 1) ? w + input_ch : w);
		PtsLinears->push_back(nn::Linear(in, w));
	}
	if (use_viewdirs) {
		ViewsLinears->push_back(nn::Linear(input_ch_views + w, w / 2));
		FeatureLinear = nn::Linear(w, w);
		AlphaLinear = nn::Linear(w, 1);
		RGBLinear = nn::Linear(w / 2, 3);
	} else {
		OutputLinear = nn::Linear(w + input_ch, output_ch);
	}
	for (size_t i = 0; i < PtsLinears->size(); i++) register_module(module_name + "_pts_linears_" + std::to_string(i), PtsLinears[i]);
	if (use_viewdirs) {
		for (size_t i = 0; i < ViewsLinears->size(); i++) register_module(module_name + "_views_linears_" + std::to_string(i), ViewsLinears[i]);
		register_module(module_name + "_feature_linear", FeatureLinear);
		register_module(module_name + "_alpha_linear", AlphaLinear);
		register_module(module_name + "_rgb_linear", RGBLinear);
	} else {
		register_module(module_name + "_output_linear", OutputLinear);
	}
}

bool NeRFImpl::FusedShape() const
{
	const nrf_mlp_nerf_shape s{D, W, InputCh, InputChViews, Skips.size() == 1 ? *Skips.begin() : -1, UseViewDirs ? 1 : 0};
	return nrf_mlp_nerf_packed_bytes(&s) > 0;   // the library answers -1 for shapes it was not built for
}

std::vector<Tensor> NeRFImpl::FusedParams()
{
	std::vector<Tensor> params;
	for (size_t i = 0; i < PtsLinears->size(); i++) {
		params.push_back(PtsLinears[i]->as<nn::Linear>()->weight);
		params.push_back(PtsLinears[i]->as<nn::Linear>()->bias);
	}
	nn::LinearImpl* tail[4] = {FeatureLinear.get(), AlphaLinear.get(), ViewsLinears[0]->as<nn::Linear>(), RGBLinear.get()};
	for (nn::LinearImpl* l : tail) {
		params.push_back(l->weight);
		params.push_back(l->bias);
	}
	return params;
}

namespace {
// 24 tensors in FusedParams() order -> the C ABI's pointer structs (weights and gradients share the field layout)
template <typename S, typename P>
S NerfPointers(const std::vector<Tensor>& t)
{
	S w{};
	for (int i = 0; i < 8; i++) { w.pts_w[i] = (P)t[2 * i].data_ptr<float>(); w.pts_b[i] = (P)t[2 * i + 1].data_ptr<float>(); }
	w.feature_w = (P)t[16].data_ptr<float>(); w.feature_b = (P)t[17].data_ptr<float>();
	w.alpha_w = (P)t[18].data_ptr<float>();   w.alpha_b = (P)t[19].data_ptr<float>();
	w.views_w = (P)t[20].data_ptr<float>();   w.views_b = (P)t[21].data_ptr<float>();
	w.rgb_w = (P)t[22].data_ptr<float>();     w.rgb_b = (P)t[23].data_ptr<float>();
	return w;
}

Tensor NerfPack(const nrf_mlp_nerf_shape& s, const std::vector<Tensor>& params, bool train)
{
	std::vector<Tensor> dense;
	for (const Tensor& t : params) dense.push_back(nrfhost::Dense(t.detach(), torch::kFloat32, "NeRF parameter"));
	const nrf_mlp_nerf_weights w = NerfPointers<nrf_mlp_nerf_weights, const float*>(dense);
	Tensor blob = torch::empty({nrf_mlp_nerf_packed_bytes(&s)}, torch::TensorOptions().dtype(torch::kUInt8).device(dense[0].device()));
	nrfhost::Check((train ? nrf_mlp_nerf_pack_train : nrf_mlp_nerf_pack)(&s, &w, blob.data_ptr(), nrfhost::Stream()), "nrf_mlp_nerf_pack");
	return blob;
}

// NeRFImpl::forward with the backward LibTorch autograd would derive (src/NeRF.cpp:92-126), as three tcgen05 kernels:
// nrf_mlp_nerf_fwd_train (stores every layer's input, bf16) and nrf_mlp_nerf_bwd (gradient chain + weight gradients).
struct NerfTrainFunction : public torch::autograd::Function<NerfTrainFunction> {
	static variable_list forward(AutogradContext* ctx, at::TensorList in, int64_t samples_per_ray, std::vector<double> freqs_pts,
		std::vector<double> freqs_views)
	{
		// in = 24 parameters (FusedParams order) + x [N, 90], or + points [N,3], dirs [R,3] when samples_per_ray > 0;
		// passed as ONE TensorList so that every entry is an autograd input
		const nrf_mlp_nerf_shape s{8, 256, 63, 27, 4, 1};
		std::vector<Tensor> params(in.begin(), in.begin() + 24);
		const Tensor flat = nrfhost::Dense(in[24].detach(), torch::kFloat32, "NeRF input");
		const int64_t n = flat.size(0);
		Tensor blob = NerfPack(s, params, true);
		Tensor out = torch::empty({n, 4}, nrfhost::F32Like(flat));
		Tensor saved = torch::empty({std::max<int64_t>(nrf_mlp_nerf_saved_bytes(&s, n), 0)}, blob.options());
		if (samples_per_ray > 0) {
			const Tensor dirs = nrfhost::Dense(in[25].detach(), torch::kFloat32, "NeRF view directions");
			const std::vector<float> fp(freqs_pts.begin(), freqs_pts.end()), fv(freqs_views.begin(), freqs_views.end());
			nrfhost::Check(nrf_mlp_nerf_fwd_train_points(&s, blob.data_ptr(), flat.data_ptr<float>(), dirs.data_ptr<float>(), int32_t(samples_per_ray),
				fp.data(), int32_t(fp.size()), fv.data(), int32_t(fv.size()), n, out.data_ptr<float>(), saved.data_ptr(), nrfhost::Stream()),
				"nrf_mlp_nerf_fwd_train_points");
		} else
		nrfhost::Check(nrf_mlp_nerf_fwd_train(&s, blob.data_ptr(), flat.data_ptr<float>(), n, out.data_ptr<float>(), saved.data_ptr(), nrfhost::Stream()),
			"nrf_mlp_nerf_fwd_train");
		ctx->save_for_backward({blob, saved});
		std::vector<std::vector<int64_t>> shapes;
		for (const Tensor& t : params) shapes.push_back(t.sizes().vec());
		ctx->saved_data["n"] = n;
		ctx->saved_data["shapes"] = shapes;
		ctx->saved_data["extra_inputs"] = int64_t(in.size()) - 24;
		return {out};
	}
	static variable_list backward(AutogradContext* ctx, variable_list grad_out)
	{
		const nrf_mlp_nerf_shape s{8, 256, 63, 27, 4, 1};
		const auto saved = ctx->get_saved_variables();
		const int64_t n = ctx->saved_data["n"].toInt();
		const auto shapes = ctx->saved_data["shapes"].toListRef();
		const Tensor g = nrfhost::Dense(grad_out[0], torch::kFloat32, "NeRF output gradient");
		std::vector<Tensor> grads;
		for (const auto& sh : shapes) grads.push_back(torch::zeros(sh.toIntVector(), nrfhost::F32Like(g)));
		if (n > 0) {
			Tensor ws = torch::empty({nrf_mlp_nerf_bwd_workspace_bytes(&s, n)}, saved[0].options());
			const nrf_mlp_nerf_grads gp = NerfPointers<nrf_mlp_nerf_grads, float*>(grads);
			nrfhost::Check(nrf_mlp_nerf_bwd(&s, saved[0].data_ptr(), saved[1].data_ptr(), g.data_ptr<float>(), n, ws.data_ptr(), &gp, nrfhost::Stream()),
				"nrf_mlp_nerf_bwd");
		}
		variable_list ret(grads.begin(), grads.end());
		// x (or points, dirs): the positional embedder has no parameters (src/NeRF.cpp:4-39), no gradient is produced; then the 3 plain arguments
		for (int64_t i = 0; i < ctx->saved_data["extra_inputs"].toInt() + 3; i++) ret.push_back(Tensor());
		return ret;
	}
};
}  // namespace

Tensor NeRFImpl::ForwardFused(const Tensor& x)
{
	torch::NoGradGuard no_grad;
	const nrf_mlp_nerf_shape s{D, W, InputCh, InputChViews, *Skips.begin(), 1};
	// re-pack only when a parameter changed (data pointer or version counter)
	const std::vector<Tensor> params = FusedParams();
	std::vector<std::pair<const void*, uint32_t>> key;
	for (const Tensor& t : params) key.emplace_back(t.data_ptr(), t._version());
	if (!PackedBlob.defined() || key != PackedKey) {
		PackedBlob = NerfPack(s, params, false);
		PackedKey = key;
	}
	std::vector<int64_t> shape = x.sizes().vec();
	Tensor flat = nrfhost::Dense(x, torch::kFloat32, "NeRF input").reshape({-1, int64_t(InputCh + InputChViews)});
	Tensor out = torch::empty({flat.size(0), 4}, nrfhost::F32Like(flat));
	nrfhost::Check(nrf_mlp_nerf_fwd(&s, PackedBlob.data_ptr(), flat.data_ptr<float>(), flat.size(0), out.data_ptr<float>(), nrfhost::Stream()), "nrf_mlp_nerf_fwd");
	shape.back() = 4;
	return out.view(shape);
}

Tensor NeRFImpl::ForwardFusedTrain(const Tensor& x)
{
	variable_list in = FusedParams();
	std::vector<int64_t> shape = x.sizes().vec();
	in.push_back(x.reshape({-1, int64_t(InputCh + InputChViews)}));
	shape.back() = 4;
	return NerfTrainFunction::apply(at::TensorList(in), int64_t(0), std::vector<double>{}, std::vector<double>{})[0].view(shape);
}

bool NeRFImpl::FusedEmbeddingShape(EmbedderImpl& e_pts, EmbedderImpl& e_dirs) const
{
	return FusedEmbedding && FusedShape() && e_pts.GetIncludeInput() && e_dirs.GetIncludeInput() && e_pts.GetInputDims() == 3 && e_dirs.GetInputDims() == 3 &&
	       e_pts.GetFreqBands().size() == 10 && e_dirs.GetFreqBands().size() == 4 && e_pts.GetOutputDims() == InputCh && e_dirs.GetOutputDims() == InputChViews;
}

Tensor NeRFImpl::ForwardPoints(const Tensor& points, const Tensor& view_dirs, int samples_per_ray, const std::vector<float>& freqs_pts,
	const std::vector<float>& freqs_views)
{
	const Tensor pts = nrfhost::Dense(points.detach(), torch::kFloat32, "NeRF sample positions"), dirs = nrfhost::Dense(view_dirs.detach(), torch::kFloat32, "NeRF view directions");
	TORCH_CHECK(pts.dim() == 2 && pts.size(1) == 3 && dirs.dim() == 2 && dirs.size(1) == 3 && pts.size(0) == dirs.size(0) * samples_per_ray,
		"NeRF::ForwardPoints: points [R*S,3], view_dirs [R,3]");
	TORCH_CHECK(!torch::GradMode::is_enabled() || FusedTraining,
		"NeRF::ForwardPoints: autograd is recording but FusedTraining is off - embed and call forward() instead (RunNetwork does)");
	if (torch::GradMode::is_enabled()) {
		variable_list in = FusedParams();
		in.push_back(pts);
		in.push_back(dirs);
		return NerfTrainFunction::apply(at::TensorList(in), int64_t(samples_per_ray), std::vector<double>(freqs_pts.begin(), freqs_pts.end()),
			std::vector<double>(freqs_views.begin(), freqs_views.end()))[0];
	}
	torch::NoGradGuard no_grad;
	const nrf_mlp_nerf_shape s{D, W, InputCh, InputChViews, *Skips.begin(), 1};
	const std::vector<Tensor> params = FusedParams();
	std::vector<std::pair<const void*, uint32_t>> key;
	for (const Tensor& t : params) key.emplace_back(t.data_ptr(), t._version());
	if (!PackedBlob.defined() || key != PackedKey) {
		PackedBlob = NerfPack(s, params, false);
		PackedKey = key;
	}
	Tensor out = torch::empty({pts.size(0), 4}, nrfhost::F32Like(pts));
	nrfhost::Check(nrf_mlp_nerf_fwd_points(&s, PackedBlob.data_ptr(), pts.data_ptr<float>(), dirs.data_ptr<float>(), samples_per_ray, freqs_pts.data(),
		int32_t(freqs_pts.size()), freqs_views.data(), int32_t(freqs_views.size()), pts.size(0), out.data_ptr<float>(), nrfhost::Stream()), "nrf_mlp_nerf_fwd_points");
	return out;
}

Tensor NeRFImpl::forward(Tensor x)
{
	const bool built = x.size(-1) == InputCh + InputChViews && FusedShape();
	TORCH_CHECK(!built || x.is_cuda(), "NeRF: the input must be a CUDA tensor (the sm_100a path has no CPU fallback)");
	const bool fusable = built;
	if (fusable && !torch::GradMode::is_enabled()) return ForwardFused(x);
	// training: the fused backward yields parameter gradients only — an input that itself requires a gradient keeps the ATen path
	if (fusable && FusedTraining && !x.requires_grad()) return ForwardFusedTrain(x);
	Tensor pts = x.narrow(-1, 0, InputCh), views = x.narrow(-1, InputCh, InputChViews);
	Tensor h = pts;
	for (size_t i = 0; i < PtsLinears->size(); i++) {
		h = torch::relu(PtsLinears[i]->as<nn::Linear>()->forward(h));
		if (Skips.count(int(i))) h = torch::cat({pts, h}, -1);             // src/NeRF.cpp:103-104
	}
	if (!UseViewDirs) return OutputLinear(torch::cat({h, pts}, -1));        // :122-123
	Tensor alpha = AlphaLinear(h);                                          // :110
	h = torch::cat({FeatureLinear(h), views}, -1);                          // :111-112
	for (size_t i = 0; i < ViewsLinears->size(); i++) h = torch::relu(ViewsLinears[i]->as<nn::Linear>()->forward(h));
	return torch::cat({RGBLinear(h), alpha}, -1);                           // :119-120
}

// ------------------------------------------------------------------------------------------------ NeRFSmall
NeRFSmallImpl::NeRFSmallImpl(const int num_layers, const int hidden_dim, const int geo_feat_dim, const int num_layers_color,
	const int hidden_dim_color, const bool use_pred_normal, const int num_layers_normals, const int hidden_dim_normals,
	const int input_ch, const int input_ch_views, const std::string module_name)
	: BaseNeRFImpl(module_name), InputCh(input_ch), InputChViews(input_ch_views), NumLayers(num_layers), HiddenDim(hidden_dim),
	  GeoFeatDim(geo_feat_dim), NumLayersColor(num_layers_color), HiddenDimColor(hidden_dim_color),
	  NumLayersNormals(num_layers_normals), HiddenDimNormals(hidden_dim_normals), UsePredNormal(use_pred_normal)
{
	auto chain = [](nn::ModuleList& list, int n, int first_in, int hidden, int last_out) {
		for (int l = 0; l < n; l++)
			list->push_back(nn::Linear(nn::LinearOptions(l == 0 ? first_in : hidden, l == n - 1 ? last_out : hidden).bias(false)));
	};
	chain(SigmaNet, NumLayers, InputCh, HiddenDim, 1 + GeoFeatDim);                       // src/NeRF.cpp:338-339
	chain(ColorNet, NumLayersColor, InputChViews + GeoFeatDim, HiddenDimColor, 3);        // :341-342
	if (UsePredNormal) chain(NormalsNet, NumLayersNormals, 1 + GeoFeatDim + InputCh, HiddenDimNormals, 3);   // :344-348
	for (size_t i = 0; i < SigmaNet->size(); i++) register_module(module_name + "_sigma_net_" + std::to_string(i), SigmaNet[i]);
	for (size_t i = 0; i < ColorNet->size(); i++) register_module(module_name + "_color_net_" + std::to_string(i), ColorNet[i]);
	for (size_t i = 0; i < NormalsNet->size(); i++) register_module(module_name + "_normals_net_" + std::to_string(i), NormalsNet[i]);
}

nrf_mlp_small_shape NeRFSmallImpl::Shape() const
{
	nrf_mlp_small_shape s{};
	s.input_ch = InputCh; s.input_ch_views = InputChViews; s.hidden_dim = HiddenDim; s.geo_feat_dim = GeoFeatDim;
	s.hidden_dim_color = HiddenDimColor; s.num_layers = NumLayers; s.num_layers_color = NumLayersColor;
	return s;
}

bool NeRFSmallImpl::Fused() const
{
	if (UsePredNormal) return false;
	const nrf_mlp_small_shape s = Shape();
	// the library answers -1 for shapes it was not built for; other view widths exist there as NRF_MLP_IN_ENC16_RAYBIAS (a per-ray view term,
	// nerfpp_b200.pipeline), which this per-point module interface ([N, input_ch + input_ch_views] rows) cannot express: ATen formulation
	return s.input_ch_views == 16 && nrf_mlp_small_param_count(&s) > 0;
}

bool NeRFSmallImpl::FusedPerRay() const
{
	if (UsePredNormal) return false;
	const nrf_mlp_small_shape s = Shape();
	return nrf_mlp_small_param_count(&s) > 0;
}

std::vector<Tensor> NeRFSmallImpl::Weights()
{
	std::vector<Tensor> w;
	for (size_t i = 0; i < SigmaNet->size(); i++) w.push_back(SigmaNet[i]->as<nn::Linear>()->weight);
	for (size_t i = 0; i < ColorNet->size(); i++) w.push_back(ColorNet[i]->as<nn::Linear>()->weight);
	return w;
}

Tensor NeRFSmallImpl::Packed()
{
	std::vector<Tensor> w = Weights();
	std::vector<std::pair<const void*, uint32_t>> key;
	for (const Tensor& t : w) key.emplace_back(t.data_ptr(), t._version());
	if (!PackedBlob.defined() || key != PackedKey) {
		torch::NoGradGuard no_grad;
		std::vector<Tensor> flat;
		for (const Tensor& t : w) flat.push_back(nrfhost::Dense(t, torch::kFloat32, "NeRFSmall weight").reshape({-1}));
		FlatParams = torch::cat(flat).contiguous();
		const nrf_mlp_small_shape s = Shape();
		// a NEW blob per re-pack: a blob saved for a pending backward stays valid while the optimizer moves on
		PackedBlob = torch::empty({nrf_mlp_small_packed_bytes(&s)}, torch::TensorOptions().dtype(torch::kUInt8).device(FlatParams.device()));
		nrfhost::Check(nrf_mlp_small_pack(&s, FlatParams.data_ptr<float>(), PackedBlob.data_ptr(), nrfhost::Stream()), "nrf_mlp_small_pack");
		PackedKey = key;
	}
	return PackedBlob;
}

namespace {

variable_list SplitFlatGrad(const Tensor& flat, const std::vector<Tensor>& weights)
{
	variable_list out;
	int64_t off = 0;
	for (const Tensor& w : weights) {
		out.push_back(flat.narrow(0, off, w.numel()).view(w.sizes()));
		off += w.numel();
	}
	return out;
}

// NeRFSmall::forward on a dense [N, in+views] fp32 input
struct MlpSmallFn : public torch::autograd::Function<MlpSmallFn> {
	static Tensor forward(AutogradContext* ctx, Tensor x, Tensor w0, Tensor w1, Tensor w2, Tensor w3, Tensor w4, int64_t module)
	{
		auto* model = reinterpret_cast<NeRFSmallImpl*>(module);
		Tensor in = nrfhost::Dense(x, torch::kFloat32, "NeRFSmall input");
		Tensor packed = model->Packed();
		const nrf_mlp_small_shape s = model->Shape();
		Tensor out = torch::empty({in.size(0), 4}, nrfhost::F32Like(in));
		nrfhost::Check(nrf_mlp_small_fwd(&s, packed.data_ptr(), NRF_MLP_IN_F32_CAT, in.data_ptr(), nullptr, 1, nullptr, in.size(0),
			nrfhost::Ptr<float>(out), nrfhost::Stream()), "nrf_mlp_small_fwd");
		ctx->save_for_backward({in, packed});
		ctx->saved_data["module"] = module;
		ctx->saved_data["need_dx"] = x.requires_grad();
		return out;
	}

	static variable_list backward(AutogradContext* ctx, variable_list grad_out)
	{
		auto* model = reinterpret_cast<NeRFSmallImpl*>(ctx->saved_data["module"].toInt());
		auto saved = ctx->get_saved_variables();
		Tensor in = saved[0], packed = saved[1];
		const nrf_mlp_small_shape s = model->Shape();
		Tensor g = nrfhost::Dense(grad_out[0], torch::kFloat32, "grad of NeRFSmall output");
		Tensor flat = torch::zeros({nrf_mlp_small_param_count(&s)}, nrfhost::F32Like(in));
		Tensor gx = ctx->saved_data["need_dx"].toBool() ? torch::empty_like(in) : Tensor();
		nrfhost::Check(nrf_mlp_small_bwd(&s, packed.data_ptr(), NRF_MLP_IN_F32_CAT, in.data_ptr(), nullptr, 1, nullptr, in.size(0),
			nrfhost::CPtr<float>(g), gx.defined() ? gx.data_ptr() : nullptr, flat.data_ptr<float>(), nrfhost::Stream()), "nrf_mlp_small_bwd");
		variable_list grads = {gx};
		for (Tensor& t : SplitFlatGrad(flat, model->Weights())) grads.push_back(t);
		grads.push_back(Tensor());
		return grads;
	}
};

// hash encode (fp16 rows) -> fused MLP with per-ray SH and keep mask: what RunNetwork computes (src/NeRFRenderer.h:164-194)
struct HashNeRFNetworkFn : public torch::autograd::Function<HashNeRFNetworkFn> {
	static Tensor forward(AutogradContext* ctx, Tensor embeddings, Tensor w0, Tensor w1, Tensor w2, Tensor w3, Tensor w4,
		Tensor points, Tensor ray_sh, int64_t samples_per_ray, int64_t hash_module, int64_t mlp_module)
	{
		auto* hash = reinterpret_cast<CuHashEmbedderImpl*>(hash_module);
		auto* model = reinterpret_cast<NeRFSmallImpl*>(mlp_module);
		Tensor pts = nrfhost::Dense(points, torch::kFloat32, "points");
		Tensor sh = nrfhost::Dense(ray_sh, torch::kFloat32, "per-ray SH");
		auto [enc, keep] = hash->EncodeF16(pts);
		Tensor packed = model->Packed();
		const nrf_mlp_small_shape s = model->Shape();
		Tensor raw = torch::empty({pts.size(0), 4}, nrfhost::F32Like(pts));
		nrfhost::Check(nrf_mlp_small_fwd(&s, packed.data_ptr(), NRF_MLP_IN_ENC16_RAYDIRS, enc.data_ptr(), nrfhost::CPtr<float>(sh),
			int32_t(samples_per_ray), reinterpret_cast<const uint8_t*>(keep.data_ptr()), pts.size(0), nrfhost::Ptr<float>(raw),
			nrfhost::Stream()), "nrf_mlp_small_fwd");
		ctx->save_for_backward({pts, enc, keep, sh, packed});
		ctx->saved_data["spr"] = samples_per_ray;
		ctx->saved_data["hash"] = hash_module;
		ctx->saved_data["mlp"] = mlp_module;
		return raw;
	}

	static variable_list backward(AutogradContext* ctx, variable_list grad_out)
	{
		auto* hash = reinterpret_cast<CuHashEmbedderImpl*>(ctx->saved_data["hash"].toInt());
		auto* model = reinterpret_cast<NeRFSmallImpl*>(ctx->saved_data["mlp"].toInt());
		auto saved = ctx->get_saved_variables();
		Tensor pts = saved[0], enc = saved[1], keep = saved[2], sh = saved[3], packed = saved[4];
		const nrf_mlp_small_shape s = model->Shape();
		Tensor g = nrfhost::Dense(grad_out[0], torch::kFloat32, "grad of raw").reshape({-1, 4});
		Tensor flat = torch::zeros({nrf_mlp_small_param_count(&s)}, nrfhost::F32Like(pts));
		Tensor g_enc = torch::empty({pts.size(0), enc.size(1)}, torch::TensorOptions().dtype(torch::kBFloat16).device(pts.device()));
		nrfhost::Check(nrf_mlp_small_bwd(&s, packed.data_ptr(), NRF_MLP_IN_ENC16_RAYDIRS, enc.data_ptr(), nrfhost::CPtr<float>(sh),
			int32_t(ctx->saved_data["spr"].toInt()), reinterpret_cast<const uint8_t*>(keep.data_ptr()), pts.size(0),
			nrfhost::CPtr<float>(g), g_enc.data_ptr(), flat.data_ptr<float>(), nrfhost::Stream()), "nrf_mlp_small_bwd");
		Tensor grad_table = torch::zeros_like(hash->Embeddings);
		hash->Backward(pts, g_enc, grad_table);
		variable_list grads = {grad_table};
		for (Tensor& t : SplitFlatGrad(flat, model->Weights())) grads.push_back(t);
		for (int i = 0; i < 5; i++) grads.push_back(Tensor());
		return grads;
	}
};

}  // namespace

Tensor nrfhost::MlpSmallF32(NeRFSmallImpl& model, const Tensor& x)
{
	std::vector<Tensor> w = model.Weights();
	return MlpSmallFn::apply(x, w[0], w[1], w[2], w[3], w[4], reinterpret_cast<int64_t>(&model));
}

Tensor nrfhost::HashNeRFNetwork(CuHashEmbedderImpl& hash, NeRFSmallImpl& model, const Tensor& points, const Tensor& ray_sh,
	int samples_per_ray)
{
	std::vector<Tensor> w = model.Weights();
	return HashNeRFNetworkFn::apply(hash.Embeddings, w[0], w[1], w[2], w[3], w[4], points, ray_sh, int64_t(samples_per_ray),
		reinterpret_cast<int64_t>(&hash), reinterpret_cast<int64_t>(&model));
}

Tensor NeRFSmallImpl::forward(Tensor x)
{
	TORCH_CHECK(x.dim() == 2 && x.size(1) == InputCh + InputChViews, "NeRFSmall: input must be [N,", InputCh + InputChViews, "]");
	if (Fused()) return nrfhost::MlpSmallF32(*this, x);
	// other shapes (normals head, different widths): same maths through torch::linear (src/NeRF.cpp:363-412)
	Tensor pts = x.narrow(-1, 0, InputCh), views = x.narrow(-1, InputCh, InputChViews);
	auto run = [](nn::ModuleList& net, Tensor h) {
		for (size_t i = 0; i < net->size(); i++) {
			h = net[i]->as<nn::Linear>()->forward(h);
			if (i + 1 != net->size()) h = torch::relu(h);                  // final activations live in RawToOutputs
		}
		return h;
	};
	Tensor h = run(SigmaNet, pts);
	Tensor sigma = h.narrow(-1, 0, 1), geo = h.narrow(-1, 1, h.size(-1) - 1);
	Tensor color = run(ColorNet, torch::cat({views, geo}, -1));             // views first (:383)
	if (!UsePredNormal) return torch::cat({color, sigma}, -1);              // :408
	Tensor normals = run(NormalsNet, torch::cat({sigma, geo, pts}, -1));    // :394
	return torch::cat({color, sigma, normals}, -1);
}
