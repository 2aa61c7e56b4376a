// TEST INFRASTRUCTURE ONLY — never imported by the product path (nerfpp_b200/).
//
// pybind11 front-end over the UNMODIFIED reference sources, compiled in place from /root/reference/src by
// oracle/Makefile into oracle/_ref/nerfpp_ref*.so.  It lets tests/ pin the numpy/torch restatement
// (oracle/restate.py) against the reference's own code, generates the golden fixtures under tests/golden/
// (tests/golden/make_golden.py), and is the `--impl reference` / cpu_baseline arm of bench.py.
//
// Nothing here restates reference code: every function forwards to a reference symbol, cited beside it.
#include <torch/extension.h>
#include <chrono>

#include "NeRF.h"          // Embedder, NeRF, NeRFSmall, HashEmbedder, SHEncoder  (src/NeRF.h)
#include "CustomOps.h"     // TruncExp                                             (src/CustomOps.h)
#include "Sampler.h"       // SamplePDF                                            (src/Sampler.h:6)
#include "RayUtils.h"      // GetRays, IntersectWithAABB                           (src/RayUtils.h:23,87)
#include "NeRFRenderer.h"  // NeRFRenderer<>                                       (src/NeRFRenderer.h:88)
#include "LeRF.h"          // LeRF                                                 (src/LeRF.h:6)
#include "LeRFRenderer.h"  // RenderCLIPEmbedding, LeRFRenderer                    (src/LeRFRenderer.h:45,58)
#ifdef NRF_REF_CUDA
#include "CuHashEmbedder.h"  // src/CuHashEmbedder.h
#include "CuSHEncoder.h"     // src/CuSHEncoder.h
#endif

namespace py = pybind11;
using torch::Tensor;

// Exposes the protected virtual stages of the reference renderer (src/NeRFRenderer.h:96,99).
template <class E, class D, class N>
struct OpenRenderer : public NeRFRenderer<E, D, N> {
	using Base = NeRFRenderer<E, D, N>;
	using Base::Base;
	using Base::RunNetwork;
	using Base::RawToOutputs;
	using Base::NeRF;
	using Base::EmbedFn;
	using Base::EmbeddirsFn;
};

static py::dict OutputsToDict(const NeRFRendererOutputs& o)
{
	py::dict d;
	d["rgb"] = o.RGBMap; d["disp"] = o.DispMap; d["acc"] = o.AccMap; d["weights"] = o.Weights; d["depth"] = o.DepthMap;
	return d;
}

static NeRFRenderParams MakeParams(int n_samples, int n_importance, int chunk, bool white_bkgr, bool use_viewdirs,
	Tensor bbox, bool return_raw, bool lin_disp)
{
	NeRFRenderParams p;
	p.NSamples = n_samples; p.NImportance = n_importance; p.Chunk = chunk; p.ReturnRaw = return_raw;
	p.LinDisp = lin_disp; p.Perturb = 0.f; p.WhiteBkgr = white_bkgr; p.RawNoiseStd = 0.f; p.Ndc = false;
	p.UseViewdirs = use_viewdirs; p.ReturnWeights = true; p.ThinRay = true;   // parity config, SURVEY §9-Q4
	p.BoundingBox = bbox; p.StochasticPreconditioningAlpha = 0.f;
	return p;
}

// One wrapper for every <embedder, dir-embedder, model> triple the reference instantiates on this path.
template <class E, class D, class N>
struct RefPipeline {
	E embed{nullptr};
	D embeddirs{nullptr};
	N model{nullptr};
	std::unique_ptr<OpenRenderer<E, D, N>> renderer;
	std::unique_ptr<torch::optim::Adam> opt;
	Tensor bbox;

	void Finish(torch::Device dev)
	{
		embed->to(dev); embeddirs->to(dev); model->to(dev);
		renderer = std::make_unique<OpenRenderer<E, D, N>>(embed, embeddirs, model);
	}
	std::vector<Tensor> EmbedParams() { return embed->parameters(); }
	std::vector<Tensor> ModelParams() { return model->parameters(); }
	std::vector<std::string> ModelParamNames()
	{
		std::vector<std::string> r;
		for (auto& p : model->named_parameters()) r.push_back(p.key());
		return r;
	}
	std::vector<std::string> EmbedBufferNames()
	{
		std::vector<std::string> r;
		for (auto& p : embed->named_buffers()) r.push_back(p.key());
		return r;
	}
	std::vector<Tensor> EmbedBuffers() { return embed->buffers(); }
	void InitModel() { Trainable::Initialize(model); }                    // src/LibTorchTraining/Trainable.h:32
	std::pair<Tensor, Tensor> Embed(Tensor x) { return embed->forward(x); }
	Tensor EmbedDirs(Tensor d) { return embeddirs->forward(d).first; }
	Tensor Model(Tensor x) { return model->forward(x); }
	Tensor RunNetwork(Tensor pts, Tensor viewdirs)                       // src/NeRFRenderer.h:164
	{
		return renderer->RunNetwork(pts, viewdirs, renderer->NeRF, renderer->EmbedFn, renderer->EmbeddirsFn);
	}
	py::dict RawToOutputs(Tensor raw, Tensor z, Tensor rays_d, float noise, bool white)   // src/NeRFRenderer.h:199
	{
		return OutputsToDict(renderer->RawToOutputs(raw, Tensor(), z, rays_d, noise, white));
	}
	// src/NeRFRenderer.h:366; ray_batch = [o(3) d(3) near far viewdirs(3)]
	py::dict RenderRays(Tensor ray_batch, int n_samples, int n_importance, bool white, bool return_raw)
	{
		auto r = renderer->RenderRays(ray_batch, Tensor(), n_samples, return_raw, false, 0.f, n_importance, white, 0.f, 0.f, bbox, true);
		py::dict d = OutputsToDict(r.Outputs);
		if (return_raw) d["raw"] = r.Raw;
		return d;
	}
	// src/NeRFRenderer.h:530 with a ray batch (training call, src/NeRFExecutor.h:876)
	py::dict Render(Tensor rays_o, Tensor rays_d, int n_samples, int n_importance, int chunk, bool white, bool use_viewdirs)
	{
		auto p = MakeParams(n_samples, n_importance, chunk, white, use_viewdirs, bbox, false, false);
		auto r = renderer->Render(0, 0, Tensor(), p, {rays_o, rays_d, Tensor()}, Tensor(), Tensor());
		py::dict d = OutputsToDict(r.Outputs);
		d["near"] = r.Near; d["far"] = r.Far;
		return d;
	}
	// src/NeRFRenderer.h:530 with a camera pose (full-image call, src/NeRFExecutor.h:684)
	py::dict RenderImage(int h, int w, Tensor k, Tensor c2w, int n_samples, int n_importance, int chunk, bool white, bool use_viewdirs)
	{
		torch::NoGradGuard ng;
		auto p = MakeParams(n_samples, n_importance, chunk, white, use_viewdirs, bbox, false, false);
		auto r = renderer->Render(h, w, k, p, {Tensor(), Tensor(), Tensor()}, c2w, Tensor());
		py::dict d = OutputsToDict(r.Outputs);
		d["near"] = r.Near; d["far"] = r.Far;
		return d;
	}
	// The training lines of NeRFExecutor::Train that touch this path (src/NeRFExecutor.h:539,868,876-890,923,986,
	// 992-996) driven verbatim: Adam(lr, betas (0.9,0.99), eps 1e-15), zero_grad, Render, huber, backward, step,
	// lr decay.  Returns (per-step seconds, per-step loss).
	std::pair<std::vector<double>, std::vector<float>> TrainSteps(Tensor rays_o, Tensor rays_d, Tensor target,
		int n_steps, int n_samples, int n_importance, int chunk, bool use_viewdirs, float lr, int lrate_decay)
	{
		if (!opt)
		{
			std::vector<Tensor> gv;
			for (auto& p : embed->parameters()) gv.push_back(p);
			for (auto& p : model->parameters()) gv.push_back(p);
			opt = std::make_unique<torch::optim::Adam>(gv, torch::optim::AdamOptions(lr).eps(1e-15).betas(std::make_tuple(0.9, 0.99)));
			global_step = 0;
		}
		std::vector<double> secs; std::vector<float> losses;
		auto p = MakeParams(n_samples, n_importance, chunk, false, use_viewdirs, bbox, false, false);
		for (int i = 0; i < n_steps; i++)
		{
			auto t0 = std::chrono::steady_clock::now();
			opt->zero_grad();
			auto r = renderer->Render(0, 0, Tensor(), p, {rays_o, rays_d, Tensor()}, Tensor(), Tensor());
			auto loss = torch::nn::functional::huber_loss(r.Outputs.RGBMap, target.detach());
			loss.backward();
			opt->step();
			float new_lr = lr * powf(0.1f, (float)global_step / (lrate_decay * 1000));
			for (auto& g : opt->param_groups()) g.options().set_lr(new_lr);
			global_step++;   // src/NeRFExecutor.h:1047 (after the lr update)
			float lv = loss.template item<float>();   // forces completion on CUDA as well
			secs.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
			losses.push_back(lv);
		}
		return {secs, losses};
	}
	int global_step = 0;
};

using HashCpuPipe = RefPipeline<HashEmbedder, SHEncoder, NeRFSmall>;
using ClassicPipe = RefPipeline<Embedder, Embedder, NeRF>;
#ifdef NRF_REF_CUDA
using CuHashPipe = RefPipeline<CuHashEmbedder, CuSHEncoder, NeRFSmall>;
#endif

template <class P>
static void BindPipe(py::module_& m, const char* name)
{
	py::class_<P>(m, name, py::module_local())
		.def("embed_params", &P::EmbedParams).def("model_params", &P::ModelParams)
		.def("model_param_names", &P::ModelParamNames)
		.def("embed_buffers", &P::EmbedBuffers).def("embed_buffer_names", &P::EmbedBufferNames)
		.def("init_model", &P::InitModel)
		.def("embed", &P::Embed).def("embed_dirs", &P::EmbedDirs).def("model", &P::Model)
		.def("run_network", &P::RunNetwork).def("raw_to_outputs", &P::RawToOutputs)
		.def("render_rays", &P::RenderRays).def("render", &P::Render).def("render_image", &P::RenderImage)
		.def("train_steps", &P::TrainSteps, py::call_guard<py::gil_scoped_release>())
		// NeRFExecutor::SaveCheckpoint / the restore branch of Initialize (src/NeRFExecutor.h:1054-1068, 546-553): same file names
		.def("save_checkpoint", [](P& p, const std::string& dir) {
			torch::save(p.embed, dir + "/embedder_checkpoint.pt");
			torch::save(p.model, dir + "/model_checkpoint.pt");
		})
		.def("load_checkpoint", [](P& p, const std::string& dir) {
			torch::load(p.embed, dir + "/embedder_checkpoint.pt");
			torch::load(p.model, dir + "/model_checkpoint.pt");
		});
}

// Exposes the protected virtual RawToLEOutputs of the reference LeRF renderer (src/LeRFRenderer.h:68).
struct OpenLeRFRenderer : public LeRFRenderer {
	using LeRFRenderer::LeRFRenderer;
	using LeRFRenderer::RawToLEOutputs;
	using LeRFRenderer::RunLENetwork;
};

#ifdef NRF_REF_CUDA
// The language branch of NeRFExecutor (src/NeRFExecutor.h:458-470, 504-523, 531-535, 957-983) on the reference's own classes:
// CuHashEmbedder("lang_embedder", ...) + LeRF(..., "lang_model") + LeRFRenderer.  Same call surface as LerfPipe in
// nerfpp_b200/host/bindings.cpp.  Relevancy comes from stub/RuCLIPProcessor.h (undefined): prompts are dummies.
static py::dict LeOutputsToDict(const LeRFRendererOutputs& o)
{
	py::dict d;
	d["rendered"] = o.RenderedLangEmbedding; d["embedding"] = o.LangEmbedding; d["weights"] = o.WeightsLE; d["depth"] = o.DepthMapLE;
	d["disp"] = o.DispMapLE; d["acc"] = o.AccMapLE; d["relevancy"] = o.Relevancy;
	return d;
}

struct LerfRefPipe {
	CuHashEmbedder embed{nullptr};
	LeRF model{nullptr};
	std::unique_ptr<OpenLeRFRenderer> renderer;
	std::unique_ptr<torch::optim::Adam> opt;
	Tensor bbox;

	void Finish()
	{
		embed->to(torch::kCUDA); model->to(torch::kCUDA);
		const int d = model->GetLangEmbedDim();
		renderer = std::make_unique<OpenLeRFRenderer>(embed, model, torch::zeros({1, d}), torch::zeros({1, d}));
	}
	std::vector<Tensor> EmbedParams() { return embed->parameters(); }
	std::vector<Tensor> ModelParams() { return model->parameters(); }
	std::vector<std::string> ModelParamNames() { std::vector<std::string> r; for (auto& p : model->named_parameters()) r.push_back(p.key()); return r; }
	void InitModel() { Trainable::Initialize(model); }
	Tensor Model(Tensor x) { return model->forward(x); }
	Tensor RunLENetwork(Tensor pts) { return renderer->RunLENetwork(pts, model, embed); }                         // src/LeRFRenderer.cpp:5
	py::dict RawToLEOutputs(Tensor raw_le, Tensor z, Tensor rays_d) { return LeOutputsToDict(renderer->RawToLEOutputs(raw_le, z, rays_d, model->GetLangEmbedDim(), 0.f)); }
	py::dict Render(Tensor rays_o, Tensor rays_d, int n_samples, int n_importance, int chunk, bool /*materialize*/, bool return_raw)   // src/LeRFRenderer.cpp:266
	{
		auto p = MakeParams(n_samples, n_importance, chunk, false, false, bbox, return_raw, false);
		auto r = renderer->Render(0, 0, Tensor(), p, {rays_o, rays_d, Tensor()}, Tensor(), Tensor());
		py::dict d = LeOutputsToDict(r.Outputs);
		d["near"] = r.Near; d["far"] = r.Far; d["raw"] = r.Raw;
		return d;
	}
	// src/NeRFExecutor.h:957-983, 986 (the language loss is back-propagated on its own)
	std::vector<float> TrainSteps(Tensor rays_o, Tensor rays_d, Tensor target, int n_steps, int n_samples, int n_importance, int chunk, float lr)
	{
		if (!opt) {
			std::vector<Tensor> gv;
			for (auto& p : embed->parameters()) gv.push_back(p);
			for (auto& p : model->parameters()) gv.push_back(p);
			opt = std::make_unique<torch::optim::Adam>(gv, torch::optim::AdamOptions(lr).eps(1e-15).betas(std::make_tuple(0.9, 0.99)));
		}
		std::vector<float> losses;
		auto p = MakeParams(n_samples, n_importance, chunk, false, false, bbox, false, false);
		for (int i = 0; i < n_steps; i++) {
			opt->zero_grad();
			auto r = renderer->Render(0, 0, Tensor(), p, {rays_o, rays_d, Tensor()}, Tensor(), Tensor());
			auto loss = torch::nn::functional::huber_loss(r.Outputs.RenderedLangEmbedding, target.detach(),
				torch::nn::functional::HuberLossFuncOptions().reduction(torch::kNone).delta(1.25)).sum(-1).nanmean();
			loss.backward();
			opt->step();
			losses.push_back(loss.item<float>());
		}
		return losses;
	}
};
#endif

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
	m.doc() = "reference (DeliriumV01D/NeRFpp) hot-path symbols, unmodified — test oracle only";
#ifdef NRF_REF_CUDA
	m.attr("has_cuda") = true;
#else
	m.attr("has_cuda") = false;
#endif
	m.def("manual_seed", [](int64_t s) { torch::manual_seed(s); });   // src/main.cpp:174
	m.def("set_num_threads", [](int n) { at::set_num_threads(n); });
	m.def("get_num_threads", []() { return at::get_num_threads(); });
	m.def("trunc_exp", [](Tensor x) { return torch::autograd::TruncExp::apply(x)[0]; });   // src/CustomOps.cpp:5
	m.def("sample_pdf", &SamplePDF);                                                          // src/Sampler.h:6
	m.def("intersect_aabb", [](Tensor o, Tensor d, Tensor bbox, float near_plane) { return IntersectWithAABB(o, d, bbox, near_plane); });   // src/RayUtils.h:87
	m.def("get_rays", [](int h, int w, Tensor k, Tensor c2w) { return GetRays(h, w, k, c2w); });   // src/RayUtils.h:23
	m.def("embedder", [](Tensor x, int multires) { Embedder e("embedder", multires); return e->forward(x).first; });   // src/NeRF.cpp:22
	m.def("sh_encoder", [](Tensor x, int degree) { SHEncoder e("embeddirs", 3, degree); return e->forward(x).first; }); // src/NeRF.cpp:131
	m.def("sort_merge", [](Tensor z, Tensor zs) { return std::get<0>(torch::sort(torch::cat({z, zs}, -1), -1)); });      // src/NeRFRenderer.h:431

	// src/LeRF.cpp:3-26 (constructor as src/NeRFExecutor.h:507-514), :28-111 (forward).  The weights are copied into the module's own
	// registered parameters ("<name>_sigma_le_net_<i>.weight", "<name>_le_net_<i>.weight"); the names are returned for the checkpoint test.
	m.def("lerf_forward", [](Tensor x, std::vector<Tensor> sigma_w, std::vector<Tensor> le_w, int geo_feat, int hidden, int lang_dim) {
		LeRF mod(geo_feat, static_cast<int>(sigma_w.size()), hidden, lang_dim, static_cast<int>(x.size(-1)), "lang_model");
		torch::NoGradGuard ng;
		auto np = mod->named_parameters();
		std::vector<std::string> names;
		for (size_t i = 0; i < sigma_w.size(); i++) np["lang_model_sigma_le_net_" + std::to_string(i) + ".weight"].copy_(sigma_w[i]);
		for (size_t i = 0; i < le_w.size(); i++) np["lang_model_le_net_" + std::to_string(i) + ".weight"].copy_(le_w[i]);
		for (auto& kv : np) names.push_back(kv.key());
		mod->to(x.device());
		return std::make_pair(mod->forward(x), names);
	});
	m.def("render_clip_embedding", [](Tensor e, Tensor w) { return RenderCLIPEmbedding(e, w); });   // src/LeRFRenderer.h:45-54
	// The language branch of one training iteration on already-encoded samples, through the reference's own modules and LibTorch autograd:
	// LeRF::forward (src/LeRF.cpp:28) -> RawToLEOutputs (src/LeRFRenderer.cpp:27) -> huber(delta 1.25).sum(-1).nanmean() (src/NeRFExecutor.h:964-968)
	// -> backward (:981).  x [R,S,C]; returns (loss, rendered, d loss / d {sigma_le_net_0, sigma_le_net_1, le_net_0, le_net_1, x}).
	m.def("lerf_language_grads", [](Tensor x, std::vector<Tensor> sigma_w, std::vector<Tensor> le_w, Tensor z, Tensor rays_d, Tensor target,
		int geo_feat, int hidden, int lang_dim) {
		LeRF mod(geo_feat, static_cast<int>(sigma_w.size()), hidden, lang_dim, static_cast<int>(x.size(-1)), "lang_model");
		mod->to(x.scalar_type());
		auto np = mod->named_parameters();
		{
			torch::NoGradGuard ng;
			for (size_t i = 0; i < sigma_w.size(); i++) np["lang_model_sigma_le_net_" + std::to_string(i) + ".weight"].copy_(sigma_w[i]);
			for (size_t i = 0; i < le_w.size(); i++) np["lang_model_le_net_" + std::to_string(i) + ".weight"].copy_(le_w[i]);
		}
		Tensor xin = x.detach().clone().set_requires_grad(true);
		Tensor raw = mod->forward(xin.reshape({-1, x.size(-1)})).reshape({x.size(0), x.size(1), lang_dim + 1});
		OpenLeRFRenderer r(nullptr, nullptr, torch::zeros({1, lang_dim}), torch::zeros({1, lang_dim}));
		auto o = r.RawToLEOutputs(raw, z, rays_d, lang_dim, 0.f);
		Tensor loss = torch::nn::functional::huber_loss(o.RenderedLangEmbedding, target.detach(),
			torch::nn::functional::HuberLossFuncOptions().reduction(torch::kNone).delta(1.25)).sum(-1).nanmean();
		loss.backward();
		std::vector<Tensor> grads;
		for (size_t i = 0; i < sigma_w.size(); i++) grads.push_back(np["lang_model_sigma_le_net_" + std::to_string(i) + ".weight"].grad());
		for (size_t i = 0; i < le_w.size(); i++) grads.push_back(np["lang_model_le_net_" + std::to_string(i) + ".weight"].grad());
		grads.push_back(xin.grad());
		return std::make_tuple(loss.detach(), o.RenderedLangEmbedding.detach(), grads);
	}, py::call_guard<py::gil_scoped_release>());
	// src/LeRFRenderer.cpp:27-82.  Relevancy (RuCLIP, absent) comes from the stub header and stays undefined.
	m.def("lerf_raw_to_outputs", [](Tensor raw_le, Tensor z, Tensor rays_d, int lang_dim) {
		OpenLeRFRenderer r(nullptr, nullptr, torch::zeros({1, lang_dim}), torch::zeros({1, lang_dim}));
		auto o = r.RawToLEOutputs(raw_le, z, rays_d, lang_dim, 0.f);
		py::dict d;
		d["rendered"] = o.RenderedLangEmbedding; d["weights"] = o.WeightsLE; d["depth"] = o.DepthMapLE; d["disp"] = o.DispMapLE; d["acc"] = o.AccMapLE;
		return d;
	});

	BindPipe<HashCpuPipe>(m, "HashCpuPipe");
	BindPipe<ClassicPipe>(m, "ClassicPipe");
	// src/NeRFExecutor.h:430-432,451-452,480-493 (HashEmbedder branch; CPU-capable LibTorch embedders)
	m.def("make_hash_cpu", [](Tensor bbox, int n_levels, int n_feat, int log2_t, int base_res, int finest_res, int sh_degree,
		int num_layers, int hidden, int geo_feat, int num_layers_color, int hidden_color) {
		auto p = std::make_unique<HashCpuPipe>();
		p->bbox = bbox;
		p->embed = HashEmbedder("embedder", bbox, n_levels, n_feat, log2_t, base_res, finest_res);
		p->embeddirs = SHEncoder("embeddirs", 3, sh_degree);
		p->model = NeRFSmall(num_layers, hidden, geo_feat, num_layers_color, hidden_color, false, 3, 64,
			p->embed->GetOutputDims(), p->embeddirs->GetOutputDims(), "model");
		p->Finish(torch::kCPU);
		return p;
	});
	// src/NeRFExecutor.h:428,446,478 (classic NeRF branch)
	m.def("make_classic", [](Tensor bbox, int multires, int multires_views, int depth, int width, bool use_viewdirs, bool cuda) {
		auto p = std::make_unique<ClassicPipe>();
		p->bbox = bbox;
		p->embed = Embedder("embedder", multires);
		p->embeddirs = Embedder("embeddirs", multires_views);
		p->model = NeRF(depth, width, p->embed->GetOutputDims(), p->embeddirs->GetOutputDims(), 5, std::set<int>{4}, use_viewdirs, "model");
		p->Finish(cuda ? torch::kCUDA : torch::kCPU);
		return p;
	});
#ifdef NRF_REF_CUDA
	BindPipe<CuHashPipe>(m, "CuHashPipe");
	// src/NeRFExecutor.h:432,452,481-493 (CuHashEmbedder branch) — CUDA only (src/CuHashEmbedder.cpp:24)
	m.def("make_cuhash", [](Tensor bbox, int n_levels, int n_feat, int log2_t, int base_res, int finest_res, int sh_degree,
		int num_layers, int hidden, int geo_feat, int num_layers_color, int hidden_color) {
		auto p = std::make_unique<CuHashPipe>();
		p->bbox = bbox;
		p->embed = CuHashEmbedder("embedder", bbox, n_levels, n_feat, log2_t, base_res, finest_res);
		p->embeddirs = CuSHEncoder("embeddirs", 3, sh_degree);
		p->model = NeRFSmall(num_layers, hidden, geo_feat, num_layers_color, hidden_color, false, 3, 64,
			p->embed->GetOutputDims(), p->embeddirs->GetOutputDims(), "model");
		p->Finish(torch::kCUDA);
		return p;
	});
	m.def("cu_sh_encoder", [](Tensor x, int degree) { CuSHEncoder e("embeddirs", 3, degree); return e->forward(x).first; });   // src/CuSHEncoder.cu:109
	py::class_<LerfRefPipe>(m, "LerfPipe", py::module_local())
		.def("embed_params", &LerfRefPipe::EmbedParams).def("model_params", &LerfRefPipe::ModelParams).def("model_param_names", &LerfRefPipe::ModelParamNames)
		.def("init_model", &LerfRefPipe::InitModel).def("model", &LerfRefPipe::Model).def("run_le_network", &LerfRefPipe::RunLENetwork)
		.def("raw_to_le_outputs", &LerfRefPipe::RawToLEOutputs).def("render", &LerfRefPipe::Render)
		.def("train_steps", &LerfRefPipe::TrainSteps, py::call_guard<py::gil_scoped_release>())
		// src/NeRFExecutor.h:556-560, 1062-1066
		.def("save_checkpoint", [](LerfRefPipe& p, const std::string& dir) {
			torch::save(p.embed, dir + "/lang_embedder_checkpoint.pt");
			torch::save(p.model, dir + "/lang_model_checkpoint.pt");
		})
		.def("load_checkpoint", [](LerfRefPipe& p, const std::string& dir) {
			torch::load(p.embed, dir + "/lang_embedder_checkpoint.pt");
			torch::load(p.model, dir + "/lang_model_checkpoint.pt");
		});
	// src/NeRFExecutor.h:461 and :507-514
	m.def("make_lerf", [](Tensor bbox, int n_levels, int n_feat, int log2_t, int base_res, int finest_res, int geo_feat, int num_layers, int hidden,
		int lang_dim) {
		auto p = std::make_unique<LerfRefPipe>();
		p->bbox = bbox;
		p->embed = CuHashEmbedder("lang_embedder", bbox, n_levels, n_feat, log2_t, base_res, finest_res);
		p->model = LeRF(geo_feat, num_layers, hidden, lang_dim, p->embed->GetOutputDims(), "lang_model");
		p->Finish();
		return p;
	});
#endif
}
