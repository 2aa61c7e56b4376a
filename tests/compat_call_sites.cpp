// Compile-only check of the drop-in boundary: the call sites of the reference's NeRFExecutor (src/NeRFExecutor.h, lines cited beside each
// statement) and src/main.cpp:220-221, written against the reference's OWN header names, must compile when nerfpp_b200/host/compat is
// first on the include path.  NeRFExecutor.h itself cannot be compiled here (OpenCV, COLMAP, RuCLIP, NeRFactor are absent); this unit holds
// the statements of it that touch the hot-path classes.  Never linked or run (tests/test_abi.py: g++ -fsyntax-only).
#include "BaseEmbedder.h"
#include "CuHashEmbedder.h"
#include "CuSHEncoder.h"
#include "CustomOps.h"
#include "NeRF.h"
#include "NeRFRenderer.h"
#include "RayUtils.h"
#include "Sampler.h"
#include "LeRF.h"
#include "LeRFRenderer.h"

#include <memory>
#include <type_traits>

template <typename TEmbedder, typename TEmbedDirs, typename TNeRF, typename TNeRFRenderer, typename TLeRFEmbedder, typename TLeRF, typename TLeRFRenderer>
struct ExecutorCallSites {
	TEmbedder ExecutorEmbedder = nullptr;                    // src/NeRFExecutor.h:303-310
	TEmbedDirs ExecutorEmbeddirs = nullptr;
	TNeRF Model = nullptr;
	std::unique_ptr<TNeRFRenderer> NeRFRenderer = nullptr;
	TLeRFEmbedder LangEmbedder = nullptr;
	TLeRF LangModel = nullptr;
	std::unique_ptr<TLeRFRenderer> LeRFRenderer = nullptr;
	std::vector<torch::Tensor> GradVars;
	std::unique_ptr<torch::optim::Adam> Optimizer;

	void Initialize(torch::Tensor bounding_box, torch::Device device)
	{
		int input_ch = 0, input_ch_views = 0;
		if constexpr (std::is_same_v<TEmbedder, Embedder>) ExecutorEmbedder = Embedder("embedder", 10);                                        // :428
		if constexpr (std::is_same_v<TEmbedder, CuHashEmbedder>) ExecutorEmbedder = CuHashEmbedder("embedder", bounding_box.to(device), 16, 2, 19, 16, 512);   // :432
		ExecutorEmbedder->to(device);                                                                                                       // :434
		input_ch = ExecutorEmbedder->GetOutputDims();                                                                                       // :435
		auto embp = ExecutorEmbedder->parameters();                                                                                         // :436
		GradVars.insert(GradVars.end(), std::make_move_iterator(embp.begin()), std::make_move_iterator(embp.end()));                        // :437
		(void)Trainable::ParamsCount(ExecutorEmbedder);                                                                                     // :441
		if constexpr (std::is_same_v<TEmbedDirs, Embedder>) ExecutorEmbeddirs = Embedder("embeddirs", 4);                                   // :446
		if constexpr (std::is_same_v<TEmbedDirs, CuSHEncoder>) ExecutorEmbeddirs = CuSHEncoder("embeddirs", 3, 4);                          // :452
		input_ch_views = ExecutorEmbeddirs->GetOutputDims();                                                                                // :453
		ExecutorEmbeddirs->to(device);                                                                                                      // :454
		LangEmbedder = CuHashEmbedder("lang_embedder", bounding_box.to(device), 16, 8, 19, 16, 512);                                        // :461
		LangEmbedder->to(device);                                                                                                           // :463
		if constexpr (std::is_same_v<TNeRF, NeRF>) Model = NeRF(8, 256, input_ch, input_ch_views, 5, std::set<int>{4}, true, "model");      // :478
		if constexpr (std::is_same_v<TNeRF, NeRFSmall>) Model = NeRFSmall(2, 64, 15, 3, 64, false, 3, 64, input_ch, input_ch_views, "model");   // :481-493
		Model->to(device);                                                                                                                  // :495
		Trainable::Initialize(Model);
		LangModel = TLeRF(32, 2, 256, 512, LangEmbedder->GetOutputDims(), "lang_model");                                                    // :507-514
		LangModel->to(device);                                                                                                              // :516
		auto mp = LangModel->parameters();                                                                                                  // :517
		GradVars.insert(GradVars.end(), std::make_move_iterator(mp.begin()), std::make_move_iterator(mp.end()));                            // :518
		NeRFRenderer = std::make_unique<TNeRFRenderer>(ExecutorEmbedder, ExecutorEmbeddirs, Model);                                         // :528
		LeRFRenderer = std::make_unique<TLeRFRenderer>(LangEmbedder, LangModel);                                                            // :534
		Optimizer = std::make_unique<torch::optim::Adam>(GradVars, torch::optim::AdamOptions(1e-2).eps(1e-15).betas(std::make_tuple(0.9, 0.99)));   // :539
		torch::load(LangEmbedder, "lang_embedder_checkpoint.pt");                                                                           // :559
		torch::save(LangModel, "lang_model_checkpoint.pt");                                                                                 // :1066
	}

	std::tuple<NeRFRenderResult, LeRFRenderResult> RenderView(torch::Tensor render_pose, int w, int h, torch::Tensor k, const NeRFRenderParams& rparams)
	{
		NeRFRenderResult nerf_render_result = NeRFRenderer->Render(h, w, k, rparams, {torch::Tensor(), torch::Tensor(), torch::Tensor()}, render_pose, torch::Tensor());   // :633-638
		LeRFRenderResult lerf_render_result = LeRFRenderer->Render(h, w, k, rparams, {torch::Tensor(), torch::Tensor(), torch::Tensor()}, render_pose, torch::Tensor());   // :642-647
		torch::Tensor rel = lerf_render_result.Outputs.Relevancy;                                                                           // :713
		(void)rel;
		return std::make_tuple(nerf_render_result, lerf_render_result);                                                                     // :650
	}

	torch::Tensor TrainStep(torch::Tensor rays_o, torch::Tensor rays_d, torch::Tensor cone_angle, torch::Tensor target, torch::Tensor target_lang, const NeRFRenderParams& render_params)
	{
		Optimizer->zero_grad();
		auto rgb_disp_acc_extras = NeRFRenderer->Render(0, 0, torch::Tensor(), render_params, {rays_o, rays_d, cone_angle}, torch::Tensor(), torch::Tensor());   // :876-878
		auto loss = torch::nn::functional::huber_loss(rgb_disp_acc_extras.Outputs.RGBMap, target.detach());                                 // :883-886
		loss.backward();                                                                                                                    // :923
		auto lerf_render_result = LeRFRenderer->Render(0, 0, torch::Tensor(), render_params, {rays_o, rays_d, cone_angle}, torch::Tensor(), torch::Tensor());    // :960-962
		auto lang_loss = torch::nn::functional::huber_loss(lerf_render_result.Outputs.RenderedLangEmbedding, target_lang.detach(),
			torch::nn::functional::HuberLossFuncOptions().reduction(torch::kNone).delta(1.25)).sum(-1).nanmean();                           // :964-968
		lang_loss.backward();                                                                                                               // :981
		Optimizer->step();                                                                                                                  // :986
		LeRFRenderer->SetLeRFPrompts(target_lang, target_lang);                                                                             // :760
		auto prompts = LeRFRenderer->GetLeRFPrompts();                                                                                      // :768
		(void)prompts;
		auto z = SamplePDF(target, target, 8, true);                                                                                        // src/Sampler.h:6
		auto nf = IntersectWithAABB(rays_o, rays_d, render_params.BoundingBox, 0.f);                                                        // src/RayUtils.h:87
		auto rays = GetRays(4, 4, target, target);                                                                                          // src/RayUtils.h:23
		auto te = torch::autograd::TruncExp::apply(target)[0];                                                                              // src/CustomOps.h
		auto clip = RenderCLIPEmbedding(target_lang, target);                                                                               // src/LeRFRenderer.h:45
		(void)z; (void)nf; (void)rays; (void)te; (void)clip;
		return loss + lang_loss;
	}
};

// src/main.cpp:220-221 (HashNeRF + LeRF) and the classic instantiation
template struct ExecutorCallSites<CuHashEmbedder, CuSHEncoder, NeRFSmall, NeRFRenderer<CuHashEmbedder, CuSHEncoder, NeRFSmall>, CuHashEmbedder, LeRF, LeRFRenderer>;
template struct ExecutorCallSites<Embedder, Embedder, NeRF, NeRFRenderer<Embedder, Embedder, NeRF>, CuHashEmbedder, LeRF, LeRFRenderer>;
