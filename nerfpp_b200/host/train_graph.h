// NeRFExecutor::Train's loop body for NeRFRenderer<CuHashEmbedder, CuSHEncoder, NeRFSmall> as ONE CUDA-graph replay on the C++ surface.
//
// The reference's iteration (src/NeRFExecutor.h:868-890, 923, 986-996) is
//     Optimizer->zero_grad(); Render(rays) -> huber_loss(RGBMap, target) -> loss.backward(); Optimizer->step(); set_lr(decayed)
// which, on the drop-in classes, is ~60 kernel launches behind ATen / autograd bookkeeping (1.2 ms per 4096-ray step, of which 0.8 ms is kernel
// time).  HashNeRFTrainGraph runs the SAME step — the same C-ABI kernels in the same order as the autograd path issues them — as one captured
// graph of 14 kernel nodes:
//     ray setup -> [hash encode -> NeRFSmall] coarse -> RawToOutputs -> SamplePDF + merge -> [hash encode -> NeRFSmall] fine ->
//     RawToOutputs + huber + RawToOutputs backward (one kernel) -> NeRFSmall backward -> hash scatter -> Adam (+ fp16 shadow, gradient clear,
//     LR schedule on the device) -> weight re-pack
// Parity configuration of the benchmark (ThinRay, Perturb 0, no raw noise, no stochastic preconditioning, one chunk, coarse pass not
// differentiated — SURVEY §9-Q3/Q4); anything else stays on the autograd path of renderer.h.
//
// The modules stay the owners of their parameters: at construction the embedder's Embeddings and the five NeRFSmall weights are re-pointed
// (Tensor::set_data) at slices of ONE flat fp32 vector, so parameters(), named_parameters(), torch::save / torch::load archives and the
// autograd path keep working on the same storage, while the step needs one gradient buffer and two Adam launches (table prefix the kernels can
// reach, then the 9 344 weights).  Adam moments live here; they are exposed for optimiser checkpoints (ExpAvg / ExpAvgSq / StepCount).
#pragma once
#include <ATen/cuda/CUDAGraph.h>

#include "embedders.h"
#include "models.h"

class HashNeRFTrainGraph {
public:
	HashNeRFTrainGraph(CuHashEmbedder embed, CuSHEncoder embeddirs, NeRFSmall model, torch::Tensor bounding_box, int n_samples, int n_importance,
		float learning_rate, int lrate_decay, int64_t n_rays);
	/// one optimisation step on rays_o / rays_d [n_rays,3] and target [n_rays,3] (device or pinned host tensors); returns the loss (device scalar,
	/// overwritten by the next step)
	torch::Tensor Step(const torch::Tensor& rays_o, const torch::Tensor& rays_d, const torch::Tensor& target);
	int64_t StepCount() const { return Steps; }
	int KernelsPerStep() const { return Kernels; }
	torch::Tensor ExpAvg() const { return M; }
	torch::Tensor ExpAvgSq() const { return V; }
	torch::Tensor RGB() const { return Rgb; }        ///< RGBMap of the last step [n_rays,3]

private:
	void Enqueue(bool with_optimizer);                ///< the step's kernels on the current stream (captured once)
	CuHashEmbedder Embed = nullptr;
	CuSHEncoder EmbedDirs = nullptr;
	NeRFSmall Model = nullptr;
	std::array<float, 6> Box{};
	int S, N;
	float Lr0;
	int LrateDecay;
	int64_t R, NTableAll, NTableUsed, NMlp, Steps = 0;
	int Kernels = 0;
	torch::Tensor P, G, M, V, Shadow, Packed, Sched, Loss, TVals, U, InO, InD, InT, Rgb;
	std::vector<torch::Tensor> Keep;                  ///< intermediates of the captured step (their storage belongs to the graph's pool)
	at::cuda::CUDAGraph Graph;
};
