// TEST INFRASTRUCTURE ONLY.  Stand-in for the RuCLIP header that src/LeRFRenderer.cpp:2 includes for `Relevancy`.
// RuCLIP is a sibling checkout the reference does not vendor (CMakeLists.Files.txt:8-10), so the real function is absent:
// Relevancy is PARITY UNPINNED and out of scope.  This stub only lets LeRFRenderer.cpp compile; it returns an undefined
// tensor, so LeRFRendererOutputs::Relevancy is never produced by the oracle.
#pragma once
#include <torch/torch.h>
inline torch::Tensor Relevancy(torch::Tensor, torch::Tensor, torch::Tensor) { return torch::Tensor(); }
