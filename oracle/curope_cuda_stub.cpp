// TEST INFRASTRUCTURE ONLY -- link-time stand-in for the reference's `rope_2d_cuda`
// (curope/kernels.cu:84-108), which does not compile against torch 2.11 (kernels.cu:101).
// Only the reference's CPU loop in oracle/_ref is used; reaching this is an error.
#include <torch/extension.h>
void rope_2d_cuda(torch::Tensor, const torch::Tensor, const float, const float) {
  TORCH_CHECK(false, "oracle/_ref: the reference CUDA RoPE kernel is not buildable here; CPU path only");
}
