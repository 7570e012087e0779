// TEST INFRASTRUCTURE ONLY.  Side translation unit linked next to the
// UNMODIFIED reference sources (oracle/build_ref.py).  It contains no decode
// logic: it declares the reference's own forward function, which has external
// linkage (/root/reference/torbi/csrc/viterbi.cpp:35), and exposes its two
// intermediate tensors -- the fp32 final posterior (B,S) and the int32
// backpointer trellis (B,T,S) -- so the CUDA path can be compared against them.
#include <ATen/ATen.h>
#include <torch/library.h>
#include <tuple>

namespace torbi {

void viterbi_make_trellis_cpu(
    const at::Tensor observation,
    const at::Tensor batch_frames,
    const at::Tensor transition,
    const at::Tensor initial,
    at::Tensor posterior,
    at::Tensor trellis_tensor);

static std::tuple<at::Tensor, at::Tensor> debug_trellis(
    const at::Tensor& observation,
    const at::Tensor& batch_frames,
    const at::Tensor& transition,
    const at::Tensor& initial) {
    // same allocations as viterbi_decode_cpu (viterbi.cpp:201-206)
    at::Tensor trellis = at::zeros(
        {observation.size(0), observation.size(1), observation.size(2)},
        at::dtype(at::kInt));
    at::Tensor posterior = at::zeros(
        {observation.size(0), observation.size(2)}, at::dtype(at::kFloat));
    viterbi_make_trellis_cpu(
        observation.contiguous(), batch_frames.contiguous(),
        transition.contiguous(), initial.contiguous(), posterior, trellis);
    return std::make_tuple(posterior, trellis);
}

TORCH_LIBRARY_FRAGMENT(torbi, m) {
    m.def("debug_trellis(Tensor observation, Tensor batch_frames, Tensor transition, Tensor initial) -> (Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(torbi, CPU, m) {
    m.impl("debug_trellis", &debug_trellis);
}

}  // namespace torbi
