// Tail of the segmentation head at 1/4 resolution (head_tail.cu): fusion-conv sum over the resized per-level slices, ReLU,
// conv_seg and the P2CL cosine logits as one tcgen05 kernel (reference swin_transformer.py:727-767).
#pragma once
#include "gemm.cuh"

namespace vpu {

struct HeadTailArgs {
    const __nv_bfloat16* y[4] = {nullptr, nullptr, nullptr, nullptr};   // per-level fusion-conv slices, NHWC [B, res, res, channels]
    int res[4] = {0, 0, 0, 0};                                          // res[0] = 1/4 scale, res[l] = res[0] >> l
    int B = 0, channels = 0;
    const float* bias = nullptr;        // fusion conv bias [channels]
    const float* wseg = nullptr;        // conv_seg weight [channels]
    float seg_bias = 0.f;
    const __nv_bfloat16* qn = nullptr;  // L2-normalised prompt queries [B, 64, channels] (rows >= nq unused); only read with aux_out
    int nq = 0;
    float* seg_out = nullptr;           // [B, res0, res0]
    float* aux_out = nullptr;           // [B, nq, res0, res0] or nullptr
};

int head_tail_prepare(int res0);        // builds the interpolation-matrix table of this geometry (allocates once; call before capture)
bool head_tail_supported(const HeadTailArgs& a);
int head_tail_launch(const HeadTailArgs& a, cudaStream_t stream);

}  // namespace vpu
