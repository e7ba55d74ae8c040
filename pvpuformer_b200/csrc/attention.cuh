#pragma once
#include "common.cuh"

namespace vpu {

// Maps (problem index, position within the problem's sequence) -> row of the token-major buffer.
//   mode 0: row = prob * per_prob + s                       (contiguous sequences)
//   mode 1: 224-px window regrouping of reference models_vit.py:225-255: prob = b * nw^2 + window,
//           s = i * win + j inside the window, row = b * tokens + (wi*win + i) * grid + wj*win + j
struct RowMap {
    int mode = 0;
    int per_prob = 0;
    int tokens = 0, grid = 0, win = 0;
};

struct AttnArgs {
    const __nv_bfloat16* q = nullptr;
    const __nv_bfloat16* k = nullptr;
    const __nv_bfloat16* v = nullptr;
    __nv_bfloat16* o = nullptr;
    int ldq = 0, ldk = 0, ldv = 0, ldo = 0;     // row strides (elements)
    int qoff = 0, koff = 0, voff = 0;           // column offsets of head 0 (elements)
    int Sq = 0, Sk = 0;                         // queries / keys per problem
    int heads = 0, nprob = 0;
    float scale_log2 = 0.f;                     // softmax scale * log2(e)
    RowMap qmap, kmap;                          // o uses qmap
};

int attention_launch(const AttnArgs& a, int head_dim, cudaStream_t stream);

// tcgen05 / TMEM kernel for 14x14-token windows, head_dim 64 (attention_tc.cu)
bool window_attention_tc_supported(const AttnArgs& a, int head_dim);
int window_attention_tc_launch(const AttnArgs& a, cudaStream_t stream);
// tcgen05 / TMEM kernel for 16x16-token windows, head_dim 80 (ViT-H; attention_tc.cu)
bool window_attention_tc80_supported(const AttnArgs& a, int head_dim);
int window_attention_tc80_launch(const AttnArgs& a, cudaStream_t stream);
// tcgen05 / TMEM flash kernel for un-windowed self-attention, head_dim 64, S a multiple of 112 (attention_tc.cu)
bool global_attention_tc_supported(const AttnArgs& a, int head_dim);
int global_attention_tc_launch(const AttnArgs& a, cudaStream_t stream);
// same for head_dim 80, S a multiple of 64 (ViT-H; attention_tc.cu)
bool global_attention_tc80_supported(const AttnArgs& a, int head_dim);
int global_attention_tc80_launch(const AttnArgs& a, cudaStream_t stream);
// tcgen05 / TMEM kernels for the Dual-cross Merging Attention shapes: 48 keys (prompt self-attention, image -> tokens) or
// <= 128 queries of an even number of heads (tokens -> image); head_dim 48 ... 160 (attention_dma.cu)
bool dma_attention_tc_supported(const AttnArgs& a, int head_dim);
int dma_attention_tc_launch(const AttnArgs& a, int head_dim, cudaStream_t stream);
// measurement only: CTA 0 of the next global-attention launches logs (event << 56 | clock64) into dev_buf[4 roles][cap]
void attention_debug_trace(unsigned long long* dev_buf, int cap);

}  // namespace vpu
