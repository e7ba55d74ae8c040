#pragma once
#include "common.cuh"
#include "gemm.cuh"   // GN_SUM_SCALE / GN_SQ_SCALE

namespace vpu {

struct LnArgs {
    const float* in = nullptr;          // [rows, C]
    const float* gamma = nullptr;
    const float* beta = nullptr;
    float eps = 1e-5f;
    int rows = 0;
    float* out_f32 = nullptr;           // optional
    __nv_bfloat16* out_bf16 = nullptr;  // optional
    __nv_bfloat16* out_pe_bf16 = nullptr;  // optional: bf16(y + pe[row])
    const float* pe = nullptr;
    float* rowmax = nullptr;            // optional: max_c y
    int ld_bf16 = 0;                    // row stride of the two bf16 outputs (elements); 0 = C
};
int layernorm_launch(const LnArgs& a, int C, cudaStream_t stream);

int cast_add_launch(const float* a, const float* b, __nv_bfloat16* out, size_t n, cudaStream_t stream);

constexpr int GN_MAX_CHUNKS = 1024;
// in-place GroupNorm(1, C) (+GELU) on NHWC bf16 [B, per_sample]; partial: [B, GN_MAX_CHUNKS] float2
int groupnorm_launch(__nv_bfloat16* x, int B, size_t per_sample, int C, const float* gamma, const float* beta,
                     int gelu, float2* partial, float2* mean_rstd, cudaStream_t stream);
int groupnorm_apply_launch(__nv_bfloat16* x, int B, size_t per_sample, int C, const float* gamma, const float* beta, int gelu,
                           const long long* sums, cudaStream_t stream);

int qout_gate_launch(const float* q0, const float* q1, const float* q2, const float* q3, int B, int T, int C, float* qout,
                     __nv_bfloat16* qout_bf16, float* cg, cudaStream_t stream);

struct MergeArgs {
    const float* x = nullptr;        // backbone tokens fp32 [M, C]
    const float* cg = nullptr;       // [3, B, C] channel gates (already sigmoid)
    const float* rowmax = nullptr;   // [3, rowmax_parts, M] per-token max of keys_l (pre-sigmoid), as partial maxima over column blocks
    int rowmax_parts = 1;
    __nv_bfloat16* x0 = nullptr;     // bf16(x)            [M, C]
    __nv_bfloat16* x2 = nullptr;     // merged level 1     [M, C]
    __nv_bfloat16* x3 = nullptr;     // merged level 2     [M, C]
    __nv_bfloat16* x4_s2d = nullptr; // merged level 3, space-to-depth [B*(g/2)^2, 4C], k = (kh, kw, c)
    int B = 0, N = 0, M = 0, C = 0, grid = 0;
};
int merge_launch(const MergeArgs& a, cudaStream_t stream);

struct HeadCombineArgs {
    const __nv_bfloat16* y[4] = {nullptr, nullptr, nullptr, nullptr};  // NHWC [B, res[l], res[l], 256]
    int res[4] = {0, 0, 0, 0};
    const float* bias = nullptr;     // fusion conv bias [256]
    __nv_bfloat16* out = nullptr;    // [B*res0^2, 256]
    float* rnorm = nullptr;          // [B*res0^2]
    const float* wseg = nullptr;     // conv_seg weight [256] fp32
    float seg_bias = 0.f;
    float* seg_out = nullptr;        // [B*res0^2] fp32: conv_seg(f) evaluated on the un-rounded fp32 features
    int B = 0;
};
int head_combine_launch(const HeadCombineArgs& a, cudaStream_t stream);
int head_queries_launch(const float* qe, const float* wseg, int B, int nq, __nv_bfloat16* out, cudaStream_t stream);
int upsample_ac_launch(const float* in, float* out, int h, int w, int H, int W, size_t planes, cudaStream_t stream);

}  // namespace vpu
