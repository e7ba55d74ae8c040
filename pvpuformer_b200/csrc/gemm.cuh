// GEMM interface shared by the forward orchestration and the C-ABI test entry points.
//   out[m, n] = epilogue( sum_k A[m, k] * W[n, k] )      A: bf16 [M, K] row-major, W: bf16 [N, K]
// (the nn.Linear / 1x1-conv convention of the reference, so packed weights keep their layout).
#pragma once
#include "common.cuh"

namespace vpu {

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };
enum { EPI_PLAIN = 0, EPI_PIXEL_SHUFFLE = 1, EPI_HEAD_FINAL = 2 };
// fixed-point scales of the GroupNorm accumulators: each fp32 partial (one row x four columns) is rounded to 2^-24 resp.
// 2^-16 before it is added; int64 holds |sum| < 5e11 and a sum of squares < 1.4e14 per sample
constexpr float GN_SUM_SCALE = 16777216.0f, GN_SQ_SCALE = 65536.0f;

struct Epi {
    void* out = nullptr;            // [rows, ldo]; fp32 or bf16
    int out_bf16 = 1;
    int ldo = 0;
    const float* bias = nullptr;    // [N]
    const float* bias2d = nullptr;  // [bias2d_rows, N], row = m % bias2d_rows (positional tables)
    int bias2d_rows = 0;
    int bias2d_pad_rows = 0;        // rows 0 .. pad-1 of the table are repeated after row bias2d_rows-1 (a 128-row TMA box never wraps)
    const void* res = nullptr;      // residual [M, ldr], fp32 or bf16, added before the activation is NOT
    int res_bf16 = 0;               //   applied (act and res are never combined on this path)
    int ldr = 0;
    int act = ACT_NONE;
    int mode = EPI_PLAIN;
    // EPI_PIXEL_SHUFFLE: ConvTranspose2d(k=2,s=2) as a GEMM with N = 4*cout ordered (kh, kw, co):
    // input row m = (b, i, j) on a g x g grid goes to output row (b, 2i+kh, 2j+kw), column co.
    int ps_g = 0, ps_cout = 0;
    // EPI_HEAD_FINAL: per-sample B operand (rows [b*b_rows_per_batch, +N)), NCHW fp32 stores:
    // n < nq : aux[b, n, pix] = (acc * rnorm[m] + 1) / 2 ;  n == nq : seg[b, pix] = acc + seg_bias
    int m_per_batch = 0, b_rows_per_batch = 0;
    const float* rnorm = nullptr;
    float* aux_out = nullptr;
    float* seg_out = nullptr;
    float seg_bias = 0.f;
    int nq = 0;
    // GroupNorm(1, C) fusion for the neck (reference is_vpu_model.py:55-86; statistics over all values of one sample):
    //   gn_out  [samples][2] int64 fixed point (GN_SUM_SCALE / GN_SQ_SCALE): += (sum, sum of squares) of this GEMM's fp32
    //           outputs, per sample of gn_rows M-rows.  Integer accumulation is associative, so the statistics -- and with
    //           them the whole forward -- stay bit-identical whatever the batch size, tile assignment or atomic order
    //   gn_in   the same pair for the tensor the A operand holds un-normalised: the weights bound for this GEMM are
    //           W' = W diag(gamma) and bias = W beta + b, so GN folds into  out = rstd * acc - mean * rstd * gn_wg[n] + bias[n]
    //           with gn_wg[n] = sum_k W'[n, k]  (mean / rstd from gn_in over gn_in_count values, eps 1e-5)
    long long* gn_out = nullptr;
    const long long* gn_in = nullptr;
    const float* gn_wg = nullptr;
    int gn_rows = 0;
    float gn_in_count = 0.f;
    // LayerNorm fusion for the ViT blocks (reference models_vit.py:72-75: x + attn(norm1(x)), x + mlp(norm2(x))).
    //   ln_out      [M][ln_slots] (sum, sum of squares) of the fp32 rows this GEMM writes (the residual stream), one slot per
    //               (256-column tile, epilogue warp): every slot is written exactly once, by fixed threads in a fixed order, and
    //               the reader adds the slots in index order -- no atomics, bit-identical from run to run and for any batch
    //   ln_out_bf16 bf16 copy of the same rows [M, ldo]: the A operand of the next GEMM
    //   ln_in       [M] (rstd, mean * rstd) of the rows the A operand holds un-normalised (ln_rowstats_launch turns the slots
    //               into this).  The weights bound for this GEMM are W' = W diag(gamma), bias = W beta + b,
    //               ln_s[n] = sum_k W'[n, k] (of the bf16-rounded W'), so that  LN(x) W^T + b = rstd (x W'^T) - rstd mean ln_s + bias
    float2* ln_out = nullptr;
    __nv_bfloat16* ln_out_bf16 = nullptr;
    const float2* ln_in = nullptr;
    const float* ln_s = nullptr;
    int ln_slots = 0;
};

// slots per row of Epi::ln_out for an [M, N] residual GEMM (column tiles x epilogue warps per TMEM lane quarter)
int gemm_ln_slots(int M, int N);     // for the tile width the launch will choose for this shape (256, or 128 when that leaves SMs idle)
int gemm_ln_slots_max(int N);        // buffer sizing
// slots [M][P] -> (rstd, mean * rstd) [M] of a LayerNorm over C values with the given eps
int ln_rowstats_launch(const float2* slots, int M, int P, int C, float eps, float2* out, cudaStream_t stream);

struct GemmProblem {
    const __nv_bfloat16* A = nullptr;  // [M, lda]
    const __nv_bfloat16* W = nullptr;  // [N (x batches), ldw]
    int M = 0, N = 0, K = 0;
    int lda = 0, ldw = 0;
    int w_rows = 0;                    // total rows of W (N, or batches*b_rows_per_batch)
    Epi epi;
};

// impl: 0 = tcgen05/TMA (product path), 1 = mma.sync reference kernel (debug / cross-check only)
int gemm_launch(const GemmProblem& p, cudaStream_t stream, int impl = 0);
int gemm_init();  // resolves cuTensorMapEncodeTiled, sets smem attributes; idempotent
// cached 2-D tensor map over a bf16 [rows, cols] matrix with leading dimension ld: box = 64 columns x box_rows rows, 128-byte swizzle
int gemm_tmap(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int gemm_num_sms();

// Residual-stream GEMM with a TMA-staged epilogue (gemm_res.cu): X (fp32, in place) += A W^T + bias, bf16 copy, LayerNorm slots
bool gemm_res_supported(const GemmProblem& p);
int gemm_res_launch(const GemmProblem& p, cudaStream_t stream);
// same skeleton: bf16 out = A W^T + periodic fp32 table (needs Epi::bias2d_pad_rows >= 128)
bool gemm_tab_supported(const GemmProblem& p);
int gemm_tab_launch(const GemmProblem& p, cudaStream_t stream);
// Out-projection + bf16 residual + LayerNorm in one kernel, a cluster of C / 256 CTAs per 128-row tile (gemm_ln.cu):
// out (bf16) = LayerNorm_C(A W^T + bias + res) gamma + beta; rowmax_parts[r][m] = max over CTA r's 256 columns of row m (optional)
struct GemmLn {
    const __nv_bfloat16* A = nullptr;
    const __nv_bfloat16* W = nullptr;       // [C, K]
    const float* bias = nullptr;
    const __nv_bfloat16* res = nullptr;     // [M, C]
    const float* gamma = nullptr;
    const float* beta = nullptr;
    __nv_bfloat16* out = nullptr;           // [M, C]
    float* rowmax_parts = nullptr;          // [gemm_ln_parts(C), M]
    float eps = 1e-5f;
    int M = 0, K = 0, C = 0, lda = 0, ldw = 0, ldr = 0, ldo = 0;
};
int gemm_ln_parts(int C);
bool gemm_ln_supported(const GemmLn& p);
int gemm_ln_launch(const GemmLn& p, cudaStream_t stream);
// GroupNorm-fused neck GEMM with a TMA-staged epilogue (gemm_gn.cu): bf16 out = [rstd (A W^T) - mean rstd wg] + bias, int64 statistics
bool gemm_gn_supported(const GemmProblem& p);
int gemm_gn_launch(const GemmProblem& p, cudaStream_t stream);
// ConvTranspose2d(k=2, s=2) GEMMs of the neck: pixel-shuffle store as one 5-D TMA box per 64-column group (gemm_gn.cu, PS mode)
bool gemm_ps_tma_supported(const GemmProblem& p);
int gemm_ps_tma_launch(const GemmProblem& p, cudaStream_t stream);

// Back-to-back GEMM pair of the segmentation head (gemm_b2b.cu): per pyramid level the 1x1 conv + ReLU and that level's
// [256, 256] slice of the fusion conv (reference swin_transformer.py:723-737; the slice is applied at native resolution
// because a 1x1 conv commutes with the bilinear resize),
//     Y[M, 256] = bf16( bf16(relu(A[M, K1] W1^T + b1)) W2^T ),
// with the [M, 256] intermediate living only in shared memory (it was 2 x 411 MB of HBM traffic per batch-64 step at 1/4 res).
struct GemmB2B {
    const __nv_bfloat16* A = nullptr;   // [M, lda]
    const __nv_bfloat16* W1 = nullptr;  // [256, K1]
    const __nv_bfloat16* W2 = nullptr;  // [256, 256]
    const float* bias1 = nullptr;       // [256]
    __nv_bfloat16* out = nullptr;       // [M, ldo]
    int M = 0, K1 = 0, lda = 0, ldo = 0;
};
bool gemm_b2b_supported(const GemmB2B& p, int n1, int n2);
int gemm_b2b_launch(const GemmB2B& p, cudaStream_t stream);

}  // namespace vpu
