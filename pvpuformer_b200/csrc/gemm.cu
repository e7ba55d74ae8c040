// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA (128B-swizzled tiles) -> shared memory
// -> tcgen05.mma (accumulators in TMEM, double-buffered) -> tcgen05.ld epilogue with the fused
// bias / positional-table / GELU / ReLU / residual / pixel-shuffle / head-final stores.
//
// Reference ops served (all of them GEMMs on this path): patch embeds
// (models_vit.py:94-104), qkv/proj/fc1/fc2 (models_vit.py:43-56,21-27), PPuE FFN
// (common.py:28-42), all DMA projections + MLP (transformer.py:499-521, common.py:13-26),
// neck ConvTranspose/Conv stride==kernel (is_vpu_model.py:55-86), head 1x1 convs, fusion conv,
// conv_seg and the P2CL cosine logits (swin_transformer.py:723-767).
//
// CTA = 10 warps: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..9 epilogue
// (two warps per TMEM lane quarter, each taking half of the tile's columns).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gemm.cuh"

namespace vpu {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

// Programmatic dependent launch.  VPU_PDL=1 / 0 forces it on / off; by default the forward switches it on for small batches only
// (pdl_set_auto): a batch-2 forward (one NoBRS click with flip TTA) is a chain of ~180 launch-latency-bound kernels and gains
// 16 % (2.38 -> 2.00 ms), while on the batch-64 step of the power-capped B200s of this pool closing the inter-kernel gaps lowered
// the SM clock by as much as it saved (16.54 ms at 1620 MHz with PDL, 16.15 ms at 1725 MHz without).
static thread_local bool g_pdl_auto = false;
void pdl_set_auto(bool on) { g_pdl_auto = on; }
bool pdl_enabled() {
    static const int forced = [] { const char* e = vpu_debug_env("VPU_PDL"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    return forced >= 0 ? forced == 1 : g_pdl_auto;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------
// Epilogue: CNT consecutive columns [n0, n0+CNT) of output row m.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void gn_mean_rstd(const long long* sums, int sample, float count, float& rstd, float& mean_rstd) {
    const double inv_n = 1.0 / (double)count;
    const double mean = (double)sums[2 * sample] * (1.0 / (double)GN_SUM_SCALE) * inv_n;
    double var = (double)sums[2 * sample + 1] * (1.0 / (double)GN_SQ_SCALE) * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    const double r = 1.0 / sqrt(var + 1e-5);
    rstd = (float)r;
    mean_rstd = (float)(mean * r);
}

template <int CNT>
__device__ __forceinline__ void epi_store(const Epi& e, int m, int n0, float (&v)[CNT], int N) {
    if (e.gn_in) {
        float r, mr;
        gn_mean_rstd(e.gn_in, m / e.gn_rows, e.gn_in_count, r, mr);
#pragma unroll
        for (int t = 0; t < CNT; ++t) v[t] = fmaf(v[t], r, -mr * __ldg(e.gn_wg + n0 + t));
    }
    if (e.bias) {
        if constexpr (CNT % 4 == 0) {
#pragma unroll
            for (int t = 0; t < CNT; t += 4) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + t));
                v[t] += x.x; v[t + 1] += x.y; v[t + 2] += x.z; v[t + 3] += x.w;
            }
        } else {
#pragma unroll
            for (int t = 0; t < CNT; ++t) v[t] += __ldg(e.bias + n0 + t);
        }
    }
    if (e.bias2d) {
        const float* b2 = e.bias2d + (size_t)(m % e.bias2d_rows) * N + n0;
        if constexpr (CNT % 4 == 0) {
#pragma unroll
            for (int t = 0; t < CNT; t += 4) {
                float4 x = __ldg(reinterpret_cast<const float4*>(b2 + t));
                v[t] += x.x; v[t + 1] += x.y; v[t + 2] += x.z; v[t + 3] += x.w;
            }
        } else {
#pragma unroll
            for (int t = 0; t < CNT; ++t) v[t] += __ldg(b2 + t);
        }
    }
    if (e.mode == EPI_HEAD_FINAL) {
        const int b = m / e.m_per_batch, pix = m % e.m_per_batch;
        const float rn = e.rnorm[m];
#pragma unroll
        for (int t = 0; t < CNT; ++t) {
            const int n = n0 + t;
            if (n < e.nq) {
                if (e.aux_out) e.aux_out[((size_t)b * e.nq + n) * e.m_per_batch + pix] = (v[t] * rn + 1.0f) * 0.5f;
            } else if (n == e.nq && e.seg_out) {
                e.seg_out[(size_t)b * e.m_per_batch + pix] = v[t] + e.seg_bias;
            }
        }
        return;
    }
    if (e.res) {
        if (e.res_bf16) {
            const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(e.res) + (size_t)m * e.ldr + n0;
            if constexpr (CNT % 8 == 0) {
#pragma unroll
                for (int t = 0; t < CNT; t += 8) {
                    uint4 x = *reinterpret_cast<const uint4*>(r + t);
                    float2 a = unpack_bf16(x.x), b = unpack_bf16(x.y), c = unpack_bf16(x.z), d = unpack_bf16(x.w);
                    v[t] += a.x; v[t + 1] += a.y; v[t + 2] += b.x; v[t + 3] += b.y;
                    v[t + 4] += c.x; v[t + 5] += c.y; v[t + 6] += d.x; v[t + 7] += d.y;
                }
            } else {
#pragma unroll
                for (int t = 0; t < CNT; ++t) v[t] += __bfloat162float(r[t]);
            }
        } else {
            const float* r = reinterpret_cast<const float*>(e.res) + (size_t)m * e.ldr + n0;
            if constexpr (CNT % 4 == 0) {
#pragma unroll
                for (int t = 0; t < CNT; t += 4) {
                    float4 x = *reinterpret_cast<const float4*>(r + t);
                    v[t] += x.x; v[t + 1] += x.y; v[t + 2] += x.z; v[t + 3] += x.w;
                }
            } else {
#pragma unroll
                for (int t = 0; t < CNT; ++t) v[t] += r[t];
            }
        }
    }
    if (e.gn_out) {
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int t = 0; t < CNT; ++t) { sum += v[t]; sq = fmaf(v[t], v[t], sq); }
        atomicAdd(reinterpret_cast<unsigned long long*>(e.gn_out) + 2 * (m / e.gn_rows), (unsigned long long)__float2ll_rn(sum * GN_SUM_SCALE));
        atomicAdd(reinterpret_cast<unsigned long long*>(e.gn_out) + 2 * (m / e.gn_rows) + 1, (unsigned long long)__float2ll_rn(sq * GN_SQ_SCALE));
    }
    if (e.act == ACT_GELU) {
#pragma unroll
        for (int t = 0; t < CNT; ++t) v[t] = gelu_erf(v[t]);
    } else if (e.act == ACT_RELU) {
#pragma unroll
        for (int t = 0; t < CNT; ++t) v[t] = fmaxf(v[t], 0.0f);
    }
    size_t orow = (size_t)m;
    int ocol = n0;
    if (e.mode == EPI_PIXEL_SHUFFLE) {
        const int g = e.ps_g, gg = g * g;
        const int b = m / gg, ij = m % gg, i = ij / g, j = ij % g;
        const int q = n0 / e.ps_cout;
        ocol = n0 % e.ps_cout;
        orow = ((size_t)b * 2 * g + 2 * i + (q >> 1)) * (size_t)(2 * g) + 2 * j + (q & 1);
    }
    if (e.out_bf16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + orow * e.ldo + ocol;
        if constexpr (CNT % 8 == 0) {
#pragma unroll
            for (int t = 0; t < CNT; t += 8) {
                uint4 x;
                x.x = pack_bf16(v[t], v[t + 1]); x.y = pack_bf16(v[t + 2], v[t + 3]);
                x.z = pack_bf16(v[t + 4], v[t + 5]); x.w = pack_bf16(v[t + 6], v[t + 7]);
                *reinterpret_cast<uint4*>(o + t) = x;
            }
        } else {
#pragma unroll
            for (int t = 0; t < CNT; t += 2) *reinterpret_cast<uint32_t*>(o + t) = pack_bf16(v[t], v[t + 1]);
        }
    } else {
        float* o = reinterpret_cast<float*>(e.out) + orow * e.ldo + ocol;
        if constexpr (CNT % 4 == 0) {
#pragma unroll
            for (int t = 0; t < CNT; t += 4) *reinterpret_cast<float4*>(o + t) = make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]);
        } else {
#pragma unroll
            for (int t = 0; t < CNT; t += 2) *reinterpret_cast<float2*>(o + t) = make_float2(v[t], v[t + 1]);
        }
    }
}


// ------------------------------------------------------------------------------------------
// Tile epilogue of the tcgen05 kernels.  tcgen05.ld hands every thread ONE accumulator row
// (32 consecutive columns), which makes direct global access row-strided: 32 lanes touch 32
// different lines per instruction (round-1 profile: the K=768 GEMMs were epilogue-bound).  Each
// epilogue warp therefore transposes its 32x32 chunk through a private padded shared-memory buffer
// (pitch 36 words: conflict-free 16-byte row writes and row reads) and then works row-major:
// 8 lanes x 4 columns cover one row, so a warp instruction touches 4 rows x 64-128 contiguous
// bytes, residual / table loads and stores are full-sector, and the residual loads of a whole
// chunk are in flight before the accumulator arrives.
// The epilogue is specialised at compile time (EpiKind): with every variant behind run-time flags
// the row loop was ~20 instructions per 64 elements plus uniform branches, and with only two
// epilogue warps per scheduler that issue stream -- not memory -- bounded the K=768 GEMMs.
// ------------------------------------------------------------------------------------------
struct GemmDims {
    int M, N, K;
    int stages;   // pipeline depth actually used (<= the compiled maximum; VPU_GEMM_STAGES experiment knob)
    int ablate;   // measurement only (VPU_GEMM_ABLATE): 1 = no TMA loads (stale operands), 2 = no MMA issue; results are garbage
};

enum EpiKind {
    EK_BF16 = 0,       // out bf16 = acc + bias                       (qkv, DMA / neck / head-Y projections)
    EK_BF16_GELU = 1,  // out bf16 = gelu(acc + bias)                 (ViT fc1)
    EK_BF16_RELU = 2,  // out bf16 = relu(acc + bias)                 (FFNs, head convs)
    EK_F32_RES = 3,    // out fp32 = acc + bias + residual fp32       (ViT proj / fc2 on the residual stream)
    EK_GENERIC = 4,    // everything behind run-time flags            (tables, pixel shuffle, bf16 residual ...)
    EK_HEAD = 5,       // EPI_HEAD_FINAL                              (P2CL logits, 1-CTA BN=64 only)
    EK_BF16_TAB = 6,   // out bf16 = acc + table[m % rows]            (DMA image-side K|V|Q projection with the folded key_pe)
    EK_PS = 7,         // out bf16 = acc + bias, pixel-shuffle store  (ConvTranspose2d k=2 s=2 of the neck)
    // GroupNorm-fused neck variants (Epi::gn_*): the epilogue accumulates the per-sample sum / sum of squares of its
    // fp32 outputs (so no statistics pass ever reads the tensor back), and EK_GNIN applies the producer's GroupNorm
    // algebraically (no apply pass for the GroupNorms that are not followed by GELU)
    EK_PS_ST = 8,      // EK_PS + statistics
    EK_BF16_ST = 9,    // EK_BF16 + statistics
    EK_GNIN_ST = 10,   // out bf16 = rstd*acc - mean*rstd*wg + bias, + statistics
    EK_F32_RESBF = 11, // out fp32 = acc + bias + residual bf16       (DMA image<-token out_proj on the bf16 keys)
    // LayerNorm-fused ViT variants (Epi::ln_*)
    EK_F32_RES_LNOUT = 12,    // EK_F32_RES + bf16 copy + per-row (sum, sum of squares) slots   (ViT proj / fc2)
    EK_BF16_LNIN = 13,        // out bf16 = rstd*acc - rstd*mean*s + bias                        (ViT qkv on the un-normalised tokens)
    EK_BF16_GELU_LNIN = 14    // gelu of the same                                                (ViT fc1)
};

// measurement-only ablation (VPU_GEMM_ABLATE) is compiled in with -DVPU_GEMM_DEBUG: the check sat in the MMA issue loop
#ifdef VPU_GEMM_DEBUG
#define GEMM_ABLATE(bit) (d.ablate & (bit))
#else
#define GEMM_ABLATE(bit) 0
#endif

constexpr int EPI_PITCH = 36;                        // fp32 words per staged row
constexpr int EPI_WARP_WORDS = 32 * EPI_PITCH;       // 4608 B per epilogue warp
constexpr int EPI_SMEM_BYTES = 8 * EPI_WARP_WORDS * 4;

template <int BN, int EK>
__device__ __forceinline__ void epilogue_tile(const Epi& e, const GemmDims& d, uint32_t tmem_acc, int row_base, int col_base,
                                              int quarter, int ew, int EW, int lane, float* sbuf) {   // warp ew of EW in this lane quarter
    const int r_lo = row_base + quarter * 32;
    if constexpr (EK == EK_HEAD) {   // column-major NCHW stores: lane == row is already the coalesced mapping
        const int row = r_lo + lane;
#pragma unroll 1
        for (int c = ew * 32; c < BN; c += EW * 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + c, r);
            tmem_ld_wait();
            const int n0 = col_base + c;
            if (row < d.M && n0 < d.N) {
                float v[32];
#pragma unroll
                for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
                epi_store<32>(e, row, n0, v, d.N);
            }
        }
    } else {
        constexpr bool DYN = EK == EK_GENERIC;
        constexpr bool LNOUT = EK == EK_F32_RES_LNOUT;
        constexpr bool LNIN = EK == EK_BF16_LNIN || EK == EK_BF16_GELU_LNIN;
        const bool has_res = DYN ? (e.res != nullptr) : (EK == EK_F32_RES || EK == EK_F32_RESBF || LNOUT);
        const bool res_bf16 = DYN ? (e.res_bf16 != 0) : (EK == EK_F32_RESBF);
        const bool out_bf16 = DYN ? (e.out_bf16 != 0) : (EK != EK_F32_RES && EK != EK_F32_RESBF && !LNOUT);
        const int act = DYN ? e.act : ((EK == EK_BF16_GELU || EK == EK_BF16_GELU_LNIN) ? ACT_GELU : (EK == EK_BF16_RELU ? ACT_RELU : ACT_NONE));
        const bool has_tab = DYN ? (e.bias2d != nullptr) : (EK == EK_BF16_TAB);
        const bool pshuf = DYN ? (e.mode == EPI_PIXEL_SHUFFLE) : (EK == EK_PS || EK == EK_PS_ST);
        const bool stats = DYN ? (e.gn_out != nullptr) : (EK == EK_PS_ST || EK == EK_BF16_ST || EK == EK_GNIN_ST);
        const bool gnin = DYN ? (e.gn_in != nullptr) : (EK == EK_GNIN_ST);
        // a warp's 32 rows touch at most two samples (gn_rows >= 32): A = sample of the first row, B = the next one
        long long st_s0 = 0, st_q0 = 0, st_s1 = 0, st_q1 = 0;      // fixed point: see GN_SUM_SCALE
        float rA = 1.f, mrA = 0.f, rB = 1.f, mrB = 0.f;
        int mB = 0x7fffffff, sA = 0;
        bool hasB = false;
        if ((stats || gnin) && r_lo < d.M) {
            sA = r_lo / e.gn_rows;
            mB = (sA + 1) * e.gn_rows;
            hasB = mB < d.M && mB < r_lo + 32;
            if (gnin) {
                gn_mean_rstd(e.gn_in, sA, e.gn_in_count, rA, mrA);
                if (hasB) gn_mean_rstd(e.gn_in, sA + 1, e.gn_in_count, rB, mrB);
            }
        }
        const int sub = lane >> 3, j4 = (lane & 7) * 4;
        // LayerNorm fusion: this lane's eight rows are r_lo + sub + 4*it
        float ln_r[8], ln_mr[8], ln_sum[8], ln_sq[8];
        if constexpr (LNIN) {      // (rstd, mean * rstd) per row, finalised by ln_rowstats_kernel
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int m = r_lo + 4 * it + sub;
                const float2 t = m < d.M ? __ldg(e.ln_in + m) : make_float2(0.f, 0.f);
                ln_r[it] = t.x;
                ln_mr[it] = t.y;
            }
        }
        if constexpr (LNOUT) {
#pragma unroll
            for (int it = 0; it < 8; ++it) { ln_sum[it] = 0.f; ln_sq[it] = 0.f; }
        }
        // rows of this lane are r_lo + sub + 4*it: the table row and the pixel-shuffle coordinates are divided out
        // once per tile and advanced by 4 per step (run-time divisions per row dominated these epilogues before)
        int trow0 = 0;
        if (has_tab) trow0 = (r_lo + sub) % e.bias2d_rows;
        // pixel shuffle: input row m = (b, i, j) -> output row (b, 2i + kh, 2j + kw) = ps_base + kh * 2g + kw.  The row part is
        // worked out once per tile for the lane's eight rows (they advance by 4 with at most g / 4 carries); the (kh, kw, channel)
        // part is uniform over a 32-column chunk (cout % 32 == 0).  Doing both per row and per chunk -- divergent carry loops and
        // an integer division inside the store loop -- cost 73 / 167 us on the neck's ConvTranspose GEMMs (tools/gemm_ps_ab.py).
        int ps_base[8];
        if (pshuf) {
            const int g = e.ps_g, m0 = r_lo + sub, gg = g * g, ij = m0 % gg;
            int b = m0 / gg, i = ij / g, jx = ij % g;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                ps_base[it] = (b * 2 * g + 2 * i) * (2 * g) + 2 * jx;
                jx += 4;
                while (jx >= g) { jx -= g; ++i; }      // g >= 8 on the compile-time paths: at most one carry per step
                while (i >= g) { i -= g; ++b; }
            }
        }
#pragma unroll 1
        for (int c = ew * 32; c < BN; c += EW * 32) {
            const int n0 = col_base + c;
            if (n0 >= d.N) break;                     // N is a multiple of 32: chunks are all-valid or all-outside
            uint32_t r[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + c, r);
            const int n = n0 + j4;
            int ps_qoff = 0, ps_ocol = 0;
            if (pshuf) {
                const int q = n0 / e.ps_cout;                  // warp-uniform: a chunk never straddles two (kh, kw) blocks
                ps_qoff = (q >> 1) * 2 * e.ps_g + (q & 1);
                ps_ocol = n - q * e.ps_cout;
            }
            // operands that do not depend on the accumulator are requested while the TMEM load is in flight
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), wg = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.bias) bv = __ldg(reinterpret_cast<const float4*>(e.bias + n));
            if (gnin) wg = __ldg(reinterpret_cast<const float4*>(e.gn_wg + n));
            if constexpr (LNIN) wg = __ldg(reinterpret_cast<const float4*>(e.ln_s + n));
            float4 res[8];
            if (has_res) {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int m = r_lo + 4 * it + sub;
                    res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (m < d.M) {
                        if (res_bf16) {
                            const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(e.res) + (size_t)m * e.ldr + n);
                            const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
                            res[it] = make_float4(a.x, a.y, b.x, b.y);
                        } else {
                            res[it] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.res) + (size_t)m * e.ldr + n);
                        }
                    }
                }
            }
            float4 tabv[8];
            if (has_tab) {                           // like the residual: all eight rows in flight before the accumulator is needed
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    int trow = trow0 + 4 * it;                           // table rows >= 32 on this path: at most one wrap
                    while (trow >= e.bias2d_rows) trow -= e.bias2d_rows;
                    tabv[it] = __ldg(reinterpret_cast<const float4*>(e.bias2d + (size_t)trow * d.N + n));
                }
            }
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<uint4*>(sbuf + lane * EPI_PITCH + 4 * q) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = 4 * it + sub, m = r_lo + rr;
                float4 v = *reinterpret_cast<const float4*>(sbuf + rr * EPI_PITCH + j4);
                if (m < d.M) {
                    if (gnin) {
                        const float r = m >= mB ? rB : rA, mr = m >= mB ? mrB : mrA;
                        v.x = fmaf(v.x, r, -mr * wg.x); v.y = fmaf(v.y, r, -mr * wg.y);
                        v.z = fmaf(v.z, r, -mr * wg.z); v.w = fmaf(v.w, r, -mr * wg.w);
                    }
                    if constexpr (LNIN) {
                        const float r = ln_r[it], mr = ln_mr[it];
                        v.x = fmaf(v.x, r, -mr * wg.x); v.y = fmaf(v.y, r, -mr * wg.y);
                        v.z = fmaf(v.z, r, -mr * wg.z); v.w = fmaf(v.w, r, -mr * wg.w);
                    }
                    v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                    if (stats) {
                        // one row x four fixed columns: this fp32 partial does not depend on where the sample sits in the batch
                        const long long sm = __float2ll_rn(((v.x + v.y) + (v.z + v.w)) * GN_SUM_SCALE);
                        const long long sq = __float2ll_rn(fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w))) * GN_SQ_SCALE);
                        if (m >= mB) { st_s1 += sm; st_q1 += sq; } else { st_s0 += sm; st_q0 += sq; }
                    }
                    if (has_res) { v.x += res[it].x; v.y += res[it].y; v.z += res[it].z; v.w += res[it].w; }
                    if (has_tab) { v.x += tabv[it].x; v.y += tabv[it].y; v.z += tabv[it].z; v.w += tabv[it].w; }
                    if (act == ACT_GELU) { v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w); }
                    else if (act == ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    size_t orow = (size_t)m;
                    int ocol = n;
                    if (pshuf) { orow = (size_t)(ps_base[it] + ps_qoff); ocol = ps_ocol; }
                    if constexpr (LNOUT) {
                        ln_sum[it] += (v.x + v.y) + (v.z + v.w);
                        ln_sq[it] += fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
                        *reinterpret_cast<uint2*>(e.ln_out_bf16 + orow * e.ldo + ocol) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
                    }
                    if (out_bf16)
                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.out) + orow * e.ldo + ocol) =
                            make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
                    else
                        *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + orow * e.ldo + ocol) = v;
                }
            }
            __syncwarp();
            if constexpr (LNOUT) {
                // One slot per (128-column group of the row, epilogue warp), whatever the tile width: a warp's chunks are
                // ew * 32 + 64 k, so its partial over a 128-column group is the same two chunks on 256- and 128-wide tiles and
                // the statistics do not depend on the tile choice (which depends on M: see ln_tile_bn).  The 8 lanes that share
                // a row (lane & 7) fold their partials in a fixed butterfly.
                const int cn = c + EW * 32;
                if (cn >= BN || col_base + cn >= d.N || (cn >> 7) != (c >> 7)) {
                    const int slot = ((col_base + c) >> 7) * EW + ew;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        float sm = ln_sum[it], sq = ln_sq[it];
#pragma unroll
                        for (int o = 4; o > 0; o >>= 1) { sm += __shfl_xor_sync(0xffffffffu, sm, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
                        const int m = r_lo + 4 * it + sub;
                        if ((lane & 7) == 0 && m < d.M) e.ln_out[(size_t)m * e.ln_slots + slot] = make_float2(sm, sq);
                        ln_sum[it] = 0.f;
                        ln_sq[it] = 0.f;
                    }
                }
            }
        }
        if (stats) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                st_s0 += __shfl_xor_sync(0xffffffffu, st_s0, o); st_q0 += __shfl_xor_sync(0xffffffffu, st_q0, o);
                st_s1 += __shfl_xor_sync(0xffffffffu, st_s1, o); st_q1 += __shfl_xor_sync(0xffffffffu, st_q1, o);
            }
            if (lane == 0 && r_lo < d.M) {
                unsigned long long* acc = reinterpret_cast<unsigned long long*>(e.gn_out) + 2 * sA;
                atomicAdd(acc, (unsigned long long)st_s0);
                atomicAdd(acc + 1, (unsigned long long)st_q0);
                if (hasB) {
                    atomicAdd(acc + 2, (unsigned long long)st_s1);
                    atomicAdd(acc + 3, (unsigned long long)st_q1);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------
constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 320;

template <int BN> struct TileCfg {
    static constexpr int STAGES = BN == 256 ? 3 : (BN == 192 ? 4 : (BN == 128 ? 5 : 7));
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int ACC_STRIDE = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_SMEM_BYTES + 1024;  // + alignment slack
};

template <int BN, int EK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDims d,
               const Epi e) {
    pdl_launch_dependents();
    using C = TileCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[C::STAGES], empty_bar[C::STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 8);  // one elected lane per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();     // everything above overlapped the tail of the previous kernel; global memory is touched only below

    const int n_blks = (d.N + BN - 1) / BN;
    const int m_blks = (d.M + BM - 1) / BM;
    const int tiles = n_blks * m_blks;
    const int kblks = (d.K + BK - 1) / BK;

    if (warp == 0) {
        // ---------------- TMA producer (converged warp, one elected lane issues; see gemm_tc2_kernel) ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int m_blk = tile / n_blks, n_blk = tile % n_blks;
            int brow = n_blk * BN;
            if (e.m_per_batch > 0) brow += ((m_blk * BM) / e.m_per_batch) * e.b_rows_per_batch;
            for (int kb = 0; kb < kblks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    tma_load_2d(sa + C::A_BYTES, &tmB, &full_bar[stage], kb * BK, brow);
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (converged warp, one elected lane issues) ----------------
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
        const uint64_t adesc0 = umma_desc_k_sw128(smem_base), bdesc0 = umma_desc_k_sw128(smem_base + C::A_BYTES);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * C::ACC_STRIDE;
            for (int kb = 0; kb < kblks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t soff = (uint64_t)(stage * (C::STAGE_BYTES >> 4));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(d_tmem, adesc0 + soff + 2 * k, bdesc0 + soff + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (kb + 1 == kblks) umma_commit(&tfull_bar[acc]);        // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {  // ---------------- epilogue warps ----------------
        const int quarter = warp & 3;            // TMEM lanes [32*quarter, +32) are this warp's
        const int ew = (warp - 2) >> 2;          // 32-column chunks ew, ew + 2, ... of the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int m_blk = tile / n_blks, n_blk = tile % n_blks;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            epilogue_tile<BN, EK>(e, d, tmem_base + acc * C::ACC_STRIDE, m_blk * BM, n_blk * BN, quarter, ew, 2, lane,
                              reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES) + (warp - 2) * EPI_WARP_WORDS);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}


// ------------------------------------------------------------------------------------------
// 2-CTA tcgen05 kernel: a CTA pair (cluster of 2 on one TPC) computes a 256 x BN tile with
// tcgen05.mma.cta_group::2 (M = 256).  Each CTA loads its own 128 rows of A and HALF of the B tile,
// so the L2 -> shared-memory traffic per MMA-flop drops to 2/3 of the 1-CTA kernel (the 1-CTA 128x256
// tile needs ~96 B/clk/SM, above the measured ~6.3 KB/clk chip-wide L2 cap; B300_MICROARCH.md "L2").
// Roles per CTA as above; only the leader's warp 1 issues MMAs.  Barriers:
//   full[s]   leader only, tx = both CTAs' loads (peer TMA signals the leader's barrier)
//   empty[s]  one per CTA, released by the leader's multicast tcgen05.commit
//   tfull[a]  one per CTA (multicast commit);  tempty[a] leader only, 16 arrivals (8 epilogue warps x 2 CTAs)
// ------------------------------------------------------------------------------------------
template <int BN> struct TileCfg2 {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (BN / 2) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = BN == 256 ? 5 : 7;
    static constexpr int ACC_STRIDE = BN;   // 256 or 128 columns per accumulator buffer
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_SMEM_BYTES + 1024;
};
// Epilogue warps per TMEM lane quarter.  Two suffice when the epilogue is a handful of instructions per element; the GELU
// epilogue of fc1 (9 instructions + 2 MUFU per element, 32768 elements per CTA tile) ran 9800 clk per tile on 8 warps --
// latency-bound with two warps per scheduler -- against 7000 clk of MMA work, so it gets four per quarter (and, for the
// shared-memory budget of their staging buffers, one pipeline stage less).
// The same holds for the K = 384 out_proj of the DMA image<-token attention (6 k-blocks of MMA per 128 KB of fp32 output per CTA:
// 94 -> 60 us with four).  Four warps per quarter on the neck / table epilogues measured no gain on the whole step.
#ifndef VPU_EPI_VARIANT
#define VPU_EPI_VARIANT 0
#endif
template <int EK> struct EpiWarps {
    static constexpr int N = (EK == EK_BF16_GELU || EK == EK_BF16_GELU_LNIN || EK == EK_F32_RESBF) ? 4 :
                             ((VPU_EPI_VARIANT >= 1 && EK == EK_F32_RES_LNOUT) || (VPU_EPI_VARIANT >= 2 && EK == EK_F32_RES) ||
                              (VPU_EPI_VARIANT >= 3 && EK == EK_BF16_LNIN)) ? 3 : 2;
};
template <int EK> __host__ __device__ constexpr int tc2_threads() { return (2 + 4 * EpiWarps<EK>::N) * 32; }
template <int BN, int EK> __host__ __device__ constexpr int tc2_stages() { return TileCfg2<BN>::STAGES - (EpiWarps<EK>::N > 3 ? 1 : 0); }   // 3 per quarter still fit 5 stages (220 KB)
template <int BN, int EK> __host__ __device__ constexpr int tc2_smem() { return tc2_stages<BN, EK>() * TileCfg2<BN>::STAGE_BYTES + 4 * EpiWarps<EK>::N * EPI_WARP_WORDS * 4 + 1024; }

template <int BN, int EK, int CL>   // CL = CTAs per cluster: 2 (one MMA pair) or 4 (two pairs sharing the B tile by TMA multicast)
__global__ void __launch_bounds__(tc2_threads<EK>(), 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDims d,
                const Epi e) {
    pdl_launch_dependents();
    using C = TileCfg2<BN>;
    constexpr int STAGES = tc2_stages<BN, EK>();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PP = CL / 2;                        // MMA pairs per cluster (stacked along M, same N tile)
    const uint32_t crank = cluster_ctarank();
    const int rank = (int)(crank & 1);                // position inside the MMA pair
    const int pr = (int)(crank >> 1);                 // pair inside the cluster
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PP);   // one multicast commit per pair that reads (and whose peers write) this stage
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 8 * EpiWarps<EK>::N);    // one elected lane per epilogue warp of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_2sm(&tmem_base_smem, C::TMEM_COLS);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();     // barriers of both CTAs initialised before any remote arrive / TMA signal
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();     // everything above overlapped the tail of the previous kernel; global memory is touched only below

    const int n_blks = (d.N + BN - 1) / BN;
    const int m_blks = (d.M + 2 * BM * PP - 1) / (2 * BM * PP);     // cluster tiles along M: PP * 256 rows each
    const int tiles = n_blks * m_blks;
    const int kblks = (d.K + BK - 1) / BK;
    const int pair = blockIdx.x / CL, npairs = gridDim.x / CL;      // cluster index / number of clusters
    constexpr uint16_t kAllMask = (uint16_t)((1u << CL) - 1);
    const uint16_t pair_mask = (uint16_t)(3u << (2 * pr));
    const int nst = d.stages > 0 && d.stages < STAGES ? d.stages : STAGES;

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs) ----------------
        // The producer and the MMA issuer run as whole converged warps with one ELECTed lane doing the issue: inside an
        // `if (lane == 0)` region nvcc wraps every tcgen05.mma / commit in an elect + BRA.U.ANY loop and rebuilds both
        // 64-bit descriptors per instruction (~120 SASS instructions per k-block; a lone warp retires a dependent
        // instruction every ~5-8 clk, so issuing one k-block took longer than its 512 clk of MMA work and capped the tensor
        // pipe at 67 %, ncu round 1c).  Descriptors are built once and advanced by adding to their address field.
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            const int m_blk = tile / n_blks, n_blk = tile % n_blks;
            const int arow = (m_blk * PP + pr) * 2 * BM + rank * BM;
            // this CTA fetches 1/PP of the B half its pair position needs and multicasts it to the CTAs at the same
            // position in every pair of the cluster
            const int brow = n_blk * BN + rank * (BN / 2) + pr * (BN / 2 / PP);
            const uint16_t bmask = (uint16_t)(PP == 1 ? 0 : ((1u << rank) | (1u << (2 + rank))));
            for (int kb = 0; kb < kblks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (GEMM_ABLATE(1)) {
                        if (leader) mbar_arrive(&full_bar[stage]);
                    } else {
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                        uint8_t* sa = smem + stage * C::STAGE_BYTES;
                        tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BK, arow);
                        if constexpr (PP == 1)
                            tma_load_2d_2sm(sa + C::A_BYTES, &tmB, &full_bar[stage], kb * BK, brow);
                        else
                            tma_load_2d_2sm_mc(sa + C::A_BYTES + pr * (C::B_BYTES / PP), &tmB, &full_bar[stage], kb * BK, brow, bmask);
                    }
                }
                __syncwarp();
                if (++stage == nst) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {  // ---------------- MMA issuer (leader CTA only; converged warp, one elected lane) ----------------
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
            const uint64_t adesc0 = umma_desc_k_sw128(smem_base), bdesc0 = umma_desc_k_sw128(smem_base + C::A_BYTES);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < tiles; tile += npairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C::ACC_STRIDE;
                for (int kb = 0; kb < kblks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t soff = (uint64_t)(stage * (C::STAGE_BYTES >> 4));      // descriptor address field: bytes >> 4
                        if (!GEMM_ABLATE(2)) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)                                 // +32 B along K per UMMA
                                umma_bf16_2sm(d_tmem, adesc0 + soff + 2 * k, bdesc0 + soff + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_2sm(&empty_bar[stage], kAllMask);   // this pair is done with the stage: tell every CTA that writes into it
                        if (kb + 1 == kblks) umma_commit_2sm(&tfull_bar[acc], pair_mask);   // accumulator halves complete in both CTAs of this pair
                    }
                    __syncwarp();
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {  // ---------------- epilogue warps (both CTAs: own 128 rows, all BN columns) ----------------
        const int quarter = warp & 3;
        const int ew = (warp - 2) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            const int m_blk = tile / n_blks, n_blk = tile % n_blks;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            epilogue_tile<BN, EK>(e, d, tmem_base + acc * C::ACC_STRIDE, (m_blk * PP + pr) * 2 * BM + rank * BM, n_blk * BN, quarter, ew, EpiWarps<EK>::N,
                              lane, reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES) + (warp - 2) * EPI_WARP_WORDS);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    cluster_sync_all();     // the peer's shared memory / barriers stay alive until every MMA and commit has landed
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    }
}

#ifdef VPU_DEBUG
// ------------------------------------------------------------------------------------------
// mma.sync cross-check kernel (-DVPU_DEBUG builds only: impl 1 of vpu_gemm / VPU_GEMM_IMPL=1; not part of the shipped library)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gemm_mma_kernel(const __nv_bfloat16* __restrict__ A,
                                                       const __nv_bfloat16* __restrict__ W, int lda, int ldw,
                                                       const GemmDims d, const Epi e) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) __nv_bfloat16 As[64][40];
    __shared__ __align__(16) __nv_bfloat16 Bs[64][40];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    size_t wrow0 = 0;
    if (e.m_per_batch > 0) wrow0 = (size_t)(m0 / e.m_per_batch) * e.b_rows_per_batch;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < d.K; k0 += 32) {
        for (int idx = threadIdx.x; idx < 64 * 32; idx += 128) {
            const int r = idx >> 5, c = idx & 31;
            const int k = k0 + c;
            As[r][c] = (m0 + r < d.M && k < d.K) ? A[(size_t)(m0 + r) * lda + k] : __float2bfloat16(0.f);
            Bs[r][c] = (n0 + r < d.N && k < d.K) ? W[(wrow0 + n0 + r) * ldw + k] : __float2bfloat16(0.f);
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < 32; ks += 16) {
            uint32_t a0 = *reinterpret_cast<uint32_t*>(&As[warp * 16 + g][ks + 2 * t]);
            uint32_t a1 = *reinterpret_cast<uint32_t*>(&As[warp * 16 + g + 8][ks + 2 * t]);
            uint32_t a2 = *reinterpret_cast<uint32_t*>(&As[warp * 16 + g][ks + 2 * t + 8]);
            uint32_t a3 = *reinterpret_cast<uint32_t*>(&As[warp * 16 + g + 8][ks + 2 * t + 8]);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                uint32_t b0 = *reinterpret_cast<uint32_t*>(&Bs[nt * 8 + g][ks + 2 * t]);
                uint32_t b1 = *reinterpret_cast<uint32_t*>(&Bs[nt * 8 + g][ks + 2 * t + 8]);
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                    "{%0,%1,%2,%3};"
                    : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int n = n0 + nt * 8 + 2 * t;
        if (n >= d.N) continue;
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int m = m0 + warp * 16 + g + hrow * 8;
            if (m >= d.M) continue;
            float v[2] = {acc[nt][hrow * 2], acc[nt][hrow * 2 + 1]};
            epi_store<2>(e, m, n, v, d.N);
        }
    }
}
#endif  // VPU_DEBUG

// Tile width of the LayerNorm-fused GEMMs: 256, or 128 when 256-wide tiles would occupy at most a quarter of the SMs (a batch-2
// forward: proj / fc2 with 21 pair tiles for 74 CTA pairs).  The statistics slots are per 128-column group (epilogue_tile), so the
// fp32 grouping of the row sums -- and with it every bit of the forward -- does not depend on this choice: a sample alone, first or in
// the middle of a batch gives the same bits (the lock-step NoC loop reproduces the serial one bit for bit because of that).
static int g_num_sms_for_tiles();
static int ln_tile_bn(int M, int N) {
    const long long tiles256 = (long long)((M + 2 * BM - 1) / (2 * BM)) * ((N + 255) / 256);
    return (tiles256 * 4 <= g_num_sms_for_tiles() && N % 128 == 0) ? 128 : 256;
}
int gemm_ln_slots(int, int N) { return ((N + 127) / 128) * EpiWarps<EK_F32_RES_LNOUT>::N; }
int gemm_ln_slots_max(int N) { return ((N + 127) / 128) * EpiWarps<EK_F32_RES_LNOUT>::N; }

// Row statistics of the LayerNorm fusion: the slots a residual GEMM wrote (Epi::ln_out) -> (rstd, mean * rstd) per row, added in
// slot order in double precision (one thread per row; 2.4 MB in, 0.4 MB out for ViT-B at batch 64; eight lanes per row with
// coalesced slot reads measured the same 8-9 us under the event brackets: the launch is latency, not traffic).
__global__ void __launch_bounds__(256) ln_rowstats_kernel(const float2* __restrict__ slots, int M, int P, float inv_c, float eps,
                                                          float2* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double S = 0.0, Q = 0.0;
    for (int k = 0; k < P; ++k) {
        const float2 t = __ldg(slots + (size_t)m * P + k);
        S += (double)t.x; Q += (double)t.y;
    }
    const double mean = S * (double)inv_c;
    const double var = fmax(Q * (double)inv_c - mean * mean, 0.0);
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    out[m] = make_float2(rstd, (float)mean * rstd);
}

int ln_rowstats_launch(const float2* slots, int M, int P, int C, float eps, float2* out, cudaStream_t stream) {
    VPU_REQUIRE(slots && out && M > 0 && P > 0 && C > 0, "ln_rowstats: bad argument");
    VPU_CHECK_CUDA(launch_pdl(ln_rowstats_kernel, dim3((M + 255) / 256), dim3(256), 0, stream, slots, M, P, 1.0f / (float)C, eps, out));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;
static bool g_small_tiles = true;      // VPU_GEMM_SMALL_TILES=0 (-DVPU_DEBUG builds): the round-1 tile choice for small problems
static bool g_use_2cta = true;
static int g_stages = 0;
static int g_cluster = 2;
static int g_ablate = 0;
static bool g_ragged256 = true;
static int g_gn_tma = 3;                // VPU_GEMM_GN_TMA (-DVPU_DEBUG builds): bit 0 GroupNorm-fused neck GEMMs, bit 1 pixel-shuffle GEMMs on gemm_gn.cu
static int g_res_modes = 1;             // VPU_GEMM_RES_MODES (-DVPU_DEBUG builds): 0 keeps the table GEMMs on the generic epilogue
static int g_res_kmax = 1 << 30;       // VPU_GEMM_RES_KMAX (-DVPU_DEBUG builds): largest K that takes gemm_res.cu (0 = never)
static std::mutex g_mu;

static int g_num_sms_for_tiles() { return g_small_tiles ? (g_num_sms > 0 ? g_num_sms : 148) : 0; }

struct TmKey {
    const void* p;
    uint64_t rows, cols, ld;
    uint32_t box_rows;
    bool operator==(const TmKey& o) const {
        return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct TmKeyHash {
    size_t operator()(const TmKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.p);
        h = h * 1000003u ^ k.rows; h = h * 1000003u ^ k.cols; h = h * 1000003u ^ k.ld; h = h * 1000003u ^ k.box_rows;
        return h;
    }
};
static std::unordered_map<TmKey, CUtensorMap, TmKeyHash> g_tm_cache;

template <int BN, int EK> static int attr1() {
    VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EK>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<BN>::SMEM_BYTES));
    return 0;
}
template <int BN, int EK> static int attr2() {
    VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN, EK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2_smem<BN, EK>()));
    VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN, EK, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2_smem<BN, EK>()));
    return 0;
}
template <int BN> static int attrs2_all() {
    if (int rc = attr2<BN, EK_BF16>()) return rc;
    if (int rc = attr2<BN, EK_BF16_GELU>()) return rc;
    if (int rc = attr2<BN, EK_BF16_RELU>()) return rc;
    if (int rc = attr2<BN, EK_F32_RES>()) return rc;
    if (int rc = attr2<BN, EK_BF16_TAB>()) return rc;
    if (int rc = attr2<BN, EK_PS>()) return rc;
    if (int rc = attr2<BN, EK_PS_ST>()) return rc;
    if (int rc = attr2<BN, EK_BF16_ST>()) return rc;
    if (int rc = attr2<BN, EK_GNIN_ST>()) return rc;
    if (int rc = attr2<BN, EK_F32_RESBF>()) return rc;
    if (int rc = attr2<BN, EK_F32_RES_LNOUT>()) return rc;
    if (int rc = attr2<BN, EK_BF16_LNIN>()) return rc;
    if (int rc = attr2<BN, EK_BF16_GELU_LNIN>()) return rc;
    return attr2<BN, EK_GENERIC>();
}
static int set_smem_attrs() {
    if (int rc = attr1<64, EK_HEAD>()) return rc;
    if (int rc = attr1<64, EK_GENERIC>()) return rc;
    if (int rc = attr1<128, EK_GENERIC>()) return rc;
    if (int rc = attr1<192, EK_GENERIC>()) return rc;
    if (int rc = attr1<256, EK_GENERIC>()) return rc;
    if (int rc = attrs2_all<256>()) return rc;
    return attrs2_all<128>();
}

int gemm_init() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VPU_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    VPU_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    int dev = 0;
    VPU_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    VPU_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    VPU_REQUIRE(prop.major == 10, "pvpuformer_b200 needs an sm_100a device (got sm_%d%d): no fallback path exists",
                prop.major, prop.minor);
    g_num_sms = prop.multiProcessorCount;
    if (int rc = set_smem_attrs()) return rc;
    const char* two = vpu_debug_env("VPU_GEMM_2CTA");
    g_use_2cta = !(two && two[0] == '0');
    if (const char* st = vpu_debug_env("VPU_GEMM_STAGES")) g_stages = atoi(st);
    if (const char* cl = vpu_debug_env("VPU_GEMM_CLUSTER")) g_cluster = atoi(cl) == 4 ? 4 : 2;
    if (const char* ab = vpu_debug_env("VPU_GEMM_ABLATE")) g_ablate = atoi(ab);
    if (const char* rg = vpu_debug_env("VPU_GEMM_RAGGED256")) g_ragged256 = rg[0] != '0';
    if (const char* rk = vpu_debug_env("VPU_GEMM_RES_KMAX")) g_res_kmax = atoi(rk);
    if (const char* gt = vpu_debug_env("VPU_GEMM_GN_TMA")) g_gn_tma = atoi(gt);
    if (const char* rm = vpu_debug_env("VPU_GEMM_RES_MODES")) g_res_modes = atoi(rm);
    if (const char* sm = vpu_debug_env("VPU_GEMM_SMALL_TILES")) g_small_tiles = sm[0] != '0';
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// bf16 [rows, cols] row-major with leading dimension ld (elements); box = [box_rows, 64 cols].
static int make_tmap(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    TmKey key{ptr, rows, cols, ld, box_rows};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_tm_cache.find(key);
        if (it != g_tm_cache.end()) { *tm = it->second; return 0; }
    }
    VPU_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand pointer must be 16-byte aligned");
    VPU_REQUIRE((ld * 2) % 16 == 0, "GEMM operand leading dimension (%llu) must be a multiple of 8 elements",
                (unsigned long long)ld);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPU_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows=%llu cols=%llu ld=%llu box_rows=%u)", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_tm_cache.size() > 4096) g_tm_cache.clear();
    g_tm_cache[key] = *tm;
    return 0;
}

int gemm_tmap(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    if (int rc = gemm_init()) return rc;
    return make_tmap(tm, ptr, rows, cols, ld, box_rows);
}
int gemm_num_sms() { return g_num_sms; }

template <int BN, int EK = EK_GENERIC>
static int launch_tc(const GemmProblem& p, cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap(&tmA, p.A, p.M, p.K, p.lda, BM)) return rc;
    if (int rc = make_tmap(&tmB, p.W, p.w_rows ? p.w_rows : p.N, p.K, p.ldw, BN)) return rc;
    const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;
    GemmDims d{p.M, p.N, p.K, g_stages, g_ablate};
    VPU_CHECK_CUDA(launch_pdl(gemm_tc_kernel<BN, EK>, dim3(grid), dim3(GEMM_THREADS), TileCfg<BN>::SMEM_BYTES, stream, tmA, tmB, d, p.epi));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

template <int BN, int EK, int CL>
static int launch_tc2_cl(const GemmProblem& p, cudaStream_t stream) {
    constexpr int PP = CL / 2;
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap(&tmA, p.A, p.M, p.K, p.lda, BM)) return rc;
    if (int rc = make_tmap(&tmB, p.W, p.w_rows ? p.w_rows : p.N, p.K, p.ldw, BN / 2 / PP)) return rc;
    const int tiles = ((p.M + 2 * BM * PP - 1) / (2 * BM * PP)) * ((p.N + BN - 1) / BN);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    cfg.blockDim = dim3(tc2_threads<EK>()); cfg.dynamicSmemBytes = tc2_smem<BN, EK>(); cfg.stream = stream;
    static int max_clusters = 0;      // co-resident clusters of this instantiation (a persistent grid must not exceed it by much)
    if (max_clusters == 0) {
        cfg.gridDim = dim3(CL * (g_num_sms / CL));
        int n = 0;
        VPU_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_tc2_kernel<BN, EK, CL>, &cfg));
        max_clusters = n > 0 ? n : 1;
    }
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    cfg.gridDim = dim3(CL * clusters);
    GemmDims d{p.M, p.N, p.K, g_stages, g_ablate};
    VPU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<BN, EK, CL>, tmA, tmB, d, p.epi));
    count_launch();
    return 0;
}

template <int BN, int EK>
static int launch_tc2_k(const GemmProblem& p, cudaStream_t stream) {
    // clusters of 4 (B-tile multicast between two MMA pairs) need at least two 256-row blocks along M
    if (g_cluster == 4 && p.M >= 4 * BM) return launch_tc2_cl<BN, EK, 4>(p, stream);
    return launch_tc2_cl<BN, EK, 2>(p, stream);
}

// pick the compile-time epilogue the problem's run-time flags describe
static int epi_kind(const Epi& e) {
    if (e.ln_out) return EK_F32_RES_LNOUT;                                   // shapes / flags validated in gemm_launch
    if (e.ln_in) return e.act == ACT_GELU ? EK_BF16_GELU_LNIN : EK_BF16_LNIN;
    if (e.gn_out || e.gn_in) {
        const bool simple = e.out_bf16 && !e.res && !e.bias2d && e.act == ACT_NONE && e.bias && e.gn_out;
        if (simple && e.mode == EPI_PIXEL_SHUFFLE && !e.gn_in && e.ps_g >= 8) return EK_PS_ST;
        if (simple && e.mode == EPI_PLAIN) return e.gn_in ? EK_GNIN_ST : EK_BF16_ST;
        return EK_GENERIC;
    }
    if (e.mode == EPI_PIXEL_SHUFFLE && e.out_bf16 && !e.res && !e.bias2d && e.act == ACT_NONE && e.ps_g >= 8) return EK_PS;
    if (e.mode == EPI_PLAIN && e.bias2d && e.bias2d_rows >= 32 && e.out_bf16 && !e.res && e.act == ACT_NONE) return EK_BF16_TAB;
    if (e.mode != EPI_PLAIN || e.bias2d) return EK_GENERIC;
    if (e.out_bf16 && !e.res) return e.act == ACT_GELU ? EK_BF16_GELU : (e.act == ACT_RELU ? EK_BF16_RELU : EK_BF16);
    if (!e.out_bf16 && e.res && e.act == ACT_NONE) return e.res_bf16 ? EK_F32_RESBF : EK_F32_RES;
    return EK_GENERIC;
}

template <int BN>
static int launch_tc2(const GemmProblem& p, cudaStream_t stream) {
    switch (epi_kind(p.epi)) {
        case EK_BF16: return launch_tc2_k<BN, EK_BF16>(p, stream);
        case EK_BF16_GELU: return launch_tc2_k<BN, EK_BF16_GELU>(p, stream);
        case EK_BF16_RELU: return launch_tc2_k<BN, EK_BF16_RELU>(p, stream);
        case EK_F32_RES: return launch_tc2_k<BN, EK_F32_RES>(p, stream);
        case EK_BF16_TAB: return launch_tc2_k<BN, EK_BF16_TAB>(p, stream);
        case EK_PS: return launch_tc2_k<BN, EK_PS>(p, stream);
        case EK_PS_ST: return launch_tc2_k<BN, EK_PS_ST>(p, stream);
        case EK_BF16_ST: return launch_tc2_k<BN, EK_BF16_ST>(p, stream);
        case EK_GNIN_ST: return launch_tc2_k<BN, EK_GNIN_ST>(p, stream);
        case EK_F32_RESBF: return launch_tc2_k<BN, EK_F32_RESBF>(p, stream);
        case EK_F32_RES_LNOUT: return launch_tc2_k<BN, EK_F32_RES_LNOUT>(p, stream);
        case EK_BF16_LNIN: return launch_tc2_k<BN, EK_BF16_LNIN>(p, stream);
        case EK_BF16_GELU_LNIN: return launch_tc2_k<BN, EK_BF16_GELU_LNIN>(p, stream);
        default: return launch_tc2_k<BN, EK_GENERIC>(p, stream);
    }
}

int gemm_launch(const GemmProblem& p, cudaStream_t stream, int impl) {
    if (int rc = gemm_init()) return rc;
    VPU_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "empty GEMM (M=%d N=%d K=%d)", p.M, p.N, p.K);
    VPU_REQUIRE(p.K % 8 == 0, "GEMM K (%d) must be a multiple of 8", p.K);
    if (p.epi.mode == EPI_HEAD_FINAL) {
        VPU_REQUIRE(p.N == 64 && p.epi.m_per_batch % BM == 0, "head-final GEMM needs N=64 and m_per_batch %% 128 == 0");
    } else {
        VPU_REQUIRE(p.N % 32 == 0, "GEMM N (%d) must be a multiple of 32", p.N);
    }
    if (p.epi.mode == EPI_PIXEL_SHUFFLE)
        VPU_REQUIRE(p.epi.ps_cout % 32 == 0 && p.N == 4 * p.epi.ps_cout, "pixel-shuffle GEMM needs N == 4*cout, cout %% 32 == 0");
    if (p.epi.gn_out || p.epi.gn_in) {
        VPU_REQUIRE(p.epi.gn_rows >= 32 && p.M % p.epi.gn_rows == 0, "GroupNorm-fused GEMM: gn_rows (%d) must be >= 32 and divide M", p.epi.gn_rows);
        VPU_REQUIRE(!p.epi.gn_in || (p.epi.gn_wg && p.epi.gn_in_count > 0.f), "GroupNorm-fused GEMM: gn_in needs gn_wg and gn_in_count");
        VPU_REQUIRE(p.epi.mode != EPI_HEAD_FINAL, "GroupNorm fusion is not available in the head-final epilogue");
    }
    if (p.epi.ln_out || p.epi.ln_in) {
        const Epi& e = p.epi;
        VPU_REQUIRE(impl == 0 && g_use_2cta && p.M >= 2 * BM && p.N % 256 == 0 && e.mode == EPI_PLAIN && !e.bias2d && !e.gn_in && !e.gn_out,
                    "LayerNorm-fused GEMM needs the 2-CTA kernel, N %% 256 == 0 and a plain epilogue (M=%d N=%d)", p.M, p.N);
        VPU_REQUIRE(e.ln_slots > 0 && e.bias, "LayerNorm-fused GEMM: ln_slots / bias missing");
        if (e.ln_out)
            VPU_REQUIRE(!e.ln_in && !e.out_bf16 && e.res && !e.res_bf16 && e.act == ACT_NONE && e.ln_out_bf16 && e.ln_slots == gemm_ln_slots(p.M, p.N),
                        "ln_out needs a fp32 output with fp32 residual, a bf16 copy buffer and ln_slots == %d", gemm_ln_slots(p.M, p.N));
        else
            VPU_REQUIRE(e.out_bf16 && !e.res && e.ln_s && (e.act == ACT_NONE || e.act == ACT_GELU),
                        "ln_in needs a bf16 output without residual and ln_s");
        if (ln_tile_bn(p.M, p.N) == 128) return launch_tc2<128>(p, stream);
        // HBM-bound residual GEMMs (K <= 2 N: proj 119 -> 91 us, patch-embed low part) take the TMA-staged epilogue of gemm_res.cu; fc2
        // (K = 4 N, tensor-bound) keeps the generic kernel (181 us against 188 us with a 5-stage / one-staging-pair variant there)
        if (e.ln_out && p.K <= 2 * p.N && p.K <= g_res_kmax && gemm_res_supported(p)) return gemm_res_launch(p, stream);
        return launch_tc2<256>(p, stream);
    }
    // HBM-bound GroupNorm-fused neck GEMMs (K < 2 N, bf16 output + statistics): TMA-staged epilogue of gemm_gn.cu, same bits
    if (impl == 0 && g_use_2cta && (g_gn_tma & 1) && gemm_gn_supported(p)) return gemm_gn_launch(p, stream);
    if (impl == 0 && g_use_2cta && (g_gn_tma & 2) && gemm_ps_tma_supported(p)) return gemm_ps_tma_launch(p, stream);
    // the DMA stage's image-side K|V|Q projections: positional table through the TMA-staged epilogue of gemm_res.cu
    if (impl == 0 && g_use_2cta && g_res_modes && gemm_tab_supported(p)) return gemm_tab_launch(p, stream);
    if (impl == 1) {
#ifndef VPU_DEBUG
        VPU_REQUIRE(false, "GEMM impl 1 (mma.sync cross-check) exists in -DVPU_DEBUG builds only");
#else
        GemmDims d{p.M, p.N, p.K, g_stages, g_ablate};
        dim3 grid((p.N + 63) / 64, (p.M + 63) / 64);
        if (p.epi.m_per_batch > 0) VPU_REQUIRE(p.epi.m_per_batch % 64 == 0, "m_per_batch must be a multiple of 64");
        VPU_CHECK_CUDA(launch_pdl(gemm_mma_kernel, dim3(grid), dim3(128), 0, stream, p.A, p.W, p.lda, p.ldw, d, p.epi));
        VPU_CHECK_CUDA(cudaGetLastError());
        count_launch();
        return 0;
#endif
    }
    // impl 0: 2-CTA pairs whenever the shape allows it; impl 2 forces the 1-CTA kernel (A/B comparison, tests)
    if (impl == 0 && g_use_2cta && p.epi.mode != EPI_HEAD_FINAL && p.M >= 2 * BM) {
        // few tiles (the token-side projections of the DMA stage: M = 48 B rows): 128-wide tiles put twice as many CTA pairs to work
        const long long tiles256 = (long long)((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + 255) / 256);
        if (g_small_tiles && p.N % 256 == 0 && tiles256 * 4 <= g_num_sms) return launch_tc2<128>(p, stream);
        if (p.N % 256 == 0) return launch_tc2<256>(p, stream);
        // N = 128 (2k+1), k >= 2 (the DMA image-side K|V|Q projection, N = 1152): 256-wide tiles with a half-empty last tile
        // (TMA zero-fills the missing weight rows, the epilogue skips the missing columns) waste <= 1/5 of the MMA work but keep
        // the operand traffic at 64 B/clk/SM; 128-wide tiles need 96 B/clk/SM and run at half the tensor rate (ncu round 1f)
        if (g_ragged256 && p.N % 128 == 0 && p.N >= 640) return launch_tc2<256>(p, stream);
        if (p.N % 128 == 0) return launch_tc2<128>(p, stream);
    }
    if (p.epi.mode == EPI_HEAD_FINAL) return launch_tc<64, EK_HEAD>(p, stream);
    // one row tile (a handful of click sessions): the launch is bound by how fast its few CTAs pull the weights, so use many narrow tiles
    if (g_small_tiles && p.M <= BM && p.N % 64 == 0 && p.N >= 128 && p.epi.mode == EPI_PLAIN) return launch_tc<64>(p, stream);
    if (p.N % 256 == 0) return launch_tc<256>(p, stream);
    if (p.N % 192 == 0) return launch_tc<192>(p, stream);
    if (p.N % 128 == 0) return launch_tc<128>(p, stream);
    if (p.N <= 64) return launch_tc<64>(p, stream);
    return launch_tc<128>(p, stream);  // ragged N: TMA zero-fills, epilogue predicates n0 < N
}

}  // namespace vpu
