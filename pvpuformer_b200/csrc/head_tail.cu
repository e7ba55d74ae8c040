// Tail of the segmentation head at 1/4 resolution as ONE tcgen05 kernel (reference swin_transformer.py:727-767):
//     f    = relu(b + y_0 + sum_{l=1..3} resize(y_l))        fusion conv of the concat, as a sum of per-level slices
//     seg  = <f, w_seg> + b_seg                                conv_seg
//     aux  = (<f / |f|, q_n / |q_n|> + 1) / 2                  P2CL cosine logits of the 48 prompt queries
// where y_l = W_l h_l are the [256, 256] fusion-conv slices applied at native resolution by gemm_b2b.cu (a 1x1 conv commutes
// with bilinear resizing).  Round 1 ran this as head_combine_kernel (SIMT gathers of 2 x 2 taps per level, pixel and channel:
// 395 M warp instructions, 0.49 ms per batch-64 step, 25 % of the HBM rate) + a GEMM that re-read f from HBM (0.16 ms).
//
// Bilinear resizing is a linear map over pixels, so for a tile of 8 x 16 output pixels
//     resize(y_l)[tile] = U_l  .  y_l[source patch]        U_l [128 x patch pixels]: the 2 x 2 tap weights of every output pixel
// and the whole sum is ONE accumulation on the tensor core:
//     acc[128 x 256] = [U_1 | U_2 | U_3 | I] . [y_1 patch (6 x 10) ; y_2 patch (4 x 6) ; y_3 patch (3 x 4) ; y_0 tile (8 x 16)]
// (K = 64 + 32 + 16 + 128; the tap weights are products of two multiples of 1/16, exact in bf16; the identity block adds y_0).
// The B operands are the NHWC feature maps themselves, fetched as 4-D TMA boxes (channel group, x, y, image) -- rows past the
// map border arrive as zeros and the clamped taps of align_corners=False are folded into U, which therefore comes in 9 variants
// (first / interior / last tile row x column), a 288 KB table built once per geometry.
//   warp 0   TMA producer: U variant + 12 patch boxes + 4 y_0 boxes per tile, the image's 48 normalised queries on image change
//   warp 1   MMA issuer: 60 tcgen05.mma (N = 64 channel groups, MN-major B) per tile into one of two TMEM accumulators; then
//            aux = F . Qn^T with F read from TMEM (16 tcgen05.mma, N = 48)
//   warps 2-9   epilogue A: acc -> + bias, ReLU -> |f|^2, <f, w_seg> per pixel -> bf16 F packed in place in TMEM
//   warps 10-13 epilogue B: aux accumulator -> (v / |f| + 1) / 2 -> NCHW fp32 stores
// f itself never exists in memory.  Without the aux output (NoBRS / NoC loops read 'instances' only) the kernel stops after
// epilogue A and writes seg alone.
#include <mutex>
#include <unordered_map>
#include <vector>

#include "head_tail.cuh"
#include "tc_attn.cuh"

namespace vpu {

namespace {

constexpr int TH = 8, TW = 16, TM = TH * TW;        // output pixels per tile (UMMA M)
constexpr int CH = 256, NG = CH / 64;               // channels, 64-channel groups (one swizzle span)
constexpr int PH1 = 6, PW1 = 10, PH2 = 4, PW2 = 6, PH3 = 3, PW3 = 4;     // source patches of the three coarser levels
constexpr int K1P = 64, K2P = 32, K3P = 16;         // patch pixels padded to UMMA K steps: columns [0,64) [64,96) [96,112) of U
constexpr int NQ = 48;                              // prompt queries (2 x num_max_points)
constexpr int I_OFF = 0;                            // identity [128 x 128] bf16, K-major, two 64-column chunks
constexpr int U_OFF = I_OFF + 2 * TM * 128;         // U variant [128 x 128] (112 columns used)
constexpr int P1_OFF = U_OFF + 2 * TM * 128;        // y_1 patch: 4 groups x [64 rows x 128 B]
constexpr int P2_OFF = P1_OFF + NG * K1P * 128;
constexpr int P3_OFF = P2_OFF + NG * K2P * 128;
constexpr int Y0_OFF = P3_OFF + NG * K3P * 128;     // y_0 tile: 4 groups x [128 rows x 128 B]
constexpr int QN_OFF = Y0_OFF + NG * TM * 128;      // queries: 4 k-chunks x [48 rows x 128 B]
constexpr int SMEM = QN_OFF + NG * NQ * 128 + 1024;
constexpr int OPS_TX = 2 * TM * 128 + NG * 128 * (PH1 * PW1 + PH2 * PW2 + PH3 * PW3 + TM);
constexpr int THREADS = 14 * 32;
constexpr int AUX_COL = 64;                         // aux accumulator inside the (dead) columns [64, 128) of the tile's accumulator
static_assert(SMEM <= 232448 - 6144, "dynamic shared memory limit of sm_100");
static_assert(P1_OFF % 1024 == 0 && P2_OFF % 1024 == 0 && P3_OFF % 1024 == 0 && Y0_OFF % 1024 == 0 && QN_OFF % 1024 == 0, "swizzle atoms");

// TMEM column of the packed F chunk c (64 channels = 32 columns): see gemm_b2b.cu -- each epilogue warp overwrites only columns
// it has already read
__host__ __device__ constexpr int f_col(int c) { return (c >> 1) * 128 + (c & 1) * 32; }

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

struct TailArgs {
    const float* bias;
    const float* wseg;
    float seg_bias;
    float* seg_out;
    float* aux_out;      // nullptr: seg only
    int R, TY, TX;       // 1/4-scale resolution, tiles per column / row
    int tiles;           // B * TY * TX
};

__global__ void __launch_bounds__(THREADS, 1)
head_tail_kernel(const __grid_constant__ CUtensorMap tmY0, const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2,
                 const __grid_constant__ CUtensorMap tmY3, const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmQ,
                 const TailArgs a) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t ops_full, ops_empty, qn_full, qn_empty, acc_full[2], f_full[2], aux_full[2], acc_empty[2];
    __shared__ __align__(16) float bias_s[CH], wseg_s[CH];
    __shared__ float rn_s[2][TM], red_s[2][TM][2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool want_aux = a.aux_out != nullptr;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmY0); tma_prefetch_desc(&tmY1); tma_prefetch_desc(&tmY2); tma_prefetch_desc(&tmY3);
        tma_prefetch_desc(&tmU); tma_prefetch_desc(&tmQ);
        mbar_init(&ops_full, 1);
        mbar_init(&ops_empty, 1);
        mbar_init(&qn_full, 1);
        mbar_init(&qn_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&f_full[i], 8);
            mbar_init(&aux_full[i], 1);
            mbar_init(&acc_empty[i], want_aux ? 4 : 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, 512);
        tmem_relinquish();
    }
    // the padding rows of the patch operands are never written by TMA (boxes of 60 / 24 / 12 rows in 64 / 32 / 16): clear them
    // once (0 x stale NaN bits would poison the accumulator), and build the identity block
    for (int i = threadIdx.x; i < (Y0_OFF - I_OFF) / 16; i += THREADS) reinterpret_cast<uint4*>(smem + I_OFF)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x < TM) {
        const int r = threadIdx.x;
        uint8_t* p = smem + I_OFF + (r >> 6) * (TM * 128) + r * 128 + (((((r & 63) >> 3)) ^ (r & 7)) << 4) + (r & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16(1.0f);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();
    if (threadIdx.x < CH) {
        bias_s[threadIdx.x] = a.bias[threadIdx.x];
        wseg_s[threadIdx.x] = a.wseg[threadIdx.x];
    }
    __syncthreads();

    const int tpi = a.TY * a.TX;
    const int t0 = (int)((long long)blockIdx.x * a.tiles / gridDim.x), t1 = (int)((long long)(blockIdx.x + 1) * a.tiles / gridDim.x);

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        int img = -1;
        uint32_t oph = 0, qph = 0;
        for (int t = t0; t < t1; ++t) {
            const int b = t / tpi, by = (t % tpi) / a.TX, bx = t % a.TX;
            const int cls = (by == 0 ? 0 : (by == a.TY - 1 ? 2 : 1)) * 3 + (bx == 0 ? 0 : (bx == a.TX - 1 ? 2 : 1));
            mbar_wait(&ops_empty, oph ^ 1);
            oph ^= 1;
            if (elect_one()) {
                mbar_arrive_expect_tx(&ops_full, OPS_TX);
                tma_load_2d(smem + U_OFF, &tmU, &ops_full, 0, cls * TM);
                tma_load_2d(smem + U_OFF + TM * 128, &tmU, &ops_full, 64, cls * TM);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    tma_load_4d(smem + P1_OFF + g * K1P * 128, &tmY1, &ops_full, 64 * g, 8 * bx - 1, 4 * by - 1, b);
                    tma_load_4d(smem + P2_OFF + g * K2P * 128, &tmY2, &ops_full, 64 * g, 4 * bx - 1, 2 * by - 1, b);
                    tma_load_4d(smem + P3_OFF + g * K3P * 128, &tmY3, &ops_full, 64 * g, 2 * bx - 1, by - 1, b);
                    tma_load_4d(smem + Y0_OFF + g * TM * 128, &tmY0, &ops_full, 64 * g, TW * bx, TH * by, b);
                }
            }
            __syncwarp();
            // after the tile's operands: the wait below needs the aux product of the previous tile, which the issuer starts only
            // after this tile's first product (whose operands were requested above)
            if (want_aux && b != img) {
                mbar_wait(&qn_empty, qph ^ 1);       // every aux product of the previous image has completed
                qph ^= 1;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&qn_full, NG * NQ * 128);
#pragma unroll
                    for (int c = 0; c < NG; ++c) tma_load_3d(smem + QN_OFF + c * NQ * 128, &tmQ, &qn_full, 64 * c, 0, b);
                }
                __syncwarp();
                img = b;
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (converged warp, one elected lane) ----------------
        constexpr uint32_t idesc_acc = idesc_bf16(TM, 64, true), idesc_aux = idesc_bf16(TM, NQ, false);
        const uint64_t u_a = umma_desc_k_sw128(smem_base + U_OFF), i_a = umma_desc_k_sw128(smem_base + I_OFF);
        const uint64_t p1_b = umma_desc_mn_sw128(smem_base + P1_OFF), p2_b = umma_desc_mn_sw128(smem_base + P2_OFF);
        const uint64_t p3_b = umma_desc_mn_sw128(smem_base + P3_OFF), y0_b = umma_desc_mn_sw128(smem_base + Y0_OFF);
        const uint64_t qn_b = umma_desc_k_sw128(smem_base + QN_OFF);
        uint32_t oph = 0, qph = 0, eph[2] = {0, 0}, fph[2] = {0, 0};
        int img_aux = -1;
        auto aux_product = [&](int t, int buf) {      // aux(t) = F(t) Qn^T, F packed in TMEM by epilogue A
            const int b = t / tpi;
            mbar_wait(&f_full[buf], fph[buf]);
            fph[buf] ^= 1;
            if (b != img_aux) {
                mbar_wait(&qn_full, qph);
                qph ^= 1;
                img_aux = b;
            }
            tc_fence_after();
            if (elect_one()) {
                const uint32_t acc = tmem_base + buf * 256;
#pragma unroll
                for (int c = 0; c < NG; ++c) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ts(acc + AUX_COL, acc + f_col(c) + 8 * k, qn_b + (uint64_t)(c * (NQ * 128 >> 4)) + 2 * k, idesc_aux, (c | k) != 0 ? 1u : 0u);
                }
                umma_commit(&aux_full[buf]);
                if (t + 1 >= t1 || (t + 1) / tpi != b) umma_commit(&qn_empty);      // last use of this image's queries
            }
            __syncwarp();
        };
        int i = 0;
        for (int t = t0; t < t1; ++t, ++i) {
            const int buf = i & 1;
            mbar_wait(&acc_empty[buf], eph[buf] ^ 1);
            eph[buf] ^= 1;
            mbar_wait(&ops_full, oph);
            oph ^= 1;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t acc = tmem_base + buf * 256;
#pragma unroll
                for (int ks = 0; ks < 7; ++ks) {          // U: k-steps 0-3 level 1, 4-5 level 2, 6 level 3
                    const uint64_t ad = u_a + (uint64_t)((ks >> 2) * (TM * 128 >> 4)) + 2 * (ks & 3);
                    const uint64_t bd = ks < 4 ? p1_b + 128 * ks : (ks < 6 ? p2_b + 128 * (ks - 4) : p3_b);
                    const uint32_t gstride = ks < 4 ? (K1P * 128 >> 4) : (ks < 6 ? (K2P * 128 >> 4) : (K3P * 128 >> 4));
#pragma unroll
                    for (int g = 0; g < NG; ++g) umma_bf16(acc + 64 * g, ad, bd + (uint64_t)(g * gstride), idesc_acc, ks != 0 ? 1u : 0u);
                }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {          // identity: + y_0
                    const uint64_t ad = i_a + (uint64_t)((ks >> 2) * (TM * 128 >> 4)) + 2 * (ks & 3);
#pragma unroll
                    for (int g = 0; g < NG; ++g) umma_bf16(acc + 64 * g, ad, y0_b + 128 * ks + (uint64_t)(g * (TM * 128 >> 4)), idesc_acc, 1u);
                }
                umma_commit(&acc_full[buf]);
                umma_commit(&ops_empty);
            }
            __syncwarp();
            if (want_aux && i > 0) aux_product(t - 1, buf ^ 1);
        }
        if (want_aux && i > 0) aux_product(t1 - 1, (i - 1) & 1);
    } else if (warp < 10) {
        // ---------------- epilogue A: acc -> + bias, ReLU -> |f|^2, <f, w_seg> -> bf16 F packed in place ----------------
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        uint32_t aph[2] = {0, 0};
        int i = 0;
        for (int t = t0; t < t1; ++t, ++i) {
            const int buf = i & 1;
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256 + half * 128;
            mbar_wait(&acc_full[buf], aph[buf]);
            aph[buf] ^= 1;
            tc_fence_after();
            float ss = 0.f, seg = 0.f;
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(t_acc, ra);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t* r = (j & 1) ? rb : ra;
                tmem_ld_wait();
                if (j + 1 < 4) tmem_ld_32x32(t_acc + (j + 1) * 32, (j & 1) ? ra : rb);
                const int n0 = half * 128 + j * 32;
                uint32_t pk[16];
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    const float4 bb = *reinterpret_cast<const float4*>(bias_s + n0 + k);
                    const float4 ww = *reinterpret_cast<const float4*>(wseg_s + n0 + k);
                    const float f0 = fmaxf(__uint_as_float(r[k]) + bb.x, 0.f), f1 = fmaxf(__uint_as_float(r[k + 1]) + bb.y, 0.f);
                    const float f2 = fmaxf(__uint_as_float(r[k + 2]) + bb.z, 0.f), f3 = fmaxf(__uint_as_float(r[k + 3]) + bb.w, 0.f);
                    ss = fmaf(f0, f0, ss); ss = fmaf(f1, f1, ss); ss = fmaf(f2, f2, ss); ss = fmaf(f3, f3, ss);
                    seg = fmaf(f0, ww.x, seg); seg = fmaf(f1, ww.y, seg); seg = fmaf(f2, ww.z, seg); seg = fmaf(f3, ww.w, seg);
                    pk[k / 2] = pack_bf16(f0, f1);
                    pk[k / 2 + 1] = pack_bf16(f2, f3);
                }
                if (want_aux) tmem_st_32x16(t_acc + j * 16, pk);     // channels n0 .. n0+31 -> TMEM columns half*128 + 16 j .. +15 (already read)
            }
            // the row's other 128 channels are with the partner warp of this TMEM lane quarter
            if (half == 1) {
                red_s[buf][row][0] = ss;
                red_s[buf][row][1] = seg;
            }
            named_bar_sync(1 + quarter, 64);
            if (half == 0) {
                ss += red_s[buf][row][0];
                seg += red_s[buf][row][1];
                const int b = t / tpi, by = (t % tpi) / a.TX, bx = t % a.TX;
                const int y = TH * by + (row >> 4), x = TW * bx + (row & 15);
                a.seg_out[((size_t)b * a.R + y) * a.R + x] = seg + a.seg_bias;
                rn_s[buf][row] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            }
            if (want_aux) {
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&f_full[buf]);
            } else {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
        }
    } else if (want_aux) {
        // ---------------- epilogue B: aux accumulator -> (v / |f| + 1) / 2 -> aux[b, n, y, x] ----------------
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        uint32_t xph[2] = {0, 0};
        int i = 0;
        for (int t = t0; t < t1; ++t, ++i) {
            const int buf = i & 1;
            const uint32_t t_aux = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256 + AUX_COL;
            mbar_wait(&aux_full[buf], xph[buf]);
            xph[buf] ^= 1;
            tc_fence_after();
            uint32_t v0[32], v1[16];
            tmem_ld_32x32(t_aux, v0);
            tmem_ld_32x16(t_aux + 32, v1);
            tmem_ld_wait();
            tc_fence_before();
            const float rn = rn_s[buf][row];
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            const int b = t / tpi, by = (t % tpi) / a.TX, bx = t % a.TX;
            const int y = TH * by + (row >> 4), x = TW * bx + (row & 15);
            float* dst = a.aux_out + ((size_t)b * NQ * a.R + y) * a.R + x;
            const size_t plane = (size_t)a.R * a.R;
#pragma unroll
            for (int n = 0; n < 32; ++n) dst[n * plane] = fmaf(__uint_as_float(v0[n]) * rn, 0.5f, 0.5f);
#pragma unroll
            for (int n = 0; n < 16; ++n) dst[(32 + n) * plane] = fmaf(__uint_as_float(v1[n]) * rn, 0.5f, 0.5f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- host ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::mutex g_mu;
std::unordered_map<int, __nv_bfloat16*> g_tables;     // 1/4-scale resolution -> device table of the 9 U variants

void src_index(int dst, int in, int out, int& i0, int& i1, float& lam) {     // align_corners=False, as in elementwise.cu
    float s = ((float)dst + 0.5f) * ((float)in / (float)out) - 0.5f;
    if (s < 0.f) s = 0.f;
    i0 = (int)s;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    lam = s - (float)i0;
}

int init() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VPU_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    VPU_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    VPU_CHECK_CUDA(cudaFuncSetAttribute(head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// NHWC bf16 [B, r, r, 256]: dims (channel, x, y, image), box = 64 channels x bw x bh x 1
int make_map_4d(CUtensorMap* tm, const void* ptr, int B, int r, int bw, int bh) {
    cuuint64_t gdim[4] = {(cuuint64_t)CH, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)B};
    cuuint64_t gstride[3] = {(cuuint64_t)CH * 2, (cuuint64_t)r * CH * 2, (cuuint64_t)r * r * CH * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPU_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled (4-D head map, r=%d box %dx%d) failed with %d", r, bw, bh, (int)rc);
    return 0;
}

}  // namespace

// U variants for a 1/4-scale map of R x R pixels: [9][128][128] bf16, variant = 3 * (tile row class) + tile column class
// (0 first, 1 interior, 2 last), row = output pixel of the 8 x 16 tile, column = source pixel of the level's patch.
int head_tail_prepare(int R) {
    if (int rc = init()) return rc;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_tables.count(R)) return 0;
    }
    VPU_REQUIRE(R % TW == 0 && R % 8 == 0 && R / TH >= 3 && R / TW >= 3, "head tail: unsupported 1/4-scale resolution %d", R);
    const int TY = R / TH, TX = R / TW;
    std::vector<float> u((size_t)9 * TM * 128, 0.f);
    const int pw[4] = {0, PW1, PW2, PW3}, koff[4] = {0, 0, K1P, K1P + K2P};
    for (int cy = 0; cy < 3; ++cy)
        for (int cx = 0; cx < 3; ++cx) {
            const int by = cy == 0 ? 0 : (cy == 1 ? 1 : TY - 1), bx = cx == 0 ? 0 : (cx == 1 ? 1 : TX - 1);
            float* tab = u.data() + (size_t)(cy * 3 + cx) * TM * 128;
            for (int p = 0; p < TM; ++p) {
                const int y = TH * by + p / TW, x = TW * bx + p % TW;
                for (int l = 1; l < 4; ++l) {
                    const int r = R >> l, oy = ((TH * by) >> l) - 1, ox = ((TW * bx) >> l) - 1;
                    int y0, y1, x0, x1;
                    float ly, lx;
                    src_index(y, r, R, y0, y1, ly);
                    src_index(x, r, R, x0, x1, lx);
                    auto add = [&](int yy, int xx, float w) { tab[(size_t)p * 128 + koff[l] + (yy - oy) * pw[l] + (xx - ox)] += w; };
                    add(y0, x0, (1.f - ly) * (1.f - lx));
                    add(y0, x1, (1.f - ly) * lx);
                    add(y1, x0, ly * (1.f - lx));
                    add(y1, x1, ly * lx);
                }
            }
        }
    std::vector<__nv_bfloat16> ub(u.size());
    for (size_t i = 0; i < u.size(); ++i) {
        ub[i] = __float2bfloat16(u[i]);
        VPU_REQUIRE(__bfloat162float(ub[i]) == u[i], "head tail: interpolation weight %g is not exact in bf16", (double)u[i]);
    }
    __nv_bfloat16* dev = nullptr;
    VPU_CHECK_CUDA(cudaMalloc(&dev, ub.size() * sizeof(__nv_bfloat16)));
    VPU_CHECK_CUDA(cudaMemcpy(dev, ub.data(), ub.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    std::lock_guard<std::mutex> lk(g_mu);
    g_tables[R] = dev;
    return 0;
}

bool head_tail_supported(const HeadTailArgs& a) {
    const int R = a.res[0];
    return a.channels == CH && (!a.aux_out || a.nq == NQ) && R % TW == 0 && R / TH >= 3 && R / TW >= 3 && a.res[1] * 2 == R && a.res[2] * 4 == R &&
           a.res[3] * 8 == R;
}

int head_tail_launch(const HeadTailArgs& a, cudaStream_t stream) {
    VPU_REQUIRE(head_tail_supported(a), "head tail: unsupported geometry (R=%d, channels=%d, nq=%d)", a.res[0], a.channels, a.nq);
    const int R = a.res[0];
    if (int rc = head_tail_prepare(R)) return rc;
    const __nv_bfloat16* table;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        table = g_tables.at(R);
    }
    CUtensorMap tmY0, tmY1, tmY2, tmY3, tmU, tmQ;
    if (int rc = make_map_4d(&tmY0, a.y[0], a.B, R, TW, TH)) return rc;
    if (int rc = make_map_4d(&tmY1, a.y[1], a.B, R / 2, PW1, PH1)) return rc;
    if (int rc = make_map_4d(&tmY2, a.y[2], a.B, R / 4, PW2, PH2)) return rc;
    if (int rc = make_map_4d(&tmY3, a.y[3], a.B, R / 8, PW3, PH3)) return rc;
    if (int rc = gemm_tmap(&tmU, table, 9 * TM, 128, 128, TM)) return rc;
    if (a.aux_out) {
        cuuint64_t gdim[3] = {(cuuint64_t)CH, 64, (cuuint64_t)a.B};
        cuuint64_t gstride[2] = {(cuuint64_t)CH * 2, (cuuint64_t)64 * CH * 2};
        cuuint32_t box[3] = {64, NQ, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult rc = g_encode(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(a.qn), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VPU_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled (head queries) failed with %d", (int)rc);
    } else {
        tmQ = tmU;
    }
    TailArgs t;
    t.bias = a.bias; t.wseg = a.wseg; t.seg_bias = a.seg_bias; t.seg_out = a.seg_out; t.aux_out = a.aux_out;
    t.R = R; t.TY = R / TH; t.TX = R / TW; t.tiles = a.B * t.TY * t.TX;
    const int sms = gemm_num_sms();
    const int ctas = t.tiles < sms ? t.tiles : sms;
    VPU_CHECK_CUDA(launch_pdl(head_tail_kernel, dim3(ctas), dim3(THREADS), (size_t)SMEM, stream, tmY0, tmY1, tmY2, tmY3, tmU, tmQ, t));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace vpu
