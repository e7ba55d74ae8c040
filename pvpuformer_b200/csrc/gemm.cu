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

static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------
// Epilogue: CNT consecutive columns [n0, n0+CNT) of output row m.
// ------------------------------------------------------------------------------------------
template <int CNT>
__device__ __forceinline__ void epi_store(const Epi& e, int m, int n0, float (&v)[CNT], int N) {
    if (e.bias) {
#pragma unroll
        for (int t = 0; t < CNT; ++t) v[t] += __ldg(e.bias + n0 + t);
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
// tcgen05 kernel
// ------------------------------------------------------------------------------------------
constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 320;

template <int BN> struct TileCfg {
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 192 ? 5 : (BN == 128 ? 6 : 8));
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int ACC_STRIDE = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
};

struct GemmDims {
    int M, N, K;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDims d,
               const Epi e) {
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

    const int n_blks = (d.N + BN - 1) / BN;
    const int m_blks = (d.M + BM - 1) / BM;
    const int tiles = n_blks * m_blks;
    const int kblks = (d.K + BK - 1) / BK;

    if (warp == 0) {
        if (lane == 0) {  // ---------------- TMA producer ----------------
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int m_blk = tile / n_blks, n_blk = tile % n_blks;
                int brow = n_blk * BN;
                if (e.m_per_batch > 0) brow += ((m_blk * BM) / e.m_per_batch) * e.b_rows_per_batch;
                for (int kb = 0; kb < kblks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    tma_load_2d(sa + C::A_BYTES, &tmB, &full_bar[stage], kb * BK, brow);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---------------- MMA issuer ----------------
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
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
                    const uint32_t a_addr = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        umma_bf16(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32),
                                  idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);        // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {  // ---------------- epilogue warps ----------------
        const int quarter = warp & 3;            // TMEM lanes [32*quarter, +32) are this warp's
        const int half = (warp - 2) >> 2;        // column half of the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int m_blk = tile / n_blks, n_blk = tile % n_blks;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int row = m_blk * BM + quarter * 32 + lane;
#pragma unroll 1
            for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * C::ACC_STRIDE + c, r);
                tmem_ld_wait();
                const int n0 = n_blk * BN + c;
                if (row < d.M && n0 < d.N) {
                    float v[32];
#pragma unroll
                    for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
                    epi_store<32>(e, row, n0, v, d.N);
                }
            }
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
// mma.sync cross-check kernel (debug only; never selected by the forward unless VPU_GEMM_IMPL=1)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gemm_mma_kernel(const __nv_bfloat16* __restrict__ A,
                                                       const __nv_bfloat16* __restrict__ W, int lda, int ldw,
                                                       const GemmDims d, const Epi e) {
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

// ------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;
static std::mutex g_mu;

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
    VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<64>::SMEM_BYTES));
    VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<128>::SMEM_BYTES));
    VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<192>::SMEM_BYTES));
    VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<256>::SMEM_BYTES));
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

template <int BN>
static int launch_tc(const GemmProblem& p, cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap(&tmA, p.A, p.M, p.K, p.lda, BM)) return rc;
    if (int rc = make_tmap(&tmB, p.W, p.w_rows ? p.w_rows : p.N, p.K, p.ldw, BN)) return rc;
    const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;
    GemmDims d{p.M, p.N, p.K};
    gemm_tc_kernel<BN><<<grid, GEMM_THREADS, TileCfg<BN>::SMEM_BYTES, stream>>>(tmA, tmB, d, p.epi);
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
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
    if (impl == 1) {
        GemmDims d{p.M, p.N, p.K};
        dim3 grid((p.N + 63) / 64, (p.M + 63) / 64);
        if (p.epi.m_per_batch > 0) VPU_REQUIRE(p.epi.m_per_batch % 64 == 0, "m_per_batch must be a multiple of 64");
        gemm_mma_kernel<<<grid, 128, 0, stream>>>(p.A, p.W, p.lda, p.ldw, d, p.epi);
        VPU_CHECK_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    if (p.N % 256 == 0) return launch_tc<256>(p, stream);
    if (p.N % 192 == 0) return launch_tc<192>(p, stream);
    if (p.N % 128 == 0) return launch_tc<128>(p, stream);
    if (p.N <= 64) return launch_tc<64>(p, stream);
    return launch_tc<128>(p, stream);  // ragged N: TMA zero-fills, epilogue predicates n0 < N
}

}  // namespace vpu
