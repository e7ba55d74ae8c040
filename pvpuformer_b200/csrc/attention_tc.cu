// tcgen05 / TMEM window attention for head_dim 64 (ViT-B / ViT-L blocks; reference
// models_vit.py:43-56 applied per 224-px window, models_vit.py:225-255).
//
// One persistent CTA per SM loops over (image, window, head) problems: Q, K, V are [196 x 64] bf16
// slices of the fused qkv projection output, gathered by 5-D TMA boxes (column, j, wj, grid row, image)
// so the window partition never exists in memory.  Per problem:
//   S_t = Q_t K^T        tcgen05.mma, M = 128 query rows (two row tiles: 9 and 5 window rows = 126 + 70
//                        queries), N = 208 keys (196 valid, 12 zero rows), fp32 accumulators in TMEM
//   P_t = exp2(S_t - max) two softmax warpgroups (one per row tile, thread == query row) read S_t with
//                        tcgen05.ld, write bf16 P_t back into the same TMEM columns with tcgen05.st
//   O_t = P_t V          tcgen05.mma with the A operand in TMEM and V as an MN-major shared-memory
//                        operand (no transpose copy), accumulators aliasing the dead S_t columns
//   out = O_t / rowsum   tcgen05.ld epilogue, bf16, straight into the token-major attention output
// Scores and probabilities never touch shared or global memory.  The MMA warp software-pipelines
// S(n+1) between the two P V products of problem n, and Q/K/V of the next problem are prefetched into
// the second shared-memory stage while the current one is in flight.
#include <mutex>
#include <unordered_map>

#include "attention.cuh"

namespace vpu {

namespace {

constexpr int D = 64;                 // head dim
constexpr int WIN = 14;               // tokens per window side
constexpr int SK = WIN * WIN;         // 196 keys / queries per window
constexpr int SKP = 208;              // keys padded to a multiple of 16 (UMMA N / K granularity)
constexpr int T0_IROWS = 9, T1_IROWS = 5;            // window rows per query tile
constexpr int T0_ROWS = T0_IROWS * WIN;              // 126
constexpr int T1_ROWS = T1_IROWS * WIN;              // 70
constexpr int Q_TILE_BYTES = 128 * D * 2;            // 16 KB (128-row UMMA tile; rows past the box are don't-care)
constexpr int KV_BYTES = SKP * D * 2;                // 26 KB
constexpr int STAGE_BYTES = 2 * Q_TILE_BYTES + 2 * KV_BYTES;   // 84 KB
constexpr int STAGES = 2;
constexpr int TX_BYTES = (T0_ROWS + T1_ROWS + 2 * SK) * D * 2; // bytes the four TMA boxes deliver
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr int THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2-5 softmax tile 0, warps 6-9 softmax tile 1
constexpr int TILE_COLS = 256;        // TMEM columns reserved per row tile: S [0,208), P [0,104), O [128,192)
constexpr int O_COL = 128;

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {   // one MUFU; arguments are <= 0 here, results in (0, 1]
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major 128B-swizzled operand (V: rows = keys (K), 64 head-dim elements = one 128-byte row):
// 8-row swizzle atoms 1024 B apart along K (SBO); a single 64-element block along N, so LBO is unused.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

struct WinArgs {
    __nv_bfloat16* o;
    int ldo;
    int heads, nwin_side, grid, tokens;   // 12, 2, 28, 784
    int nprob;                            // images * windows * heads
    int qcol, kcol, vcol;                 // column of head 0 in the fused projection buffer
    float scale_log2;
};

__global__ void __launch_bounds__(THREADS, 1)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmQ0,
                           const __grid_constant__ CUtensorMap tmQ1, const WinArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], s_full[2], p_full[2], o_full[2], s_empty[2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // rows 196..207 of every K / V buffer are never written by TMA: zero them once (finite scores, zero P V terms)
    for (int i = threadIdx.x; i < STAGES * 2 * (SKP - SK) * D * 2 / 16; i += THREADS) {
        const int per = (SKP - SK) * D * 2 / 16;          // 16-byte units per buffer tail
        const int buf = i / per, u = i % per;             // buf = stage * 2 + {K, V}
        uint8_t* base = smem + (buf >> 1) * STAGE_BYTES + 2 * Q_TILE_BYTES + (buf & 1) * KV_BYTES + SK * D * 2;
        reinterpret_cast<uint4*>(base)[u] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmKV);
        tma_prefetch_desc(&tmQ0);
        tma_prefetch_desc(&tmQ1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&p_full[t], 4);
            mbar_init(&o_full[t], 1);
            mbar_init(&s_empty[t], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int nw2 = a.nwin_side * a.nwin_side;

    if (warp == 0) {
        if (lane == 0) {  // ---------------- TMA producer ----------------
            int stage = 0;
            uint32_t phase = 0;
            for (int p = blockIdx.x; p < a.nprob; p += gridDim.x) {
                const int h = p % a.heads, w = (p / a.heads) % nw2, b = p / (a.heads * nw2);
                const int wi = w / a.nwin_side, wj = w % a.nwin_side;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full_bar[stage], TX_BYTES);
                uint8_t* st = smem + stage * STAGE_BYTES;
                tma_load_5d(st, &tmQ0, &full_bar[stage], a.qcol + h * D, 0, wj, wi * WIN, b);
                tma_load_5d(st + Q_TILE_BYTES, &tmQ1, &full_bar[stage], a.qcol + h * D, 0, wj, wi * WIN + T0_IROWS, b);
                tma_load_5d(st + 2 * Q_TILE_BYTES, &tmKV, &full_bar[stage], a.kcol + h * D, 0, wj, wi * WIN, b);
                tma_load_5d(st + 2 * Q_TILE_BYTES + KV_BYTES, &tmKV, &full_bar[stage], a.vcol + h * D, 0, wj, wi * WIN, b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---------------- MMA issuer ----------------
            constexpr uint32_t idesc_s = idesc_bf16(128, SKP, false);
            constexpr uint32_t idesc_o = idesc_bf16(128, D, true);
            auto issue_s = [&](int t, uint32_t st_addr) {
                const uint32_t q_addr = st_addr + t * Q_TILE_BYTES, k_addr = st_addr + 2 * Q_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    umma_bf16(tmem_base + t * TILE_COLS, umma_desc_k_sw128(q_addr + k * 32), umma_desc_k_sw128(k_addr + k * 32),
                              idesc_s, k ? 1u : 0u);
                umma_commit(&s_full[t]);
            };
            auto issue_pv = [&](int t, uint32_t st_addr) {
                const uint32_t v_addr = st_addr + 2 * Q_TILE_BYTES + KV_BYTES;
#pragma unroll
                for (int k = 0; k < SKP / 16; ++k)
                    umma_bf16_ts(tmem_base + t * TILE_COLS + O_COL, tmem_base + t * TILE_COLS + k * 8,
                                 umma_desc_mn_sw128(v_addr + k * 16 * D * 2), idesc_o, k ? 1u : 0u);
                umma_commit(&o_full[t]);
            };
            int stage = 0;
            uint32_t phase = 0;              // full / empty ring
            uint32_t tphase = 0;             // per-problem phase of s_full / p_full / o_full / s_empty
            int p = blockIdx.x;
            if (p < a.nprob) {               // prologue: S0, S1 of the first problem
                mbar_wait(&full_bar[0], 0);
                tc_fence_after();
                issue_s(0, smem_base);
                issue_s(1, smem_base);
            }
            for (; p < a.nprob; p += gridDim.x) {
                const uint32_t st_addr = smem_base + stage * STAGE_BYTES;
                const bool has_next = p + (int)gridDim.x < a.nprob;
                const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
                const uint32_t nphase = (stage + 1 == STAGES) ? phase ^ 1 : phase;
                const uint32_t nst_addr = smem_base + nstage * STAGE_BYTES;
                mbar_wait(&p_full[0], tphase);
                tc_fence_after();
                issue_pv(0, st_addr);
                if (has_next) {
                    mbar_wait(&full_bar[nstage], nphase);
                    mbar_wait(&s_empty[0], tphase);           // tile-0 columns free: epilogue 0 of this problem done
                    tc_fence_after();
                    issue_s(0, nst_addr);
                }
                mbar_wait(&p_full[1], tphase);
                tc_fence_after();
                issue_pv(1, st_addr);
                umma_commit(&empty_bar[stage]);               // every MMA reading this stage has been issued before
                if (has_next) {
                    mbar_wait(&s_empty[1], tphase);
                    tc_fence_after();
                    issue_s(1, nst_addr);
                }
                stage = nstage;
                phase = nphase;
                tphase ^= 1;
            }
        }
    } else {  // ---------------- softmax + epilogue warpgroups ----------------
        const int t = (warp - 2) >> 2;                   // row tile of this warpgroup
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;             // query row inside the tile
        const int nvalid = t == 0 ? T0_ROWS : T1_ROWS;
        const uint32_t tile_tmem = tmem_base + ((uint32_t)(quarter * 32) << 16) + t * TILE_COLS;
        const bool warp_has_rows = quarter * 32 < nvalid;
        uint32_t tphase = 0;
        for (int p = blockIdx.x; p < a.nprob; p += gridDim.x) {
            const int h = p % a.heads, w = (p / a.heads) % nw2, b = p / (a.heads * nw2);
            mbar_wait(&s_full[t], tphase);
            tc_fence_after();
            float sum = 0.f;
            if (warp_has_rows) {      // warps whose 32 rows are all past the tile's valid queries only keep the barriers moving
                // pass 1: row maximum over the 196 valid keys; the TMEM load of chunk c+1 is in flight while chunk c is reduced
                float mx = -INFINITY;
                uint32_t ra[32], rb[32];
                tmem_ld_32x32(tile_tmem, ra);
#pragma unroll 1
                for (int c = 0; c < 6; c += 2) {
                    tmem_ld_wait();
                    tmem_ld_32x32(tile_tmem + (c + 1) * 32, rb);
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(ra[i]));
                    tmem_ld_wait();
                    if (c + 2 < 6) tmem_ld_32x32(tile_tmem + (c + 2) * 32, ra);
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(rb[i]));
                }
                uint32_t rt[16];
                tmem_ld_32x16(tile_tmem + 192, rt);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < SK - 192; ++i) mx = fmaxf(mx, __uint_as_float(rt[i]));
                const float moff = mx * a.scale_log2;
                // pass 2: p = exp2(s * scale - max * scale), bf16 pairs written over the already-consumed S columns
                auto expo = [&](uint32_t sbits) { return ex2_approx(fmaf(__uint_as_float(sbits), a.scale_log2, -moff)); };
                tmem_ld_32x32(tile_tmem, ra);
#pragma unroll 1
                for (int c = 0; c < 6; c += 2) {
                    uint32_t pk[16];
                    tmem_ld_wait();
                    tmem_ld_32x32(tile_tmem + (c + 1) * 32, rb);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = expo(ra[2 * i]), p1 = expo(ra[2 * i + 1]);
                        sum += p0 + p1;
                        pk[i] = pack_bf16(p0, p1);
                    }
                    tmem_st_32x16(tile_tmem + c * 16, pk);
                    tmem_ld_wait();
                    if (c + 2 < 6) tmem_ld_32x32(tile_tmem + (c + 2) * 32, ra);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = expo(rb[2 * i]), p1 = expo(rb[2 * i + 1]);
                        sum += p0 + p1;
                        pk[i] = pack_bf16(p0, p1);
                    }
                    tmem_st_32x16(tile_tmem + (c + 1) * 16, pk);
                }
                {
                    uint32_t pk[8];
                    tmem_ld_32x16(tile_tmem + 192, rt);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float p0 = 0.f, p1 = 0.f;
                        if (2 * i < SK - 192) p0 = expo(rt[2 * i]);
                        if (2 * i + 1 < SK - 192) p1 = expo(rt[2 * i + 1]);
                        sum += p0 + p1;
                        pk[i] = pack_bf16(p0, p1);
                    }
                    tmem_st_32x8(tile_tmem + 96, pk);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
            // epilogue: O / rowsum -> bf16 -> token-major output row of this query
            mbar_wait(&o_full[t], tphase);
            tc_fence_after();
            const float inv = warp_has_rows ? 1.0f / sum : 0.f;
            const int wi = w / a.nwin_side, wj = w % a.nwin_side;
            const int s = (t == 0 ? 0 : T0_ROWS) + row, i = s / WIN, j = s % WIN;
            __nv_bfloat16* dst = a.o + ((size_t)b * a.tokens + (size_t)(wi * WIN + i) * a.grid + wj * WIN + j) * a.ldo + h * D;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tile_tmem + O_COL + c * 32, r);
                tmem_ld_wait();
                if (row < nvalid) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 v;
                        v.x = pack_bf16(__uint_as_float(r[8 * q]) * inv, __uint_as_float(r[8 * q + 1]) * inv);
                        v.y = pack_bf16(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv);
                        v.z = pack_bf16(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv);
                        v.w = pack_bf16(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv);
                        *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = v;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[t]);
            tphase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- host: 5-D tensor maps over the fused projection buffer --------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_sms = 0;
std::mutex g_mu;
struct Key {
    const void* p;
    int ld, images, grid, irows;
    bool operator==(const Key& o) const { return p == o.p && ld == o.ld && images == o.images && grid == o.grid && irows == o.irows; }
};
struct KeyHash {
    size_t operator()(const Key& k) const {
        size_t h = reinterpret_cast<size_t>(k.p);
        h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)k.images; h = h * 1000003u ^ (size_t)k.grid;
        return h * 1000003u ^ (size_t)k.irows;
    }
};
std::unordered_map<Key, CUtensorMap, KeyHash> g_cache;

int init() {
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
    VPU_REQUIRE(prop.major == 10, "window attention needs an sm_100a device");
    g_sms = prop.multiProcessorCount;
    VPU_CHECK_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// dims (innermost first): column, j inside the window, window column, grid row, image
int make_map(CUtensorMap* tm, const void* ptr, int ld, int images, int grid, int irows) {
    Key key{ptr, ld, images, grid, irows};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) { *tm = it->second; return 0; }
    }
    const cuuint64_t row_b = (cuuint64_t)ld * 2;
    cuuint64_t gdim[5] = {(cuuint64_t)ld, (cuuint64_t)WIN, (cuuint64_t)(grid / WIN), (cuuint64_t)grid, (cuuint64_t)images};
    cuuint64_t gstride[4] = {row_b, WIN * row_b, (cuuint64_t)grid * row_b, (cuuint64_t)grid * grid * row_b};
    cuuint32_t box[5] = {(cuuint32_t)D, (cuuint32_t)WIN, 1, (cuuint32_t)irows, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPU_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (5-D window map) failed with %d", (int)r);
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.size() > 1024) g_cache.clear();
    g_cache[key] = *tm;
    return 0;
}

}  // namespace

bool window_attention_tc_supported(const AttnArgs& a, int head_dim) {
    return head_dim == D && a.qmap.mode == 1 && a.qmap.win == WIN && a.qmap.grid % WIN == 0 && a.Sq == SK && a.Sk == SK &&
           a.q == a.k && a.q == a.v && a.ldq == a.ldk && a.ldq == a.ldv && a.ldq % 8 == 0 && a.ldo % 8 == 0 &&
           a.qoff % 8 == 0 && a.koff % 8 == 0 && a.voff % 8 == 0 && (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(a.o) & 15) == 0;
}

int window_attention_tc_launch(const AttnArgs& a, cudaStream_t stream) {
    if (int rc = init()) return rc;
    const int grid = a.qmap.grid, nws = grid / WIN, images = a.nprob / (nws * nws);
    VPU_REQUIRE(images * nws * nws == a.nprob, "window attention: nprob must be images * windows");
    CUtensorMap tmKV, tmQ0, tmQ1;
    if (int rc = make_map(&tmKV, a.q, a.ldq, images, grid, WIN)) return rc;
    if (int rc = make_map(&tmQ0, a.q, a.ldq, images, grid, T0_IROWS)) return rc;
    if (int rc = make_map(&tmQ1, a.q, a.ldq, images, grid, T1_IROWS)) return rc;
    WinArgs w;
    w.o = a.o; w.ldo = a.ldo; w.heads = a.heads; w.nwin_side = nws; w.grid = grid; w.tokens = a.qmap.tokens;
    w.nprob = a.nprob * a.heads; w.qcol = a.qoff; w.kcol = a.koff; w.vcol = a.voff; w.scale_log2 = a.scale_log2;
    const int ctas = w.nprob < g_sms ? w.nprob : g_sms;
    window_attention_tc_kernel<<<ctas, THREADS, SMEM_BYTES, stream>>>(tmKV, tmQ0, tmQ1, w);
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace vpu
