// Out-projection + residual + LayerNorm of the image tokens in the Dual-cross Merging Attention (reference transformer.py:456-463:
// keys = norm4(keys + cross_attn_image_to_token(...))), one kernel:
//     y[M, C] (bf16) = LayerNorm_C( A[M, K] W[C, K]^T + bias + res[M, C] (bf16) ) * gamma + beta,   rowmax_part[rank][m] = max_c y
// The unfused pair (GEMM writing fp32 rows + LayerNorm kernel reading them back) moved 10 bytes per output element through HBM; this
// one moves 4 (bf16 residual in, bf16 result out) and the fp32 rows never leave the SM.
//
// A row of C = 768 / 1024 / 1280 fp32 values does not fit one CTA's 512 TMEM columns, so a CLUSTER of NC = C / 256 CTAs shares a
// 128-row tile: CTA r computes columns [256 r, 256 r + 256) (tcgen05.mma, M = 128, N = 256, accumulator in TMEM), its epilogue threads
// (= rows) add bias + residual, write x back into the same TMEM columns and send the row's partial (sum, sum of squares) to every CTA
// of the cluster with st.async (DSMEM write that completes transaction bytes on the receiver's mbarrier: no cluster-wide barrier, no
// fences).  Every CTA then adds the NC partials in rank order (same bits everywhere), normalises its columns out of TMEM and stages
// bf16 rows for a TMA tensor store.
// The residual goes through the tensor core: its [128 x 64] bf16 chunks travel through the operand ring like A chunks and are
// multiplied by a 16 x 16 identity into the accumulator columns they belong to (16 tcgen05.mma of N = 16 per tile; bf16 x 1.0 is
// exact in the fp32 accumulator).  The producer warp therefore prefetches the residual with the operands, and the epilogue reads
// nothing but TMEM; staging it in the output tiles exposed an HBM round trip per tile (84 us against 64 us for ViT-B, batch 64;
// the unfused pair took 114 us).  What bounds it now (ncu, profiles/ncu_full_r2.md): the 3-stage operand ring -- a 1-CTA tile re-reads
// its 192 KB weight slice per 128 rows, 352 KB of TMA loads per tile with 144 KB in flight; shared memory is full.  (128 columns
// per CTA -- clusters of 6 / 8, five 32 KB stages, half the weight bytes per CTA and tile -- measured slower: 67 against 63 us for
// ViT-B, 120 against 105 us for ViT-L: the larger cluster's lockstep costs more than the deeper ring gains.)
// Two accumulators: the MMAs of the next row tile run under the epilogue of the current one.
#include "gemm.cuh"
#include "tc_attn.cuh"

namespace vpu {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2, STAGE = A_BYTES + W_BYTES;      // 16 + 32 KB
constexpr int STAGES = 3;
constexpr int TILE16 = BM * 64 * 2;             // staging tile: [128 rows x 64 bf16], 128-byte swizzle
constexpr int STG_OFF = STAGES * STAGE;
constexpr int ID_OFF = STG_OFF + 4 * TILE16;     // 16 x 64 bf16 tile, 128-byte swizzle: row n holds ones at k = n, n + 16, n + 32, n + 48
constexpr int ID_BYTES = 16 * 128;
constexpr int SMEM = ID_OFF + ID_BYTES + 1024;
constexpr int EPI_WARPS = 8;
constexpr int STORE_WARP = 2 + EPI_WARPS;
constexpr int THREADS = (STORE_WARP + 1) * 32;
constexpr int MAX_NC = 5;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
// 8-byte DSMEM store that completes 8 transaction bytes on the receiving CTA's mbarrier
__device__ __forceinline__ void st_async_f2(uint32_t remote_addr, float a, float b, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(remote_addr), "f"(a), "f"(b), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t nclusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}

struct LnFuseArgs {
    const float* bias;
    const float* gamma;
    const float* beta;
    float* rowmax_parts;          // [NC, M] or nullptr
    float eps;
    int M, K, C;
};

template <int NC>
__global__ void __launch_bounds__(THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmO, const LnFuseArgs a) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2], stg_full, stg_free, xbar[2];
    __shared__ __align__(16) float bias_s[BN], gamma_s[BN], beta_s[BN];
    __shared__ float2 loc_s[2][BM];                  // the two column halves' partials of a row, combined before they are sent
    __shared__ __align__(8) float2 part_s[2][NC][BM];   // [exchange buffer][source CTA][row]: written by the peers (st.async)
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    constexpr uint32_t XBYTES = NC * BM * 8;         // bytes one exchange delivers to a CTA

    // identity operand of the residual MMAs: element (n, k) = 1 iff k mod 16 == n, in the swizzled K-major layout (16-byte chunk c of
    // row n sits at chunk c ^ (n & 7))
    for (int i = threadIdx.x; i < 16 * 8; i += THREADS) {
        const int n = i >> 3, c = i & 7;             // chunk c = k in [8c, 8c + 8)
        uint32_t w[4] = {0, 0, 0, 0};
        const int kk = n - (c & 1) * 8;              // position of the one inside this chunk, if any
        if (kk >= 0 && kk < 8) w[kk >> 1] = (kk & 1) ? 0x3F800000u : 0x00003F80u;
        *reinterpret_cast<uint4*>(smem + ID_OFF + n * 128 + ((c ^ (n & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmO);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], EPI_WARPS);
            mbar_init(&xbar[i], 1);
        }
        mbar_init(&stg_full, EPI_WARPS);
        mbar_init(&stg_free, 1);
        fence_barrier_init();
        // armed before any peer can send (the cluster barrier below): exchanges 0 and 1
        mbar_arrive_expect_tx(&xbar[0], XBYTES);
        mbar_arrive_expect_tx(&xbar[1], XBYTES);
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();

    const int row_tiles = (a.M + BM - 1) / BM;
    const int kblks = (a.K + BK - 1) / BK;
    const int cl = (int)cluster_id_x(), ncl = (int)nclusters_x();
    const int col0 = rank * BN;                      // this CTA's columns of the row

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int t = cl; t < row_tiles; t += ncl) {
            for (int kb = 0; kb < kblks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], STAGE);
                    uint8_t* st = smem + stage * STAGE;
                    tma_load_2d(st, &tmA, &full_bar[stage], kb * BK, t * BM);
                    tma_load_2d(st + A_BYTES, &tmW, &full_bar[stage], kb * BK, col0);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            for (int j = 0; j < BN / 128; ++j) {     // the residual's four 64-column chunks, two per stage (A slot + head of the W slot)
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], 2 * A_BYTES);
                    tma_load_2d(smem + stage * STAGE, &tmR, &full_bar[stage], col0 + j * 128, t * BM);
                    tma_load_2d(smem + stage * STAGE + A_BYTES, &tmR, &full_bar[stage], col0 + j * 128 + 64, t * BM);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN), idesc_r = umma_idesc_bf16(BM, 16);
        const uint64_t adesc0 = umma_desc_k_sw128(smem_base), bdesc0 = umma_desc_k_sw128(smem_base + A_BYTES);
        const uint64_t idesc0 = umma_desc_k_sw128(smem_base + ID_OFF);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int t = cl; t < row_tiles; t += ncl) {
            mbar_wait(&acc_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < kblks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t soff = (uint64_t)(stage * (STAGE >> 4));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, adesc0 + soff + 2 * k, bdesc0 + soff + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            for (int j = 0; j < BN / 128; ++j) {     // + residual: 64-column chunk times the identity into columns [64 chunk + 16 k, + 16)
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t soff = (uint64_t)(stage * (STAGE >> 4));
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16(d_tmem + (2 * j + h) * 64 + k * 16, adesc0 + soff + (uint64_t)(h * (A_BYTES >> 4)) + 2 * k, idesc0 + 2 * k, idesc_r, 1u);
                    umma_commit(&empty_bar[stage]);
                    if (j + 1 == BN / 128) umma_commit(&acc_full[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else if (warp < STORE_WARP) {
        // ---------------- epilogue: thread = row; warp (quarter, half) owns columns [128 half, 128 half + 128) of the CTA's 256 ----------------
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane, sw = row & 7;
        const int et = threadIdx.x - 64;             // 0 .. 255
        bias_s[et] = __ldg(a.bias + col0 + et);
        gamma_s[et] = __ldg(a.gamma + col0 + et);
        beta_s[et] = __ldg(a.beta + col0 + et);
        asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
        const uint32_t part_addr0 = smem_u32(&part_s[0][rank][row]), part_addr1 = smem_u32(&part_s[1][rank][row]);
        const uint32_t xbar_addr0 = smem_u32(&xbar[0]), xbar_addr1 = smem_u32(&xbar[1]);
        int acc = 0;
        uint32_t acc_phase = 0, tile_phase = 0;
        int round = 0;                               // exchange counter: buffer round & 1, barrier phase (round >> 1) & 1
        const float inv_c = 1.0f / (float)a.C;
        for (int t = cl; t < row_tiles; t += ncl) {
            const int m = t * BM + row;
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * 128;
            // pass 1: x = accumulator (A W^T + residual) + bias; the row's partial moments over this warp's 128 columns.  The TMEM
            // load of chunk c + 1 is in flight while chunk c is reduced
            float s1 = 0.f, s2 = 0.f;
            {
                uint32_t ra[32], rb[32];
                tmem_ld_32x32(t_acc, ra);
#pragma unroll
                for (int c = 0; c < 4; c += 2) {
                    tmem_ld_wait();
                    tmem_ld_32x32(t_acc + (c + 1) * 32, rb);
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float x = __uint_as_float(ra[e]) + bias_s[half * 128 + c * 32 + e];
                        s1 += x;
                        s2 = fmaf(x, x, s2);
                    }
                    tmem_ld_wait();
                    if (c + 2 < 4) tmem_ld_32x32(t_acc + (c + 2) * 32, ra);
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float x = __uint_as_float(rb[e]) + bias_s[half * 128 + (c + 1) * 32 + e];
                        s1 += x;
                        s2 = fmaf(x, x, s2);
                    }
                }
            }
            // exchange: the halves combine in shared memory, then 128 threads send the row's (sum, sum of squares) to every CTA
            loc_s[half][row] = make_float2(s1, s2);
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
            const int buf = round & 1;
            if (half == 0) {
                const float2 p0 = loc_s[0][row], p1 = loc_s[1][row];
                const float c1 = p0.x + p1.x, c2 = p0.y + p1.y;
                const uint32_t pa = buf ? part_addr1 : part_addr0, xb = buf ? xbar_addr1 : xbar_addr0;
#pragma unroll
                for (int d = 0; d < NC; ++d) st_async_f2(mapa(pa, d), c1, c2, mapa(xb, d));
            }
            mbar_wait(&xbar[buf], (uint32_t)((round >> 1) & 1));
            float S1 = 0.f, S2 = 0.f;
#pragma unroll
            for (int src = 0; src < NC; ++src) {
                const float2 p = part_s[buf][src][row];
                S1 += p.x;
                S2 += p.y;
            }
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");      // everyone has read this buffer: re-arm it for exchange round + 2
            if (et == 0) mbar_arrive_expect_tx(&xbar[buf], XBYTES);
            ++round;
            const float mean = S1 * inv_c;
            const double var_d = fmax((double)S2 * (double)inv_c - (double)mean * (double)mean, 0.0);
            const float rstd = rsqrtf((float)var_d + a.eps);
            // pass 2: y = (x - mean) rstd gamma + beta -> bf16 rows in the staging tiles (free once the previous tile's stores have read them)
            float mx = -INFINITY;
            mbar_wait(&stg_free, tile_phase ^ 1);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(t_acc + c * 32, r);
                uint8_t* orow = smem + STG_OFF + (half * 2 + (c >> 1)) * TILE16 + row * 128;
                const float* gg = gamma_s + half * 128 + c * 32;
                const float* be = beta_s + half * 128 + c * 32;
                const float* bb = bias_s + half * 128 + c * 32;
                tmem_ld_wait();
                if (c == 3) {                        // x is in registers: the MMAs of the tile after next may overwrite the accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float y[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        y[e] = ((__uint_as_float(r[8 * u + e]) + bb[8 * u + e]) - mean) * rstd * gg[8 * u + e] + be[8 * u + e];
                        mx = fmaxf(mx, y[e]);
                    }
                    *reinterpret_cast<uint4*>(orow + ((((c & 1) * 4 + u) ^ sw) << 4)) =
                        make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&stg_full);
            if (a.rowmax_parts) {                    // per-CTA maximum of the row; the consumer (merge) takes the maximum over the NC parts
                loc_s[half][row].x = mx;
                asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
                if (half == 0 && m < a.M) a.rowmax_parts[(size_t)rank * a.M + m] = fmaxf(loc_s[0][row].x, loc_s[1][row].x);
                asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
            tile_phase ^= 1;
        }
    } else {
        // ---------------- result stores: four 64-column tiles per row tile ----------------
        uint32_t tile_phase = 0;
        for (int t = cl; t < row_tiles; t += ncl) {
            mbar_wait(&stg_full, tile_phase);
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 4; ++i) tma_store_2d(&tmO, smem + STG_OFF + i * TILE16, col0 + i * 64, t * BM);
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&stg_free);
            }
            __syncwarp();
            tile_phase ^= 1;
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NC>
int launch_ln(const GemmLn& p, cudaStream_t stream) {
    CUtensorMap tmA, tmW, tmR, tmO;
    if (int rc = gemm_tmap(&tmA, p.A, p.M, p.K, p.lda, BM)) return rc;
    if (int rc = gemm_tmap(&tmW, p.W, p.C, p.K, p.ldw, BN)) return rc;
    if (int rc = gemm_tmap(&tmR, p.res, p.M, p.C, p.ldr, BM)) return rc;
    if (int rc = gemm_tmap(&tmO, p.out, p.M, p.C, p.ldo, BM)) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln_kernel<NC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = stream;
    static int max_clusters = 0;
    if (max_clusters == 0) {
        cfg.gridDim = dim3(NC * (gemm_num_sms() / NC));
        int n = 0;
        VPU_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_ln_kernel<NC>, &cfg));
        VPU_REQUIRE(n > 0, "GEMM + LayerNorm: no cluster of %d CTAs fits the device", NC);
        max_clusters = n;
    }
    const int row_tiles = (p.M + BM - 1) / BM;
    const int clusters = row_tiles < max_clusters ? row_tiles : max_clusters;
    cfg.gridDim = dim3(NC * clusters);
    LnFuseArgs a{p.bias, p.gamma, p.beta, p.rowmax_parts, p.eps, p.M, p.K, p.C};
    VPU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_ln_kernel<NC>, tmA, tmW, tmR, tmO, a));
    count_launch();
    return 0;
}

}  // namespace

int gemm_ln_parts(int C) { return C / BN; }

bool gemm_ln_supported(const GemmLn& p) {
    return p.A && p.W && p.bias && p.res && p.gamma && p.beta && p.out && p.M > 0 && p.C % BN == 0 && p.C / BN >= 2 && p.C / BN <= MAX_NC &&
           p.K % 8 == 0 && p.lda % 8 == 0 && p.ldw % 8 == 0 && p.ldr % 8 == 0 && p.ldo % 8 == 0 &&
           ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.W) | reinterpret_cast<uintptr_t>(p.res) |
             reinterpret_cast<uintptr_t>(p.out)) & 15) == 0;
}

int gemm_ln_launch(const GemmLn& p, cudaStream_t stream) {
    if (int rc = gemm_init()) return rc;
    VPU_REQUIRE(gemm_ln_supported(p), "GEMM + LayerNorm: unsupported problem (C = %d)", p.C);
    switch (p.C / BN) {
        case 2: return launch_ln<2>(p, stream);
        case 3: return launch_ln<3>(p, stream);
        case 4: return launch_ln<4>(p, stream);
        default: return launch_ln<5>(p, stream);
    }
}

}  // namespace vpu
