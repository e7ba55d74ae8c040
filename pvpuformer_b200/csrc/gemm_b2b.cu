// Back-to-back tcgen05 GEMM pair for the segmentation head (reference swin_transformer.py:723-737):
//     H = relu(A W1^T + b1)   [M, 256]   1x1 conv + ReLU of one pyramid level (ConvModule, K1 = 128 / 256 / 512 / 1024)
//     Y = H W2^T              [M, 256]   that level's slice of the fusion conv, applied at native resolution
// as ONE persistent kernel: H never leaves the SM.  The two separate GEMMs wrote and re-read H through HBM -- 2 x 411 MB per
// batch-64 step at 1/4 resolution -- and both ran HBM-bound at 3.7 / 5.1 TB/s (round-2 per-launch profile).
//
// A CTA pair (cluster of 2, tcgen05.mma.cta_group::2, M = 256) owns 256 rows per tile; per CTA:
//   TMEM            acc1 [0, 256) fp32; the first epilogue packs relu(acc1 + b1) to bf16 IN PLACE (H chunk c -- 64 columns of H --
//                   occupies 32 TMEM columns), and the second product takes its A operand straight from TMEM, like P V in the
//                   attention kernels: H touches neither shared memory nor the generic -> async proxy fence.  acc2 [256, 512).
//   shared memory   W2 half [128 x 256] resident (64 KB, loaded once) | 3-stage ring of (A k-chunk [128 x 64] + W1 half k-chunk
//                   [128 x 64]) = 96 KB | output staging tile [128 x 256] bf16 (64 KB, four 128B-swizzled 64-column chunks)
//   warps           0 TMA producer, 1 MMA issuer (leader CTA), 2-9 first epilogue, 10-17 second epilogue (TMEM -> bf16 -> staging),
//                   18 output stores (four TMA tensor stores per tile: an epilogue thread owns a row, so direct global stores touch
//                   32 lines per instruction; measured 222 us with them against 155 us without any store at 1/4 resolution)
// Per tile: MMA1(i) -> epi1(i) chunk by chunk || MMA2(i) k-chunk by k-chunk -> MMA1(i+1) -> epi2(i) || epi1(i+1) ...
//
// Two negative results of round 2, kept out of the code:
//   * H through shared memory (generic-proxy stores of both CTAs -> fence.proxy.async -> mbarrier with .release.cluster /
//     .acquire.cluster): the cluster-scope semantics compile to MEMBAR + ERRBAR + an L1 invalidate (CCTL.IVALL) per wait
//     iteration; 258 us at 1/4 resolution against 130 us with H in TMEM.
//   * the producing branch's GroupNorm + GELU applied to the A k-chunks in shared memory by extra warps (instead of the separate
//     gn_apply pass): the kernel becomes issue-bound (7 FMA + 2 MUFU per element on top of both epilogues) and gains exactly what
//     the removed passes cost (407 us against 229 + 181 us).
#include "gemm.cuh"
#include "tc_attn.cuh"

namespace vpu {

namespace {

constexpr int BM = 128, BK = 64, NN = 256;
constexpr int CHUNK = BM * BK * 2;              // 16 KB: [128 rows x 64 columns] bf16, 128-byte swizzle
constexpr int STAGE = 2 * CHUNK;                // A k-chunk + W1 half k-chunk
constexpr int STAGES = 3;
constexpr int W2_OFF = 0, OUT_OFF = 4 * CHUNK, RING_OFF = 8 * CHUNK;
constexpr int SMEM = RING_OFF + STAGES * STAGE + 1024;
constexpr int EPI_WARPS = 8;                    // per epilogue: two warps per TMEM lane quarter, 128 columns each
constexpr int STORE_WARP = 2 + 2 * EPI_WARPS;
constexpr int THREADS = (STORE_WARP + 1) * 32;
static_assert(SMEM <= 232448 - 1280, "dynamic shared memory limit of sm_100");

// TMEM column of H chunk c (64 bf16 columns of H = 32 TMEM columns): chunks 0, 1 are packed by the warps that own acc1
// columns [0, 128), chunks 2, 3 by the owners of [128, 256) -- each warp only overwrites columns it has already read.
__host__ __device__ constexpr int h_col(int c) { return (c >> 1) * 128 + (c & 1) * 32; }

// D[tmem] (+)= A[tmem] * B[smem desc], M = 256 over the CTA pair (each CTA's TMEM holds its own 128 rows of A and D)
__device__ __forceinline__ void umma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct B2BArgs {
    const float* bias1;
    int M, K1;
};

__global__ void __launch_bounds__(THREADS, 1)
gemm_b2b_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY, const B2BArgs a) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], w2_full, acc1_full, acc2_full, h_full[4], acc2_empty, stg_full, stg_free;
    __shared__ __align__(16) float bias_s[NN];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)(cluster_ctarank() & 1);
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmY);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&w2_full, 1);
        mbar_init(&acc1_full, 1);
        mbar_init(&acc2_full, 1);
        for (int c = 0; c < 4; ++c) mbar_init(&h_full[c], 8);   // H chunk c: its four first-epilogue warps in both CTAs (leader's barrier)
        mbar_init(&acc2_empty, 2 * EPI_WARPS);                 // second-epilogue warps of both CTAs (leader's barrier)
        mbar_init(&stg_full, EPI_WARPS);
        mbar_init(&stg_free, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_2sm(&tmem_base_smem, 512);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();
    if (threadIdx.x < NN) bias_s[threadIdx.x] = a.bias1[threadIdx.x];     // read back by broadcast LDS: an L2 round trip per chunk stalled the first epilogue
    __syncthreads();

    const int tiles = (a.M + 2 * BM - 1) / (2 * BM);
    const int kblks = a.K1 / BK;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs; converged warp, one elected lane) ----------------
        if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&w2_full, 2 * 4 * CHUNK);
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_load_2d_2sm(smem + W2_OFF + c * CHUNK, &tmW2, &w2_full, c * BK, rank * BM);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            const int arow = tile * 2 * BM + rank * BM;
            for (int kb = 0; kb < kblks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE);
                    uint8_t* st = smem + RING_OFF + stage * STAGE;
                    tma_load_2d_2sm(st, &tmA, &full_bar[stage], kb * BK, arow);
                    tma_load_2d_2sm(st + CHUNK, &tmW1, &full_bar[stage], kb * BK, rank * BM);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {  // ---------------- MMA issuer (leader CTA only; converged warp, one elected lane) ----------------
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, NN);
            const uint64_t ring_a = umma_desc_k_sw128(smem_base + RING_OFF), ring_b = umma_desc_k_sw128(smem_base + RING_OFF + CHUNK);
            const uint64_t w2_b = umma_desc_k_sw128(smem_base + W2_OFF);
            const uint32_t acc1 = tmem_base, acc2 = tmem_base + NN;
            int stage = 0;
            uint32_t phase = 0;
            auto mma1 = [&]() {
                for (int kb = 0; kb < kblks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t soff = (uint64_t)(stage * (STAGE >> 4));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16_2sm(acc1, ring_a + soff + 2 * k, ring_b + soff + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_2sm(&empty_bar[stage], (uint16_t)3);
                        if (kb + 1 == kblks) umma_commit_2sm(&acc1_full, (uint16_t)3);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            };
            mbar_wait(&w2_full, 0);
            uint32_t hph = 0, eph = 0;
            if (pair < tiles) mma1();
            for (int tile = pair; tile < tiles; tile += npairs) {
                mbar_wait(&acc2_empty, eph ^ 1);    // the second epilogue of the previous tile has drained acc2
                eph ^= 1;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    mbar_wait(&h_full[c], hph);     // columns [64c, 64c+64) of H(tile) are packed in TMEM of both CTAs
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)      // 16 columns of H = 8 TMEM columns, +32 B along K in W2
                            umma_bf16_ts_2sm(acc2, acc1 + h_col(c) + 8 * k, w2_b + (uint64_t)(c * (CHUNK >> 4)) + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
                        if (c == 3) umma_commit_2sm(&acc2_full, (uint16_t)3);
                    }
                    __syncwarp();
                }
                hph ^= 1;
                // the tensor pipe executes in issue order: MMA1 of the next tile overwrites acc1 (and H in it) only after the second
                // product above has consumed it
                if (tile + npairs < tiles) mma1();
            }
        }
    } else if (warp < 2 + EPI_WARPS) {
        // ---------------- first epilogue: acc1 -> + bias, ReLU -> bf16, packed in place (A operand of the second product) ----------------
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * 128;
        uint32_t p1 = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            mbar_wait(&acc1_full, p1);
            p1 ^= 1;
            tc_fence_after();
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
                for (int i = 0; i < 32; i += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(bias_s + n0 + i);
                    pk[i / 2] = pack_bf16(fmaxf(__uint_as_float(r[i]) + b.x, 0.f), fmaxf(__uint_as_float(r[i + 1]) + b.y, 0.f));
                    pk[i / 2 + 1] = pack_bf16(fmaxf(__uint_as_float(r[i + 2]) + b.z, 0.f), fmaxf(__uint_as_float(r[i + 3]) + b.w, 0.f));
                }
                tmem_st_32x16(t_acc + j * 16, pk);      // H columns n0 .. n0+31 -> TMEM columns half*128 + 16 j .. +15 (already read)
                if (j & 1) {                              // H chunk 2 half + j / 2 complete in this warp
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&h_full[2 * half + (j >> 1)]);
                }
            }
        }
    } else if (warp < STORE_WARP) {
        // ---------------- second epilogue: acc2 -> bf16 -> staging tile (128B-swizzled 64-column chunks) ----------------
        const int quarter = warp & 3, half = (warp - 2 - EPI_WARPS) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + NN + half * 128;
        uint8_t* srow0 = smem + OUT_OFF + row * 128;
        const int sw = row & 7;
        uint32_t p2 = 0, fph = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            mbar_wait(&acc2_full, p2);
            p2 ^= 1;
            tc_fence_after();
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(t_acc, ra);
            mbar_wait(&stg_free, fph ^ 1);          // the TMA stores of the previous tile have read the staging tile
            fph ^= 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t* r = (j & 1) ? rb : ra;
                tmem_ld_wait();
                if (j + 1 < 4) {
                    tmem_ld_32x32(t_acc + (j + 1) * 32, (j & 1) ? ra : rb);
                } else {                              // acc2 is in registers: the next tile's second product may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&acc2_empty);
                }
                const int n0 = half * 128 + j * 32;
                uint8_t* srow = srow0 + (n0 >> 6) * CHUNK;
                const int u0 = (n0 & 63) >> 3;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t* q = r + 8 * u;
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(q[0]), __uint_as_float(q[1]));
                    v.y = pack_bf16(__uint_as_float(q[2]), __uint_as_float(q[3]));
                    v.z = pack_bf16(__uint_as_float(q[4]), __uint_as_float(q[5]));
                    v.w = pack_bf16(__uint_as_float(q[6]), __uint_as_float(q[7]));
                    *reinterpret_cast<uint4*>(srow + (((u0 + u) ^ sw) << 4)) = v;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&stg_full);
        }
    } else {
        // ---------------- output stores: four TMA tensor stores per tile (rows past M are clipped by the map) ----------------
        uint32_t sph = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            const int row0 = tile * 2 * BM + rank * BM;
            mbar_wait(&stg_full, sph);
            sph ^= 1;
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < 4; ++c) tma_store_2d(&tmY, smem + OUT_OFF + c * CHUNK, c * BK, row0);
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&stg_free);
            }
            __syncwarp();
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

}  // namespace

bool gemm_b2b_supported(const GemmB2B& p, int n1, int n2) {
    return n1 == NN && n2 == NN && p.K1 % BK == 0 && p.K1 >= BK && p.lda % 8 == 0 && p.ldo % 16 == 0 &&
           (reinterpret_cast<uintptr_t>(p.out) & 31) == 0 && (reinterpret_cast<uintptr_t>(p.bias1) & 15) == 0;
}

int gemm_b2b_launch(const GemmB2B& p, cudaStream_t stream) {
    if (int rc = gemm_init()) return rc;
    VPU_REQUIRE(gemm_b2b_supported(p, NN, NN), "back-to-back head GEMM: unsupported shape (K1=%d lda=%d ldo=%d)", p.K1, p.lda, p.ldo);
    CUtensorMap tmA, tmW1, tmW2, tmY;
    if (int rc = gemm_tmap(&tmA, p.A, p.M, p.K1, p.lda, BM)) return rc;
    if (int rc = gemm_tmap(&tmW1, p.W1, NN, p.K1, p.K1, BM)) return rc;
    if (int rc = gemm_tmap(&tmW2, p.W2, NN, NN, NN, BM)) return rc;
    if (int rc = gemm_tmap(&tmY, p.out, p.M, NN, p.ldo, BM)) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_b2b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    const int tiles = (p.M + 2 * BM - 1) / (2 * BM);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = stream;
    static int max_clusters = 0;
    if (max_clusters == 0) {
        cfg.gridDim = dim3(2 * (gemm_num_sms() / 2));
        int n = 0;
        VPU_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_b2b_kernel, &cfg));
        max_clusters = n > 0 ? n : 1;
    }
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    cfg.gridDim = dim3(2 * clusters);
    B2BArgs a{p.bias1, p.M, p.K1};
    VPU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_b2b_kernel, tmA, tmW1, tmW2, tmY, a));
    count_launch();
    return 0;
}

}  // namespace vpu
