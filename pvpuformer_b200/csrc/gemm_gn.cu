// GroupNorm-fused neck GEMMs (reference is_vpu_model.py:55-86: 1x1 / 2x2-stride-2 convs followed by GroupNorm(1, C)) with the
// TMA-staged epilogue of gemm_res.cu:
//     out[M, N] (bf16) = rstd_s * (A W^T) - mean_s rstd_s wg[n] + bias[n]      (the producer's GroupNorm folded in: Epi::gn_in), or
//                        A W^T + bias[n],
//     gn_out[sample] += (sum, sum of squares) of the fp32 outputs               (int64 fixed point, Epi::gn_out)
// These shapes are skinny (N = 128 ... 1024, K <= 1.5 N) and HBM-bound; the generic epilogue (8 warps, per-warp transpose, row-wise
// 8-byte stores) left d4.c (M = 802 816, N = 128, K = 192) at 181 us against an HBM floor of 80.  Here an epilogue thread owns an
// accumulator row: TMEM -> fold + bias -> statistics -> bf16 -> swizzled staging tile -> TMA tensor store by a store warp.
// Same arithmetic, operation for operation, as the generic epilogue (which still runs these layers for M < 256), and the
// statistics are integer sums of the same per-(row, 4-column) partials: bit-identical whatever the kernel or the batch.
//
// PS = true: the neck's ConvTranspose2d(k=2, s=2) layers (is_vpu_model.py:57-75) as GEMM + pixel-shuffle store.  Input row
// m = (b, i, j) of a g x g grid and column n = (kh, kw, c) go to output pixel (b, 2i + kh, 2j + kw), channel c.  A CTA takes
// rows_per_cta = the largest multiple of g <= 128 input rows (112 for g = 28 / 56; the other accumulator rows are computed and
// dropped), i.e. whole grid rows, so that a 64-column group of its tile is ONE 5-D TMA box (c, kw, j, kh, b*g + i) of the NHWC output.
#include "gemm.cuh"

namespace vpu {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 8;            // 16 (four per TMEM lane quarter) measured slower: d4.c 130 -> 176 us
constexpr int UC = 256 / EPI_WARPS;     // columns of a 64-column group per epilogue warp
static_assert(UC == 32, "tmem_ld_unit");
constexpr int STORE_WARP = 2 + EPI_WARPS;
constexpr int THREADS = (STORE_WARP + 1) * 32;
constexpr int TILE16 = BM * 64 * 2;             // staging tile: [128 rows x 64 bf16], 128-byte swizzle
constexpr int NSLOT = 4;
constexpr int MAX_N = 2560, MAX_SAMPLES = 512;      // widest layer: ViT-H d4.a, 4 x 640 columns

template <int BN> struct Cfg {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2, STAGE = A_BYTES + B_BYTES;
    static constexpr int STAGES = BN == 256 ? 4 : 5;
    static constexpr int SLOT_OFF = STAGES * STAGE;
    static constexpr int SMEM = SLOT_OFF + NSLOT * TILE16 + 1024;
    static constexpr int PAIRS = BN / 64;       // 64-column groups per tile
    static_assert(STAGE % 1024 == 0 && SMEM <= 232448 - 2 * MAX_N * 4 - MAX_SAMPLES * 8 - 512, "shared memory");
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_unit(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }

struct GnArgs {
    const float* bias;
    const float* wg;              // gn_in: row sums of the folded weight
    const long long* gn_in;       // statistics of the A operand's tensor, or nullptr
    float gn_in_count;
    long long* gn_out;            // nullptr: no statistics
    int gn_rows;
    int M, N, K;
    int rows_per_cta;             // 128, or (PS) the largest multiple of ps_g <= 128
    int ps_g, ps_cout;            // PS: input grid side, output channels (N = 4 ps_cout)
};

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}

__device__ __forceinline__ void mean_rstd(const long long* sums, int sample, float count, float& rstd, float& mean_rstd_) {
    const double inv_n = 1.0 / (double)count;          // gemm.cu: gn_mean_rstd
    const double mean = (double)sums[2 * sample] * (1.0 / (double)GN_SUM_SCALE) * inv_n;
    double var = (double)sums[2 * sample + 1] * (1.0 / (double)GN_SQ_SCALE) * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    const double r = 1.0 / sqrt(var + 1e-5);
    rstd = (float)r;
    mean_rstd_ = (float)(mean * r);
}

template <int BN, bool PS>
__global__ void __launch_bounds__(THREADS, 1)
gemm_gn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
               const GnArgs a) {
    using C = Cfg<BN>;
    constexpr int STAGES = C::STAGES, PAIRS = C::PAIRS;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2], slot_done[NSLOT], slot_free[NSLOT];
    __shared__ __align__(16) float bias_s[MAX_N], wg_s[MAX_N];      // whole bias / folded-weight row sums: loaded once
    __shared__ float2 fold_s[MAX_SAMPLES];                          // (rstd, mean * rstd) of every sample of the A operand
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)(cluster_ctarank() & 1);
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmO);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 2 * EPI_WARPS);
        }
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&slot_done[i], EPI_WARPS);
            mbar_init(&slot_free[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_2sm(&tmem_base_smem, 2 * BN);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();

    const int n_blks = a.N / BN;
    const int rpc = a.rows_per_cta;
    const int m_blks = (a.M + 2 * rpc - 1) / (2 * rpc);
    const int tiles = n_blks * m_blks;
    const int kblks = (a.K + BK - 1) / BK;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs) ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            const int m_blk = tile / n_blks, n_blk = tile % n_blks;
            const int arow = (m_blk * 2 + rank) * rpc, brow = n_blk * BN + rank * (BN / 2);
            for (int kb = 0; kb < kblks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (rpc * BK * 2 + C::B_BYTES));
                    uint8_t* st = smem + stage * C::STAGE;
                    tma_load_2d_2sm(st, &tmA, &full_bar[stage], kb * BK, arow);
                    tma_load_2d_2sm(st + C::A_BYTES, &tmW, &full_bar[stage], kb * BK, brow);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {  // ---------------- MMA issuer (leader CTA only) ----------------
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
            const uint64_t adesc0 = umma_desc_k_sw128(smem_base), bdesc0 = umma_desc_k_sw128(smem_base + C::A_BYTES);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = pair; tile < tiles; tile += npairs) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < kblks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t soff = (uint64_t)(stage * (C::STAGE >> 4));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16_2sm(d_tmem, adesc0 + soff + 2 * k, bdesc0 + soff + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_2sm(&empty_bar[stage], (uint16_t)3);
                        if (kb + 1 == kblks) umma_commit_2sm(&acc_full[acc], (uint16_t)3);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp < STORE_WARP) {
        // ---------------- epilogue: thread = accumulator row; warp (quarter, sub) takes the UC-column unit `sub` of every 64-column group ----------------
        const int quarter = warp & 3, sub = (warp - 2) >> 2;
        const int row = quarter * 32 + lane, sw = row & 7;
        const int et = threadIdx.x - 64;             // 0 .. 511
        // per-launch constants: nothing below waits on global memory except the accumulator itself
        for (int i = et; i < a.N / 4; i += EPI_WARPS * 32) {
            *reinterpret_cast<float4*>(bias_s + 4 * i) = __ldg(reinterpret_cast<const float4*>(a.bias + 4 * i));
            if (a.gn_in) *reinterpret_cast<float4*>(wg_s + 4 * i) = __ldg(reinterpret_cast<const float4*>(a.wg + 4 * i));
        }
        if (a.gn_in)
            for (int s = et; s < a.M / a.gn_rows; s += EPI_WARPS * 32) {
                float r, mr;
                mean_rstd(a.gn_in, s, a.gn_in_count, r, mr);
                fold_s[s] = make_float2(r, mr);
            }
        asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
        int acc = 0, g = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            const int m_blk = n_blks == 1 ? tile : tile / n_blks, n_blk = n_blks == 1 ? 0 : tile - m_blk * n_blks;
            const int m0 = (m_blk * 2 + rank) * rpc, m = m0 + row;
            const bool live = row < rpc && m < a.M;      // PS: rows past rows_per_cta are computed and dropped
            const int sA = m0 / a.gn_rows, mB = (sA + 1) * a.gn_rows;      // rows >= mB belong to the next sample (gn_rows >= 128)
            const bool inB = m >= mB;
            float rr = 1.f, mr = 0.f;
            if (a.gn_in && live) {
                const float2 f = fold_s[sA + (inB ? 1 : 0)];
                rr = f.x; mr = f.y;
            }
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + sub * UC;
            long long st_s = 0, st_q = 0;            // this row's fixed-point partial sums (one sample per row)
#pragma unroll 1
            for (int p = 0; p < PAIRS; ++p, ++g) {
                const int slot = g % NSLOT;
                uint32_t r[UC];
                tmem_ld_unit(t_acc + p * 64, r);
                mbar_wait(&slot_free[slot], (uint32_t)(((g / NSLOT) & 1) ^ 1));      // the store of this slot's previous use has read it
                tmem_ld_wait();
                if (p == PAIRS - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&acc_empty[acc]);
                }
                uint8_t* orow = smem + C::SLOT_OFF + slot * TILE16 + row * 128;
                const float* bb = bias_s + n_blk * BN + p * 64 + sub * UC;
                const float* ww = wg_s + n_blk * BN + p * 64 + sub * UC;
                uint32_t pk[UC / 2];
#pragma unroll
                for (int u = 0; u < UC / 4; ++u) {
                    float4 v = make_float4(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 1]), __uint_as_float(r[4 * u + 2]), __uint_as_float(r[4 * u + 3]));
                    if (a.gn_in) {
                        const float4 w = *reinterpret_cast<const float4*>(ww + 4 * u);
                        v.x = fmaf(v.x, rr, -mr * w.x); v.y = fmaf(v.y, rr, -mr * w.y);
                        v.z = fmaf(v.z, rr, -mr * w.z); v.w = fmaf(v.w, rr, -mr * w.w);
                    }
                    const float4 bv = *reinterpret_cast<const float4*>(bb + 4 * u);
                    v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                    st_s += __float2ll_rn(((v.x + v.y) + (v.z + v.w)) * GN_SUM_SCALE);
                    st_q += __float2ll_rn(fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w))) * GN_SQ_SCALE);
                    pk[2 * u] = pack_bf16(v.x, v.y);
                    pk[2 * u + 1] = pack_bf16(v.z, v.w);
                }
#pragma unroll
                for (int u = 0; u < UC / 8; ++u)
                    *reinterpret_cast<uint4*>(orow + ((((UC / 8) * sub + u) ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&slot_done[slot]);
            }
            // per-sample sums of this warp's 32 rows: integer adds, any order.  Warp sums by REDUX on 16 / 16 / 32-bit pieces of the
            // 64-bit values (exact: 32 x 16-bit pieces fit 21 bits, the signed top words stay far below 2^26); the second sample's
            // sums only when the warp has such rows (a 64-bit shuffle butterfly for all four sums cost ~100 instructions per tile)
            if (!a.gn_out) { acc ^= 1; if (acc == 0) acc_phase ^= 1; continue; }
            auto warp_sum64 = [](long long v) -> long long {
                const unsigned lo = (unsigned)v & 0xffffu, mid = ((unsigned)v >> 16) & 0xffffu;
                const int hi = (int)(v >> 32);
                const unsigned slo = __reduce_add_sync(0xffffffffu, lo), smid = __reduce_add_sync(0xffffffffu, mid);
                const int shi = __reduce_add_sync(0xffffffffu, hi);
                return (long long)(((unsigned long long)(long long)shi << 32) + ((unsigned long long)smid << 16) + slo);
            };
            const bool anyB = __any_sync(0xffffffffu, live && inB);
            const long long s0 = warp_sum64((live && !inB) ? st_s : 0), q0 = warp_sum64((live && !inB) ? st_q : 0);
            long long s1 = 0, q1 = 0;
            if (anyB) { s1 = warp_sum64((live && inB) ? st_s : 0); q1 = warp_sum64((live && inB) ? st_q : 0); }
            if (lane == 0 && m0 + quarter * 32 < a.M) {
                unsigned long long* accp = reinterpret_cast<unsigned long long*>(a.gn_out) + 2 * sA;
                if (s0 | q0) { atomicAdd(accp, (unsigned long long)s0); atomicAdd(accp + 1, (unsigned long long)q0); }
                if (s1 | q1) { atomicAdd(accp + 2, (unsigned long long)s1); atomicAdd(accp + 3, (unsigned long long)q1); }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {
        // ---------------- output stores: one TMA tensor store per 64-column group (rows past M are clipped by the map) ----------------
        int g = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            const int row0 = ((tile / n_blks) * 2 + rank) * rpc, col0 = (tile % n_blks) * BN;
            for (int p = 0; p < PAIRS; ++p, ++g) {
                const int slot = g % NSLOT;
                mbar_wait(&slot_done[slot], (uint32_t)((g / NSLOT) & 1));
                if (elect_one()) {
                    if constexpr (PS) {
                        const int n0 = col0 + p * 64, q = n0 / a.ps_cout;      // (kh, kw) block: a 64-column group never straddles two
                        tma_store_5d(&tmO, smem + C::SLOT_OFF + slot * TILE16, n0 - q * a.ps_cout, q & 1, 0, q >> 1, row0 / a.ps_g);
                    } else {
                        tma_store_2d(&tmO, smem + C::SLOT_OFF + slot * TILE16, col0 + p * 64, row0);
                    }
                    tma_store_commit();
                    if (g > 0) {                     // one store stays in flight: the previous one has read its tile
                        tma_store_wait_read1();
                        mbar_arrive(&slot_free[(g - 1) % NSLOT]);
                    }
                }
                __syncwarp();
            }
        }
        if (elect_one()) {
            tma_store_wait_read();
            if (g > 0) mbar_arrive(&slot_free[(g - 1) % NSLOT]);
        }
        __syncwarp();
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 2 * BN);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NHWC output [B, 2g, 2g, cout] of the pixel-shuffle store as (c, kw, j, kh, b*g + i); box = (64, 1, g, 1, rows_per_cta / g)
int make_ps_map(CUtensorMap* tm, void* out, long long images, int g, int cout, int rows_per_cta) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        VPU_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        VPU_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t cb = (cuuint64_t)cout * 2;
    cuuint64_t gdim[5] = {(cuuint64_t)cout, 2, (cuuint64_t)g, 2, (cuuint64_t)(images * g)};
    cuuint64_t gstride[4] = {cb, 2 * cb, 2 * (cuuint64_t)g * cb, 4 * (cuuint64_t)g * cb};
    cuuint32_t box[5] = {64, 1, (cuuint32_t)g, 1, (cuuint32_t)(rows_per_cta / g)};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, out, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPU_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (pixel-shuffle store) failed with %d", (int)r);
    return 0;
}

int ps_rows_per_cta(int g) { return g <= BM ? (BM / g) * g : 0; }

template <int BN, bool PS>
int launch_gn(const GemmProblem& p, cudaStream_t stream) {
    using C = Cfg<BN>;
    const int rpc = PS ? ps_rows_per_cta(p.epi.ps_g) : BM;
    CUtensorMap tmA, tmW, tmO;
    if (int rc = gemm_tmap(&tmA, p.A, p.M, p.K, p.lda, rpc)) return rc;
    if (int rc = gemm_tmap(&tmW, p.W, p.N, p.K, p.ldw, BN / 2)) return rc;
    if (PS) {
        if (int rc = make_ps_map(&tmO, p.epi.out, p.M / ((long long)p.epi.ps_g * p.epi.ps_g), p.epi.ps_g, p.epi.ps_cout, rpc)) return rc;
    } else {
        if (int rc = gemm_tmap(&tmO, p.epi.out, p.M, p.N, p.epi.ldo, BM)) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_gn_kernel<BN, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    const int tiles = ((p.M + 2 * rpc - 1) / (2 * rpc)) * (p.N / BN);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = stream;
    static int max_clusters = 0;
    if (max_clusters == 0) {
        cfg.gridDim = dim3(2 * (gemm_num_sms() / 2));
        int n = 0;
        VPU_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_gn_kernel<BN, PS>, &cfg));
        max_clusters = n > 0 ? n : 1;
    }
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    cfg.gridDim = dim3(2 * clusters);
    GnArgs a{p.epi.bias, p.epi.gn_wg, p.epi.gn_in, p.epi.gn_in_count, p.epi.gn_out, p.epi.gn_out ? p.epi.gn_rows : (1 << 30), p.M, p.N, p.K,
             rpc, p.epi.ps_g, p.epi.ps_cout};
    VPU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_gn_kernel<BN, PS>, tmA, tmW, tmO, a));
    count_launch();
    return 0;
}

}  // namespace

bool gemm_gn_supported(const GemmProblem& p) {
    const Epi& e = p.epi;
    return e.gn_out && e.out && e.out_bf16 && e.bias && !e.res && !e.bias2d && e.act == ACT_NONE && e.mode == EPI_PLAIN && !e.ln_out && !e.ln_in &&
           (p.N == 128 || p.N % 256 == 0) && p.N <= MAX_N && p.M / (e.gn_rows > 0 ? e.gn_rows : 1) <= MAX_SAMPLES && p.K % 8 == 0 && 2 * p.K < 4 * p.N && e.ldo == p.N && e.gn_rows >= BM && p.M % e.gn_rows == 0 &&
           p.M >= 2 * BM && (!e.gn_in || (e.gn_wg && e.gn_in_count > 0.f)) && (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0 && (!e.gn_wg || (reinterpret_cast<uintptr_t>(e.gn_wg) & 15) == 0);
}

// ConvTranspose2d(k=2, s=2) as GEMM + pixel-shuffle TMA store, with or without GroupNorm statistics of the output
bool gemm_ps_tma_supported(const GemmProblem& p) {
    const Epi& e = p.epi;
    if (e.mode != EPI_PIXEL_SHUFFLE || !e.out || !e.out_bf16 || !e.bias || e.res || e.bias2d || e.act != ACT_NONE || e.ln_out || e.ln_in || e.gn_in)
        return false;
    const int g = e.ps_g, rpc = g > 0 ? ps_rows_per_cta(g) : 0;
    if (rpc < 96 || e.ps_cout % 64 != 0 || p.N != 4 * e.ps_cout || p.N % 256 != 0 || p.N > MAX_N || e.ldo != e.ps_cout) return false;
    if (p.M % (g * g) != 0 || p.M % rpc != 0 || p.M < 2 * BM || p.K % 8 != 0 || p.K > p.N) return false;
    if (e.gn_out && (e.gn_rows < BM || p.M % e.gn_rows != 0 || p.M / e.gn_rows > MAX_SAMPLES)) return false;
    return (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0;
}

int gemm_gn_launch(const GemmProblem& p, cudaStream_t stream) {
    if (int rc = gemm_init()) return rc;
    VPU_REQUIRE(gemm_gn_supported(p), "GroupNorm-fused GEMM (TMA epilogue): unsupported problem");
    return p.N == 128 ? launch_gn<128, false>(p, stream) : launch_gn<256, false>(p, stream);
}

int gemm_ps_tma_launch(const GemmProblem& p, cudaStream_t stream) {
    if (int rc = gemm_init()) return rc;
    VPU_REQUIRE(gemm_ps_tma_supported(p), "pixel-shuffle GEMM (TMA epilogue): unsupported problem");
    return launch_gn<256, true>(p, stream);
}

}  // namespace vpu
