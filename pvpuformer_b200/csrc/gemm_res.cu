// Residual-stream GEMM of the ViT blocks (reference models_vit.py:72-75, x + proj(attn(..)) and x + fc2(..)):
//     X[M, N] (fp32, in place) = X + A[M, K] W[N, K]^T + bias,   Xn = bf16(X),   LayerNorm statistics slots of the new rows
// as a kernel whose epilogue moves every byte with TMA.  The generic epilogue (gemm.cu, EK_F32_RES_LNOUT) has its eight warps
// request the fp32 residual rows, transpose the accumulator through shared memory and store fp32 + bf16 rows themselves; proj
// (K = N) ran at 63 % of the HBM rate with it although simple streaming kernels reach 95 % on the same board.  Here
//   * a loader / store warp keeps the residual of the NEXT column groups in flight ahead of the MMAs (TMA boxes of 32 fp32
//     columns x 128 rows into 128-byte-swizzled staging tiles), independent of the epilogue warps' progress,
//   * an epilogue thread (= accumulator row, straight out of TMEM) adds its staged residual row and the bias, writes the fp32
//     result back in place and the bf16 copy into a second staging tile, and accumulates the row's sum / sum of squares,
//   * the store warp sends fp32 and bf16 tiles back with TMA tensor stores and re-arms the staging pair with the next residual.
// 2-CTA pairs (tcgen05.mma.cta_group::2, 256-row tiles, 256 columns), two TMEM accumulators, a 4-stage operand ring.
// Statistics slots: one per (128-column group, column half of the 64-column pairs), computed in the generic epilogue's grouping
// and order, so that both kernels -- the choice depends on M -- produce the same bits.
// A second mode shares the skeleton (Dual-cross Merging Attention, image side, reference transformer.py:444-449, 456-458):
//   MODE_TAB    out (bf16) = A W^T + table[m mod R]      the fused K|V|Q projection of the image tokens, whose positional term
//               key_pe W^T + b is a precomputed fp32 table (R + 128 rows: the first 128 repeated, so one box never wraps);
//               106 -> 87 us at 50176 x 1152 x 768.  (A third mode for the image -> tokens out-projection, fp32 out = A W^T +
//               bias + bf16 residual with the residual tile loaded by TMA, measured 59.9 against 59.6 us for the generic epilogue
//               and was removed.)
#include "gemm.cuh"
#include "tc_attn.cuh"

namespace vpu {

namespace {

constexpr int BM = 128, BK = 64, BN = 256;
constexpr int CHUNK = BM * BK * 2;              // 16 KB operand tile: [128 rows x 64 bf16], 128-byte swizzle
constexpr int STAGE = 2 * CHUNK;                // A k-chunk + B half k-chunk (128 of the tile's 256 weight rows)
constexpr int UNIT = BM * 32 * 4;               // 16 KB: [128 rows x 32 fp32]
constexpr int PAIR = 2 * UNIT + BM * 64 * 2;    // two fp32 units + one bf16 tile [128 x 64]: 48 KB
constexpr int RING_OFF = 0;
// <STAGES, NPAIR> = <4, 2>: two staging pairs in flight and a 4-stage operand ring (3 stages starved the MMAs at K = 1280: 306 us
// against 270).  <5, 1> was tried for the tensor-bound fc2 (K = 4 N): 188 us against 181 us for the generic kernel, and for the
// ViT-H proj (K = N = 1280): 345 us against 267 us -- the staging pairs (bytes of residual in flight), not the ring, set the pace.
// One loader / store warp per staging pair (so that a warp blocked in wait_group.read cannot delay the other pair) was slower too:
// 101.5 against 93 us (ViT-B), 272 against 267 us (ViT-H).
template <int STAGES, int NPAIR> __host__ __device__ constexpr int res_smem() { return STAGES * STAGE + NPAIR * PAIR + 1024; }
constexpr int EPI_WARPS = 8;
constexpr int STORE_WARP = 2 + EPI_WARPS;
constexpr int THREADS = (STORE_WARP + 1) * 32;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct ResArgs {
    const float* bias;
    float2* ln_out;
    int ln_slots;
    int M, N, K;
    int tab_rows;                 // MODE_TAB: period R of the table
};
enum { MODE_RES = 0, MODE_TAB = 1 };

// tmX: fp32 [*, N] in boxes of 32 columns (MODE_RES: residual stream in / out; MODE_TAB: the table);
// tmXn: bf16 [M, N] in boxes of 64 columns (MODE_RES: bf16 copy; MODE_TAB: the output)
template <int STAGES, int NPAIR, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
gemm_res_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmXn, const ResArgs a) {
    static_assert(res_smem<STAGES, NPAIR>() <= 232448 - 1536, "dynamic shared memory limit of sm_100");
    constexpr int PAIR_OFF = STAGES * STAGE;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2], res_full[NPAIR], pair_done[NPAIR];
    __shared__ __align__(16) float bias_s[BN];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)(cluster_ctarank() & 1);
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmXn);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 2 * EPI_WARPS);       // epilogue warps of both CTAs (leader's barrier)
        }
        for (int i = 0; i < NPAIR; ++i) {
            mbar_init(&res_full[i], 1);
            mbar_init(&pair_done[i], EPI_WARPS);
        }
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

    const int n_blks = (a.N + BN - 1) / BN;         // MODE_TAB: N = 1152 leaves a half tile (TMA zero-fills the loads and clips the stores)
    const int m_blks = (a.M + 2 * BM - 1) / (2 * BM);
    const int tiles = n_blks * m_blks;
    const int kblks = (a.K + BK - 1) / BK;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        // ---------------- TMA producer of the operands (both CTAs) ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair; tile < tiles; tile += npairs) {
            const int m_blk = tile / n_blks, n_blk = tile % n_blks;
            const int arow = m_blk * 2 * BM + rank * BM, brow = n_blk * BN + rank * (BN / 2);
            for (int kb = 0; kb < kblks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE);
                    uint8_t* st = smem + RING_OFF + stage * STAGE;
                    tma_load_2d_2sm(st, &tmA, &full_bar[stage], kb * BK, arow);
                    tma_load_2d_2sm(st + CHUNK, &tmW, &full_bar[stage], kb * BK, brow);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {  // ---------------- MMA issuer (leader CTA only) ----------------
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
            const uint64_t adesc0 = umma_desc_k_sw128(smem_base + RING_OFF), bdesc0 = umma_desc_k_sw128(smem_base + RING_OFF + CHUNK);
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
                        const uint64_t soff = (uint64_t)(stage * (STAGE >> 4));
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
        // ---------------- epilogue: thread = accumulator row; warp (quarter, half) takes the 32-column unit `half` of every 64-column pair ----------------
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane, sw = row & 7;
        int acc = 0, it = 0;
        uint32_t acc_phase = 0;
        int g = 0;                                   // running 64-column pair index of this CTA: staging slot g & 1, phase (g >> 1) & 1
        float4 bias_next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == MODE_RES && (warp == 2 || warp == 3) && pair < tiles)
            bias_next = __ldg(reinterpret_cast<const float4*>(a.bias + (pair % n_blks) * BN + (warp - 2) * 128 + lane * 4));
        for (int tile = pair; tile < tiles; tile += npairs, ++it) {
            const int m_blk = tile / n_blks, n_blk = tile % n_blks;
            const int m = m_blk * 2 * BM + rank * BM + row;
            // this tile's 256 bias values: every epilogue warp has finished the previous tile (first barrier) before they are replaced
            // (the values were requested one tile earlier: no global-memory latency between the two barriers)
            float* bs = bias_s;
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
            if (warp == 2 || warp == 3) *reinterpret_cast<float4*>(bs + (warp - 2) * 128 + lane * 4) = bias_next;
            asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
            if (MODE == MODE_RES && (warp == 2 || warp == 3) && tile + npairs < tiles)
                bias_next = __ldg(reinterpret_cast<const float4*>(a.bias + ((tile + npairs) % n_blks) * BN + (warp - 2) * 128 + lane * 4));
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * 32;
            // row statistics in exactly the grouping and order of the generic epilogue (gemm.cu): per 4-column piece u of a 32-column
            // chunk a partial, the two chunks of a 128-column group added per piece, the eight pieces folded as the xor-4 / 2 / 1
            // butterfly does -- small batches run the generic kernel (128-wide tiles) and must produce the same bits
            float s8[8], q8[8];
#pragma unroll 1
            for (int p = 0; p < 4; ++p, ++g) {
                const int slot = g % NPAIR;
                uint8_t* ps = smem + PAIR_OFF + slot * PAIR;
                uint32_t r[32];
                tmem_ld_32x32(t_acc + p * 64, r);
                mbar_wait(&res_full[slot], (uint32_t)((g / NPAIR) & 1));      // both fp32 residual units of this pair have landed
                tmem_ld_wait();
                if (p == 3) {                        // the accumulator is in registers: the MMAs of the tile after next may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&acc_empty[acc]);
                }
                uint8_t* frow = ps + half * UNIT + row * 128;
                uint8_t* brow = ps + 2 * UNIT + row * 128;
                const float* bb = bs + p * 64 + half * 32;
                if constexpr (MODE == MODE_RES) {
                    uint32_t pk[16];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float4* q = reinterpret_cast<float4*>(frow + ((u ^ sw) << 4));
                        float4 x = *q;
                        const float4 bv = *reinterpret_cast<const float4*>(bb + 4 * u);
                        x.x += __uint_as_float(r[4 * u]) + bv.x;
                        x.y += __uint_as_float(r[4 * u + 1]) + bv.y;
                        x.z += __uint_as_float(r[4 * u + 2]) + bv.z;
                        x.w += __uint_as_float(r[4 * u + 3]) + bv.w;
                        *q = x;
                        const float ps_ = (x.x + x.y) + (x.z + x.w), pq_ = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, x.w * x.w)));
                        s8[u] = (p & 1) ? s8[u] + ps_ : ps_;
                        q8[u] = (p & 1) ? q8[u] + pq_ : pq_;
                        pk[2 * u] = pack_bf16(x.x, x.y);
                        pk[2 * u + 1] = pack_bf16(x.z, x.w);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        *reinterpret_cast<uint4*>(brow + (((4 * half + u) ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                } else if constexpr (MODE == MODE_TAB) {
                    uint32_t pk[16];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 x = *reinterpret_cast<const float4*>(frow + ((u ^ sw) << 4));     // table row (accumulator + table: the generic order)
                        pk[2 * u] = pack_bf16(__uint_as_float(r[4 * u]) + x.x, __uint_as_float(r[4 * u + 1]) + x.y);
                        pk[2 * u + 1] = pack_bf16(__uint_as_float(r[4 * u + 2]) + x.z, __uint_as_float(r[4 * u + 3]) + x.w);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        *reinterpret_cast<uint4*>(brow + (((4 * half + u) ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pair_done[slot]);
                if (MODE == MODE_RES && (p & 1)) {   // a 128-column group of the row is complete: its slot
                    const float sum = ((s8[0] + s8[4]) + (s8[2] + s8[6])) + ((s8[1] + s8[5]) + (s8[3] + s8[7]));
                    const float sq = ((q8[0] + q8[4]) + (q8[2] + q8[6])) + ((q8[1] + q8[5]) + (q8[3] + q8[7]));
                    if (m < a.M) a.ln_out[(size_t)m * a.ln_slots + (n_blk * 2 + (p >> 1)) * 2 + half] = make_float2(sum, sq);
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {
        // ---------------- residual loads and output stores (converged warp, one elected lane) ----------------
        // pair index g -> (tile iteration g / 4, 64-column pair g % 4); the residual of pair g + 2 is requested as soon as the stores
        // of pair g have read their staging tiles
        int my_tiles = 0;
        for (int tile = pair; tile < tiles; tile += npairs) ++my_tiles;
        const int npairs_total = 4 * my_tiles;
        auto coords = [&](int g, int& row0, int& col0) {
            const int tile = pair + (g >> 2) * npairs;
            row0 = (tile / n_blks) * 2 * BM + rank * BM;
            col0 = (tile % n_blks) * BN + (g & 3) * 64;
        };
        auto load_res = [&](int g) {
            int row0, col0;
            coords(g, row0, col0);
            uint8_t* ps = smem + PAIR_OFF + (g % NPAIR) * PAIR;
            if constexpr (MODE == MODE_TAB) row0 %= a.tab_rows;
            mbar_arrive_expect_tx(&res_full[g % NPAIR], 2 * UNIT);
            tma_load_2d(ps, &tmX, &res_full[g % NPAIR], col0, row0);
            tma_load_2d(ps + UNIT, &tmX, &res_full[g % NPAIR], col0 + 32, row0);
        };
        if (elect_one()) {
#pragma unroll
            for (int i = 0; i < NPAIR; ++i)
                if (npairs_total > i) load_res(i);
        }
        __syncwarp();
        for (int g = 0; g < npairs_total; ++g) {
            mbar_wait(&pair_done[g % NPAIR], (uint32_t)((g / NPAIR) & 1));
            if (elect_one()) {
                int row0, col0;
                coords(g, row0, col0);
                uint8_t* ps = smem + PAIR_OFF + (g % NPAIR) * PAIR;
                if constexpr (MODE == MODE_RES) {
                    tma_store_2d(&tmX, ps, col0, row0);
                    tma_store_2d(&tmX, ps + UNIT, col0 + 32, row0);
                }
                tma_store_2d(&tmXn, ps + 2 * UNIT, col0, row0);
                tma_store_commit();
                tma_store_wait_read();
                if (g + NPAIR < npairs_total) load_res(g + NPAIR);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int make_map(CUtensorMap* tm, const void* ptr, CUtensorMapDataType dt, int elem, int rows, int cols, int ld, int box_cols) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        VPU_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        VPU_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * elem};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(tm, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPU_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (residual GEMM) failed with %d", (int)r);
    return 0;
}

}  // namespace

bool gemm_res_supported(const GemmProblem& p) {
    const Epi& e = p.epi;
    return e.ln_out && e.ln_out_bf16 && e.res == e.out && !e.res_bf16 && !e.out_bf16 && e.bias && p.N % BN == 0 && p.K % 8 == 0 && e.ldo == e.ldr &&
           e.ln_slots == (p.N / 128) * 2 && e.ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(e.ln_out_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0;
}

static bool plain_epilogue(const Epi& e) {
    return e.mode == EPI_PLAIN && e.act == ACT_NONE && !e.ln_out && !e.ln_in && !e.gn_in && !e.gn_out && e.out;
}

// out (bf16) = A W^T + table[m mod R]; the table holds R + 128 rows (Epi::bias2d_pad_rows)
bool gemm_tab_supported(const GemmProblem& p) {
    const Epi& e = p.epi;
    return plain_epilogue(e) && e.bias2d && e.bias2d_rows > 0 && e.bias2d_pad_rows >= BM && e.out_bf16 && !e.bias && !e.res && p.N % 64 == 0 &&
           p.K % 8 == 0 && p.K <= 2 * p.N && p.M >= 2 * BM && e.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(e.bias2d) & 15) == 0;
}

template <int STAGES, int NPAIR, int MODE>
static int launch_res(const GemmProblem& p, cudaStream_t stream) {
    const Epi& e = p.epi;
    CUtensorMap tmA, tmW, tmX, tmXn;
    if (int rc = gemm_tmap(&tmA, p.A, p.M, p.K, p.lda, BM)) return rc;
    if (int rc = gemm_tmap(&tmW, p.W, p.N, p.K, p.ldw, BM)) return rc;
    if (MODE == MODE_RES) {
        if (int rc = make_map(&tmX, e.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.M, p.N, e.ldo, 32)) return rc;
        if (int rc = make_map(&tmXn, e.ln_out_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.M, p.N, e.ldo, 64)) return rc;
    } else {
        if (int rc = make_map(&tmX, e.bias2d, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, e.bias2d_rows + e.bias2d_pad_rows, p.N, p.N, 32)) return rc;
        if (int rc = make_map(&tmXn, e.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.M, p.N, e.ldo, 64)) return rc;
    }
    constexpr int SMEM = res_smem<STAGES, NPAIR>();
    static bool attr_set = false;
    if (!attr_set) {
        VPU_CHECK_CUDA(cudaFuncSetAttribute(gemm_res_kernel<STAGES, NPAIR, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + BN - 1) / BN);
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
        VPU_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_res_kernel<STAGES, NPAIR, MODE>, &cfg));
        max_clusters = n > 0 ? n : 1;
    }
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    cfg.gridDim = dim3(2 * clusters);
    ResArgs a{e.bias, e.ln_out, e.ln_slots, p.M, p.N, p.K, e.bias2d_rows};
    VPU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_res_kernel<STAGES, NPAIR, MODE>, tmA, tmW, tmX, tmXn, a));
    count_launch();
    return 0;
}

int gemm_res_launch(const GemmProblem& p, cudaStream_t stream) {
    if (int rc = gemm_init()) return rc;
    VPU_REQUIRE(gemm_res_supported(p), "residual GEMM: unsupported problem");
    return launch_res<4, 2, MODE_RES>(p, stream);
}

int gemm_tab_launch(const GemmProblem& p, cudaStream_t stream) {
    if (int rc = gemm_init()) return rc;
    VPU_REQUIRE(gemm_tab_supported(p), "table GEMM: unsupported problem");
    return launch_res<4, 2, MODE_TAB>(p, stream);
}

}  // namespace vpu
