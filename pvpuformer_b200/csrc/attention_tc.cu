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
//   out = O_t / rowsum   tcgen05.ld epilogue -> bf16 rows in a 128-byte-swizzled staging tile -> one TMA tensor store per tile
//                        (5-D window box, issued by the store warp), so the window un-partition never exists either
// Scores and probabilities never touch shared or global memory.  The MMA warp software-pipelines
// S(n+1) between the two P V products of problem n, and Q/K/V of the next problem are prefetched into
// the second shared-memory stage while the current one is in flight.
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "attention.cuh"
#include "tc_attn.cuh"

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
constexpr int OUT_STAGE_BYTES = 128 * D * 2;                    // 16 KB: one query tile of bf16 output rows, staged for the TMA store
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * OUT_STAGE_BYTES + 1024;
constexpr int THREADS = 352;          // warp 0 TMA, warp 1 MMA, warps 2-5 softmax tile 0, warps 6-9 softmax tile 1, warp 10 output TMA stores
constexpr int STORE_WARP = 10;
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
// shared -> global tile store through a tensor map (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
static unsigned long long* g_trace = nullptr;     // vpu_debug_attention_trace
static int g_trace_cap = 0;

struct WinArgs {
    __nv_bfloat16* o;
    int ldo;
    int heads, nwin_side, grid, tokens;   // 12, 2, 28, 784
    int nprob;                            // images * windows * heads
    int qcol, kcol, vcol;                 // column of head 0 in the fused projection buffer
    float scale_log2;
    unsigned long long* trace;            // measurement only (-DVPU_ATTN_DEBUG + vpu_debug_attention_trace): CTA 0 logs (event << 56 | clock)
    int trace_cap;
};
#ifdef VPU_ATTN_DEBUG
#define W_TRACE(role, code)                                                                              \
    do {                                                                                                 \
        if (a.trace && blockIdx.x == 0 && tr_n < a.trace_cap)                                            \
            a.trace[(role) * a.trace_cap + tr_n++] = ((unsigned long long)(code) << 56) | (clock64() & 0xFFFFFFFFFFFFFFull); \
    } while (0)
#else
#define W_TRACE(role, code) do { } while (0)
#endif

__global__ void __launch_bounds__(THREADS, 1)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmQ0,
                           const __grid_constant__ CUtensorMap tmQ1, const __grid_constant__ CUtensorMap tmO0,
                           const __grid_constant__ CUtensorMap tmO1, const WinArgs a) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], s_full[2], p_full[2], o_full[2], s_empty[2], stg_full[2], stg_free[2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int tr_n = 0;
    (void)tr_n;

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
        tma_prefetch_desc(&tmO0);
        tma_prefetch_desc(&tmO1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&p_full[t], 4);
            mbar_init(&o_full[t], 1);
            mbar_init(&s_empty[t], 4);
            mbar_init(&stg_full[t], 4);
            mbar_init(&stg_free[t], 1);
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
    pdl_wait();     // the set-up above overlapped the tail of the previous kernel; global memory is touched only below
    const int nw2 = a.nwin_side * a.nwin_side;

    if (warp == 0) {
        // ---------------- TMA producer (converged warp, one elected lane issues; see common.cuh elect_one) ----------------
        int stage = 0;
        uint32_t phase = 0;
        for (int p = blockIdx.x; p < a.nprob; p += gridDim.x) {
            const int h = p % a.heads, w = (p / a.heads) % nw2, b = p / (a.heads * nw2);
            const int wi = w / a.nwin_side, wj = w % a.nwin_side;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[stage], TX_BYTES);
                uint8_t* st = smem + stage * STAGE_BYTES;
                tma_load_5d(st, &tmQ0, &full_bar[stage], a.qcol + h * D, 0, wj, wi * WIN, b);
                tma_load_5d(st + Q_TILE_BYTES, &tmQ1, &full_bar[stage], a.qcol + h * D, 0, wj, wi * WIN + T0_IROWS, b);
                tma_load_5d(st + 2 * Q_TILE_BYTES, &tmKV, &full_bar[stage], a.kcol + h * D, 0, wj, wi * WIN, b);
                tma_load_5d(st + 2 * Q_TILE_BYTES + KV_BYTES, &tmKV, &full_bar[stage], a.vcol + h * D, 0, wj, wi * WIN, b);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (converged warp, one elected lane; descriptors advanced by adds) ----------------
        // 34 tcgen05.mma per problem: under `if (lane == 0)` each cost ~130 clk of issue (elect / BRA.U.ANY loop + descriptor
        // rebuild from a lone warp), more than the softmax of the problem.
        constexpr uint32_t idesc_s = idesc_bf16(128, SKP, false);
        constexpr uint32_t idesc_o = idesc_bf16(128, D, true);
        const uint64_t qdesc0 = umma_desc_k_sw128(smem_base), kdesc0 = umma_desc_k_sw128(smem_base + 2 * Q_TILE_BYTES);
        const uint64_t vdesc0 = umma_desc_mn_sw128(smem_base + 2 * Q_TILE_BYTES + KV_BYTES);
        auto issue_s = [&](int t, int st) {
            if (elect_one()) {
                const uint64_t so = (uint64_t)(st * (STAGE_BYTES >> 4));
                const uint64_t ad = qdesc0 + so + (uint64_t)(t * (Q_TILE_BYTES >> 4)), bd = kdesc0 + so;
#pragma unroll
                for (int k = 0; k < D / 16; ++k) umma_bf16(tmem_base + t * TILE_COLS, ad + 2 * k, bd + 2 * k, idesc_s, k ? 1u : 0u);
                umma_commit(&s_full[t]);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t, int st, bool release_stage) {
            if (elect_one()) {
                const uint64_t bd = vdesc0 + (uint64_t)(st * (STAGE_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < SKP / 16; ++k)     // 16 keys = 2048 B of the MN-major V tile, 8 TMEM columns of P
                    umma_bf16_ts(tmem_base + t * TILE_COLS + O_COL, tmem_base + t * TILE_COLS + k * 8, bd + 128 * k, idesc_o, k ? 1u : 0u);
                umma_commit(&o_full[t]);
                if (release_stage) umma_commit(&empty_bar[st]);   // every MMA reading this stage has been issued before
            }
            __syncwarp();
        };
        int stage = 0;
        uint32_t phase = 0;              // full / empty ring
        uint32_t tphase = 0;             // per-problem phase of s_full / p_full / o_full / s_empty
        int p = blockIdx.x;
        if (p < a.nprob) {               // prologue: S0, S1 of the first problem
            mbar_wait(&full_bar[0], 0);
            tc_fence_after();
            issue_s(0, 0);
            issue_s(1, 0);
        }
        for (; p < a.nprob; p += gridDim.x) {
            const bool has_next = p + (int)gridDim.x < a.nprob;
            const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
            const uint32_t nphase = (stage + 1 == STAGES) ? phase ^ 1 : phase;
            mbar_wait(&p_full[0], tphase);
            if (lane == 0) W_TRACE(1, 1);
            tc_fence_after();
            issue_pv(0, stage, false);
            if (lane == 0) W_TRACE(1, 2);
            if (has_next) {
                mbar_wait(&full_bar[nstage], nphase);
                if (lane == 0) W_TRACE(1, 10);
                mbar_wait(&s_empty[0], tphase);           // tile-0 columns free: epilogue 0 of this problem done
                if (lane == 0) W_TRACE(1, 3);
                tc_fence_after();
                issue_s(0, nstage);
                if (lane == 0) W_TRACE(1, 4);
            }
            mbar_wait(&p_full[1], tphase);
            if (lane == 0) W_TRACE(1, 5);
            tc_fence_after();
            issue_pv(1, stage, true);
            if (lane == 0) W_TRACE(1, 6);
            if (has_next) {
                mbar_wait(&s_empty[1], tphase);
                if (lane == 0) W_TRACE(1, 7);
                tc_fence_after();
                issue_s(1, nstage);
                if (lane == 0) W_TRACE(1, 8);
            }
            stage = nstage;
            phase = nphase;
            tphase ^= 1;
        }
    } else if (warp == STORE_WARP) {
        // ---------------- output stores: one TMA tensor store per (problem, tile), converged warp, one elected lane ----------------
        uint32_t tphase = 0;
        for (int p = blockIdx.x; p < a.nprob; p += gridDim.x) {
            const int h = p % a.heads, w = (p / a.heads) % nw2, b = p / (a.heads * nw2);
            const int wi = w / a.nwin_side, wj = w % a.nwin_side;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                mbar_wait(&stg_full[t], tphase);
                if (elect_one()) {
                    tma_store_5d(t == 0 ? &tmO0 : &tmO1, smem + STAGES * STAGE_BYTES + t * OUT_STAGE_BYTES, h * D, 0, wj,
                                 wi * WIN + (t == 0 ? 0 : T0_IROWS), b);
                    tma_store_commit();
                    tma_store_wait_read();
                    mbar_arrive(&stg_free[t]);
                }
                __syncwarp();
            }
            tphase ^= 1;
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
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
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 1);
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
                if (quarter == 2 && lane == 0) W_TRACE(2 + t, 2);
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
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 3);
            // epilogue: O / rowsum -> bf16 -> token-major output row of this query.  O goes to registers and the tile's TMEM
            // columns are released BEFORE the rows are written out: the write-out (1300 clk as direct stores, 700 through the
            // staging tile) is then off the MMA warp's critical path (next S product), clock64 trace of round 1i
            mbar_wait(&o_full[t], tphase);
            tc_fence_after();
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 4);
            uint32_t o0[32], o1[32];
            tmem_ld_32x32(tile_tmem + O_COL, o0);
            tmem_ld_32x32(tile_tmem + O_COL + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[t]);
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 7);
            // rows -> bf16 -> this tile's staging buffer in the 128-byte-swizzled layout; the store warp sends it with ONE TMA
            // tensor store through the same 5-D window map shape as the loads.  A thread owns a row, so direct global stores
            // touch 32 lines per instruction and the L1 store path charges per line: 1300 clk per tile on the serial chain
            // (clock64 trace, tools/attn_trace_win.py).
            const float inv = warp_has_rows ? 1.0f / sum : 0.f;
            uint8_t* stg = smem + STAGES * STAGE_BYTES + t * OUT_STAGE_BYTES;
            mbar_wait(&stg_free[t], tphase ^ 1);          // the TMA engine has read the previous problem's rows
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 8);
            if (row < nvalid) {
                uint8_t* rowp = stg + row * (D * 2);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t* r = (c < 4 ? o0 : o1) + (c & 3) * 8;
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
                    v.y = pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
                    v.z = pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
                    v.w = pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
                    *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = v;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&stg_full[t]);
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 5);
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

// =====================================================================================================
// Global (un-windowed) self-attention of the ViT blocks 6 / 12 (/ 18 / 24): S = 784 tokens, head_dim 64
// (reference models_vit.py:43-56 on the full token sequence, models_vit.py:274-286).
//
// Flash-style: the 784 x 784 score matrix never exists; a CTA owns two 128-query tiles of one (image, head)
// and streams the keys / values in 7 blocks of 112 (784 = 7 x 112: no key masking anywhere).
//   work unit   (image, head, query-tile pair); 7 query tiles per problem -> 4 pairs, the last holds one tile
//   per block j S_t = Q_t K_j^T     tcgen05.mma M=128 N=112 K=64, fp32 in TMEM columns [0,112) of tile t
//               P_t = exp2(S_t - m) softmax warpgroup t (thread == query row) -> bf16 in TMEM columns [112,168)
//               O_t += P_t V_j      tcgen05.mma, A operand from TMEM, V MN-major from shared memory, O in [192,256)
// Online softmax with the lazy rescale of FlashAttention-4: the running maximum m is only raised (and O, l
// rescaled in TMEM) when a block's maximum exceeds it by more than 2^8, which is exact -- the final O / l is
// independent of m -- and makes the correction a rare warp-uniform branch instead of a per-block O round trip.
// The softmax warpgroup pulls a whole S row block into registers and releases the S columns at once, so the
// MMA warp computes S(j+1) under the exponentials of block j (a warpgroup never waits for the tensor pipe in
// steady state; the kernel is bound by the 16 ex2/clk/SM of the MUFU pipe, not by the MMAs), and the two
// query tiles of the CTA keep that pipe busy from two independent chains.
// =====================================================================================================
constexpr int G_KB = 112;                            // keys per block
constexpr int G_QT = 128;                            // query rows per tile
constexpr int G_Q_BYTES = G_QT * D * 2;              // 16 KB
constexpr int G_KV_BYTES = G_KB * D * 2;             // 14 KB
constexpr int G_STAGE_BYTES = 2 * G_KV_BYTES;        // K block + V block
constexpr int G_STAGES = 4;
constexpr int G_SMEM_BYTES = 2 * 2 * G_Q_BYTES + G_STAGES * G_STAGE_BYTES + 1024;
constexpr int G_P_COL = 112, G_O_COL = 192;
static_assert(G_SMEM_BYTES <= 232448 && SMEM_BYTES <= 232448, "dynamic shared memory limit of sm_100");
constexpr int G_MMA1_WARP = 10;                      // warp 0 TMA, 1 MMA tile 0, 2-5 / 6-9 softmax tile 0 / 1, 10 MMA tile 1
constexpr int G_THREADS = 352;

struct GlobArgs {
    __nv_bfloat16* o;
    int ldo;
    int heads, S;                 // tokens per problem (multiple of 112)
    int nblocks;                  // S / 112
    int ntiles, npairs;           // query tiles per problem, tile pairs per problem
    int nunits;                   // problems * heads * npairs
    int qcol, kcol, vcol;
    float scale_log2;
    int ablate;                   // measurement only (VPU_ATTN_ABLATE): 1 = no ex2, 2 = no PV MMAs, 4 = no S MMAs; results are garbage
    unsigned long long* trace;    // measurement only (vpu_debug_attention_trace): CTA 0 logs (event << 56 | clock) per role
    int trace_cap;
};

// Measurement hooks (clock64 trace of CTA 0, ablation of the exponentials / MMAs) are compiled in only with
// -DVPU_ATTN_DEBUG: a lone warp retires one dependent instruction per ~10 clk, so even untaken checks cost time here.
#ifdef VPU_ATTN_DEBUG
#define G_TRACE(role, code)                                                                              \
    do {                                                                                                 \
        if (a.trace && blockIdx.x == 0 && tr_n < a.trace_cap)                                            \
            a.trace[(role) * a.trace_cap + tr_n++] = ((unsigned long long)(code) << 56) | (clock64() & 0xFFFFFFFFFFFFFFull); \
    } while (0)
#define G_ABLATE(bit) (a.ablate & (bit))
#else
#define G_TRACE(role, code) do { } while (0)
#define G_ABLATE(bit) 0
#endif

__global__ void __launch_bounds__(G_THREADS, 1)
global_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const GlobArgs a) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t q_full[2], q_empty[2], kv_full[G_STAGES], kv_empty[G_STAGES], s_full[2], s_free[2], p_full[2], pv_done[2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t KV_OFF = 4 * G_Q_BYTES;       // after the two double-buffered Q tile pairs
    int tr_n = 0;
    (void)tr_n;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 2);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&pv_done[i], 1);
        }
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 2);     // one commit per MMA warp (tile)
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
    pdl_wait();     // the set-up above overlapped the tail of the previous kernel; global memory is touched only below
    const int last_tile_pair = a.npairs - 1;
    const bool odd_tiles = (a.ntiles & 1) != 0;      // the last pair of every problem holds a single tile

    if (warp == 0) {
        // ---------------- TMA producer (converged warp, one elected lane issues) ----------------
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int u = blockIdx.x; u < a.nunits; u += gridDim.x, ++it) {
            const int pr = u % a.npairs, bh = u / a.npairs, h = bh % a.heads, b = bh / a.heads;
            const bool two = !(odd_tiles && pr == last_tile_pair);
            const int qb = it & 1;
            mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&q_full[qb], (two ? 2 : 1) * G_Q_BYTES);
                uint8_t* qs = smem + qb * 2 * G_Q_BYTES;
                tma_load_2d(qs, &tmQ, &q_full[qb], a.qcol + h * D, b * a.S + (2 * pr) * G_QT);
                if (two) tma_load_2d(qs + G_Q_BYTES, &tmQ, &q_full[qb], a.qcol + h * D, b * a.S + (2 * pr + 1) * G_QT);
            }
            __syncwarp();
            for (int j = 0; j < a.nblocks; ++j) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                G_TRACE(0, 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&kv_full[stage], G_STAGE_BYTES);
                    uint8_t* st = smem + KV_OFF + stage * G_STAGE_BYTES;
                    tma_load_2d(st, &tmK, &kv_full[stage], a.kcol + h * D, b * a.S + j * G_KB);
                    tma_load_2d(st + G_KV_BYTES, &tmV, &kv_full[stage], a.vcol + h * D, b * a.S + j * G_KB);
                }
                __syncwarp();
                if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 || warp == G_MMA1_WARP) {
        // ---------------- MMA issuers: one warp per query tile ----------------
        // A lone warp retires a dependent instruction every ~10 clk, so everything this role executes per key block
        // sits on the kernel's critical path (clock64 trace, tools/attn_trace.py: one warp issuing both tiles' 22 MMAs
        // through `if (lane == 0)` -- where nvcc wraps every tcgen05.mma in an elect / BRA.U.ANY loop and rebuilds the
        // descriptors on the uniform datapath -- needed 3500 clk per block, twice the exponentials).  Hence: one issuer
        // warp per tile, converged, one ELECTed lane, descriptors built once and advanced by adds.
        const int t = warp == 1 ? 0 : 1;
        constexpr uint32_t idesc_s = idesc_bf16(128, G_KB, false);
        constexpr uint32_t idesc_o = idesc_bf16(128, D, true);
        const uint32_t s_tmem = tmem_base + t * TILE_COLS, p_tmem = s_tmem + G_P_COL, o_tmem = s_tmem + G_O_COL;
        const uint64_t qdesc0 = umma_desc_k_sw128(smem_base + t * G_Q_BYTES);             // + qb * (2 Q tiles)
        const uint64_t kdesc0 = umma_desc_k_sw128(smem_base + KV_OFF);                     // + stage * (K + V block)
        const uint64_t vdesc0 = umma_desc_mn_sw128(smem_base + KV_OFF + G_KV_BYTES);
        auto issue_s = [&](int qb, int st) {
            if (elect_one()) {
                const uint64_t ad = qdesc0 + (uint64_t)(qb * (2 * G_Q_BYTES >> 4)), bd = kdesc0 + (uint64_t)(st * (G_STAGE_BYTES >> 4));
                if (!G_ABLATE(4)) {
#pragma unroll
                    for (int k = 0; k < D / 16; ++k) umma_bf16(s_tmem, ad + 2 * k, bd + 2 * k, idesc_s, k ? 1u : 0u);   // +32 B along K
                }
                umma_commit(&s_full[t]);
            }
            __syncwarp();
        };
        int stage = 0;
        uint32_t phase = 0, pph = 0, fph = 0;          // K/V ring, p_full[t], s_free[t]
        int it = 0;
        int u = blockIdx.x;
        if (u < a.nunits && (t == 0 || !(odd_tiles && (u % a.npairs) == last_tile_pair))) {   // prologue: S(0) of the first unit
            mbar_wait(&q_full[0], 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
        }
        for (; u < a.nunits; u += gridDim.x, ++it) {
            const bool two = !(odd_tiles && (u % a.npairs) == last_tile_pair);
            const bool valid = t == 0 || two;
            const int un = u + (int)gridDim.x;
            const bool valid_next = un < a.nunits && (t == 0 || !(odd_tiles && (un % a.npairs) == last_tile_pair));
            const int ncommit = (t == 0 && !two) ? 2 : 1;     // K/V and Q buffers are released by two arrivals
            for (int j = 0; j < a.nblocks; ++j) {
                const int nstage = (stage + 1 == G_STAGES) ? 0 : stage + 1;
                const uint32_t nphase = (stage + 1 == G_STAGES) ? phase ^ 1 : phase;
                const bool last = j + 1 == a.nblocks;
                // (1) S of the successor block as soon as the softmax warpgroup holds S(j) in registers: the score MMA of
                //     block j+1 runs under the exponentials of block j
                if (valid) {
                    mbar_wait(&s_free[t], fph);
                    fph ^= 1;
                    if (lane == 0 && t == 0) G_TRACE(1, 1 + t);
                }
                if (last ? valid_next : valid) {
                    if (last) mbar_wait(&q_full[(it + 1) & 1], ((it + 1) >> 1) & 1);
                    mbar_wait(&kv_full[nstage], nphase);
                    tc_fence_after();
                    issue_s(last ? (it + 1) & 1 : it & 1, nstage);
                    if (lane == 0 && t == 0) G_TRACE(1, 7 + t);
                }
                // (2) O += P(j) V(j) once the probabilities are in TMEM; the commits behind it also release the K / V stage
                //     and, on the last block, the Q buffer (every MMA of this tile that reads them is older)
                if (valid) {
                    mbar_wait(&p_full[t], pph);
                    pph ^= 1;
                    tc_fence_after();
                    if (lane == 0 && t == 0) G_TRACE(1, 5 + t);
                    if (elect_one()) {
                        if (!G_ABLATE(2)) {
                            const uint64_t bd = vdesc0 + (uint64_t)(stage * (G_STAGE_BYTES >> 4));
#pragma unroll
                            for (int k = 0; k < G_KB / 16; ++k)   // 16 keys = 2048 B of the MN-major V block, 8 TMEM columns of P
                                umma_bf16_ts(o_tmem, p_tmem + k * 8, bd + 128 * k, idesc_o, (j == 0 && k == 0) ? 0u : 1u);
                        }
                        umma_commit(&pv_done[t]);
                        umma_commit(&kv_empty[stage]);
                        if (ncommit == 2) umma_commit(&kv_empty[stage]);
                        if (last) {
                            umma_commit(&q_empty[it & 1]);
                            if (ncommit == 2) umma_commit(&q_empty[it & 1]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0 && t == 0) G_TRACE(1, 9 + t);
                }
                stage = nstage;
                phase = nphase;
            }
        }
    } else {  // ---------------- softmax + epilogue warpgroups ----------------
        const int t = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tile_tmem = tmem_base + ((uint32_t)(quarter * 32) << 16) + t * TILE_COLS;
        uint32_t sph = 0, vph = 0;
        for (int u = blockIdx.x; u < a.nunits; u += gridDim.x) {
            const int pr = u % a.npairs, bh = u / a.npairs, h = bh % a.heads, b = bh / a.heads;
            if (t == 1 && odd_tiles && pr == last_tile_pair) continue;
            const int q0 = (2 * pr + t) * G_QT;               // first query of this tile inside the problem
            const bool warp_has_rows = q0 + quarter * 32 < a.S;
            float m = 0.f, l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
            for (int j = 0; j < a.nblocks; ++j) {
                mbar_wait(&s_full[t], sph);
                sph ^= 1;
                tc_fence_after();
                if (quarter == 2 && lane == 0) G_TRACE(2 + t, 1);
                uint32_t r0[32], r1[32], r2[32], r3[16];
                if (warp_has_rows) {
                    tmem_ld_32x32(tile_tmem, r0);
                    tmem_ld_32x32(tile_tmem + 32, r1);
                    tmem_ld_32x32(tile_tmem + 64, r2);
                    tmem_ld_32x16(tile_tmem + 96, r3);
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_free[t]);      // S(j) is in registers: the MMA warp may overwrite it with S(j+1)
                if (quarter == 2 && lane == 0) G_TRACE(2 + t, 2);
                bool rescale = false;
                float alpha = 1.0f;
                uint32_t pk[16];
                const float sc = a.scale_log2;
                auto expo = [&](uint32_t sbits) {
                    const float x = fmaf(__uint_as_float(sbits), sc, -m);
                    return G_ABLATE(1) ? x : ex2_approx(x);
                };
                if (warp_has_rows) {
                    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r0[i + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r0[i + 2]), __uint_as_float(r0[i + 3])));
                        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(r1[i]), __uint_as_float(r1[i + 1])));
                        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(r1[i + 2]), __uint_as_float(r1[i + 3])));
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r2[i]), __uint_as_float(r2[i + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r2[i + 2]), __uint_as_float(r2[i + 3])));
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(r3[i]), __uint_as_float(r3[i + 1])));
                        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(r3[i + 2]), __uint_as_float(r3[i + 3])));
                    }
                    const float mb = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * a.scale_log2;
                    if (j == 0) {
                        m = mb;
                    } else {
                        const bool need = mb - m > 8.0f;
                        rescale = __any_sync(0xffffffffu, need) != 0;   // rare: raise the maximum, rescale l here and O below
                        if (rescale) {
                            alpha = need ? ex2_approx(m - mb) : 1.0f;
                            if (need) m = mb;
                            l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float p0 = expo(r0[2 * i]), p1 = expo(r0[2 * i + 1]), p2 = expo(r0[2 * i + 2]), p3 = expo(r0[2 * i + 3]);
                        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
                        pk[i] = pack_bf16(p0, p1); pk[i + 1] = pack_bf16(p2, p3);
                    }
                }
                if (quarter == 2 && lane == 0) G_TRACE(2 + t, 3);
                if (j > 0) {     // PV(j-1) complete (issued more than a quarter block of exponentials ago): P may be overwritten, O rescaled
                    mbar_wait(&pv_done[t], vph);
                    vph ^= 1;
                    tc_fence_after();
                }
                if (warp_has_rows) {
                    if (rescale) {
#pragma unroll 1
                        for (int c = 0; c < 2; ++c) {
                            uint32_t o[32];
                            tmem_ld_32x32(tile_tmem + G_O_COL + c * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st_32x32(tile_tmem + G_O_COL + c * 32, o);
                        }
                    }
                    tmem_st_32x16(tile_tmem + G_P_COL, pk);
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float p0 = expo(r1[2 * i]), p1 = expo(r1[2 * i + 1]), p2 = expo(r1[2 * i + 2]), p3 = expo(r1[2 * i + 3]);
                        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
                        pk[i] = pack_bf16(p0, p1); pk[i + 1] = pack_bf16(p2, p3);
                    }
                    tmem_st_32x16(tile_tmem + G_P_COL + 16, pk);
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float p0 = expo(r2[2 * i]), p1 = expo(r2[2 * i + 1]), p2 = expo(r2[2 * i + 2]), p3 = expo(r2[2 * i + 3]);
                        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
                        pk[i] = pack_bf16(p0, p1); pk[i + 1] = pack_bf16(p2, p3);
                    }
                    tmem_st_32x16(tile_tmem + G_P_COL + 32, pk);
                    uint32_t pk8[8];
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const float p0 = expo(r3[2 * i]), p1 = expo(r3[2 * i + 1]), p2 = expo(r3[2 * i + 2]), p3 = expo(r3[2 * i + 3]);
                        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
                        pk8[i] = pack_bf16(p0, p1); pk8[i + 1] = pack_bf16(p2, p3);
                    }
                    tmem_st_32x8(tile_tmem + G_P_COL + 48, pk8);
                }
                if (quarter == 2 && lane == 0) G_TRACE(2 + t, 4);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[t]);
                if (quarter == 2 && lane == 0) G_TRACE(2 + t, 5);
            }
            // epilogue: O / l -> bf16 -> token-major output row of this query
            mbar_wait(&pv_done[t], vph);
            vph ^= 1;
            tc_fence_after();
            const int q = q0 + row;
            const float inv = 1.0f / ((l0 + l1) + (l2 + l3));
            __nv_bfloat16* dst = a.o + ((size_t)b * a.S + q) * a.ldo + h * D;
            if (warp_has_rows) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(tile_tmem + G_O_COL + c * 32, r);
                    tmem_ld_wait();
                    if (q < a.S) {
                        store16_bf16(dst + c * 32, r, inv);
                        store16_bf16(dst + c * 32 + 16, r + 16, inv);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =====================================================================================================
// ViT-H window attention: 16 x 16-token windows (256 queries = keys per problem), head_dim 80
// (reference models_vit.py:43-56 per 224-px window with patch 14, models_vit.py:225-255, :317-319).
//
// Same dataflow as window_attention_tc_kernel (S and P live only in TMEM, O = P V with the A operand in TMEM), with
// the two differences head_dim 80 forces:
//   * 80 columns are not a swizzle span.  Every operand is staged as a 64-column part (128-byte swizzle) plus a
//     16-column part (32-byte swizzle), each by its own TMA box; S = Q K^T takes four K = 16 steps on the main
//     parts and a fifth on the tails, O = P V is an N = 64 and an N = 16 product per 16 keys (MN-major V parts).
//   * a problem needs 120 KB of operands, so the ring is split: Q / K double-buffered (their last reader is the
//     S product, issued one problem ahead), V single-buffered (reloaded as soon as the second P V product of the
//     previous problem has completed; the softmax of the next problem covers the load).
// Two full 128-query tiles per problem (8 window rows each) and 256 valid keys: no masking, no padded rows.
// TMEM per tile (256 columns): S [0,256) fp32 -> P [0,128) bf16 in place -> O [128,192) + [192,208).
// =====================================================================================================
namespace h80 {
constexpr int HD = 80, HM = 64, HT = 16;                 // head dim = main + tail columns
constexpr int HWIN = 16, HSK = HWIN * HWIN, HQT = 128, HQ_IROWS = HQT / HWIN;
constexpr int Q_MAIN = HQT * HM * 2, Q_TAIL = HQT * HT * 2, Q_TILE = Q_MAIN + Q_TAIL;       // 16 + 4 KB
constexpr int KV_MAIN = HSK * HM * 2, KV_TAIL = HSK * HT * 2, KV_BUF = KV_MAIN + KV_TAIL;   // 32 + 8 KB
constexpr int QK_STAGE = 2 * Q_TILE + KV_BUF;            // 80 KB: two Q tiles + K
constexpr int V_OFF = 2 * QK_STAGE;
constexpr int OUT_OFF = V_OFF + KV_BUF;                  // one output staging tile (main + tail part, the Q tile layout) shared by both tiles
constexpr int SMEM = OUT_OFF + Q_TILE + 1024;            // 221 KB
constexpr int O_MAIN_COL = 128, O_TAIL_COL = 192;
static_assert(Q_MAIN % 1024 == 0 && Q_TILE % 1024 == 0 && KV_MAIN % 1024 == 0 && KV_BUF % 1024 == 0, "swizzle atoms must stay aligned");
static_assert(SMEM <= 232448, "dynamic shared memory limit of sm_100");
}  // namespace h80

__global__ void __launch_bounds__(THREADS, 1)
window_attention_tc80_kernel(const __grid_constant__ CUtensorMap tmQm, const __grid_constant__ CUtensorMap tmQt,
                             const __grid_constant__ CUtensorMap tmKm, const __grid_constant__ CUtensorMap tmKt,
                             const __grid_constant__ CUtensorMap tmOm, const __grid_constant__ CUtensorMap tmOt, const WinArgs a) {
    using namespace h80;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t qk_full[2], qk_empty[2], v_full, v_empty, s_full[2], p_full[2], o_full[2], s_empty[2], stg_full, stg_free;
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int tr_n = 0;
    (void)tr_n;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQm);
        tma_prefetch_desc(&tmQt);
        tma_prefetch_desc(&tmKm);
        tma_prefetch_desc(&tmKt);
        tma_prefetch_desc(&tmOm);
        tma_prefetch_desc(&tmOt);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&qk_full[s], 1);
            mbar_init(&qk_empty[s], 1);
            mbar_init(&s_full[s], 1);
            mbar_init(&p_full[s], 4);
            mbar_init(&o_full[s], 1);
            mbar_init(&s_empty[s], 4);
        }
        mbar_init(&v_full, 1);
        mbar_init(&v_empty, 1);
        mbar_init(&stg_full, 4);
        mbar_init(&stg_free, 1);
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
    pdl_wait();
    const int nw2 = a.nwin_side * a.nwin_side;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        int stage = 0;
        uint32_t phase = 0, vph = 0;
        for (int p = blockIdx.x; p < a.nprob; p += gridDim.x) {
            const int h = p % a.heads, w = (p / a.heads) % nw2, b = p / (a.heads * nw2);
            const int wi = w / a.nwin_side, wj = w % a.nwin_side, r0 = wi * HWIN;
            const int qc = a.qcol + h * HD, kc = a.kcol + h * HD, vc = a.vcol + h * HD;
            mbar_wait(&qk_empty[stage], phase ^ 1);
            if (lane == 0) W_TRACE(0, 1);
            if (elect_one()) {
                uint64_t* bar = &qk_full[stage];
                mbar_arrive_expect_tx(bar, QK_STAGE);
                uint8_t* st = smem + stage * QK_STAGE;
                tma_load_5d(st, &tmQm, bar, qc, 0, wj, r0, b);
                tma_load_5d(st + Q_MAIN, &tmQt, bar, qc + HM, 0, wj, r0, b);
                tma_load_5d(st + Q_TILE, &tmQm, bar, qc, 0, wj, r0 + HQ_IROWS, b);
                tma_load_5d(st + Q_TILE + Q_MAIN, &tmQt, bar, qc + HM, 0, wj, r0 + HQ_IROWS, b);
                tma_load_5d(st + 2 * Q_TILE, &tmKm, bar, kc, 0, wj, r0, b);
                tma_load_5d(st + 2 * Q_TILE + KV_MAIN, &tmKt, bar, kc + HM, 0, wj, r0, b);
            }
            __syncwarp();
            mbar_wait(&v_empty, vph ^ 1);           // second P V product of the previous problem complete
            if (lane == 0) W_TRACE(0, 2);
            if (elect_one()) {
                mbar_arrive_expect_tx(&v_full, KV_BUF);
                tma_load_5d(smem + V_OFF, &tmKm, &v_full, vc, 0, wj, r0, b);
                tma_load_5d(smem + V_OFF + KV_MAIN, &tmKt, &v_full, vc + HM, 0, wj, r0, b);
            }
            __syncwarp();
            vph ^= 1;
            if (++stage == 2) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (converged warp, one elected lane) ----------------
        constexpr uint32_t idesc_s = idesc_bf16(128, HSK, false);
        constexpr uint32_t idesc_om = idesc_bf16(128, HM, true), idesc_ot = idesc_bf16(128, HT, true);
        const uint64_t qm0 = umma_desc_k_sw128(smem_base), qt0 = umma_desc_sw32(smem_base + Q_MAIN);
        const uint64_t km0 = umma_desc_k_sw128(smem_base + 2 * Q_TILE), kt0 = umma_desc_sw32(smem_base + 2 * Q_TILE + KV_MAIN);
        const uint64_t vm0 = umma_desc_mn_sw128(smem_base + V_OFF), vt0 = umma_desc_sw32(smem_base + V_OFF + KV_MAIN);
        auto issue_s = [&](int t, int st, bool release_stage) {
            if (elect_one()) {
                const uint64_t so = (uint64_t)(st * (QK_STAGE >> 4)), to = so + (uint64_t)(t * (Q_TILE >> 4));
                const uint32_t d = tmem_base + t * TILE_COLS;
#pragma unroll
                for (int k = 0; k < HM / 16; ++k) umma_bf16(d, qm0 + to + 2 * k, km0 + so + 2 * k, idesc_s, k ? 1u : 0u);
                umma_bf16(d, qt0 + to, kt0 + so, idesc_s, 1u);
                umma_commit(&s_full[t]);
                if (release_stage) umma_commit(&qk_empty[st]);   // both S products of the stage have been issued
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t, bool release_v) {
            if (elect_one()) {
                const uint32_t pt = tmem_base + t * TILE_COLS;
#pragma unroll
                for (int k = 0; k < HSK / 16; ++k) {     // 16 keys: 8 TMEM columns of P, 2048 B of the main and 512 B of the tail V part
                    umma_bf16_ts(pt + O_MAIN_COL, pt + k * 8, vm0 + 128 * k, idesc_om, k ? 1u : 0u);
                    umma_bf16_ts(pt + O_TAIL_COL, pt + k * 8, vt0 + 32 * k, idesc_ot, k ? 1u : 0u);
                }
                umma_commit(&o_full[t]);
                if (release_v) umma_commit(&v_empty);
            }
            __syncwarp();
        };
        int stage = 0;
        uint32_t phase = 0, tphase = 0;
        int p = blockIdx.x;
        if (p < a.nprob) {
            mbar_wait(&qk_full[0], 0);
            tc_fence_after();
            issue_s(0, 0, false);
            issue_s(1, 0, true);
        }
        for (; p < a.nprob; p += gridDim.x) {
            const bool has_next = p + (int)gridDim.x < a.nprob;
            const int nstage = stage ^ 1;
            const uint32_t nphase = stage ? phase ^ 1 : phase;
            mbar_wait(&p_full[0], tphase);
            if (lane == 0) W_TRACE(1, 1);
            mbar_wait(&v_full, tphase);
            if (lane == 0) W_TRACE(1, 9);
            tc_fence_after();
            issue_pv(0, false);
            if (lane == 0) W_TRACE(1, 2);
            if (has_next) {
                mbar_wait(&qk_full[nstage], nphase);
                if (lane == 0) W_TRACE(1, 10);
                mbar_wait(&s_empty[0], tphase);
                if (lane == 0) W_TRACE(1, 3);
                tc_fence_after();
                issue_s(0, nstage, false);
                if (lane == 0) W_TRACE(1, 4);
            }
            mbar_wait(&p_full[1], tphase);
            if (lane == 0) W_TRACE(1, 5);
            tc_fence_after();
            issue_pv(1, true);
            if (lane == 0) W_TRACE(1, 6);
            if (has_next) {
                mbar_wait(&s_empty[1], tphase);
                if (lane == 0) W_TRACE(1, 7);
                tc_fence_after();
                issue_s(1, nstage, true);
                if (lane == 0) W_TRACE(1, 8);
            }
            stage = nstage;
            phase = nphase;
            tphase ^= 1;
        }
    } else if (warp == STORE_WARP) {
        // ---------------- output stores: two TMA tensor stores per (problem, tile), converged warp, one elected lane ----------------
        uint32_t fph = 0;
        for (int p = blockIdx.x; p < a.nprob; p += gridDim.x) {
            const int h = p % a.heads, w = (p / a.heads) % nw2, b = p / (a.heads * nw2);
            const int wi = w / a.nwin_side, wj = w % a.nwin_side;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                mbar_wait(&stg_full, fph);
                fph ^= 1;
                if (elect_one()) {
                    tma_store_5d(&tmOm, smem + OUT_OFF, h * HD, 0, wj, wi * HWIN + t * HQ_IROWS, b);
                    tma_store_5d(&tmOt, smem + OUT_OFF + Q_MAIN, h * HD + HM, 0, wj, wi * HWIN + t * HQ_IROWS, b);
                    tma_store_commit();
                    tma_store_wait_read();
                    mbar_arrive(&stg_free);
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else {  // ---------------- softmax + epilogue warpgroups ----------------
        const int t = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tile_tmem = tmem_base + ((uint32_t)(quarter * 32) << 16) + t * TILE_COLS;
        uint32_t tphase = 0;
        for (int p = blockIdx.x; p < a.nprob; p += gridDim.x) {
            const int h = p % a.heads, w = (p / a.heads) % nw2, b = p / (a.heads * nw2);
            mbar_wait(&s_full[t], tphase);
            tc_fence_after();
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 1);
            float sum0 = 0.f, sum1 = 0.f;
            {
                float mx0 = -INFINITY, mx1 = -INFINITY;
                uint32_t ra[32], rb[32];
                tmem_ld_32x32(tile_tmem, ra);
#pragma unroll 1
                for (int c = 0; c < 8; c += 2) {
                    tmem_ld_wait();
                    tmem_ld_32x32(tile_tmem + (c + 1) * 32, rb);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        mx0 = fmaxf(mx0, __uint_as_float(ra[i]));
                        mx1 = fmaxf(mx1, __uint_as_float(ra[i + 1]));
                    }
                    tmem_ld_wait();
                    if (c + 2 < 8) tmem_ld_32x32(tile_tmem + (c + 2) * 32, ra);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        mx0 = fmaxf(mx0, __uint_as_float(rb[i]));
                        mx1 = fmaxf(mx1, __uint_as_float(rb[i + 1]));
                    }
                }
                const float moff = fmaxf(mx0, mx1) * a.scale_log2;
                if (quarter == 2 && lane == 0) W_TRACE(2 + t, 2);
                auto expo = [&](uint32_t sbits) { return ex2_approx(fmaf(__uint_as_float(sbits), a.scale_log2, -moff)); };
                tmem_ld_32x32(tile_tmem, ra);
#pragma unroll 1
                for (int c = 0; c < 8; c += 2) {
                    uint32_t pk[16];
                    tmem_ld_wait();
                    tmem_ld_32x32(tile_tmem + (c + 1) * 32, rb);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = expo(ra[2 * i]), p1 = expo(ra[2 * i + 1]);
                        sum0 += p0; sum1 += p1;
                        pk[i] = pack_bf16(p0, p1);
                    }
                    tmem_st_32x16(tile_tmem + c * 16, pk);
                    tmem_ld_wait();
                    if (c + 2 < 8) tmem_ld_32x32(tile_tmem + (c + 2) * 32, ra);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = expo(rb[2 * i]), p1 = expo(rb[2 * i + 1]);
                        sum0 += p0; sum1 += p1;
                        pk[i] = pack_bf16(p0, p1);
                    }
                    tmem_st_32x16(tile_tmem + (c + 1) * 16, pk);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 3);
            // epilogue: O / rowsum -> bf16 -> token-major output row of this query (80 columns = ten 16-byte stores).
            // O is pulled into registers and the tile's TMEM columns are handed back BEFORE the rows are written out: the
            // write-out (2300 clk as direct 16-byte stores, 640 through the staging tile) is then off the MMA warp's critical
            // path to the next S product (clock64 trace of round 1i).
            mbar_wait(&o_full[t], tphase);
            tc_fence_after();
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 4);
            uint32_t o0[32], o1[32], o2[16];
            tmem_ld_32x32(tile_tmem + O_MAIN_COL, o0);
            tmem_ld_32x32(tile_tmem + O_MAIN_COL + 32, o1);
            tmem_ld_32x16(tile_tmem + O_TAIL_COL, o2);
            tmem_ld_wait();
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 6);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[t]);
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 7);
            // rows -> bf16 -> staging tile (64-column part 128-byte-swizzled, 16-column part 32-byte-swizzled: the Q tile layout);
            // the store warp sends it with two TMA tensor stores.  A thread owns a row, so direct global stores touch 32 lines
            // per instruction and the L1 store path charges per line (1500 clk per tile on the serial chain, clock64 trace).
            // 201 + 20 KB is all that fits: the two query tiles, which run half a period apart and strictly alternate, share
            // ONE staging tile (use k = 2 * problem + tile waits for k releases: parity (tile & 1) ^ 1).
            const float inv = 1.0f / (sum0 + sum1);
            uint8_t* stg = smem + OUT_OFF;
            mbar_wait(&stg_free, (uint32_t)(t ^ 1));
            {
                uint8_t* rowm = stg + row * (HM * 2);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t* r = (c < 4 ? o0 : o1) + (c & 3) * 8;
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
                    v.y = pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
                    v.z = pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
                    v.w = pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
                    *reinterpret_cast<uint4*>(rowm + ((c ^ (row & 7)) << 4)) = v;
                }
                uint8_t* rowt = stg + Q_MAIN + row * (HT * 2);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t* r = o2 + c * 8;
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
                    v.y = pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
                    v.z = pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
                    v.w = pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
                    *reinterpret_cast<uint4*>(rowt + ((c ^ ((row >> 2) & 1)) << 4)) = v;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&stg_full);
            if (quarter == 2 && lane == 0) W_TRACE(2 + t, 5);
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

// =====================================================================================================
// ViT-H global self-attention (blocks 8 / 16 / 24 / 32): S = 1024 tokens, head_dim 80.
// The flash kernel above with the 64 + 16 column operand split of the ViT-H window kernel: key blocks of 64
// (1024 = 16 x 64, no masking; S 64 + P 32 + O 80 TMEM columns per query tile), six K/V stages.
// =====================================================================================================
namespace g80 {
constexpr int KB = 64, QT = 128, HD = 80, HM = 64, HT = 16;
constexpr int Q_MAIN = QT * HM * 2, Q_TAIL = QT * HT * 2, Q_TILE = Q_MAIN + Q_TAIL;          // 16 + 4 KB
constexpr int KV_MAIN = KB * HM * 2, KV_TAIL = KB * HT * 2, KV_BLK = KV_MAIN + KV_TAIL;      // 8 + 2 KB
constexpr int STG_B = 2 * KV_BLK, NSTG = 6;
constexpr int KV_OFF = 4 * Q_TILE;
constexpr int SMEM = KV_OFF + NSTG * STG_B + 1024;
constexpr int P_COL = 64, O_MAIN_COL = 96, O_TAIL_COL = 160;
static_assert(Q_TILE % 1024 == 0 && KV_MAIN % 1024 == 0 && KV_BLK % 1024 == 0, "swizzle atoms must stay aligned");
static_assert(SMEM <= 232448, "dynamic shared memory limit of sm_100");
}  // namespace g80

__global__ void __launch_bounds__(G_THREADS, 1)
global_attention_tc80_kernel(const __grid_constant__ CUtensorMap tmQm, const __grid_constant__ CUtensorMap tmQt,
                             const __grid_constant__ CUtensorMap tmKm, const __grid_constant__ CUtensorMap tmKt, const GlobArgs a) {
    using namespace g80;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t q_full[2], q_empty[2], kv_full[NSTG], kv_empty[NSTG], s_full[2], s_free[2], p_full[2], pv_done[2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQm);
        tma_prefetch_desc(&tmQt);
        tma_prefetch_desc(&tmKm);
        tma_prefetch_desc(&tmKt);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 2);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&pv_done[i], 1);
        }
        for (int s = 0; s < NSTG; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 2);
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
    pdl_wait();
    const int last_tile_pair = a.npairs - 1;
    const bool odd_tiles = (a.ntiles & 1) != 0;

    if (warp == 0) {
        // ---------------- TMA producer (converged warp, one elected lane) ----------------
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int u = blockIdx.x; u < a.nunits; u += gridDim.x, ++it) {
            const int pr = u % a.npairs, bh = u / a.npairs, h = bh % a.heads, b = bh / a.heads;
            const bool two = !(odd_tiles && pr == last_tile_pair);
            const int qb = it & 1;
            const int qc = a.qcol + h * HD, kc = a.kcol + h * HD, vc = a.vcol + h * HD;
            mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&q_full[qb], (two ? 2 : 1) * Q_TILE);
                uint8_t* qs = smem + qb * 2 * Q_TILE;
                const int r0 = b * a.S + (2 * pr) * QT;
                tma_load_2d(qs, &tmQm, &q_full[qb], qc, r0);
                tma_load_2d(qs + Q_MAIN, &tmQt, &q_full[qb], qc + HM, r0);
                if (two) {
                    tma_load_2d(qs + Q_TILE, &tmQm, &q_full[qb], qc, r0 + QT);
                    tma_load_2d(qs + Q_TILE + Q_MAIN, &tmQt, &q_full[qb], qc + HM, r0 + QT);
                }
            }
            __syncwarp();
            for (int j = 0; j < a.nblocks; ++j) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&kv_full[stage], STG_B);
                    uint8_t* st = smem + KV_OFF + stage * STG_B;
                    const int r = b * a.S + j * KB;
                    tma_load_2d(st, &tmKm, &kv_full[stage], kc, r);
                    tma_load_2d(st + KV_MAIN, &tmKt, &kv_full[stage], kc + HM, r);
                    tma_load_2d(st + KV_BLK, &tmKm, &kv_full[stage], vc, r);
                    tma_load_2d(st + KV_BLK + KV_MAIN, &tmKt, &kv_full[stage], vc + HM, r);
                }
                __syncwarp();
                if (++stage == NSTG) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 || warp == G_MMA1_WARP) {
        // ---------------- MMA issuers: one warp per query tile ----------------
        const int t = warp == 1 ? 0 : 1;
        constexpr uint32_t idesc_s = idesc_bf16(128, KB, false);
        constexpr uint32_t idesc_om = idesc_bf16(128, HM, true), idesc_ot = idesc_bf16(128, HT, true);
        const uint32_t s_tmem = tmem_base + t * TILE_COLS, p_tmem = s_tmem + P_COL;
        const uint64_t qm0 = umma_desc_k_sw128(smem_base + t * Q_TILE), qt0 = umma_desc_sw32(smem_base + t * Q_TILE + Q_MAIN);
        const uint64_t km0 = umma_desc_k_sw128(smem_base + KV_OFF), kt0 = umma_desc_sw32(smem_base + KV_OFF + KV_MAIN);
        const uint64_t vm0 = umma_desc_mn_sw128(smem_base + KV_OFF + KV_BLK), vt0 = umma_desc_sw32(smem_base + KV_OFF + KV_BLK + KV_MAIN);
        auto issue_s = [&](int qb, int st) {
            if (elect_one()) {
                const uint64_t qo = (uint64_t)(qb * (2 * Q_TILE >> 4)), ko = (uint64_t)(st * (STG_B >> 4));
#pragma unroll
                for (int k = 0; k < HM / 16; ++k) umma_bf16(s_tmem, qm0 + qo + 2 * k, km0 + ko + 2 * k, idesc_s, k ? 1u : 0u);
                umma_bf16(s_tmem, qt0 + qo, kt0 + ko, idesc_s, 1u);
                umma_commit(&s_full[t]);
            }
            __syncwarp();
        };
        int stage = 0;
        uint32_t phase = 0, pph = 0, fph = 0;
        int it = 0;
        int u = blockIdx.x;
        if (u < a.nunits && (t == 0 || !(odd_tiles && (u % a.npairs) == last_tile_pair))) {
            mbar_wait(&q_full[0], 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
        }
        for (; u < a.nunits; u += gridDim.x, ++it) {
            const bool two = !(odd_tiles && (u % a.npairs) == last_tile_pair);
            const bool valid = t == 0 || two;
            const int un = u + (int)gridDim.x;
            const bool valid_next = un < a.nunits && (t == 0 || !(odd_tiles && (un % a.npairs) == last_tile_pair));
            const int ncommit = (t == 0 && !two) ? 2 : 1;
            for (int j = 0; j < a.nblocks; ++j) {
                const int nstage = (stage + 1 == NSTG) ? 0 : stage + 1;
                const uint32_t nphase = (stage + 1 == NSTG) ? phase ^ 1 : phase;
                const bool last = j + 1 == a.nblocks;
                if (valid) {
                    mbar_wait(&s_free[t], fph);
                    fph ^= 1;
                }
                if (last ? valid_next : valid) {
                    if (last) mbar_wait(&q_full[(it + 1) & 1], ((it + 1) >> 1) & 1);
                    mbar_wait(&kv_full[nstage], nphase);
                    tc_fence_after();
                    issue_s(last ? (it + 1) & 1 : it & 1, nstage);
                }
                if (valid) {
                    mbar_wait(&p_full[t], pph);
                    pph ^= 1;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t vo = (uint64_t)(stage * (STG_B >> 4));
#pragma unroll
                        for (int k = 0; k < KB / 16; ++k) {
                            const uint32_t acc = (j == 0 && k == 0) ? 0u : 1u;
                            umma_bf16_ts(s_tmem + O_MAIN_COL, p_tmem + k * 8, vm0 + vo + 128 * k, idesc_om, acc);
                            umma_bf16_ts(s_tmem + O_TAIL_COL, p_tmem + k * 8, vt0 + vo + 32 * k, idesc_ot, acc);
                        }
                        umma_commit(&pv_done[t]);
                        umma_commit(&kv_empty[stage]);
                        if (ncommit == 2) umma_commit(&kv_empty[stage]);
                        if (last) {
                            umma_commit(&q_empty[it & 1]);
                            if (ncommit == 2) umma_commit(&q_empty[it & 1]);
                        }
                    }
                    __syncwarp();
                }
                stage = nstage;
                phase = nphase;
            }
        }
    } else {  // ---------------- softmax + epilogue warpgroups ----------------
        const int t = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tile_tmem = tmem_base + ((uint32_t)(quarter * 32) << 16) + t * TILE_COLS;
        uint32_t sph = 0, vph = 0;
        for (int u = blockIdx.x; u < a.nunits; u += gridDim.x) {
            const int pr = u % a.npairs, bh = u / a.npairs, h = bh % a.heads, b = bh / a.heads;
            if (t == 1 && odd_tiles && pr == last_tile_pair) continue;
            const int q0 = (2 * pr + t) * QT;
            float m = 0.f, l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
            for (int j = 0; j < a.nblocks; ++j) {
                mbar_wait(&s_full[t], sph);
                sph ^= 1;
                tc_fence_after();
                uint32_t r0[32], r1[32];
                tmem_ld_32x32(tile_tmem, r0);
                tmem_ld_32x32(tile_tmem + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_free[t]);      // S(j) is in registers: the MMA warp may overwrite it with S(j+1)
                bool rescale = false;
                float alpha = 1.0f;
                uint32_t pk[16];
                const float sc = a.scale_log2;
                auto expo = [&](uint32_t sbits) { return ex2_approx(fmaf(__uint_as_float(sbits), sc, -m)); };
                {
                    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r0[i + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r0[i + 2]), __uint_as_float(r0[i + 3])));
                        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(r1[i]), __uint_as_float(r1[i + 1])));
                        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(r1[i + 2]), __uint_as_float(r1[i + 3])));
                    }
                    const float mb = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
                    if (j == 0) {
                        m = mb;
                    } else {
                        const bool need = mb - m > 8.0f;
                        rescale = __any_sync(0xffffffffu, need) != 0;
                        if (rescale) {
                            alpha = need ? ex2_approx(m - mb) : 1.0f;
                            if (need) m = mb;
                            l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float p0 = expo(r0[2 * i]), p1 = expo(r0[2 * i + 1]), p2 = expo(r0[2 * i + 2]), p3 = expo(r0[2 * i + 3]);
                        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
                        pk[i] = pack_bf16(p0, p1); pk[i + 1] = pack_bf16(p2, p3);
                    }
                }
                if (j > 0) {     // P V(j-1) complete: P may be overwritten, O rescaled
                    mbar_wait(&pv_done[t], vph);
                    vph ^= 1;
                    tc_fence_after();
                }
                if (rescale) {
#pragma unroll 1
                    for (int c = 0; c < 3; ++c) {
                        uint32_t o[32];
                        if (c < 2) {
                            tmem_ld_32x32(tile_tmem + O_MAIN_COL + c * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st_32x32(tile_tmem + O_MAIN_COL + c * 32, o);
                        } else {
                            uint32_t o16[16];
                            tmem_ld_32x16(tile_tmem + O_TAIL_COL, o16);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o16[i] = __float_as_uint(__uint_as_float(o16[i]) * alpha);
                            tmem_st_32x16(tile_tmem + O_TAIL_COL, o16);
                        }
                    }
                }
                tmem_st_32x16(tile_tmem + P_COL, pk);
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float p0 = expo(r1[2 * i]), p1 = expo(r1[2 * i + 1]), p2 = expo(r1[2 * i + 2]), p3 = expo(r1[2 * i + 3]);
                    l0 += p0; l1 += p1; l2 += p2; l3 += p3;
                    pk[i] = pack_bf16(p0, p1); pk[i + 1] = pack_bf16(p2, p3);
                }
                tmem_st_32x16(tile_tmem + P_COL + 16, pk);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[t]);
            }
            // epilogue: O / l -> bf16 -> token-major output row of this query
            mbar_wait(&pv_done[t], vph);
            vph ^= 1;
            tc_fence_after();
            const float inv = 1.0f / ((l0 + l1) + (l2 + l3));
            __nv_bfloat16* dst = a.o + ((size_t)b * a.S + q0 + row) * a.ldo + h * HD;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tile_tmem + O_MAIN_COL + c * 32, r);
                tmem_ld_wait();
                store16_bf16(dst + c * 32, r, inv);
                store16_bf16(dst + c * 32 + 16, r + 16, inv);
            }
            {
                uint32_t r[16];
                tmem_ld_32x16(tile_tmem + O_TAIL_COL, r);
                tmem_ld_wait();
                store16_bf16(dst + HM, r, inv);
            }
            tc_fence_before();
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
    VPU_CHECK_CUDA(cudaFuncSetAttribute(global_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
    VPU_CHECK_CUDA(cudaFuncSetAttribute(window_attention_tc80_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h80::SMEM));
    VPU_CHECK_CUDA(cudaFuncSetAttribute(global_attention_tc80_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g80::SMEM));
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// dims (innermost first): column, j inside the window, window column, grid row, image; box = cols x win x 1 x irows x 1
int make_map_w(CUtensorMap* tm, const void* ptr, int ld, int images, int grid, int win, int irows, int cols, CUtensorMapSwizzle swz) {
    Key key{ptr, ld, images, grid, irows | (cols << 8) | (win << 16) | ((int)swz << 24)};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) { *tm = it->second; return 0; }
    }
    const cuuint64_t row_b = (cuuint64_t)ld * 2;
    cuuint64_t gdim[5] = {(cuuint64_t)ld, (cuuint64_t)win, (cuuint64_t)(grid / win), (cuuint64_t)grid, (cuuint64_t)images};
    cuuint64_t gstride[4] = {row_b, win * row_b, (cuuint64_t)grid * row_b, (cuuint64_t)grid * grid * row_b};
    cuuint32_t box[5] = {(cuuint32_t)cols, (cuuint32_t)win, 1, (cuuint32_t)irows, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPU_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (5-D window map) failed with %d", (int)r);
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.size() > 1024) g_cache.clear();
    g_cache[key] = *tm;
    return 0;
}
int make_map(CUtensorMap* tm, const void* ptr, int ld, int images, int grid, int irows) {
    return make_map_w(tm, ptr, ld, images, grid, WIN, irows, D, CU_TENSOR_MAP_SWIZZLE_128B);
}

// plain 2-D map over a token-major [rows, ld] bf16 buffer: box = cols columns x box_rows rows
int make_map_2d_w(CUtensorMap* tm, const void* ptr, int ld, long long rows, int box_rows, int cols, CUtensorMapSwizzle swz) {
    Key key{ptr, ld, (int)rows, box_rows, -1 - (cols << 4) - (int)swz};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) { *tm = it->second; return 0; }
    }
    cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPU_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (2-D attention map) failed with %d", (int)r);
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.size() > 1024) g_cache.clear();
    g_cache[key] = *tm;
    return 0;
}
int make_map_2d(CUtensorMap* tm, const void* ptr, int ld, long long rows, int box_rows) {
    return make_map_2d_w(tm, ptr, ld, rows, box_rows, D, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace

bool global_attention_tc_supported(const AttnArgs& a, int head_dim) {
    return head_dim == D && a.qmap.mode == 0 && a.kmap.mode == 0 && a.Sq == a.Sk && a.Sk % G_KB == 0 && a.Sk >= 2 * G_KB &&
           a.qmap.per_prob == a.Sq && a.kmap.per_prob == a.Sk && a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 &&
           a.ldo % 8 == 0 && a.qoff % 8 == 0 && a.koff % 8 == 0 && a.voff % 8 == 0 &&
           ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v) |
             reinterpret_cast<uintptr_t>(a.o)) & 15) == 0 &&
           a.ldo % 16 == 0 && (reinterpret_cast<uintptr_t>(a.o) & 31) == 0;      // 32-byte epilogue stores
}

void attention_debug_trace(unsigned long long* dev_buf, int cap) { g_trace = dev_buf; g_trace_cap = cap; }

int global_attention_tc_launch(const AttnArgs& a, cudaStream_t stream) {
    if (int rc = init()) return rc;
    const long long rows = (long long)a.nprob * a.Sq;
    CUtensorMap tmQ, tmK, tmV;
    if (int rc = make_map_2d(&tmQ, a.q, a.ldq, rows, G_QT)) return rc;
    if (int rc = make_map_2d(&tmK, a.k, a.ldk, rows, G_KB)) return rc;
    if (int rc = make_map_2d(&tmV, a.v, a.ldv, rows, G_KB)) return rc;
    GlobArgs g;
    g.o = a.o; g.ldo = a.ldo; g.heads = a.heads; g.S = a.Sq; g.nblocks = a.Sk / G_KB;
    g.ntiles = (a.Sq + G_QT - 1) / G_QT; g.npairs = (g.ntiles + 1) / 2;
    g.nunits = a.nprob * a.heads * g.npairs;
    g.qcol = a.qoff; g.kcol = a.koff; g.vcol = a.voff; g.scale_log2 = a.scale_log2;
    static const int ablate = [] { const char* e = vpu_debug_env("VPU_ATTN_ABLATE"); return e ? atoi(e) : 0; }();
    g.ablate = ablate;
    g.trace = g_trace; g.trace_cap = g_trace_cap;
    const int ctas = g.nunits < g_sms ? g.nunits : g_sms;
    VPU_CHECK_CUDA(launch_pdl(global_attention_tc_kernel, dim3(ctas), dim3(G_THREADS), G_SMEM_BYTES, stream, tmQ, tmK, tmV, g));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

bool window_attention_tc_supported(const AttnArgs& a, int head_dim) {
    return head_dim == D && a.qmap.mode == 1 && a.qmap.win == WIN && a.qmap.grid % WIN == 0 && a.Sq == SK && a.Sk == SK &&
           a.q == a.k && a.q == a.v && a.ldq == a.ldk && a.ldq == a.ldv && a.ldq % 8 == 0 && a.ldo % 8 == 0 &&
           a.qoff % 8 == 0 && a.koff % 8 == 0 && a.voff % 8 == 0 && (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 &&
           a.ldo % 16 == 0 && (reinterpret_cast<uintptr_t>(a.o) & 31) == 0;      // 32-byte epilogue stores
}

int window_attention_tc_launch(const AttnArgs& a, cudaStream_t stream) {
    if (int rc = init()) return rc;
    const int grid = a.qmap.grid, nws = grid / WIN, images = a.nprob / (nws * nws);
    VPU_REQUIRE(images * nws * nws == a.nprob, "window attention: nprob must be images * windows");
    CUtensorMap tmKV, tmQ0, tmQ1;
    if (int rc = make_map(&tmKV, a.q, a.ldq, images, grid, WIN)) return rc;
    if (int rc = make_map(&tmQ0, a.q, a.ldq, images, grid, T0_IROWS)) return rc;
    if (int rc = make_map(&tmQ1, a.q, a.ldq, images, grid, T1_IROWS)) return rc;
    CUtensorMap tmO0, tmO1;          // output rows of the two query tiles, column 0 = head 0 (a.o has no column offset)
    if (int rc = make_map(&tmO0, a.o, a.ldo, images, grid, T0_IROWS)) return rc;
    if (int rc = make_map(&tmO1, a.o, a.ldo, images, grid, T1_IROWS)) return rc;
    WinArgs w;
    w.o = a.o; w.ldo = a.ldo; w.heads = a.heads; w.nwin_side = nws; w.grid = grid; w.tokens = a.qmap.tokens;
    w.nprob = a.nprob * a.heads; w.qcol = a.qoff; w.kcol = a.koff; w.vcol = a.voff; w.scale_log2 = a.scale_log2;
    w.trace = g_trace; w.trace_cap = g_trace_cap;
    const int ctas = w.nprob < g_sms ? w.nprob : g_sms;
    VPU_CHECK_CUDA(launch_pdl(window_attention_tc_kernel, dim3(ctas), dim3(THREADS), SMEM_BYTES, stream, tmKV, tmQ0, tmQ1, tmO0, tmO1, w));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

bool window_attention_tc80_supported(const AttnArgs& a, int head_dim) {
    return head_dim == h80::HD && a.qmap.mode == 1 && a.qmap.win == h80::HWIN && a.qmap.grid % h80::HWIN == 0 && a.Sq == h80::HSK &&
           a.Sk == h80::HSK && a.q == a.k && a.q == a.v && a.ldq == a.ldk && a.ldq == a.ldv && a.ldq % 8 == 0 && a.ldo % 8 == 0 &&
           a.qoff % 8 == 0 && a.koff % 8 == 0 && a.voff % 8 == 0 && (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 &&
           a.ldo % 16 == 0 && (reinterpret_cast<uintptr_t>(a.o) & 31) == 0;      // 32-byte epilogue stores
}

int window_attention_tc80_launch(const AttnArgs& a, cudaStream_t stream) {
    using namespace h80;
    if (int rc = init()) return rc;
    const int grid = a.qmap.grid, nws = grid / HWIN, images = a.nprob / (nws * nws);
    VPU_REQUIRE(images * nws * nws == a.nprob, "window attention: nprob must be images * windows");
    CUtensorMap tmQm, tmQt, tmKm, tmKt;
    if (int rc = make_map_w(&tmQm, a.q, a.ldq, images, grid, HWIN, HQ_IROWS, HM, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_w(&tmQt, a.q, a.ldq, images, grid, HWIN, HQ_IROWS, HT, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
    if (int rc = make_map_w(&tmKm, a.q, a.ldq, images, grid, HWIN, HWIN, HM, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_w(&tmKt, a.q, a.ldq, images, grid, HWIN, HWIN, HT, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
    CUtensorMap tmOm, tmOt;          // output rows of a query tile: 64-column and 16-column part
    if (int rc = make_map_w(&tmOm, a.o, a.ldo, images, grid, HWIN, HQ_IROWS, HM, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_w(&tmOt, a.o, a.ldo, images, grid, HWIN, HQ_IROWS, HT, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
    WinArgs w;
    w.o = a.o; w.ldo = a.ldo; w.heads = a.heads; w.nwin_side = nws; w.grid = grid; w.tokens = a.qmap.tokens;
    w.nprob = a.nprob * a.heads; w.qcol = a.qoff; w.kcol = a.koff; w.vcol = a.voff; w.scale_log2 = a.scale_log2;
    w.trace = g_trace; w.trace_cap = g_trace_cap;
    const int ctas = w.nprob < g_sms ? w.nprob : g_sms;
    VPU_CHECK_CUDA(launch_pdl(window_attention_tc80_kernel, dim3(ctas), dim3(THREADS), SMEM, stream, tmQm, tmQt, tmKm, tmKt, tmOm, tmOt, w));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

bool global_attention_tc80_supported(const AttnArgs& a, int head_dim) {
    return head_dim == g80::HD && a.qmap.mode == 0 && a.kmap.mode == 0 && a.Sq == a.Sk && a.Sk % g80::QT == 0 && a.k == a.v && a.ldk == a.ldv &&
           a.qmap.per_prob == a.Sq && a.kmap.per_prob == a.Sk && a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 &&
           a.ldo % 8 == 0 && a.qoff % 8 == 0 && a.koff % 8 == 0 && a.voff % 8 == 0 &&
           ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v) |
             reinterpret_cast<uintptr_t>(a.o)) & 15) == 0 &&
           a.ldo % 16 == 0 && (reinterpret_cast<uintptr_t>(a.o) & 31) == 0;      // 32-byte epilogue stores
}

int global_attention_tc80_launch(const AttnArgs& a, cudaStream_t stream) {
    using namespace g80;
    if (int rc = init()) return rc;
    const long long rows = (long long)a.nprob * a.Sq;
    CUtensorMap tmQm, tmQt, tmKm, tmKt;
    VPU_REQUIRE(a.k == a.v && a.ldk == a.ldv, "ViT-H global attention expects K and V in one buffer");
    if (int rc = make_map_2d_w(&tmQm, a.q, a.ldq, rows, QT, HM, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_2d_w(&tmQt, a.q, a.ldq, rows, QT, HT, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
    if (int rc = make_map_2d_w(&tmKm, a.k, a.ldk, rows, KB, HM, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_2d_w(&tmKt, a.k, a.ldk, rows, KB, HT, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
    GlobArgs g;
    g.o = a.o; g.ldo = a.ldo; g.heads = a.heads; g.S = a.Sq; g.nblocks = a.Sk / KB;
    g.ntiles = (a.Sq + QT - 1) / QT; g.npairs = (g.ntiles + 1) / 2;
    g.nunits = a.nprob * a.heads * g.npairs;
    g.qcol = a.qoff; g.kcol = a.koff; g.vcol = a.voff; g.scale_log2 = a.scale_log2;
    g.ablate = 0; g.trace = nullptr; g.trace_cap = 0;
    const int ctas = g.nunits < g_sms ? g.nunits : g_sms;
    VPU_CHECK_CUDA(launch_pdl(global_attention_tc80_kernel, dim3(ctas), dim3(G_THREADS), SMEM, stream, tmQm, tmQt, tmKm, tmKt, g));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace vpu
