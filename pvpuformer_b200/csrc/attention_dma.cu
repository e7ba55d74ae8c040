// tcgen05 / TMEM / TMA attention cores of the Dual-cross Merging Attention stage (reference
// isegm/model/modeling/transformer.py:499-521, called from :432-463 and :375-379): the three shapes that the
// TwoWayAttentionBlock runs on 48 prompt tokens and N image tokens,
//     prompt self-attention        48 x 48,  head_dim C/8      (96 / 128 / 160)
//     tokens -> image              48 x N,   head_dim C/16     (48 / 64 / 80)
//     image -> tokens              N x 48,   head_dim C/16
// One kernel template covers them; it is the flash dataflow of the ViT global-attention kernels (attention_tc.cu):
//   S = Q K^T      tcgen05.mma, both operands K-major from shared memory (TMA), fp32 in TMEM
//   P = exp2(S-m)  softmax warpgroup, thread == query row, S pulled out of TMEM into registers, bf16 P written back to TMEM
//   O += P V       tcgen05.mma with the A operand in TMEM, V an MN-major shared-memory operand
// generalised in three directions:
//   * head_dim = 64 NM + 16 NT: every operand is staged as NM 64-column parts (128-byte swizzle) and NT 16-column parts
//     (32-byte swizzle), each by its own TMA box -- 48 = 3 tails, 96 = 1 + 2, 160 = 2 + 2 (the ViT-H kernels' 64 + 16 split);
//   * 3-D tensor maps (column, row inside the problem, problem): a 128-row query box over 48 prompt rows, or over the last
//     16 image rows of an image, is zero-filled past the problem's end by the TMA unit, so nothing is masked in the kernel
//     and nothing is read across problems; stores are predicated on the row index;
//   * the two query tiles a CTA keeps in flight are either two 128-row tiles of one (image, head) sharing K / V blocks
//     ("few keys": Sk = 48, one key block, no online rescale ever taken), or -- HP, "few queries": Sq <= 128 -- the same
//     48 query rows of two adjacent HEADS with their own K / V blocks, so that both softmax warpgroups, both MMA issuers and
//     twice the TMA bytes are in flight for a shape that would otherwise leave half of the CTA idle.
// Online softmax with the lazy rescale of the global kernels (only when a block raises the row maximum by > 2^8).
// Warps whose 32 TMEM lanes hold no real query row (rows 64..127 of a 48-row problem) skip the loads, exponentials and
// stores and only keep the barrier protocol going: the exponentials are the MUFU-bound part of these shapes.
#include <mutex>
#include <unordered_map>

#include "attention.cuh"
#include "tc_attn.cuh"

namespace vpu {

namespace {

constexpr int QT = 128;               // query rows per tile (UMMA M)
constexpr int TILE_COLS = 256;        // TMEM columns per query tile
constexpr int THREADS = 352;          // warp 0 TMA, 1 MMA tile 0, 2-5 / 6-9 softmax tile 0 / 1, 10 MMA tile 1
constexpr int MMA1_WARP = 10;
constexpr int SMEM_LIMIT = 232448 - 2048;   // dynamic shared memory of sm_100 minus the static barriers' share

template <int NM, int NT, int KB_, bool HP_>
struct Cfg {
    static constexpr int KB = KB_;
    static constexpr bool HP = HP_;
    static constexpr int D = 64 * NM + 16 * NT;
    static constexpr int NH = HP ? 2 : 1;                       // heads whose K / V blocks share a stage
    static constexpr int QM_B = QT * 128, QT_B = QT * 32;       // bytes of a 64-column / 16-column query part
    static constexpr int KM_B = KB * 128, KT_B = KB * 32;
    static constexpr int Q_TILE = NM * QM_B + NT * QT_B;
    static constexpr int QBUF = 2 * Q_TILE;                     // both tiles: [t0 mains | t1 mains | t0 tails | t1 tails]
    static constexpr int Q_TAILS = 2 * NM * QM_B;
    static constexpr int KV_OPER = NM * KM_B + NT * KT_B;       // K (or V) block of one head
    static constexpr int STAGE_TX = NH * 2 * KV_OPER;
    static constexpr int STAGE = (STAGE_TX + 1023) / 1024 * 1024;   // [mains of (head, K|V) ... | tails ...]
    static constexpr int S_TAILS = NH * 2 * NM * KM_B;
    static constexpr int KV_OFF = 2 * QBUF;                     // after the double-buffered query tiles
    static constexpr int NSTG_FIT = (SMEM_LIMIT - 1024 - KV_OFF) / STAGE;
    static constexpr int NSTG = NSTG_FIT > 4 ? 4 : NSTG_FIT;
    static constexpr int SMEM = KV_OFF + NSTG * STAGE + 1024;
    static constexpr int P_COL = KB;                            // TMEM: S [0, KB) fp32, P [KB, KB + KB/2) bf16 pairs, O [O_COL, O_COL + D)
    static constexpr int O_COL = TILE_COLS - (D + 31) / 32 * 32;
    static_assert(KB % 16 == 0 && KB >= 16 && KB <= 112, "key block");
    static_assert(NSTG >= 2, "at least two K/V stages must fit");
    static_assert(P_COL + KB / 2 <= O_COL, "TMEM columns of a tile overlap");
    static_assert(KM_B % 1024 == 0 && KT_B % 256 == 0, "swizzle atoms must stay aligned");
    __host__ __device__ static constexpr int q_main(int t, int i) { return (t * NM + i) * QM_B; }
    __host__ __device__ static constexpr int q_tail(int t, int i) { return Q_TAILS + (t * NT + i) * QT_B; }
    __host__ __device__ static constexpr int kv_main(int hh, int o, int i) { return ((hh * 2 + o) * NM + i) * KM_B; }
    __host__ __device__ static constexpr int kv_tail(int hh, int o, int i) { return S_TAILS + ((hh * 2 + o) * NT + i) * KT_B; }
};

struct DmaArgs {
    __nv_bfloat16* o;
    int ldo;
    int heads, Sq, Sk;
    int nblocks;                  // Sk / KB
    int ntiles, npairs;           // query tiles / tile pairs per (image, head); HP: npairs = head pairs per image
    int nunits;
    int qcol, kcol, vcol;
    float scale_log2;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

struct Unit {
    int b, h0, pr;
    bool two;
};

template <class C>
__device__ __forceinline__ Unit decode_unit(const DmaArgs& a, int u) {
    Unit r;
    if (C::HP) {
        r.b = u / a.npairs;
        r.h0 = 2 * (u % a.npairs);
        r.pr = 0;
        r.two = true;
    } else {
        r.pr = u % a.npairs;
        const int bh = u / a.npairs;
        r.h0 = bh % a.heads;
        r.b = bh / a.heads;
        r.two = 2 * r.pr + 1 < a.ntiles;
    }
    return r;
}

template <int NM, int NT, int KB, bool HP>
__global__ void __launch_bounds__(THREADS, 1)
dma_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQm, const __grid_constant__ CUtensorMap tmQt,
                        const __grid_constant__ CUtensorMap tmKm, const __grid_constant__ CUtensorMap tmKt,
                        const __grid_constant__ CUtensorMap tmVm, const __grid_constant__ CUtensorMap tmVt, const DmaArgs a) {
    using C = Cfg<NM, NT, KB, HP>;
    constexpr int D = C::D, NSTG = C::NSTG;
    constexpr int NMa = NM > 0 ? NM : 1, NTa = NT > 0 ? NT : 1;      // array extents (zero-length arrays are not C++)
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t q_full[2], q_empty[2], kv_full[NSTG], kv_empty[NSTG], s_full[2], s_free[2], p_full[2], pv_done[2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        if (NM) { tma_prefetch_desc(&tmQm); tma_prefetch_desc(&tmKm); tma_prefetch_desc(&tmVm); }
        if (NT) { tma_prefetch_desc(&tmQt); tma_prefetch_desc(&tmKt); tma_prefetch_desc(&tmVt); }
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

    if (warp == 0) {
        // ---------------- TMA producer (converged warp, one elected lane) ----------------
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int u = blockIdx.x; u < a.nunits; u += gridDim.x, ++it) {
            const Unit un = decode_unit<C>(a, u);
            const int qb = it & 1;
            mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                uint64_t* bar = &q_full[qb];
                mbar_arrive_expect_tx(bar, (un.two ? 2 : 1) * C::Q_TILE);
                uint8_t* qs = smem + qb * C::QBUF;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (t == 1 && !un.two) break;
                    const int head = HP ? un.h0 + t : un.h0;
                    const int r0 = HP ? 0 : (2 * un.pr + t) * QT;
                    const int col = a.qcol + head * D;
#pragma unroll
                    for (int i = 0; i < NM; ++i) tma_load_3d(qs + C::q_main(t, i), &tmQm, bar, col + 64 * i, r0, un.b);
#pragma unroll
                    for (int i = 0; i < NT; ++i) tma_load_3d(qs + C::q_tail(t, i), &tmQt, bar, col + 64 * NM + 16 * i, r0, un.b);
                }
            }
            __syncwarp();
            for (int j = 0; j < a.nblocks; ++j) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                if (elect_one()) {
                    uint64_t* bar = &kv_full[stage];
                    mbar_arrive_expect_tx(bar, C::STAGE_TX);
                    uint8_t* st = smem + C::KV_OFF + stage * C::STAGE;
                    const int r = j * KB;
#pragma unroll
                    for (int hh = 0; hh < C::NH; ++hh) {
                        const int kc = a.kcol + (un.h0 + hh) * D, vc = a.vcol + (un.h0 + hh) * D;
#pragma unroll
                        for (int i = 0; i < NM; ++i) {
                            tma_load_3d(st + C::kv_main(hh, 0, i), &tmKm, bar, kc + 64 * i, r, un.b);
                            tma_load_3d(st + C::kv_main(hh, 1, i), &tmVm, bar, vc + 64 * i, r, un.b);
                        }
#pragma unroll
                        for (int i = 0; i < NT; ++i) {
                            tma_load_3d(st + C::kv_tail(hh, 0, i), &tmKt, bar, kc + 64 * NM + 16 * i, r, un.b);
                            tma_load_3d(st + C::kv_tail(hh, 1, i), &tmVt, bar, vc + 64 * NM + 16 * i, r, un.b);
                        }
                    }
                }
                __syncwarp();
                if (++stage == NSTG) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 || warp == MMA1_WARP) {
        // ---------------- MMA issuers: one warp per query tile ----------------
        const int t = warp == 1 ? 0 : 1;
        const int hh = HP ? t : 0;
        constexpr uint32_t idesc_s = idesc_bf16(128, KB, false);
        constexpr uint32_t idesc_om = idesc_bf16(128, 64, true), idesc_ot = idesc_bf16(128, 16, true);
        const uint32_t s_tmem = tmem_base + t * TILE_COLS, p_tmem = s_tmem + C::P_COL, o_tmem = s_tmem + C::O_COL;
        uint64_t qm[NMa], km[NMa], vm[NMa], qt[NTa], kt[NTa], vt[NTa];
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            qm[i] = umma_desc_k_sw128(smem_base + C::q_main(t, i));
            km[i] = umma_desc_k_sw128(smem_base + C::KV_OFF + C::kv_main(hh, 0, i));
            vm[i] = umma_desc_mn_sw128(smem_base + C::KV_OFF + C::kv_main(hh, 1, i));
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            qt[i] = umma_desc_sw32(smem_base + C::q_tail(t, i));
            kt[i] = umma_desc_sw32(smem_base + C::KV_OFF + C::kv_tail(hh, 0, i));
            vt[i] = umma_desc_sw32(smem_base + C::KV_OFF + C::kv_tail(hh, 1, i));
        }
        auto issue_s = [&](int qb, int st) {
            if (elect_one()) {
                const uint64_t qo = (uint64_t)(qb * (C::QBUF >> 4)), ko = (uint64_t)(st * (C::STAGE >> 4));
                uint32_t acc = 0u;
#pragma unroll
                for (int i = 0; i < NM; ++i) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_bf16(s_tmem, qm[i] + qo + 2 * k, km[i] + ko + 2 * k, idesc_s, acc);
                        acc = 1u;
                    }
                }
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    umma_bf16(s_tmem, qt[i] + qo, kt[i] + ko, idesc_s, acc);
                    acc = 1u;
                }
                umma_commit(&s_full[t]);
            }
            __syncwarp();
        };
        int stage = 0;
        uint32_t phase = 0, pph = 0, fph = 0;
        int it = 0;
        int u = blockIdx.x;
        if (u < a.nunits && (t == 0 || decode_unit<C>(a, u).two)) {
            mbar_wait(&q_full[0], 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
        }
        for (; u < a.nunits; u += gridDim.x, ++it) {
            const bool two = decode_unit<C>(a, u).two;
            const bool valid = t == 0 || two;
            const int un = u + (int)gridDim.x;
            const bool valid_next = un < a.nunits && (t == 0 || decode_unit<C>(a, un).two);
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
                        const uint64_t vo = (uint64_t)(stage * (C::STAGE >> 4));
#pragma unroll
                        for (int k = 0; k < KB / 16; ++k) {     // 16 keys: 8 TMEM columns of P, 2048 B of a main and 512 B of a tail V part
                            const uint32_t acc = (j == 0 && k == 0) ? 0u : 1u;
#pragma unroll
                            for (int i = 0; i < NM; ++i) umma_bf16_ts(o_tmem + 64 * i, p_tmem + k * 8, vm[i] + vo + 128 * k, idesc_om, acc);
#pragma unroll
                            for (int i = 0; i < NT; ++i)
                                umma_bf16_ts(o_tmem + 64 * NM + 16 * i, p_tmem + k * 8, vt[i] + vo + 32 * k, idesc_ot, acc);
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
        constexpr int NC = KB / 16;
        uint32_t sph = 0, vph = 0;
        for (int u = blockIdx.x; u < a.nunits; u += gridDim.x) {
            const Unit un = decode_unit<C>(a, u);
            if (t == 1 && !un.two) continue;
            const int head = HP ? un.h0 + t : un.h0;
            const int q0 = HP ? 0 : (2 * un.pr + t) * QT;
            const int q = q0 + row;
            const bool active = q0 + quarter * 32 < a.Sq;        // warp-uniform: this warp's lanes hold at least one real query row
            float m = 0.f, l[4] = {0.f, 0.f, 0.f, 0.f};
            const float sc = a.scale_log2;
            for (int j = 0; j < a.nblocks; ++j) {
                mbar_wait(&s_full[t], sph);
                sph ^= 1;
                tc_fence_after();
                uint32_t s[KB];
                if (active) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) tmem_ld_32x16(tile_tmem + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[c * 16]));
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_free[t]);      // S(j) is in registers: the MMA warp may overwrite it with S(j+1)
                bool rescale = false;
                float alpha = 1.0f;
                uint32_t pk[8];
                auto expo = [&](uint32_t sbits) { return ex2_approx(fmaf(__uint_as_float(sbits), sc, -m)); };
                auto chunk = [&](int c) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float p0 = expo(s[c * 16 + 2 * i]), p1 = expo(s[c * 16 + 2 * i + 1]);
                        l[(2 * i) & 3] += p0;
                        l[(2 * i + 1) & 3] += p1;
                        pk[i] = pack_bf16(p0, p1);
                    }
                };
                if (active) {
                    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int i = 0; i < KB; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(s[i]));
                    const float mb = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sc;
                    if (j == 0) {
                        m = mb;
                    } else {
                        const bool need = mb - m > 8.0f;
                        rescale = __any_sync(0xffffffffu, need) != 0;
                        if (rescale) {
                            alpha = need ? ex2_approx(m - mb) : 1.0f;
                            if (need) m = mb;
#pragma unroll
                            for (int i = 0; i < 4; ++i) l[i] *= alpha;
                        }
                    }
                    chunk(0);
                }
                if (j > 0) {     // P V(j-1) complete: P may be overwritten, O rescaled
                    mbar_wait(&pv_done[t], vph);
                    vph ^= 1;
                    tc_fence_after();
                }
                if (active) {
                    if (rescale) {
#pragma unroll 1
                        for (int c = 0; c < D / 16; ++c) {
                            uint32_t o16[16];
                            tmem_ld_32x16(tile_tmem + C::O_COL + c * 16, o16);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o16[i] = __float_as_uint(__uint_as_float(o16[i]) * alpha);
                            tmem_st_32x16(tile_tmem + C::O_COL + c * 16, o16);
                        }
                    }
                    tmem_st_32x8(tile_tmem + C::P_COL, pk);
#pragma unroll
                    for (int c = 1; c < NC; ++c) {
                        chunk(c);
                        tmem_st_32x8(tile_tmem + C::P_COL + c * 8, pk);
                    }
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[t]);
            }
            // epilogue: O / l -> bf16 -> token-major output row of this query
            mbar_wait(&pv_done[t], vph);
            vph ^= 1;
            tc_fence_after();
            if (active) {
                const float inv = 1.0f / ((l[0] + l[1]) + (l[2] + l[3]));
                __nv_bfloat16* dst = a.o + ((size_t)un.b * a.Sq + q) * a.ldo + head * D;
#pragma unroll
                for (int c = 0; c < D / 16; ++c) {
                    uint32_t r[16];
                    tmem_ld_32x16(tile_tmem + C::O_COL + c * 16, r);
                    tmem_ld_wait();
                    if (q < a.Sq) store16_bf16(dst + c * 16, r, inv);
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

// ---- host --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_sms = 0;
std::mutex g_mu;
struct Key {
    const void* p;
    int ld, rows, nprob, box;
    bool operator==(const Key& o) const { return p == o.p && ld == o.ld && rows == o.rows && nprob == o.nprob && box == o.box; }
};
struct KeyHash {
    size_t operator()(const Key& k) const {
        size_t h = reinterpret_cast<size_t>(k.p);
        h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.nprob;
        return h * 1000003u ^ (size_t)k.box;
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
    VPU_REQUIRE(prop.major == 10, "DMA attention needs an sm_100a device");
    g_sms = prop.multiProcessorCount;
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// dims (innermost first): column, row inside the problem, problem; box = cols x box_rows x 1.  Rows past the problem's end are
// out of bounds for the map and arrive as zeros.
int make_map_3d(CUtensorMap* tm, const void* ptr, int ld, int rows, int nprob, int box_rows, int cols, CUtensorMapSwizzle swz) {
    Key key{ptr, ld, rows, nprob, box_rows | (cols << 8) | ((int)swz << 16)};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) { *tm = it->second; return 0; }
    }
    const cuuint64_t row_b = (cuuint64_t)ld * 2;
    cuuint64_t gdim[3] = {(cuuint64_t)ld, (cuuint64_t)rows, (cuuint64_t)nprob};
    cuuint64_t gstride[2] = {row_b, (cuuint64_t)rows * row_b};
    cuuint32_t box[3] = {(cuuint32_t)cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPU_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D DMA attention map) failed with %d", (int)r);
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.size() > 1024) g_cache.clear();
    g_cache[key] = *tm;
    return 0;
}

template <int NM, int NT, int KB, bool HP>
int launch_cfg(const AttnArgs& a, cudaStream_t stream) {
    using C = Cfg<NM, NT, KB, HP>;
    if (int rc = init()) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        VPU_CHECK_CUDA(cudaFuncSetAttribute(dma_attention_tc_kernel<NM, NT, KB, HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    CUtensorMap tmQm, tmQt, tmKm, tmKt, tmVm, tmVt;
    memset(&tmQm, 0, sizeof(CUtensorMap));
    tmQt = tmKm = tmKt = tmVm = tmVt = tmQm;
    if (NM) {
        if (int rc = make_map_3d(&tmQm, a.q, a.ldq, a.Sq, a.nprob, QT, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
        if (int rc = make_map_3d(&tmKm, a.k, a.ldk, a.Sk, a.nprob, KB, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
        if (int rc = make_map_3d(&tmVm, a.v, a.ldv, a.Sk, a.nprob, KB, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
    if (NT) {
        if (int rc = make_map_3d(&tmQt, a.q, a.ldq, a.Sq, a.nprob, QT, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
        if (int rc = make_map_3d(&tmKt, a.k, a.ldk, a.Sk, a.nprob, KB, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
        if (int rc = make_map_3d(&tmVt, a.v, a.ldv, a.Sk, a.nprob, KB, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
    }
    DmaArgs g;
    g.o = a.o; g.ldo = a.ldo; g.heads = a.heads; g.Sq = a.Sq; g.Sk = a.Sk; g.nblocks = a.Sk / KB;
    g.ntiles = (a.Sq + QT - 1) / QT;
    g.npairs = HP ? a.heads / 2 : (g.ntiles + 1) / 2;
    g.nunits = HP ? a.nprob * g.npairs : a.nprob * a.heads * g.npairs;
    g.qcol = a.qoff; g.kcol = a.koff; g.vcol = a.voff; g.scale_log2 = a.scale_log2;
    const int ctas = g.nunits < g_sms ? g.nunits : g_sms;
    VPU_CHECK_CUDA(launch_pdl(dma_attention_tc_kernel<NM, NT, KB, HP>, dim3(ctas), dim3(THREADS), (size_t)C::SMEM, stream, tmQm, tmQt,
                              tmKm, tmKt, tmVm, tmVt, g));
    VPU_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

constexpr int KEYS_FEW = 48;      // the prompt-token side: 2 x num_max_points rows per image

bool few_keys(const AttnArgs& a, int hd) {
    return a.Sk == KEYS_FEW && (hd == 48 || hd == 64 || hd == 80 || hd == 96 || hd == 128 || hd == 160);
}
bool few_queries(const AttnArgs& a, int hd) {
    if (a.Sq > QT || a.heads % 2 != 0) return false;
    if (hd == 48 || hd == 64) return a.Sk % 112 == 0 || a.Sk % 64 == 0;
    return hd == 80 && a.Sk % 64 == 0;
}

}  // namespace

bool dma_attention_tc_supported(const AttnArgs& a, int head_dim) {
    const bool aligned = a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.qoff % 8 == 0 && a.koff % 8 == 0 && a.voff % 8 == 0 &&
                         ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) == 0 &&
                         a.ldo % 16 == 0 && (reinterpret_cast<uintptr_t>(a.o) & 31) == 0;      // 32-byte epilogue stores
    return aligned && a.qmap.mode == 0 && a.kmap.mode == 0 && a.qmap.per_prob == a.Sq && a.kmap.per_prob == a.Sk &&
           (few_keys(a, head_dim) || few_queries(a, head_dim));
}

int dma_attention_tc_launch(const AttnArgs& a, int head_dim, cudaStream_t stream) {
    if (few_keys(a, head_dim)) {
        switch (head_dim) {
            case 48: return launch_cfg<0, 3, 48, false>(a, stream);
            case 64: return launch_cfg<1, 0, 48, false>(a, stream);
            case 80: return launch_cfg<1, 1, 48, false>(a, stream);
            case 96: return launch_cfg<1, 2, 48, false>(a, stream);
            case 128: return launch_cfg<2, 0, 48, false>(a, stream);
            case 160: return launch_cfg<2, 2, 48, false>(a, stream);
        }
    }
    if (few_queries(a, head_dim)) {
        const bool kb112 = a.Sk % 112 == 0;
        switch (head_dim) {
            case 48: return kb112 ? launch_cfg<0, 3, 112, true>(a, stream) : launch_cfg<0, 3, 64, true>(a, stream);
            case 64: return kb112 ? launch_cfg<1, 0, 112, true>(a, stream) : launch_cfg<1, 0, 64, true>(a, stream);
            case 80: return launch_cfg<1, 1, 64, true>(a, stream);
        }
    }
    VPU_REQUIRE(false, "DMA attention: unsupported shape Sq=%d Sk=%d head_dim=%d heads=%d", a.Sq, a.Sk, head_dim, a.heads);
    return 1;
}

}  // namespace vpu
